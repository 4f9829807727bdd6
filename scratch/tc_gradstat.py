import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from linna_b200 import engine, synthetic, arch
from oracle.oracle import Oracle
p = synthetic.make_problem(30, 500, seed=0)
e = engine.engine_from_problem(p, with_likelihood=False)
m0 = e.predict(np.asarray(p.theta0, np.float32)[None, :], engine.LINNA_OUT_M)[0]
p.set_data_from_prediction(m0)
e.set_likelihood(p.priors, np.asarray(p.data, np.float32), p.inv_cov, 1.0)
n = 100000
u = synthetic.walkers(n, 30, scale=0.3, seed=1)
ud = torch.from_numpy(u).cuda()
e.set_path("tc"); la, ga = e.lnp_grad(ud); la2, ga2 = e.lnp_grad(ud)
print("deterministic:", bool(torch.equal(ga, ga2)))
e.set_path("ffma"); lf, gf = e.lnp_grad(ud)
ga, gf = ga.cpu().numpy().astype(np.float64), gf.cpu().numpy().astype(np.float64)
rel = np.max(np.abs(ga - gf), axis=1) / np.max(np.abs(gf), axis=1)
print("quantiles", np.quantile(rel, [0.5, 0.9, 0.99, 0.999, 0.9999, 1.0]))
bad = np.where(rel > 1e-4)[0]
print("rows > 1e-4:", len(bad), bad[:20], "mod 256:", bad[:20] % 256)
if len(bad):
    idx = bad[:8]
    ref = Oracle(p, arch).lnp(u[idx], np.float64, grad=True)["grad"]
    for k, i in enumerate(idx):
        d = np.max(np.abs(ref[k]))
        print(i, "tc vs f64 %.3e  ffma vs f64 %.3e" % (np.max(np.abs(ga[i] - ref[k])) / d, np.max(np.abs(gf[i] - ref[k])) / d))
