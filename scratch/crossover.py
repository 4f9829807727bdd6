import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from linna_b200 import engine, synthetic
shape = (30, 500) if len(sys.argv) < 2 or sys.argv[1] == "c3" else (50, 1500)
p = synthetic.make_problem(*shape, seed=0)
e = engine.engine_from_problem(p, with_likelihood=False)
m0 = e.predict(np.asarray(p.theta0, np.float32)[None, :], engine.LINNA_OUT_M)[0]
p.set_data_from_prediction(m0)
e.set_likelihood(p.priors, np.asarray(p.data, np.float32), p.inv_cov, 1.0)
def t(call, u):
    for _ in range(3): call(u)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): call(u)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / 10
for n in (64, 256, 512, 1024, 2048, 4096, 8192):
    u = torch.from_numpy(synthetic.walkers(n, shape[0], scale=0.3, seed=1)).cuda()
    r = []
    for path in ("ffma", "tc"):
        e.set_path(path)
        r.append((t(e.lnp, u), t(e.lnp_grad, u)))
    print("n %5d  lnp ffma %.3f tc %.3f ms | grad ffma %.3f tc %.3f ms" % (n, r[0][0], r[1][0], r[0][1], r[1][1]), flush=True)
