import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from linna_b200 import engine, synthetic
from linna.HMCSampler import HMCSampler
shape = (50, 1500)
p = synthetic.make_problem(*shape, seed=0)
e = engine.engine_from_problem(p, with_likelihood=False)
m0 = e.predict(np.asarray(p.theta0, np.float32)[None, :], engine.LINNA_OUT_M)[0]
p.set_data_from_prediction(m0)
e.set_likelihood(p.priors, np.asarray(p.data, np.float32), p.inv_cov, 1.0)
class LP:
    def value_and_grad(self, x): return e.lnp_grad(x.contiguous())
    def __call__(self, x): return e.lnp(x)
for C in (1250, 10000):
    x0 = torch.from_numpy(synthetic.walkers(C, shape[0], scale=0.05, seed=2)).cuda()
    s = HMCSampler(LP(), x0, torch.ones(shape[0]), device="cuda")
    s.sample_chains(3, 5, 0.01)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    xs, ls, acc = s.sample_chains(20, 5, 0.01)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
    u = x0.contiguous(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): e.lnp_grad(u)
    torch.cuda.synchronize(); dg = (time.perf_counter() - t0) / 20
    print("C4 HMC %5d chains, 5 leapfrog steps: %.3f ms per sample of every chain (%.3g chain-samples/s), gradient launch %.3f ms x 5 = %.0f%% of it, acceptance %.2f" % (C, dt * 1e3, C / dt, dg * 1e3, 500 * dg / dt, acc))
