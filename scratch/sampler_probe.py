import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from linna_b200 import engine, synthetic, sampler
p = synthetic.make_problem(30, 500, seed=0)
e = engine.engine_from_problem(p, with_likelihood=False)
m0 = e.predict(np.asarray(p.theta0, np.float32)[None, :], engine.LINNA_OUT_M)[0]
p.set_data_from_prediction(m0)
e.set_likelihood(p.priors, np.asarray(p.data, np.float32), p.inv_cov, 1.0)
for W in (1024, 16384, 100000):
    es = sampler.EnsembleSampler(W, 30, lambda x: e.lnp(x.contiguous()), seed=1)
    x0 = torch.from_numpy(synthetic.walkers(W, 30, scale=0.1, seed=3)).cuda()
    es.run_mcmc(x0, 5, store=False)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    es.run_mcmc(x0, 20, store=False)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
    u = x0[: W // 2].contiguous()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): e.lnp(u)
    torch.cuda.synchronize(); dl = (time.perf_counter() - t0) / 20
    print("W %6d: %.3f ms per ensemble iteration (2 half-steps), of which likelihood 2 x %.3f ms; acceptance %.2f; %.3g walker-updates/s" % (
        W, dt * 1e3, dl * 1e3, es.acceptance_fraction.mean(), W / dt))
