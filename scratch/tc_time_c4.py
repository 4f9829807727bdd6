import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from linna_b200 import engine, synthetic
p = synthetic.make_problem(50, 1500, seed=0)
e = engine.engine_from_problem(p, with_likelihood=False)
m0 = e.predict(np.asarray(p.theta0, np.float32)[None, :], engine.LINNA_OUT_M)[0]
p.set_data_from_prediction(m0)
e.set_likelihood(p.priors, np.asarray(p.data, np.float32), p.inv_cov, 1.0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
u = torch.from_numpy(synthetic.walkers(n, 50, scale=0.3, seed=1)).cuda()
for _ in range(3): e.lnp_grad(u)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): e.lnp_grad(u)
b.record(); torch.cuda.synchronize()
print("C4 grad n", n, "ms %.3f" % (a.elapsed_time(b) / 5), e.last_kernel())
