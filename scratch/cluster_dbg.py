"""Per-step cycle counters of the cluster kernel (LINNA_CLUSTER_DEBUG=1)."""
import sys, os
os.environ["LINNA_CLUSTER_DEBUG"] = "1"
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linna_b200 import engine, synthetic
n_in, n_out = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (33, 33)
p = synthetic.make_problem(n_in, n_out, seed=0)
e = engine.engine_from_problem(p, with_likelihood=False)
m0 = e.predict(np.asarray(p.theta0, np.float32)[None, :], engine.LINNA_OUT_M)[0]
p.set_data_from_prediction(m0)
e.set_likelihood(p.priors, np.asarray(p.data, np.float32), p.inv_cov, 1.0)
e.set_path("cluster")
u = torch.from_numpy(synthetic.walkers(4, p.n_in, scale=0.3, seed=1)).cuda()
for _ in range(10): e.lnp(u)
torch.cuda.synchronize()
engine.cluster_counters()
iters = 50
for _ in range(iters): e.lnp(u)
torch.cuda.synchronize()
c = engine.cluster_counters() / iters
print("step  kloop  reduce  epilogue  barrier")
for i in range(16):
    if c[i].sum() > 0: print(i, c[i].round(0), "epilogue: after partial sums / after bias / after first LDS", c[64 + i][:3].round(0))
print("steps total cycles", c[:64].sum(), "per column", c[:64].sum(0))
print("outside the step loop: set-up + ring prefill, first cluster barrier, all tiles, last barrier:", c[64 + 60].round(0))
