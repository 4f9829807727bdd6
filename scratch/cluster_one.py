"""A few launches of the cluster kernel at the README shape (for ncu)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linna_b200 import engine, synthetic
p = synthetic.make_problem(33, 33, seed=0)
e = engine.engine_from_problem(p, with_likelihood=False)
m0 = e.predict(np.asarray(p.theta0, np.float32)[None, :], engine.LINNA_OUT_M)[0]
p.set_data_from_prediction(m0)
e.set_likelihood(p.priors, np.asarray(p.data, np.float32), p.inv_cov, 1.0)
e.set_path("cluster")
u = torch.from_numpy(synthetic.walkers(4, p.n_in, scale=0.3, seed=1)).cuda()
for _ in range(8): e.lnp(u)
torch.cuda.synchronize()
print("ok", e.last_kernel())
