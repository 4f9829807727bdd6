"""One C3 lnP launch and one lnP+grad launch of the tensor-core kernel at 1e5 walkers (for ncu)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
D = bench.Dist()
p, eng, data = bench.make_engine(D, "c3", "auto")
u = torch.from_numpy(bench.synthetic.walkers(100000, p.n_in, scale=0.3, seed=1)).cuda()
for _ in range(3):
    eng.lnp(u); eng.lnp_grad(u)
torch.cuda.synchronize()
print("ok", eng.last_kernel())
