"""Latency of the small-batch cluster kernel against the tiled FFMA kernel (device time per call, CUDA events)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linna_b200 import engine, synthetic

def make(n_in, n_out):
    p = synthetic.make_problem(n_in, n_out, seed=0)
    e = engine.engine_from_problem(p, with_likelihood=False)
    m0 = e.predict(np.asarray(p.theta0, np.float32)[None, :], engine.LINNA_OUT_M)[0]
    p.set_data_from_prediction(m0)
    e.set_likelihood(p.priors, np.asarray(p.data, np.float32), p.inv_cov, 1.0)
    return p, e

def t(fn, iters=200):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3

for shape in ((33, 33), (30, 500), (50, 1500)):
    p, e = make(*shape)
    for n in (4, 8, 16, 64, 128, 255, 512, 1184):
        u = torch.from_numpy(synthetic.walkers(n, p.n_in, scale=0.3, seed=1)).cuda()
        row = []
        for path in ("cluster", "ffma"):
            e.set_path(path)
            try:
                row.append("%s lnp %.1f us grad %.1f us" % (path, t(lambda: e.lnp(u)), t(lambda: e.lnp_grad(u))))
            except Exception as ex:
                row.append("%s: %s" % (path, str(ex)[:80]))
        print(shape, "n=%d" % n, " | ".join(row), flush=True)
    e.close()
