"""Device time of the tensor-core lnP launch against the number of walkers (`device`: the staircase in rounds; one or
two walker pairs per CTA pair), or the pageable end-to-end time of one 10^5-walker call (MODE=lnp|grad)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

D = bench.Dist()
p, eng, data = bench.make_engine(D, "c3")
mode = os.environ.get("MODE", "lnp")
call = eng.lnp_grad if mode == "grad" else eng.lnp
u_all = torch.from_numpy(bench.synthetic.walkers(100000, 30, scale=0.3, seed=1)).cuda()
if len(sys.argv) > 1 and sys.argv[1] == "device":
    for n in (9472, 18944, 28416, 37888, 56832, 75776, 94720, 100000):
        u = u_all[:n].contiguous()
        for _ in range(3):
            call(u)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            call(u)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print("%s n=%6d  %.4f ms  %.1f M evals/s" % (mode, n, ms, n / ms / 1e3))
else:
    n = 100000
    u = [bench.synthetic.walkers(n, 30, scale=0.3, seed=100 + b) for b in range(4)]
    for w in range(3):
        call(u[w % 4])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s in range(20):
        call(u[s % 4])
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 20
    print("%s: pageable e2e %.3f ms per 1e5 walkers = %.1f M evals/s" % (mode, dt * 1e3, n / dt / 1e6))
