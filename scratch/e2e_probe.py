"""Where the pageable end-to-end time of linna_lnp_host goes (GPU box): staging-thread sweep, raw host memcpy rate,
driver-staged pageable cudaMemcpy, pinned H2D."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

D = bench.Dist()
n = 100000
u = [bench.synthetic.walkers(n, 30, scale=0.3, seed=100 + b) for b in range(4)]
# raw host copy rates
dst = np.empty_like(u[0])
for _ in range(3):
    np.copyto(dst, u[0])
t0 = time.perf_counter()
for i in range(20):
    np.copyto(dst, u[i % 4])
dt = (time.perf_counter() - t0) / 20
print("numpy copy 12 MB, one thread: %.3f ms = %.1f GB/s" % (dt * 1e3, 12e6 / dt / 1e9))
pin = torch.empty(n, 30, dtype=torch.float32).pin_memory()
pn = pin.numpy()
for _ in range(3):
    np.copyto(pn, u[0])
t0 = time.perf_counter()
for i in range(20):
    np.copyto(pn, u[i % 4])
dt = (time.perf_counter() - t0) / 20
print("numpy copy pageable -> pinned, one thread: %.3f ms = %.1f GB/s" % (dt * 1e3, 12e6 / dt / 1e9))
dev = torch.empty(n, 30, dtype=torch.float32, device="cuda")
for src, name in ((pin, "pinned"), (torch.from_numpy(u[0]), "pageable")):
    for _ in range(3):
        dev.copy_(src)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(20):
        dev.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 20
    print("H2D 12 MB from %s: %.3f ms = %.1f GB/s" % (name, dt * 1e3, 12e6 / dt / 1e9))
for thr in sys.argv[1:] or ["-"]:   # (the staging pool and its LINNA_STAGE_THREADS switch are gone: one helper thread now)
    p, eng, data = bench.make_engine(D, "c3")
    for w in range(3):
        eng.lnp(u[w % 4])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s in range(20):
        eng.lnp(u[s % 4])
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 20
    print("pageable e2e %.3f ms per 1e5 walkers = %.1f M evals/s" % (dt * 1e3, n / dt / 1e6))
    eng.close()
