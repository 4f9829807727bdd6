"""Where the time of a tensor-core training step goes (GPU box): per-layer cycle stamps (LINNA_TG_DEBUG) and the
step time with / without split-K."""
import os
import sys
import time

os.environ.setdefault("LINNA_TG_DEBUG", "1")
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

D = bench.Dist()
r = bench.measure_train(D, 500, 30, 5, want_e2e=False)
print("step ms", r["ms_per_step"], r["kernel_path"], "launches/step", r["launches"] / 30)
from linna_b200 import engine  # noqa: E402
from linna_b200.train import FusedTrainer  # noqa: E402
import linna.nn as N  # noqa: E402
p, xt, yt, loss_fn, X, Y = bench.train_setup(D)
torch.manual_seed(1234)
model = N.ChtoModelv2(30, 500, None)
tr = FusedTrainer(model, xt, yt, loss_fn.auxileryfunction, 500, device_index=0, lr=1e-3)
cmd = tr.chisq_md(X, Y)
for _ in range(3):
    tr.step(X[:500], Y[:500], cmd[:500])
torch.cuda.synchronize()
c = tr.engine.tg_counters()
print("step  setup  wait_prev  first_seg  k_loop  epilogue  teardown  (cycles of CTA 0)")
for i, row in enumerate(c):
    if row[0] == 0:
        continue
    if i == 40:
        print("weight gradients + AdamW (CTA 0):")
    print("%3d  %6d  %8d  %8d  %7d  %7d  %6d   total %7d" % (i, row[1] - row[0], row[2] - row[1], row[3] - row[2], row[4] - row[2],
                                                               row[5] - row[4], row[6] - row[5], row[6] - row[0]))
# host enqueue cost of one step
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    tr.step(X[:500], Y[:500], cmd[:500])
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host enqueue per step %.1f us, with drain %.1f us" % ((t1 - t0) / 50 * 1e6, (t2 - t0) / 50 * 1e6))
