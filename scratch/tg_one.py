"""A few C5 training steps on the tensor-core training kernels (for ncu)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
D = bench.Dist()
r = bench.measure_train(D, 500, int(os.environ.get("TG_ONE_STEPS", "4")), 3, want_e2e=False)
print("ok", r["kernel_path"], r["ms_per_step"])
