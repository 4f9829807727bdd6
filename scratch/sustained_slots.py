"""Sustained C3 lnP throughput (2.5 s loops, power-capped regime) with two and with one walker pair per CTA pair."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

D = bench.Dist()
p, eng, data = bench.make_engine(D, "c3")
for slots in ("2", "1", "2", "1"):
    os.environ["LINNA_TC_SLOTS"] = slots
    s = bench.measure_sustained(D, p, eng, 100000, "lnp", seconds=2.5)
    print("slots", slots, "ms/step %.4f  %.1f M evals/s" % (s["ms_per_step"], 100000 / s["ms_per_step"] / 1e3), s["clocks"], flush=True)
