"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (the files under profiles/)."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r[4].split("(")[0][:70]
    t = tot[name]
    t[0] += 1
    t[1] += float(r[14]) / 1e3
all_us = sum(v[1] for v in tot.values())
print("# " + (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]))
print("kernel,launches,total_us,share_pct")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%s,%d,%.1f,%.2f" % (k, v[0], v[1], 100 * v[1] / all_us))
