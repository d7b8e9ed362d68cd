"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py: launches per step, mean
duration and share per kernel, between the first and the last adam_flat launch of the capture (whole steps).

    python tools/launch_summary.py gpurun_out/launches.csv > profiles/rN_launch_summary.txt
"""
import csv
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        rows.append((r["Kernel Name"].split("(")[0], float(r["Metric Value"].replace(",", "")) / 1e3))
adam = [i for i, (k, _) in enumerate(rows) if "adam_flat" in k]
if len(adam) < 2:
    sys.exit("need at least two adam_flat launches in the capture")
sel = rows[adam[0] + 1:adam[-1] + 1]
steps = len(adam) - 1
agg = OrderedDict()
for k, us in sel:
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values())
print("# ncu launch list (gpu__time_duration.sum), %d whole steps between adam_flat launches; cold-cache, serialised: compare SHARES" % steps)
print("# columns: kernel | launches/step | mean us | us/step | share\n")
for k, (n, us) in sorted(agg.items(), key=lambda t: -t[1][1]):
    print("%-74s %6.2f %9.1f %9.1f %5.1f%%" % (k[:74], n / steps, us / n, us / steps, 100 * us / tot))
print("\ntotal per step: %.1f us in %.1f launches" % (tot / steps, len(sel) / steps))
