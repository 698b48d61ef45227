"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: count, total and mean duration, share.
usage: python tools/summarize_launches.py gpurun_out/launches_X.csv > profiles/X_launches_summary.txt"""
import collections
import csv
import sys

path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
r = list(csv.reader(lines))
h = r[0]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in r[1:]:
    try:
        v = float(row[vi].replace(",", ""))
    except (ValueError, IndexError):
        continue
    n = row[ki]
    n = n.split("(")[0][:110] if ("matgcn" in n or "kernel" in n) else n[:110]
    agg[n][0] += 1
    agg[n][1] += v
    tot += v
print("# %s: %d launches, %.1f us of kernel time (ncu per-launch times are cold-cache and serialised: compare SHARES)"
      % (path.split("/")[-1], sum(a[0] for a in agg.values()), tot / 1e3))
for n, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%9.1f us %5d x %8.1f us  %5.1f%%  %s" % (v / 1e3, c, v / 1e3 / c, 100 * v / tot, n))
