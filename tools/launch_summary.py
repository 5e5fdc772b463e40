#!/usr/bin/env python
"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list (cold-cache, serialised times:
compare shares, not absolutes).  Usage: python tools/launch_summary.py <launches.csv> [--tail-from <kernel substring>]"""
import collections
import csv
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
names = [(r["Kernel Name"], float(r["Metric Value"]) / 1e3) for r in rows]
if len(sys.argv) > 3 and sys.argv[2] == "--tail-from":
    idx = [i for i, (n, _) in enumerate(names) if sys.argv[3] in n]
    tail = names[idx[-1]:]
    tot = sum(v for _, v in tail)
    print(f"# the last occurrence of '{sys.argv[3]}' to the end of the run: one unit of work, {len(tail)} launches, {tot:.1f} us")
    for n, v in tail:
        print("%-100s %9.2f us  %5.1f %%" % (n[:100], v, 100 * v / tot))
else:
    agg = collections.OrderedDict()
    for n, v in names:
        a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    for n, (c, t) in agg.items():
        print("%-100s n=%4d avg_us=%9.2f share=%5.3f" % (n[:100], c, t / c, t / tot))
    print("sum of all launches us %.1f" % tot)
