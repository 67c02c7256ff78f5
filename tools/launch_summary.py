#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (`--metrics gpu__time_duration.sum --csv`) as a markdown table.
Usage: python tools/launch_summary.py <launches.csv> [top_n]"""
import collections
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 20
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
tot = collections.OrderedDict()
for r in rows[1:]:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    unit = r[ix["Metric Unit"]]
    v = float(r[ix["Metric Value"]].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[unit]
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
    n, t = tot.get(name, (0, 0.0))
    tot[name] = (n + 1, t + v)
total = sum(t for _, t in tot.values())
ours = sum(t for k, (_, t) in tot.items() if "gs2m::" in k)
print("| kernel | launches | total ms | share |\n|---|---|---|---|")
for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:top]:
    print("| `%s` | %d | %.3f | %.1f %% |" % (k[:70], n, t, 100 * t / total))
print("\n%d launches, %.3f ms of device time; kernels of this library (`gs2m::`): %.1f %%" %
      (sum(n for n, _ in tot.values()), total, 100 * ours / total))
