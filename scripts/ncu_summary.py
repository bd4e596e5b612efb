#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel name."""
import collections
import csv
import io
import sys

lines = open(sys.argv[1], errors="replace").read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r["Kernel Name"].split("(")[0].replace("void <unnamed>::", "").replace("void ", "")
    agg[name][0] += 1
    agg[name][1] += float(r["Metric Value"]) / 1e6
tot = sum(v[1] for v in agg.values())
print(f"# {sys.argv[1]}: {len(rows)} launches, {tot:.3f} ms of device time (cold-cache, serialised)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:34s} calls={v[0]:4d} total_ms={v[1]:10.3f} avg_ms={v[1]/v[0]:9.4f} share={v[1]/tot*100:5.1f}%")
