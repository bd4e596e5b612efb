#!/usr/bin/env python
"""Warp-stall sample totals of an .ncu-rep (source page), as percentages."""
import collections, csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True,
                     text=True, errors="replace").stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
st = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter()
for r in rows[2:]:
    if len(r) < len(hdr): continue
    for h in st:
        try: tot[h] += int(r[idx[h]])
        except ValueError: pass
s = sum(tot.values())
print(" ".join(f"{h[6:]}={v/s*100:.1f}%" for h, v in tot.most_common(9)))
