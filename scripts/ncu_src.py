#!/usr/bin/env python
"""Instruction counts per CUDA-C source line of one kernel in an .ncu-rep (needs -lineinfo and --import-source on).
usage: ncu_src.py rep kernel-regex [launch-skip] [top]"""
import csv, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kre,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True,
                     errors="replace").stdout
rows = list(csv.reader(raw.splitlines()))
# find sections: header rows start with "Address" (sass) or "#"/"Line" (source)
secs = []
for i, r in enumerate(rows):
    if r and r[0] in ("Address", "#", "Line", "File Path"):
        secs.append(i)
print([rows[i][:3] for i in secs][:6])
for si in secs:
    hdr = rows[si]
    if hdr[0] == "Address":
        continue
    ix = {h: k for k, h in enumerate(hdr)}
    body = []
    for r in rows[si+1:]:
        if len(r) < len(hdr) or r[0] in ("Address", "#", "Line", "File Path"):
            break
        body.append(r)
    key = "Instructions Executed"
    tot = sum(int(r[ix[key]] or 0) for r in body)
    print("total warp-inst", tot)
    order = sorted(range(len(body)), key=lambda i: -int(body[i][ix[key]] or 0))[:top]
    for i in sorted(order):
        r = body[i]
        print(f"{r[0]:>6s} {int(r[ix[key]] or 0)*100/max(tot,1):5.1f}%  {r[ix['Source']][:110]}")
