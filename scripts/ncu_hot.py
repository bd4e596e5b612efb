#!/usr/bin/env python
"""Hottest SASS lines of one kernel in an .ncu-rep: samples, executions, avg active threads.
usage: ncu_hot.py rep kernel-regex [launch-skip] [top]"""
import csv, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True,
                     errors="replace").stdout
rows = list(csv.reader(raw.splitlines()))
print(rows[0][1] if len(rows[0]) > 1 else rows[0])
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
body = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    if r[ix["# Samples"]] == "# Samples":      # a second view of the same kernel follows
        break
    body.append(r)
tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
texe = sum(int(r[ix["Instructions Executed"]] or 0) for r in body)
tthr = sum(int(r[ix["Thread Instructions Executed"]] or 0) for r in body)
print(f"samples {tot}  warp-inst {texe}  thread-inst {tthr}  avg active {tthr/max(texe,1):.1f}")
stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
st = {h: sum(int(r[ix[h]] or 0) for r in body) for h in stall}
print(" ".join(f"{h[6:]}={v*100/max(tot,1):.1f}%" for h, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
order = sorted(range(len(body)), key=lambda i: -int(body[i][ix["# Samples"]] or 0))[:top]
for i in sorted(order):
    r = body[i]
    s = int(r[ix["# Samples"]] or 0)
    dom = max(stall, key=lambda h: int(r[ix[h]] or 0))
    print(f"{i:5d} {s*100/max(tot,1):5.1f}% exe={r[ix['Instructions Executed']]:>9s} thr={r[ix['Avg. Threads Executed']]:>5s} "
          f"{dom[6:]:>10s}  {r[ix['Source']].strip()[:90]}")
