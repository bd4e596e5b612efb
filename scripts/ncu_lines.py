#!/usr/bin/env python
"""Instructions executed / stall samples per CUDA source line from an .ncu-rep captured with
--import-source on (ncu -i rep --page source --csv --print-source cuda,sass)."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True, errors="replace").stdout
rows = list(csv.reader(raw.splitlines()))
cur_file, cur_line, cur_src = "?", "?", ""
inst = collections.Counter(); tinst = collections.Counter(); samp = collections.Counter()
src = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r[0] in ("Function Name", "Line No"):
        continue
    if r[0] != "":
        cur_line = r[0]; cur_src = ",".join(r[1:]).strip()[:90]
        src[(cur_file, cur_line)] = cur_src
        continue
    if len(r) > 8 and r[2].startswith("0x"):
        try:
            key = (cur_file, cur_line)
            inst[key] += int(r[7]); tinst[key] += int(r[8]); samp[key] += int(r[6])
        except ValueError:
            pass
tot = sum(inst.values()); ts = sum(samp.values())
print(f"# total warp instructions {tot:,}  thread instructions {sum(tinst.values()):,}  samples {ts:,}")
for key, v in inst.most_common(top):
    print(f"{key[0]:22s}:{key[1]:>4s} inst {v/tot*100:5.1f}%  lanes {tinst[key]/max(v,1):5.1f}  "
          f"samples {samp[key]/max(ts,1)*100:5.1f}%  | {src.get(key,'')}")
