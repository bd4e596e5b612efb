#!/usr/bin/env python
"""Executed warp-instructions and stall samples per source line: joins the SASS page of an .ncu-rep
with nvdisasm -g line info of the object the kernel came from (same build!).
usage: ncu_lines2.py rep object.o kernel-regex [launch-skip] [top]"""
import collections, csv, os, re, subprocess, sys, tempfile
rep, obj, kre = sys.argv[1], os.path.abspath(sys.argv[2]), sys.argv[3]
skip = sys.argv[4] if len(sys.argv) > 4 else "0"
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True,
                     errors="replace").stdout
rows = list(csv.reader(raw.splitlines()))
kname = rows[0][1]
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
body = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    if r[ix["# Samples"]] == "# Samples": break
    body.append(r)
# mangled name fragment to find the function in the disassembly
frag = re.findall(r"(k_\w+)", kname)[0]
tmpl = re.search(r"<\((?:int|bool)\)(\d+)>", kname)
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=d, capture_output=True)
cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout.splitlines()
lines, cur, infn = [], ("?", 0), False
for l in dis:
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
    if m:
        name = m.group(1)
        infn = frag in name and (tmpl is None or ("ILi%sE" % tmpl.group(1)) in name or ("ILb%sE" % tmpl.group(1)) in name)
        continue
    if not infn: continue
    m = re.match(r'\s*//## File "(.*)", line (\d+)(.*)', l)
    if m:
        if "inlined at" in l and cur != ("?", 0): pass
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
print(kname, "sass rows", len(body), "disasm rows", len(lines))
n = min(len(body), len(lines))
exe = collections.Counter(); smp = collections.Counter()
for i in range(n):
    exe[lines[i]] += int(body[i][ix["Instructions Executed"]] or 0)
    smp[lines[i]] += int(body[i][ix["# Samples"]] or 0)
te, ts = sum(exe.values()), sum(smp.values())
src = {}
for (f, ln), v in sorted(exe.items(), key=lambda kv: -kv[1])[:top]:
    if f not in src:
        for base in ("/root/repo/dextractor_b200/csrc/", "/root/repo/include/"):
            if os.path.exists(base + f): src[f] = open(base + f).read().splitlines()
    text = src[f][ln-1].strip()[:80] if f in src and ln-1 < len(src[f]) else ""
    print(f"{f}:{ln:<5d} exe {v*100/te:5.1f}%  samples {smp[(f,ln)]*100/max(ts,1):5.1f}%   {text}")
