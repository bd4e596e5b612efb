#!/usr/bin/env python
"""dexqv_mg (one .quiva on N GPUs, C + NCCL) against the reference dexqv.
usage: mg_check.py N [big_gb] [ref_max_gb]
  * small crafted files (entries longer than a shard, '@'-starting QV lines, a file whose run
    characters appear late): dexqv_mg -gN output == reference dexqv output, byte for byte;
  * one big file of big_gb GB written to /dev/shm: dexqv_mg -g1 ... -gN must give the same sha256;
    up to ref_max_gb GB it is compared with the reference tool as well; -c (every rank decodes its
    shard again) is on.
Prints one JSON object per case."""
import hashlib
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

from dextractor_b200 import synth
from oracle import orc

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2
BIG = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
REFMAX = float(sys.argv[3]) if len(sys.argv) > 3 else 4.0
TOOL = os.path.join(ROOT, "tools", "bin", "dexqv_mg")
WORK = "/dev/shm/dxmg"
os.makedirs(WORK, exist_ok=True)


def sha_file(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        while True:
            b = f.read(1 << 24)
            if not b:
                break
            h.update(b)
    return h.hexdigest()


def run_mg(path, n, check=True):
    out = path[:-6] + ".dexqv"
    if os.path.exists(out):
        os.unlink(out)
    t0 = time.perf_counter()
    p = subprocess.run([TOOL, "-vk" + ("c" if check else ""), f"-g{n}", path], capture_output=True, text=True)
    dt = time.perf_counter() - t0
    if p.returncode != 0:
        raise SystemExit(f"dexqv_mg -g{n} failed rc={p.returncode}: {p.stderr[-2000:]}")
    return out, dt, p.stderr


def ref_dexqv(path):
    """reference tool on a copy (it writes next to its input)"""
    if orc.have_ref():
        cp = path[:-6] + "_ref.quiva"
        subprocess.check_call(["cp", path, cp])
        t0 = time.perf_counter()
        subprocess.check_call([os.path.join(orc.REF_DIR, "dexqv"), "-k", cp])
        dt = time.perf_counter() - t0
        os.unlink(cp)
        return cp[:-6] + ".dexqv", dt, "reference"
    data = open(path, "rb").read()
    t0 = time.perf_counter()
    enc = orc.dexqv(data)
    dt = time.perf_counter() - t0
    out = path[:-6] + "_ref.dexqv"
    open(out, "wb").write(enc)
    return out, dt, "port"


def at_sign_hook(i, streams):
    for k in (0, 2, 3, 4):
        streams[k][0] = ord("@")
    if len(streams[2]) > 12:
        streams[2][:12] = np.frombuffer(b"@m/1/0_9 RQ=", dtype=np.uint8)


def late_n_hook(i, s):
    if i < 9:
        s[1][:] = ord("a")


small = {
    "three_long_entries": synth.make_quiva(1, [60000, 500, 45000]),
    "at_sign_lines": synth.make_quiva(2, list(np.random.default_rng(2).integers(200, 9000, size=90)),
                                      stream_hook=at_sign_hook),
    "late_run_chars": synth.make_quiva(3, [9000] * 40, stream_hook=late_n_hook),
    "short_file_no_subchar": synth.make_quiva(4, [700] * 30),
    "mid_8mb": synth.make_quiva(5, synth.lengths_for_bytes(np.random.default_rng(5), 8_000_000, 5.0)),
}
ok_all = True
for name, text in small.items():
    path = os.path.join(WORK, name + ".quiva")
    open(path, "wb").write(text)
    want = orc.ref_tool("dexqv", text)[0] if orc.have_ref() else orc.dexqv(text)
    res = {"case": name, "bytes": len(text), "gpus": []}
    for n in sorted({1, 2, N} if N >= 2 else {1}):
        out, dt, _ = run_mg(path, n, check=False)      # (crafted tags do not round-trip to the identity)
        same = open(out, "rb").read() == want
        res["gpus"].append({"n": n, "equal_to_reference": same})
        ok_all &= same
    print(json.dumps(res), flush=True)

# ---- the big file: pieces generated on GPU 0, appended to one file --------------------------------
import torch

from dextractor_b200 import synth_torch
big = os.path.join(WORK, "big.quiva")
dev = torch.device("cuda", 0)
piece = min(BIG, 4.0)
t0 = time.perf_counter()
with open(big, "wb") as f:
    done, well, k = 0.0, 0, 0
    while done < BIG - 1e-9:
        g = min(piece, BIG - done)
        t, nent, npos = synth_torch.make_quiva_device(1000 + k, int(g * 1e9), dev, well_base=well)
        torch.cuda.synchronize()
        f.write(t.cpu().numpy().tobytes())
        well += 40 * nent + 1
        done += g; k += 1
        del t
torch.cuda.empty_cache()
size = os.path.getsize(big)
gen_s = time.perf_counter() - t0
res = {"case": "big", "bytes": size, "generate_s": round(gen_s, 1), "runs": []}
shas = set()
ns = [n for n in (1, 2, 4, 8) if n <= N]
if size > 40e9:
    ns = [n for n in ns if n >= 2]          # a 64 GB file does not fit one GPU with its scratch
for n in ns:
    # -c holds text + image + decoded text + decode scratch on the device: up to ~20 GB shards
    out, dt, err = run_mg(big, n, check=(size / n < 20e9))
    s = sha_file(out)
    shas.add(s)
    phases = [l.strip() for l in err.splitlines() if "max over ranks" in l]
    res["runs"].append({"n": n, "wall_s": round(dt, 2), "sha256": s, "out_bytes": os.path.getsize(out),
                        "GBps_wall": round(size / dt / 1e9, 2), "phases": phases[0] if phases else ""})
res["all_equal"] = len(shas) == 1
if size <= REFMAX * 1e9:
    rout, rdt, kind = ref_dexqv(big)
    res["reference"] = {"kind": kind, "wall_s": round(rdt, 1), "sha256": sha_file(rout)}
    res["equal_to_reference"] = res["reference"]["sha256"] in shas and len(shas) == 1
    ok_all &= res["equal_to_reference"]
ok_all &= res["all_equal"]
print(json.dumps(res), flush=True)
subprocess.call(["rm", "-rf", WORK])
print("MG_CHECK", "OK" if ok_all else "FAILED")
sys.exit(0 if ok_all else 1)
