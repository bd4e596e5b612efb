#!/usr/bin/env python
"""undexqv with known entry offsets under each decoder route, timed per kernel.
usage: dec_probe.py [quiva_gb] [routes...]   (route = name=value[,name=value])"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import dextractor_b200 as dx
from dextractor_b200 import synth_torch

size = float(sys.argv[1]) if len(sys.argv) > 1 else 0.5
routes = sys.argv[2:] or ["decoder=5", "decoder=6", "default=0"]
dev = torch.device("cuda", 0)
ctx = dx.Context(0)
text, nent, npos = synth_torch.make_quiva_device(101, int(size * 1e9), dev)
torch.cuda.synchronize()
U = text.numel()
enc = torch.empty(U // 2 + (1 << 20), dtype=torch.uint8, device=dev)
back = torch.empty(U + 4096, dtype=torch.uint8, device=dev)
st = ctx.qv_scan_dev(text.data_ptr(), U, None)
cd = dx.lib.make_coding(st, False)
prefix = bytes(text[:200].cpu().numpy().tobytes()); prefix = prefix[: prefix.index(b"/", 1)]
hdr = b"\xaa\x55" + dx.lib.write_coding(cd, prefix)
ctx.h2d(enc.data_ptr(), hdr)
body, _, offs = ctx.qv_encode_dev(text.data_ptr(), U, cd, False, 0, enc.data_ptr() + len(hdr),
                                  enc.numel() - len(hdr), want_offsets=nent)
n = len(hdr) + body
offs = offs + len(hdr)
print(f"U={U} C={n} entries={nent}")
for r in routes:
    ctx.route("default")
    for kv in r.split(","):
        k, v = kv.split("=")
        if k != "default":
            ctx.route(k, int(v))
    for known in (True, False):
        back.zero_()
        for _ in range(2):
            m = ctx.undexqv_dev(enc.data_ptr(), n, False, back.data_ptr(), back.numel(),
                                entry_off=offs if known else None)
        ok = m == U and bool(torch.equal(back[:U], text))
        ctx.profile(True); ctx.profile_report()
        m = ctx.undexqv_dev(enc.data_ptr(), n, False, back.data_ptr(), back.numel(),
                            entry_off=offs if known else None)
        rep = ctx.profile_report(); ctx.profile(False)
        tot = sum(v[1] for v in rep.values())
        top = sorted(rep.items(), key=lambda kv: -kv[1][1])[:4]
        print(f"{r:28s} known={known!s:5s} ok={ok} kernels {tot:8.3f} ms  " +
              "  ".join(f"{k}={v[1]:.3f}" for k, v in top), flush=True)
ctx.close()
