#!/usr/bin/env python
"""dx_dexqv_host / dx_undexqv_host wall time on pinned buffers: serial copies against the pipeline with
different window counts.  usage: e2e_probe.py [GB]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dextractor_b200 as dx
from dextractor_b200 import synth_torch
size = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
dev = torch.device("cuda", 0)
ctx = dx.Context(0)
text, nent, npos = synth_torch.make_quiva_device(100, int(size * 1e9), dev)
torch.cuda.synchronize()
U = text.numel()
h_text = torch.empty(U, dtype=torch.uint8).pin_memory(); h_text.copy_(text)
h_enc = torch.empty(U // 2 + (1 << 20), dtype=torch.uint8).pin_memory()
h_back = torch.empty(U + 4096, dtype=torch.uint8).pin_memory()
torch.cuda.synchronize()
def best(fn, reps=4):
    b = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); r = fn(); b = min(b, time.perf_counter() - t0)
    return b, r
te, n1 = best(lambda: ctx.dexqv_host_ptr(h_text.data_ptr(), U, False, h_enc.data_ptr(), h_enc.numel()))
print(f"dexqv_host   {te*1e3:7.1f} ms  ({U} -> {n1} bytes)")
for name, routes in (("serial", {"serial_io": 1}), ("pipe K=2", {"pipe_chunk": n1 // 2}), ("pipe K=4", {"pipe_chunk": n1 // 4}),
                     ("pipe K=8", {"pipe_chunk": n1 // 8}), ("pipe K=12", {"pipe_chunk": n1 // 12}),
                     ("pipe K=16", {"pipe_chunk": n1 // 16}), ("default", {})):
    ctx.route("default")
    for k, v in routes.items():
        ctx.route(k, v)
    h_back.zero_()
    td, n2 = best(lambda: ctx.undexqv_host_ptr(h_enc.data_ptr(), n1, False, h_back.data_ptr(), h_back.numel()))
    ok = n2 == U and bool(torch.equal(h_back[:U], h_text))
    print(f"undexqv_host {td*1e3:7.1f} ms  {name:20s} ok={ok}  e2e {2*U/(te+td)/1e9:.1f} GB/s", flush=True)
    ctx.profile(True); ctx.profile_report()
    t0 = time.perf_counter()
    ctx.undexqv_host_ptr(h_enc.data_ptr(), n1, False, h_back.data_ptr(), h_back.numel())
    dt = time.perf_counter() - t0
    rep = ctx.profile_report(); ctx.profile(False)
    print(f"    profiled call {dt*1e3:.1f} ms; kernels: " + "  ".join(f"{k}={v[0]}x{v[1]/v[0]:.3f}" for k, v in sorted(rep.items(), key=lambda kv: -kv[1][1])[:5]), flush=True)
ctx.close()
