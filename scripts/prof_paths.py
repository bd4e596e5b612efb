#!/usr/bin/env python
"""Per-path kernel tables (CUDA events inside the library) and wall-clock of every whole-file call.
usage: python scripts/prof_paths.py [size_gb] [fasta_gb]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import dextractor_b200 as dx
from dextractor_b200 import synth_torch

size = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
fsize = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
dev = torch.device("cuda", 0)
ctx = dx.Context(0)


def run(name, fn, reps=3):
    for _ in range(2):
        fn()
    ctx.sync(); torch.cuda.synchronize()
    ctx.profile(True); ctx.profile_report()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    ctx.sync(); torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / reps * 1e3
    prof = ctx.profile_report(); ctx.profile(False)
    ksum = sum(v[1] for v in prof.values()) / reps
    print(f"== {name}: wall {wall:.3f} ms/call, kernels {ksum:.3f} ms/call")
    for k, (c, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
        print(f"     {k:24s} {c // reps:4d} x  {ms / reps:9.4f} ms")
    return wall


if size > 0:
    text, nent, npos = synth_torch.make_quiva_device(101, int(size * 1e9), dev)
    U = text.numel()
    enc = torch.empty(U // 2 + (1 << 20), dtype=torch.uint8, device=dev)
    back = torch.empty(U + 4096, dtype=torch.uint8, device=dev)
    st = {}

    def f_enc():
        st["n"] = ctx.dexqv_dev(text.data_ptr(), U, False, enc.data_ptr(), enc.numel())

    w = run(f"dexqv_dev {U/1e9:.2f} GB", f_enc)
    print(f"   -> {U / w / 1e6:.1f} GB/s, ratio {U / st['n']:.3f}")

    def f_dec():
        st["m"] = ctx.undexqv_dev(enc.data_ptr(), st["n"], False, back.data_ptr(), back.numel())

    w = run("undexqv_dev (entries discovered)", f_dec)
    print(f"   -> {U / w / 1e6:.1f} GB/s")
    assert st["m"] == U and bool(torch.equal(back[:U], text)), "round trip differs"
    del text, enc, back
    torch.cuda.empty_cache()

if fsize > 0:
    for kind, name in ((dx.FASTA, "fasta"), (dx.ARROW, "arrow")):
        if kind == dx.FASTA:
            fa, nfa = synth_torch.make_fasta_device(7, int(fsize * 1e9), dev)
        elif hasattr(synth_torch, "make_arrow_device"):
            fa, nfa = synth_torch.make_arrow_device(7, int(fsize * 1e9), dev)
        else:
            continue
        UF = fa.numel()
        pk = torch.empty(UF // 3 + (1 << 20), dtype=torch.uint8, device=dev)
        un = torch.empty(UF + 4096, dtype=torch.uint8, device=dev)
        st = {}

        def f_pack():
            st["m"] = ctx.dexta_dev(kind, fa.data_ptr(), UF, pk.data_ptr(), pk.numel())

        w = run(f"dexta_dev {name} {UF/1e9:.2f} GB", f_pack)
        print(f"   -> {UF / w / 1e6:.1f} GB/s  ({(UF + st['m']) / w / 1e6:.1f} GB/s of traffic)")

        def f_unpack():
            st["k"] = ctx.undexta_dev(kind, pk.data_ptr(), st["m"], 80, False, un.data_ptr(), un.numel())

        w = run(f"undexta_dev {name}", f_unpack)
        print(f"   -> {UF / w / 1e6:.1f} GB/s")
        if kind == dx.FASTA:
            assert st["k"] == UF and bool(torch.equal(un[:UF], fa)), "round trip differs"
        del fa, pk, un
        torch.cuda.empty_cache()
ctx.close()
