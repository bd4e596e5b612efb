#!/usr/bin/env python
"""One launch of every hot kernel inside a cudaProfilerStart/Stop range (run under
`ncu --profile-from-start off`).  usage: ncu_once.py [quiva_gb] [fasta_gb]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import dextractor_b200 as dx
from dextractor_b200 import synth_torch

size = float(sys.argv[1]) if len(sys.argv) > 1 else 0.25
fsize = float(sys.argv[2]) if len(sys.argv) > 2 else 0.25
dev = torch.device("cuda", 0)
ctx = dx.Context(0)
rt = torch.cuda.cudart()

text, nent, npos = synth_torch.make_quiva_device(101, int(size * 1e9), dev)
U = text.numel()
enc = torch.empty(U // 2 + (1 << 20), dtype=torch.uint8, device=dev)
back = torch.empty(U + 4096, dtype=torch.uint8, device=dev)
fa, nfa = synth_torch.make_fasta_device(7, int(fsize * 1e9), dev)
UF = fa.numel()
pk = torch.empty(UF // 3 + (1 << 20), dtype=torch.uint8, device=dev)
un = torch.empty(UF + 4096, dtype=torch.uint8, device=dev)


def once():
    n = ctx.dexqv_dev(text.data_ptr(), U, False, enc.data_ptr(), enc.numel())
    m = ctx.undexqv_dev(enc.data_ptr(), n, False, back.data_ptr(), back.numel())
    assert m == U
    k = ctx.dexta_dev(dx.FASTA, fa.data_ptr(), UF, pk.data_ptr(), pk.numel())
    j = ctx.undexta_dev(dx.FASTA, pk.data_ptr(), k, 80, False, un.data_ptr(), un.numel())
    assert j == UF


for _ in range(2):
    once()
ctx.sync(); torch.cuda.synchronize()
rt.cudaProfilerStart()
once()
ctx.sync(); torch.cuda.synchronize()
rt.cudaProfilerStop()
assert bool(torch.equal(back[:U], text)) and bool(torch.equal(un[:UF], fa))
ctx.close()
print("ok")
