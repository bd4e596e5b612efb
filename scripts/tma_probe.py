#!/usr/bin/env python
"""The newline index of a 2 GB text two ways: vector loads (k_pred_slots) and cp.async.bulk tiles
(k_pred_slots_bulk, route index_bulk = CTAs per SM).  usage: tma_probe.py [GB]"""
import os, sys
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
ref = None
for mode in (0, 2, 4, 6, 0, 4):
    ctx.route("index_bulk", mode)
    for _ in range(2):
        nl, _ = ctx.text_lines_dev(text.data_ptr(), U, 0)
    ctx.profile(True); ctx.profile_report()
    for _ in range(5):
        nl, _ = ctx.text_lines_dev(text.data_ptr(), U, 0)
    prof = ctx.profile_report(); ctx.profile(False)
    name = "k_pred_slots_bulk" if mode else "k_pred_slots"
    ms = prof[name][1] / prof[name][0]
    ref = nl if ref is None else ref
    print(f"index_bulk={mode}: {name:18s} {ms:.4f} ms = {U / ms / 1e6:.0f} GB/s   lines {nl} equal {nl == ref}")
ctx.route("default")
