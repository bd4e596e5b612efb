#!/usr/bin/env python
"""k_qv_hist_run counting variants side by side (route hist_mode) on a 2 GB bench file.
usage: hist_probe.py [GB]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dextractor_b200 as dx
from dextractor_b200 import synth_torch
size = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
dev = torch.device("cuda", 0)
ctx = dx.Context(0)
text, nent, npos = synth_torch.make_quiva_device(100, int(size * 1e9), dev)
torch.cuda.synchronize()
U = text.numel()
ref = None
for mode in (1, 0, 2, 3, 4):
    ctx.route("hist_mode", mode)
    for _ in range(2):
        st = ctx.qv_scan_dev(text.data_ptr(), U, None)
    ctx.profile(True); ctx.profile_report()
    for _ in range(3):
        st = ctx.qv_scan_dev(text.data_ptr(), U, None)
    prof = ctx.profile_report(); ctx.profile(False)
    h = np.frombuffer(bytes(st.hist), dtype=np.uint64).copy()
    if ref is None:
        ref = h
    print(f"mode {mode}: k_qv_hist_run {prof['k_qv_hist_run'][1]/prof['k_qv_hist_run'][0]:.3f} ms  "
          f"k_qv_hist_plain {prof['k_qv_hist_plain'][1]/prof['k_qv_hist_plain'][0]:.3f} ms  equal {bool((h == ref).all())}")
ctx.route("default")
