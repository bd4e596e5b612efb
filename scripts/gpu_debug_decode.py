import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DEXB200_DEBUG"] = "1"
import torch, numpy as np
import dextractor_b200 as dx
from dextractor_b200 import synth_torch, lib as dxl
dev = torch.device("cuda", 0)
ctx = dx.Context(0)
text, nent, npos = synth_torch.make_quiva_device(100, int(0.5e9), dev)
U = text.numel()
enc = torch.empty(U // 2 + (1 << 20), dtype=torch.uint8, device=dev)
back = torch.empty(U + 4096, dtype=torch.uint8, device=dev)
m = ctx.dexqv_dev(text.data_ptr(), U, False, enc.data_ptr(), enc.numel())
k = ctx.undexqv_dev(enc.data_ptr(), m, False, back.data_ptr(), back.numel())
print("round trip ok:", bool(torch.equal(back[:k], text)))
