#!/bin/bash
# six 2 GB bench files of different seeds / well ranges: does the discovered-entry decode hold its assumed
# layout (decoded in place) on each of them?  Prints the library's debug line per file.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
DEXB200_DEBUG=1 timeout 600 python - <<'PY' 2>&1 | grep -v "host phases" | tail -30
import torch, numpy as np
import dextractor_b200 as dx
from dextractor_b200 import synth_torch, lib as dxl
dev = torch.device("cuda", 0)
ctx = dx.Context(0)
for seed, base in ((100, 0), (101, 2_000_000), (102, 4_000_000), (103, 6_000_000), (104, 8_000_000), (105, 10_000_000)):
    text, nent, npos = synth_torch.make_quiva_device(seed, int(2e9), dev, well_base=base)
    torch.cuda.synchronize()
    U = text.numel()
    enc = torch.empty(U // 2 + (1 << 20), dtype=torch.uint8, device=dev)
    back = torch.empty(U + 4096, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    n = ctx.dexqv_dev(text.data_ptr(), U, False, enc.data_ptr(), enc.numel())
    print("seed", seed, "entries", nent, flush=True)
    m = ctx.undexqv_dev(enc.data_ptr(), n, False, back.data_ptr(), back.numel())
    print("  ok", m == U and bool(torch.equal(back[:U], text)), flush=True)
    del text, enc, back
    torch.cuda.empty_cache()
PY
