import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time, torch
import dextractor_b200 as dx
from dextractor_b200 import synth_torch
dev = torch.device("cuda", 0)
ctx = dx.Context(0)
GB = 1e9
import numpy as np
rs = np.random.default_rng(5)
npos_t = int(0.5 * GB / 5.02)
dists = {"loguniform": lambda k: np.exp(rs.uniform(np.log(500), np.log(50000), size=k)),
         "short90": lambda k: np.where(rs.random(k) < 0.9, 500, 50000),
         "long90": lambda k: np.where(rs.random(k) < 0.1, 500, 50000)}
for mix, draw in dists.items():
    Ls = np.asarray(draw(200000), dtype=np.int64)
    Ls = Ls[: int(np.searchsorted(np.cumsum(Ls), npos_t)) + 1]
    text, nent, npos = synth_torch.make_quiva_device(50, 0, dev, lengths=Ls)
    U = text.numel()
    enc = torch.empty(U // 2 + (1 << 20), dtype=torch.uint8, device=dev)
    back = torch.empty(U + 4096, dtype=torch.uint8, device=dev)
    n = ctx.dexqv_dev(text.data_ptr(), U, False, enc.data_ptr(), enc.numel())
    for env in ({}, {"DEXB200_DECODER": "v4"}):
        for k, v in env.items(): os.environ[k] = v
        for _ in range(2): m = ctx.undexqv_dev(enc.data_ptr(), n, False, back.data_ptr(), back.numel())
        ctx.sync(); torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(3): m = ctx.undexqv_dev(enc.data_ptr(), n, False, back.data_ptr(), back.numel())
        ctx.sync(); torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
        print(mix, nent, env, f"{U/dt/1e9:.1f} GB/s", bool(torch.equal(back[:U], text)))
        for k in env: del os.environ[k]
ctx.close()
