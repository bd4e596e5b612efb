"""Synthetic PacBio-shaped text generated ON THE DEVICE with torch (plumbing for bench.py).

Same formats and value distributions as dextractor_b200/synth.py, but a
2 GB .quiva is produced in about a second instead of minutes.  The random streams differ from
the numpy generator's, so these buffers are benchmark inputs, not parity fixtures; bench.py checks
them through size-independent properties (decode(encode(x)) == x, header statistics).
"""
from __future__ import annotations

import numpy as np
import torch

from .synth import MOVIE, draw_lengths


def _lengths(rng, target_bytes, bytes_per_base, lo, hi):
    out, tot = [], 0
    while tot < target_bytes:
        L = draw_lengths(rng, 4096, lo=lo, hi=hi)
        out.append(L)
        tot += int(L.sum() * bytes_per_base) + 100 * len(L)
    L = np.concatenate(out)
    cs = np.cumsum(L * bytes_per_base + 100)
    return L[: int(np.searchsorted(cs, target_bytes)) + 1]


def _geom(n, p, cap, gen, device):
    g = torch.empty(n, dtype=torch.float32, device=device).geometric_(p, generator=gen)
    return torch.clamp(g, max=cap).to(torch.uint8)


def make_quiva_device(seed: int, target_bytes: int, device, well_base: int = 0,
                      lo: int = 500, hi: int = 60000, lengths=None):
    """-> (uint8 CUDA tensor holding a .quiva text, number of entries, number of positions)"""
    rng = np.random.default_rng(seed)
    L = np.asarray(lengths, dtype=np.int64) if lengths is not None else \
        _lengths(rng, target_bytes, 5.0, lo, hi)
    n = len(L)
    wells = well_base + np.cumsum(rng.integers(0, 40, size=n))
    begs = rng.integers(0, 20000, size=n)
    rq = rng.integers(750, 900, size=n)
    hdrs = [f"@{MOVIE}/{wells[i]}/{begs[i]}_{begs[i] + L[i]} RQ=0.{rq[i]}\n".encode()
            for i in range(n)]
    hlen = np.array([len(h) for h in hdrs], dtype=np.int64)
    body = 5 * (L + 1)
    ent_start = np.concatenate([[0], np.cumsum(hlen + body)])
    total = int(ent_start[-1])
    text = torch.empty(total, dtype=torch.uint8, device=device)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)

    # headers: one flat copy with a destination index
    hflat = torch.from_numpy(np.frombuffer(b"".join(hdrs), dtype=np.uint8).copy()).to(device)
    hdst = torch.repeat_interleave(torch.from_numpy(ent_start[:-1]).to(device),
                                   torch.from_numpy(hlen).to(device))
    hcum = np.concatenate([[0], np.cumsum(hlen)])[:-1]
    hdst += torch.arange(hflat.numel(), device=device) - \
        torch.repeat_interleave(torch.from_numpy(hcum).to(device), torch.from_numpy(hlen).to(device))
    text[hdst] = hflat
    del hflat, hdst

    # bodies, in chunks of entries to bound the index tensors
    step = 4096
    for a in range(0, n, step):
        b = min(n, a + step)
        Lc = torch.from_numpy(L[a:b]).to(device)
        P = int(L[a:b].sum())
        pstart = torch.from_numpy(np.concatenate([[0], np.cumsum(L[a:b])])[:-1]).to(device)
        ent = torch.repeat_interleave(torch.arange(b - a, device=device), Lc)
        col = torch.arange(P, device=device) - pstart[ent]
        line0 = torch.from_numpy(ent_start[a:b] + hlen[a:b]).to(device)[ent] + col
        stride = (Lc + 1)[ent]

        is_run = torch.rand(P, device=device, generator=gen) < 0.88
        dele = torch.where(is_run, torch.full((P,), 50, dtype=torch.uint8, device=device),
                           33 + _geom(P, 0.18, 16, gen, device))
        acgt = torch.tensor(list(b"acgt"), dtype=torch.uint8, device=device)
        tag = torch.where(is_run, torch.full((P,), ord("n"), dtype=torch.uint8, device=device),
                          acgt[torch.randint(0, 4, (P,), device=device, generator=gen)])
        text[line0] = dele
        text[line0 + stride] = tag
        del dele, tag, is_run
        text[line0 + 2 * stride] = 33 + _geom(P, 0.12, 93, gen, device)
        text[line0 + 3 * stride] = 33 + _geom(P, 0.04, 93, gen, device)
        is_srun = torch.rand(P, device=device, generator=gen) < 0.80
        sub = torch.where(is_srun, torch.full((P,), 63, dtype=torch.uint8, device=device),
                          33 + _geom(P, 0.08, 29, gen, device))
        text[line0 + 4 * stride] = sub
        del sub, is_srun, ent, col, line0, stride
        # the five newlines of every entry
        nl0 = torch.from_numpy(ent_start[a:b] + hlen[a:b]).to(device) + Lc
        for k in range(5):
            text[nl0 + k * (Lc + 1)] = 10
    # the library runs on its own non-blocking stream: the text must be complete before it is handed over
    torch.cuda.synchronize(device)
    return text, n, int(L.sum())


def make_fasta_device(seed: int, target_bytes: int, device, arrow: bool = False,
                      lo: int = 500, hi: int = 60000, width: int = 80):
    """-> (uint8 CUDA tensor holding a .fasta / .arrow text, number of entries)"""
    rng = np.random.default_rng(seed)
    L = _lengths(rng, target_bytes, 1.0 + 1.0 / width, lo, hi)
    n = len(L)
    wells = np.cumsum(rng.integers(0, 40, size=n))
    begs = rng.integers(0, 20000, size=n)
    if arrow:
        snr = rng.integers(400, 1501, size=(n, 4)) / 100.0
        hdrs = [(f">{MOVIE}/{wells[i]}/{begs[i]}_{begs[i] + L[i]} "
                 f"SN={snr[i,0]:.2f},{snr[i,1]:.2f},{snr[i,2]:.2f},{snr[i,3]:.2f}\n").encode()
                for i in range(n)]
    else:
        rq = rng.integers(750, 900, size=n)
        hdrs = [f">{MOVIE}/{wells[i]}/{begs[i]}_{begs[i] + L[i]} RQ=0.{rq[i]}\n".encode()
                for i in range(n)]
    hlen = np.array([len(h) for h in hdrs], dtype=np.int64)
    body = L + (L + width - 1) // width
    ent_start = np.concatenate([[0], np.cumsum(hlen + body)])
    total = int(ent_start[-1])
    text = torch.full((total,), 10, dtype=torch.uint8, device=device)      # newlines everywhere
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    hflat = torch.from_numpy(np.frombuffer(b"".join(hdrs), dtype=np.uint8).copy()).to(device)
    hl = torch.from_numpy(hlen).to(device)
    hdst = torch.repeat_interleave(torch.from_numpy(ent_start[:-1]).to(device), hl)
    hcum = torch.from_numpy(np.concatenate([[0], np.cumsum(hlen)])[:-1]).to(device)
    hdst += torch.arange(hflat.numel(), device=device) - torch.repeat_interleave(hcum, hl)
    text[hdst] = hflat
    del hflat, hdst
    if arrow:
        alpha = torch.tensor(list(b"1234"), dtype=torch.uint8, device=device)
        probs = torch.tensor([.45, .30, .15, .10], device=device)
    else:
        alpha = torch.tensor(list(b"acgt"), dtype=torch.uint8, device=device)
    step = 16384
    for a in range(0, n, step):
        b = min(n, a + step)
        Lc = torch.from_numpy(L[a:b]).to(device)
        P = int(L[a:b].sum())
        pstart = torch.from_numpy(np.concatenate([[0], np.cumsum(L[a:b])])[:-1]).to(device)
        ent = torch.repeat_interleave(torch.arange(b - a, device=device), Lc)
        col = torch.arange(P, device=device) - pstart[ent]
        dst = torch.from_numpy(ent_start[a:b] + hlen[a:b]).to(device)[ent] + col + col // width
        if arrow:
            sym = alpha[torch.multinomial(probs, P, replacement=True, generator=gen)]
        else:
            sym = alpha[torch.randint(0, 4, (P,), device=device, generator=gen)]
        text[dst] = sym
        del ent, col, dst, sym
    torch.cuda.synchronize(device)                   # see make_quiva_device
    return text, n
