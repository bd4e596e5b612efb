"""Seeded synthetic PacBio-shaped inputs (.fasta / .arrow / .quiva text).

The reference ships no sample data, so every test and benchmark input is
generated here, following the text formats written by the reference's extractor
(reference dextract.c:36-125) and the value distributions fixed in SURVEY.md section 8(d):

  * one movie name, wells non-decreasing (delta uniform 0..39), beg uniform [0,20000),
    end = beg + L, RQ uniform 750..899 (3 digits, so "RQ=0.%d" round-trips);
  * L ~ lognormal(mu=9.1, sigma=0.5) clipped to [500, 60000] unless explicit lengths are given;
  * fasta: lower-case acgt uniform, 80 columns;
  * arrow: SN=a,b,c,d with 2-decimal SNRs in [4,15], pulse widths '1'..'4' p=(.45,.30,.15,.10);
  * quiva (RS II-like): delQV '2' with tag 'n' at 88 % of positions, else 33+min(16,Geom(.18))
    with a uniform acgt tag; insQV 33+min(93,Geom(.12)); mergeQV 33+min(93,Geom(.04));
    subQV '?' at 80 %, else 33+min(29,Geom(.08)).

Everything here is host-side numpy; it is plumbing for tests/bench, not part of the hot path.
"""
from __future__ import annotations

import numpy as np

MOVIE = "m140913_050931_42139_c100713652400000001823152404301535_s1_p0"


def draw_lengths(rng: np.random.Generator, n: int, lo: int = 500, hi: int = 60000,
                 mu: float = 9.1, sigma: float = 0.5) -> np.ndarray:
    L = np.exp(rng.normal(mu, sigma, size=n))
    return np.clip(L, lo, hi).astype(np.int64)


def lengths_for_bytes(rng: np.random.Generator, target_bytes: int, bytes_per_base: float,
                      **kw) -> np.ndarray:
    """Draw lognormal lengths until about target_bytes of text would be produced."""
    out = []
    tot = 0
    while tot < target_bytes:
        L = draw_lengths(rng, 256, **kw)
        out.append(L)
        tot += int(L.sum() * bytes_per_base) + 100 * len(L)
    L = np.concatenate(out)
    cs = np.cumsum(L * bytes_per_base + 100)
    k = int(np.searchsorted(cs, target_bytes)) + 1
    return L[:k]


def _coords(rng, n, max_well_delta=39):
    wells = np.cumsum(rng.integers(0, max_well_delta + 1, size=n))
    begs = rng.integers(0, 20000, size=n)
    return wells, begs


def _wrap(seq: np.ndarray, width: int) -> bytes:
    """seq: uint8 array of symbols -> lines of `width` chars each ending in \\n."""
    n = len(seq)
    if n == 0:
        return b""
    nl = (n + width - 1) // width
    out = np.full(n + nl, 10, dtype=np.uint8)
    idx = np.arange(n)
    out[idx + idx // width] = seq
    return out.tobytes()


def make_fasta(seed: int, lengths, width: int = 80, max_well_delta: int = 39,
               alphabet: bytes = b"acgt", with_rq: bool = True, movie: str = MOVIE) -> bytes:
    rng = np.random.default_rng(seed)
    lengths = np.asarray(lengths, dtype=np.int64)
    n = len(lengths)
    wells, begs = _coords(rng, n, max_well_delta)
    rq = rng.integers(750, 900, size=n)
    alpha = np.frombuffer(alphabet, dtype=np.uint8)
    parts = []
    for i in range(n):
        L = int(lengths[i])
        hdr = f">{movie}/{wells[i]}/{begs[i]}_{begs[i] + L}"
        if with_rq:
            hdr += f" RQ=0.{rq[i]}"
        parts.append(hdr.encode() + b"\n")
        seq = alpha[rng.integers(0, len(alpha), size=L)]
        parts.append(_wrap(seq, width))
    return b"".join(parts)


def make_arrow(seed: int, lengths, width: int = 80, max_well_delta: int = 39,
               movie: str = MOVIE) -> bytes:
    rng = np.random.default_rng(seed)
    lengths = np.asarray(lengths, dtype=np.int64)
    n = len(lengths)
    wells, begs = _coords(rng, n, max_well_delta)
    alpha = np.frombuffer(b"1234", dtype=np.uint8)
    parts = []
    for i in range(n):
        L = int(lengths[i])
        snr = rng.integers(400, 1501, size=4) / 100.0
        hdr = (f">{movie}/{wells[i]}/{begs[i]}_{begs[i] + L} "
               f"SN={snr[0]:.2f},{snr[1]:.2f},{snr[2]:.2f},{snr[3]:.2f}")
        parts.append(hdr.encode() + b"\n")
        seq = alpha[rng.choice(4, size=L, p=(.45, .30, .15, .10))]
        parts.append(_wrap(seq, width))
    return b"".join(parts)


def quiva_streams(rng: np.random.Generator, L: int, p_run_del: float = 0.88,
                  p_run_sub: float = 0.80, no_n_tags: bool = False):
    """Five uint8 arrays (del, tag, ins, mrg, sub) of length L."""
    acgt = np.frombuffer(b"acgt", dtype=np.uint8)
    is_run = rng.random(L) < p_run_del
    dele = np.where(is_run, 33 + 17,
                    33 + np.minimum(16, rng.geometric(0.18, size=L))).astype(np.uint8)
    tag = np.where(is_run, ord("n"), acgt[rng.integers(0, 4, size=L)]).astype(np.uint8)
    if no_n_tags:
        tag = acgt[rng.integers(0, 4, size=L)]
    ins = (33 + np.minimum(93, rng.geometric(0.12, size=L))).astype(np.uint8)
    mrg = (33 + np.minimum(93, rng.geometric(0.04, size=L))).astype(np.uint8)
    is_srun = rng.random(L) < p_run_sub
    sub = np.where(is_srun, 33 + 30,
                   33 + np.minimum(29, rng.geometric(0.08, size=L))).astype(np.uint8)
    return dele, tag, ins, mrg, sub


def make_quiva(seed: int, lengths, max_well_delta: int = 39, p_run_del: float = 0.88,
               p_run_sub: float = 0.80, no_n_tags: bool = False, movie: str = MOVIE,
               stream_hook=None) -> bytes:
    """stream_hook(i, streams) may edit the 5 arrays of entry i in place (edge-case tests)."""
    rng = np.random.default_rng(seed)
    lengths = np.asarray(lengths, dtype=np.int64)
    n = len(lengths)
    wells, begs = _coords(rng, n, max_well_delta)
    rq = rng.integers(750, 900, size=n)
    parts = []
    for i in range(n):
        L = int(lengths[i])
        hdr = f"@{movie}/{wells[i]}/{begs[i]}_{begs[i] + L} RQ=0.{rq[i]}\n"
        parts.append(hdr.encode())
        streams = list(quiva_streams(rng, L, p_run_del, p_run_sub, no_n_tags))
        if stream_hook is not None:
            stream_hook(i, streams)
        for s in streams:
            parts.append(s.tobytes() + b"\n")
    return b"".join(parts)
