"""Seeded random input shapes shared by the oracle fuzz test (CPU, against the reference tools) and
the CUDA fuzz test (GPU, against the oracle): alphabets, skews, run densities, file sizes around the
100 000 / 200 000 position thresholds of QV.c:1005-1013 and 1100-1104, line widths, well gaps.
Every file stays inside SURVEY.md Appendix B's generator constraints (7-bit non-zero symbols,
>= 2 distinct non-run symbols per stream, non-decreasing wells)."""
import numpy as np

from dextractor_b200 import synth

PRINTABLE = np.array([c for c in range(33, 127)], dtype=np.uint8)


def _random_line(rng, L, nsym, skew, avoid=-1):
    """L symbols over a random nsym-letter alphabet with Zipf-like weights rank^-skew."""
    alpha = rng.choice(PRINTABLE[PRINTABLE != avoid], size=nsym, replace=False)
    w = 1.0 / np.arange(1, nsym + 1) ** skew
    line = alpha[rng.choice(nsym, size=L, p=w / w.sum())]
    if L >= 2:                                   # two distinct symbols somewhere in the file
        line[0], line[1] = alpha[0], alpha[1]
    return line


def fuzz_quiva(seed):
    rng = np.random.default_rng(5000 + seed)
    target = int(rng.choice([400, 5000, 90_000, 130_000, 260_000]))
    lengths = []
    while sum(lengths) < target:
        lengths.append(int(min(60000, max(1, rng.lognormal(rng.uniform(3, 9), 0.8)))))
    nsym = [int(rng.integers(2, 60)) for _ in range(5)]
    skew = [float(rng.uniform(0.0, 4.0)) for _ in range(5)]
    p_del = float(rng.choice([0.0, 0.3, 0.88, 0.995]))
    p_sub = float(rng.choice([0.0, 0.45, 0.8, 0.995]))
    tags_follow = bool(rng.integers(0, 2))

    def hook(i, st):
        L = len(st[0])
        r = np.random.default_rng(seed * 1000 + i)
        d = _random_line(r, L, nsym[0], skew[0], avoid=50)
        run = r.random(L) < p_del
        run[:2] = False                          # Appendix B.9: two distinct non-run symbols
        d = np.where(run, 50, d).astype(np.uint8)
        tag = np.frombuffer(b"acgt", dtype=np.uint8)[r.integers(0, 4, size=L)]
        if tags_follow:
            tag = np.where(d == 50, ord("n"), tag).astype(np.uint8)
        elif L > 3:
            tag[r.integers(0, L)] = ord("N" if i % 2 else "n")
        st[0][:] = d
        st[1][:] = tag
        st[2][:] = _random_line(r, L, nsym[2], skew[2])
        st[3][:] = _random_line(r, L, nsym[3], skew[3])
        s = _random_line(r, L, nsym[4], skew[4], avoid=63)
        srun = r.random(L) < p_sub
        srun[:2] = False
        st[4][:] = np.where(srun, 63, s).astype(np.uint8)

    text = synth.make_quiva(seed, lengths, max_well_delta=int(rng.choice([1, 39, 700])),
                            stream_hook=hook)
    return text, tags_follow




def fuzz_fasta_arrow(seed):
    """-> (fasta, arrow, undexta width to test)"""
    rng = np.random.default_rng(7000 + seed)
    n = int(rng.integers(1, 60))
    lengths = [int(min(70000, max(1, rng.lognormal(rng.uniform(1, 9), 1.0)))) for _ in range(n)]
    width = int(rng.choice([1, 7, 60, 80, 81, 200]))
    delta = int(rng.choice([1, 39, 254, 255, 256, 3000]))
    alphabet = [b"acgt", b"ACGT", b"acgtnN", b"ACGTacgtRYKM-*"][int(rng.integers(0, 4))]
    fa = synth.make_fasta(seed, lengths, width=width, max_well_delta=delta, alphabet=alphabet,
                          with_rq=bool(rng.integers(0, 2)))
    ar = synth.make_arrow(seed, lengths, width=width, max_well_delta=delta)
    return fa, ar, int(rng.choice([1, 13, 80, 99, 1000]))
