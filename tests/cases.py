"""Seeded input cases shared by the oracle tests (CPU) and the CUDA parity tests (GPU).

Each case is (name, bytes).  Sizes are chosen so the oracle finishes in well under a second.
The edge sweeps follow SURVEY.md section 4: lengths around the 4-symbols-per-byte, 32-bit word,
80-column and 255/256 run-bucket boundaries; scheme shapes with and without run characters.
"""
from __future__ import annotations

import numpy as np

from dextractor_b200 import synth

EDGE_LENGTHS = [1, 2, 3, 4, 5, 6, 7, 8, 9, 15, 16, 17, 31, 32, 33, 79, 80, 81, 159, 160, 161,
                254, 255, 256, 257, 511, 512, 513, 1023, 1024, 1025, 4095, 4096, 4097]


def fasta_cases():
    rng = np.random.default_rng(11)
    yield "edge_lengths", synth.make_fasta(1, EDGE_LENGTHS)
    yield "lognormal_40", synth.make_fasta(2, synth.draw_lengths(rng, 40))
    yield "width_60", synth.make_fasta(3, [100, 61, 60, 59, 1200], width=60)
    yield "width_1", synth.make_fasta(4, [5, 1, 9], width=1)
    yield "upper_and_n", synth.make_fasta(5, [300, 77, 4000], alphabet=b"ACGTNacgtnRy-")
    yield "no_rq", synth.make_fasta(6, [120, 80, 3], with_rq=False)
    yield "big_well_gaps", synth.make_fasta(7, [50] * 30, max_well_delta=1200)
    yield "long_65536", synth.make_fasta(8, [65534, 65535, 65536, 65537])
    yield "one_entry", synth.make_fasta(9, [1000])
    # ragged hand-written layout: uneven line widths, empty sequence lines, empty entry
    yield "ragged", (b">mv/1/0_10 RQ=0.851\nacg\n\ntacgtac\n>mv/1/20_20 RQ=0.800\n"
                     b">mv/7/5_9\nAC\nGT\n>mv/300/0_7 RQ=0.7\ngattaca\n")
    # header length disagrees with the sequence: the encoder packs what is there (encode only)
    yield "enc_only_len_mismatch", b">mv/3/0_3 RQ=0.75\ngattaca\n>mv/4/0_9\nacgtacgta\n"
    yield "many_short", synth.make_fasta(10, rng.integers(1, 200, size=500))


def arrow_cases():
    rng = np.random.default_rng(12)
    yield "edge_lengths", synth.make_arrow(1, EDGE_LENGTHS)
    yield "lognormal_40", synth.make_arrow(2, synth.draw_lengths(rng, 40))
    yield "big_well_gaps", synth.make_arrow(3, [50] * 30, max_well_delta=1200)
    yield "odd_symbols", (b">mv/1/0_12 SN=6.97,11.03,150.5,0.00\n12341234G0x5\n"
                          b">mv/2/0_3 SN=99.99,99.98,100.00,4.5\n421\n")


def _hook_symbol255(i, streams):
    if i % 3 == 0 and len(streams[2]) > 4:
        streams[2][::5] = 255
        streams[3][1::7] = 255


def _hook_long_runs(i, streams):
    L = len(streams[0])
    streams[0][:] = 50          # all-run deletion line (a single run item)
    streams[1][:] = ord("n")
    if i % 2 == 0 and L > 3:
        streams[0][L // 2] = 40  # one break in the middle
        streams[1][L // 2] = ord("t")
        streams[0][L - 1] = 41   # and one at the very end
        streams[1][L - 1] = ord("g")


def _hook_rare_symbols(i, streams):
    # a few very rare symbols force > 16-bit codes, i.e. type-2 (escape) tables
    if i == 0:
        for k, s in enumerate(streams):
            if k != 1:
                s[: min(30, len(s))] = np.arange(70, 70 + min(30, len(s)), dtype=np.uint8)


def _hook_rare_everywhere(i, streams):
    if i == 3:
        for k, s in enumerate(streams):
            if k != 1:
                m = min(40, len(s))
                s[100:100 + m] = np.arange(80, 80 + m, dtype=np.uint8)
                streams[1][100:100 + m] = ord("c")


def _hook_halving(i, streams):
    # symbol k with probability 2^-(k+1): Huffman depths beyond 16 in del, mrg and sub tables
    rng = np.random.default_rng(1000 + i)
    L = len(streams[0])
    for k in (0, 3, 4):
        v = np.minimum(rng.geometric(0.5, size=L) - 1, 21).astype(np.uint8) + 70
        if k == 3:
            streams[k][:] = v
        else:
            keep = streams[k] == (50 if k == 0 else 63)
            streams[k][:] = np.where(keep, streams[k], v)


def quiva_cases():
    rng = np.random.default_rng(13)
    # > 200000 positions: both run characters active, type-2 tables
    yield "lognormal_40", synth.make_quiva(1, synth.draw_lengths(rng, 40))
    # short file: subChar stays -1 (totChar < 200000)
    yield "short_file", synth.make_quiva(2, synth.draw_lengths(rng, 6, hi=3000))
    # totChar between 100000 and 200000: subChar found then dropped again
    yield "mid_file", synth.make_quiva(3, [30000] * 5)
    yield "edge_lengths", synth.make_quiva(4, EDGE_LENGTHS * 3)
    yield "no_n_tags", synth.make_quiva(5, synth.draw_lengths(rng, 30), no_n_tags=True)
    yield "sub_not_dominant", synth.make_quiva(6, synth.draw_lengths(rng, 30), p_run_sub=0.3)
    yield "symbol_255", synth.make_quiva(7, synth.draw_lengths(rng, 30),
                                         stream_hook=_hook_symbol255)
    yield "long_runs", synth.make_quiva(8, [300, 255, 256, 257, 254, 65535, 1, 2, 65000, 512]
                                        + [20000] * 8, stream_hook=_hook_long_runs)
    yield "rare_symbols", synth.make_quiva(9, synth.draw_lengths(rng, 40),
                                           stream_hook=_hook_rare_symbols)
    yield "big_well_gaps", synth.make_quiva(10, [400] * 600, max_well_delta=1500)
    yield "late_n", synth.make_quiva(11, [9000] * 30, no_n_tags=False,
                                     stream_hook=lambda i, s: (s[1].__setitem__(
                                         slice(None), ord("a")) if i < 7 else None))
    yield "dense_runs_99", synth.make_quiva(12, synth.draw_lengths(rng, 30), p_run_del=0.999,
                                            p_run_sub=0.999)
    # ~3 M positions: count ratios above 2^16 put every table (run tables too) into type 2
    yield "big_type2", synth.make_quiva(14, synth.draw_lengths(rng, 260),
                                        stream_hook=_hook_rare_everywhere)
    yield "deep_codes", synth.make_quiva(15, [20000] * 40, p_run_del=0.3, p_run_sub=0.55,
                                         stream_hook=_hook_halving)
    yield "no_runs", synth.make_quiva(13, synth.draw_lengths(rng, 30), p_run_del=0.0,
                                      p_run_sub=0.5)


# cases whose tags are not all 'n' exactly where del == delChar: the reference itself does not
# round-trip them to the identity (dropped tags come back as 'n'); parity is still exact.
QUIVA_NOT_IDENTITY = {"late_n", "rare_symbols"}


def all_cases():
    return {"fasta": dict(fasta_cases()), "arrow": dict(arrow_cases()),
            "quiva": dict(quiva_cases())}
