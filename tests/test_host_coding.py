"""Host logic of the product (no GPU): Huffman code-length assignment and the coding header,
dextractor_b200/csrc/dx_coding.cpp, against the oracle and the golden reference files.
The statistics come from the oracle here; on the GPU they come from the CUDA scan kernels
(tests/test_gpu_parity.py)."""
import ctypes as C

import numpy as np
import pytest

from dextractor_b200 import lib as dxl
from tests import cases, fuzz

QUIVA = dict(cases.quiva_cases())


def stats_from_oracle(st) -> dxl.Stats:
    """oracle statistics -> the product's dx_qv_stats (run histograms without the +1 start)."""
    out = dxl.Stats()
    for k, name in enumerate(["del_", "ins", "mrg", "sub"]):
        for i in range(256):
            out.hist[k][i] = getattr(st, name)[i]
    for i in range(256):
        out.hist[4][i] = st.delrun[i] - 1
        out.hist[5][i] = st.subrun[i] - 1
    out.totchar, out.nentries = st.totchar, st.nentries
    out.delchar, out.subchar = st.delchar, st.subchar
    return out


@pytest.mark.parametrize("lossy", [False, True])
@pytest.mark.parametrize("name", sorted(QUIVA))
def test_coding_header_matches_reference_file(orc, name, lossy):
    text = QUIVA[name]
    want = orc.dexqv(text, lossy=lossy)            # oracle == reference (test_oracle_vs_ref)
    st = orc.qv_scan(text)
    cd = dxl.make_coding(stats_from_oracle(st), lossy)
    prefix = text[: text.index(b"/", 1)]
    hdr = b"\xaa\x55" + dxl.write_coding(cd, prefix)
    assert want[: len(hdr)] == hdr
    # and the product's reader gives the same tables back
    cd2, pre2, used = dxl.read_coding(hdr[2:])
    assert pre2 == prefix and used == len(hdr) - 2
    assert cd2.delchar == cd.delchar and cd2.subchar == cd.subchar
    for k in range(6):
        if (k == 1 and cd.delchar < 0) or (k == 5 and cd.subchar < 0):
            continue
        assert list(cd2.tab[k].lens) == list(cd.tab[k].lens)
        assert list(cd2.tab[k].bits) == list(cd.tab[k].bits)
        assert cd2.tab[k].type == cd.tab[k].type


@pytest.mark.parametrize("seed", range(24))
def test_coding_header_of_random_shapes(orc, seed):
    """the same header check on the randomly shaped files of tests/fuzz.py, lossless and lossy"""
    text, _ = fuzz.fuzz_quiva(seed)
    st = stats_from_oracle(orc.qv_scan(text))
    prefix = text[: text.index(b"/", 1)]
    for lossy in (False, True):
        want = orc.dexqv(text, lossy=lossy)
        hdr = b"\xaa\x55" + dxl.write_coding(dxl.make_coding(st, lossy), prefix)
        assert want[: len(hdr)] == hdr


def test_huffman_tie_breaks_random_histograms(orc):
    rng = np.random.default_rng(5)
    for trial in range(200):
        nsym = int(rng.integers(2, 257))
        syms = rng.choice(256, size=nsym, replace=False)
        hist = np.zeros(256, dtype=np.uint64)
        mode = trial % 4
        if mode == 0:
            hist[syms] = rng.integers(1, 5, size=nsym)             # many ties
        elif mode == 1:
            hist[syms] = rng.integers(1, 1 << 30, size=nsym)
        elif mode == 2:
            hist[syms] = (2.0 ** rng.uniform(0, 40, size=nsym)).astype(np.uint64) + 1   # deep
        else:
            hist[syms] = 1
        first = orc.huffman(hist)
        want = orc.huffman(hist, first) if first.type else first
        st = dxl.Stats()
        for i in range(256):
            st.hist[1][i] = int(hist[i])
            st.hist[0][i] = st.hist[2][i] = st.hist[3][i] = 1
        cd = dxl.make_coding(st, False)
        got = cd.tab[2]
        assert got.type == want.type
        assert list(got.lens) == list(want.lens) and list(got.bits) == list(want.bits)


def test_single_symbol_stream_is_rejected():
    st = dxl.Stats()
    for i in range(256):
        st.hist[0][i] = st.hist[1][i] = st.hist[2][i] = st.hist[3][i] = 0
    st.hist[0][40] = st.hist[1][40] = st.hist[2][40] = st.hist[3][40] = 10
    with pytest.raises(dxl.DexError) as e:
        dxl.make_coding(st, False)
    assert e.value.code == -11


@pytest.mark.parametrize("seed", range(4))
def test_header_reader_survives_truncated_and_corrupt_headers(orc, seed):
    """dx_qv_read_coding (host) on every truncation of a real header -- each must be refused -- and
    on randomly corrupted headers, which must be refused or parsed but never read out of bounds
    (the reference reads through fread and simply fails, QV.c:1214-1320)."""
    text, _ = fuzz.fuzz_quiva(seed)
    data = orc.dexqv(text)
    _, _, used = dxl.read_coding(data[2:])
    hdr = data[2:2 + used]
    for k in range(len(hdr)):
        with pytest.raises(dxl.DexError):
            dxl.read_coding(hdr[:k])
    rng = np.random.default_rng(seed)
    for _ in range(500):
        b = bytearray(hdr)
        for _ in range(int(rng.integers(1, 4))):
            b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        try:
            _, _, n = dxl.read_coding(bytes(b))
            assert 0 < n <= len(b)
        except dxl.DexError:
            pass
