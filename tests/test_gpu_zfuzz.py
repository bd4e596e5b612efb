"""CUDA path (through the C ABI) against the oracle on the randomly shaped inputs of tests/fuzz.py
-- the same files tests/test_oracle_fuzz.py pins the oracle with against the reference tools.
Bit-exact.  (Named to run after the targeted parity tests.)"""
import pytest

import dextractor_b200 as dx
from tests import fuzz
from tests.test_gpu_parity import first_diff

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = dx.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("seed", range(24))
def test_fuzz_dexqv_undexqv(ctx, orc, seed):
    text, _ = fuzz.fuzz_quiva(seed)
    for lossy in (False, True):
        want = orc.dexqv(text, lossy=lossy)
        got = ctx.dexqv(text, lossy=lossy)
        assert got == want, (lossy, first_diff(got, want))
        back = ctx.undexqv(want)
        exp = orc.undexqv(want)
        assert back == exp, (lossy, first_diff(back, exp))


@pytest.mark.parametrize("seed", range(16))
def test_fuzz_dexta_dexar(ctx, orc, seed):
    fa, ar, w2 = fuzz.fuzz_fasta_arrow(seed)
    want = orc.dexta(fa)
    got = ctx.dexta(fa)
    assert got == want, first_diff(got, want)
    for upper in (False, True):
        back = ctx.undexta(want, width=w2, upper=upper)
        exp = orc.undexta(want, width=w2, upper=upper)
        assert back == exp, (w2, upper, first_diff(back, exp))
    want = orc.dexta(ar, arrow=True)
    got = ctx.dexta(ar, kind=dx.ARROW)
    assert got == want, first_diff(got, want)
    back = ctx.undexta(want, kind=dx.ARROW, width=w2)
    exp = orc.undexta(want, arrow=True, width=w2)
    assert back == exp, first_diff(back, exp)
