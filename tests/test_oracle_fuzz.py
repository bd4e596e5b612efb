"""Seeded random differential test: oracle (oracle/dx_oracle.c) against the reference tools
(oracle/_ref) on inputs whose shape is drawn at random (tests/fuzz.py).  CPU only."""
import pytest

from tests import fuzz


@pytest.mark.parametrize("seed", range(24))
def test_fuzz_dexqv_undexqv(ref, seed):
    text, _ = fuzz.fuzz_quiva(seed)
    for lossy in (False, True):
        flags = ("-l",) if lossy else ()
        want, _ = ref.ref_tool("dexqv", text, *flags)
        assert ref.dexqv(text, lossy=lossy) == want
        back, _ = ref.ref_tool("undexqv", want)
        assert ref.undexqv(want) == back
        n = text.count(b"\n") // 6
        offs = ref.dexqv_offsets(want, n)
        assert len(offs) == n + 1 and offs[-1] == len(want)


@pytest.mark.parametrize("seed", range(16))
def test_fuzz_dexta_dexar(ref, seed):
    fa, ar, w2 = fuzz.fuzz_fasta_arrow(seed)
    want, _ = ref.ref_tool("dexta", fa)
    assert ref.dexta(fa) == want
    for args, kw in ((("-w%d" % w2,), dict(width=w2)), (("-U", "-w%d" % w2), dict(width=w2, upper=True))):
        back, _ = ref.ref_tool("undexta", want, *args)
        assert ref.undexta(want, **kw) == back
    want, _ = ref.ref_tool("dexar", ar)
    assert ref.dexta(ar, arrow=True) == want
    back, _ = ref.ref_tool("undexar", want, "-w%d" % w2)
    assert ref.undexta(want, arrow=True, width=w2) == back
