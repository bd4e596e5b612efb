"""Function-level pin of the oracle: the per-read DB.h routines of oracle/dx_oracle.c beside the
reference's own functions (DB.c compiled into oracle/_ref/libdbqv_ref.so), on whole buffers --
including the bytes the reference touches past `len` (Compress_Read leaves
s[len] = 0, Uncompress_Read writes up to 3 bytes past len and s[len] = 4).  CPU only."""
import ctypes
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LENGTHS = [0, 1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 63, 64, 65, 255, 256, 257, 1000, 4099, 65537]


@pytest.fixture(scope="module")
def both(orc):
    path = os.path.join(ROOT, "oracle", "_ref", "libdbqv_ref.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libdbqv_ref.so not built (no /root/reference here)")
    return ctypes.CDLL(path), orc.lib()


def _call(fn, buf, *pre):
    b = ctypes.create_string_buffer(bytes(buf), len(buf))
    fn(*pre, b)
    return b.raw


@pytest.mark.parametrize("L", LENGTHS)
def test_number_compress_uncompress_letter(both, L):
    R, O = both
    rng = np.random.default_rng(L)
    pad = 8
    for alphabet, number, o_number, letters in (
            (b"acgtACGTnNxy-", "Number_Read", "orc_number_read",
             (("Lower_Read", "orc_lower_read"), ("Upper_Read", "orc_upper_read"))),
            (b"1234G0x5", "Number_Arrow", "orc_number_arrow", (("Letter_Arrow", "orc_letter_arrow"),))):
        alpha = np.frombuffer(alphabet, dtype=np.uint8)
        text = alpha[rng.integers(0, len(alpha), size=L)].tobytes() + b"\0" + bytes([0x5a] * pad)
        a = _call(getattr(R, number), text)
        b = _call(getattr(O, o_number), text)
        assert a == b                                            # numeric codes + the 4 terminator
        ca = _call(R.Compress_Read, a, ctypes.c_int(L))
        cb = _call(O.orc_compress_read, b, ctypes.c_int(L))
        assert ca == cb                                          # packed bytes AND what is left behind
        ua = _call(R.Uncompress_Read, ca, ctypes.c_int(L))
        ub = _call(O.orc_uncompress_read, cb, ctypes.c_int(L))
        assert ua == ub                                          # including the spill past len
        for rname, oname in letters:
            assert _call(getattr(R, rname), ua) == _call(getattr(O, oname), ub)
