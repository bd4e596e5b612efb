"""libdexcompat_i.so (the -DINTERACTIVE build of the reference-symbol layer) on a GPU box: the
success path gives the reference's results, a malformed file gives the documented error value
and a message in Ebuffer instead of ending the process (QV.h:20-27, 56-58).  Runs in a child
process, after everything else."""
import json
import os
import subprocess
import sys
import textwrap

import pytest

from dextractor_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = textwrap.dedent("""
    import ctypes, json, sys
    L = ctypes.CDLL(sys.argv[1])
    libc = ctypes.CDLL(None)
    libc.fopen.restype = ctypes.c_void_p
    libc.fopen.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
    L.QVcoding_Scan.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    class QVcoding(ctypes.Structure):
        _fields_ = [("schemes", ctypes.c_void_p * 6), ("delChar", ctypes.c_int),
                    ("subChar", ctypes.c_int), ("flip", ctypes.c_int), ("prefix", ctypes.c_char_p)]
    L.Create_QVcoding.restype = ctypes.POINTER(QVcoding)
    ebuf = (ctypes.c_char * 1000).in_dll(L, "Ebuffer")
    out = {}
    f = libc.fopen(sys.argv[2].encode(), b"r")
    out["good_entries"] = L.QVcoding_Scan(f, 2**31 - 1, None)
    c = L.Create_QVcoding(0)
    out["coding"] = bool(c)
    if c:
        out["delChar"], out["subChar"] = c.contents.delChar, c.contents.subChar
    g = libc.fopen(sys.argv[3].encode(), b"r")
    ebuf.value = b""
    out["bad_return"] = L.QVcoding_Scan(g, 2**31 - 1, None)
    out["bad_message"] = ebuf.value.decode(errors="replace")
    out["alive"] = True
    print(json.dumps(out))
""")


def test_interactive_library_on_the_gpu(orc, tmp_path):
    text = synth.make_quiva(3, synth.draw_lengths(__import__("numpy").random.default_rng(3), 30))
    good, bad = tmp_path / "good.quiva", tmp_path / "bad.quiva"
    good.write_bytes(text)
    bad.write_bytes(b"@m/1/0_4 RQ=0.8\nabcd\nacgt\nabcd\nabc\nabcd\n")
    lib = os.path.join(ROOT, "dextractor_b200", "libdexcompat_i.so")
    r = subprocess.run([sys.executable, "-c", CHILD, lib, str(good), str(bad)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    st = orc.qv_scan(text)
    assert out["good_entries"] == st.nentries and out["coding"]
    assert (out["delChar"], out["subChar"]) == (st.delchar, st.subchar)
    assert out["bad_return"] == -1 and out["alive"]
    assert "same length" in out["bad_message"]
