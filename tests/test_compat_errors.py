"""CPU-side checks of the reference-symbol layer (libdexcompat.so / libdexcompat_i.so):
both libraries export every QV.h / DB.h symbol of the path, the line reader follows QV.c:751-798,
and errors follow the reference's two conventions (DB.h:28-47): batch = message on stderr + exit,
-DINTERACTIVE = message in Ebuffer + documented error value.  No codec work runs here: without a
GPU the codec calls must fail loudly in either convention."""
import ctypes
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "dextractor_b200")

QV_H = ["Read_Lines", "QVentry", "Set_QV_Line", "Get_QV_Line", "QVcoding_Scan", "QVcoding_Scan1",
        "Create_QVcoding", "Read_QVcoding", "Write_QVcoding", "Free_QVcoding",
        "Compress_Next_QVentry", "Compress_Next_QVentry1", "Uncompress_Next_QVentry"]      # QV.h:48-97
DB_H = ["Compress_Read", "Uncompress_Read", "Lower_Read", "Upper_Read", "Number_Read",
        "Change_Read", "Letter_Arrow", "Number_Arrow",                                       # DB.h:255-267
        "Malloc", "Realloc", "Strdup", "Fopen", "PathTo", "Root", "Catenate", "Numbered_Suffix",
        "Prog_Name"]                                                                         # DB.h:71,235-247


def _load(name):
    path = os.path.join(LIBDIR, name)
    if not os.path.exists(path):
        pytest.fail(f"{name} is not built (python -c 'import __graft_entry__ as g; g.build()')")
    return ctypes.CDLL(path)


@pytest.mark.parametrize("name", ["libdexcompat.so", "libdexcompat_i.so"])
def test_reference_symbols_exported(name):
    L = _load(name)
    missing = [s for s in QV_H + DB_H if not hasattr(L, s)]
    assert not missing, missing
    assert hasattr(L, "Ebuffer") == name.endswith("_i.so")


class _Shim:
    def __init__(self):
        self.L = _load("libdexcompat_i.so")
        self.libc = ctypes.CDLL(None)
        self.libc.fopen.restype = ctypes.c_void_p
        self.libc.fopen.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
        self.libc.fclose.argtypes = [ctypes.c_void_p]
        self.L.Read_Lines.argtypes = [ctypes.c_void_p, ctypes.c_int]
        self.L.QVentry.restype = ctypes.c_char_p
        self.L.QVcoding_Scan.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        self.L.Create_QVcoding.restype = ctypes.c_void_p
        self.ebuf = (ctypes.c_char * 1000).in_dll(self.L, "Ebuffer")

    def open(self, path):
        f = self.libc.fopen(str(path).encode(), b"r")
        assert f
        return f

    def message(self):
        return self.ebuf.value.decode()


@pytest.fixture(scope="module")
def shim():
    return _Shim()


def test_read_lines_follows_the_reference(shim, tmp_path):
    p = tmp_path / "a.quiva"
    p.write_bytes(b"@m/1/0_4 RQ=0.850\nabcd\nefgh\nijkl\nmnop\nqrst\n@m/2/0_3 RQ=0.8\nabc\nab\n")
    f = shim.open(p)
    shim.L.Set_QV_Line(0)
    assert shim.L.Read_Lines(f, 1) == len("@m/1/0_4 RQ=0.850")
    assert shim.L.QVentry().startswith(b"@m/1/0_4")
    assert shim.L.Read_Lines(f, 5) == 4
    assert shim.L.Get_QV_Line() == 6
    assert shim.L.Read_Lines(f, 1) == len("@m/2/0_3 RQ=0.8")
    assert shim.L.Read_Lines(f, 5) == -2                      # QV.c:792-795
    assert "Lines for an entry are not the same length" in shim.message()
    assert shim.message().startswith("Line 9:")
    shim.libc.fclose(f)

    p.write_bytes(b"abcd\nefgh")
    f = shim.open(p)
    assert shim.L.Read_Lines(f, 2) == -2                      # QV.c:778-781
    assert "Last line does not end with a newline" in shim.message()
    assert shim.L.Read_Lines(f, 1) == -1                      # end of input before any line
    shim.libc.fclose(f)

    p.write_bytes(b"abcd\nefgh\n")
    f = shim.open(p)
    assert shim.L.Read_Lines(f, 5) == -2                      # QV.c:787-791
    assert "incomplete last entry" in shim.message()
    shim.libc.fclose(f)


def test_read_lines_grows_its_buffer(shim, tmp_path):
    L = 180_000                                               # beyond the first 50 000-byte slots
    p = tmp_path / "long.quiva"
    p.write_bytes(b"".join(bytes([65 + k]) * L + b"\n" for k in range(5)))
    f = shim.open(p)
    assert shim.L.Read_Lines(f, 5) == L
    assert shim.L.QVentry()[:3] == b"AAA"
    shim.libc.fclose(f)


def _no_gpu():
    import torch
    return not torch.cuda.is_available()


def test_interactive_codec_calls_fail_with_a_message_without_a_gpu(shim, tmp_path):
    if not _no_gpu():
        pytest.skip("a GPU is present")
    p = tmp_path / "b.quiva"
    p.write_bytes(b"@m/1/0_4 RQ=0.850\nabcd\nnnnn\nijkl\nmnop\nqrst\n")
    f = shim.open(p)
    assert shim.L.QVcoding_Scan(f, 2**31 - 1, None) == -1     # QV.h:56-58
    assert "no usable CUDA device" in shim.message()
    assert shim.L.Create_QVcoding(0) is None                  # QV.h:62-65
    shim.libc.fclose(f)


def test_batch_codec_calls_exit_without_a_gpu(tmp_path):
    if not _no_gpu():
        pytest.skip("a GPU is present")
    p = tmp_path / "c.quiva"
    p.write_bytes(b"@m/1/0_4 RQ=0.850\nabcd\nnnnn\nijkl\nmnop\nqrst\n")
    code = textwrap.dedent(f"""
        import ctypes
        L = ctypes.CDLL({os.path.join(LIBDIR, 'libdexcompat.so')!r})
        libc = ctypes.CDLL(None)
        libc.fopen.restype = ctypes.c_void_p
        L.QVcoding_Scan.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        f = libc.fopen({str(p).encode()!r}, b"r")
        L.QVcoding_Scan(f, 2**31 - 1, None)
        print("survived")
    """)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 1 and "survived" not in r.stdout
    assert "no usable CUDA device" in r.stderr
