"""ctypes view of oracle/libdxoracle.so (dx_oracle.c) plus runners for the reference tools
compiled into oracle/_ref/.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdxoracle.so")
REF_DIR = os.path.join(HERE, "_ref")

ERRORS = {-1: "format", -2: "capacity", -3: "truncated", -4: "key", -5: "linelen", -6: "toolong"}


class OracleError(RuntimeError):
    def __init__(self, code):
        super().__init__(f"oracle error {code} ({ERRORS.get(code, '?')})")
        self.code = code


class Stats(C.Structure):
    _fields_ = [("del_", C.c_uint64 * 256), ("ins", C.c_uint64 * 256), ("mrg", C.c_uint64 * 256),
                ("sub", C.c_uint64 * 256), ("delrun", C.c_uint64 * 256),
                ("subrun", C.c_uint64 * 256), ("totchar", C.c_uint64), ("delchar", C.c_int32),
                ("subchar", C.c_int32), ("nentries", C.c_int64)]


class Scheme(C.Structure):
    _fields_ = [("type", C.c_int32), ("bits", C.c_uint32 * 256), ("lens", C.c_int32 * 256)]


class Coding(C.Structure):
    _fields_ = [("tab", Scheme * 6), ("delchar", C.c_int32), ("subchar", C.c_int32)]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        u8p, i64 = C.c_char_p, C.c_int64
        L.orc_dexta.restype = i64
        L.orc_dexta.argtypes = [u8p, i64, C.c_int, C.c_void_p, i64]
        L.orc_undexta.restype = i64
        L.orc_undexta.argtypes = [u8p, i64, C.c_int, C.c_int, C.c_int, C.c_void_p, i64]
        L.orc_dexqv.restype = i64
        L.orc_dexqv.argtypes = [u8p, i64, C.c_int, C.c_void_p, i64]
        L.orc_undexqv.restype = i64
        L.orc_undexqv.argtypes = [u8p, i64, C.c_int, C.c_void_p, i64]
        L.orc_qv_scan.restype = C.c_int
        L.orc_qv_scan.argtypes = [u8p, i64, C.POINTER(Stats)]
        L.orc_qv_create.restype = C.c_int
        L.orc_qv_create.argtypes = [C.POINTER(Stats), C.c_int, C.POINTER(Coding)]
        L.orc_huffman.restype = None
        L.orc_huffman.argtypes = [C.POINTER(C.c_uint64), C.POINTER(Scheme), C.POINTER(Scheme)]
        L.orc_write_coding.restype = i64
        L.orc_write_coding.argtypes = [C.POINTER(Coding), u8p, C.c_int, C.c_void_p, i64]
        L.orc_encode_stream.restype = i64
        L.orc_encode_stream.argtypes = [C.POINTER(Scheme), C.POINTER(Scheme), C.c_int, u8p,
                                        C.c_int, C.c_void_p, i64]
        L.orc_dexqv_offsets.restype = i64
        L.orc_dexqv_offsets.argtypes = [u8p, i64, C.POINTER(i64), i64]
        L.orc_compress_read.restype = None
        L.orc_compress_read.argtypes = [C.c_int, C.c_void_p]
        L.orc_uncompress_read.restype = None
        L.orc_uncompress_read.argtypes = [C.c_int, C.c_void_p]
        for name in ("orc_number_read", "orc_number_arrow", "orc_lower_read", "orc_upper_read",
                     "orc_letter_arrow"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _run(fn, data: bytes, cap: int, *mid):
    out = C.create_string_buffer(max(cap, 16))
    r = fn(data, len(data), *mid, out, cap)
    if r < 0:
        raise OracleError(r)
    return out.raw[:r]


def dexta(text: bytes, arrow: bool = False) -> bytes:
    try:
        return _run(lib().orc_dexta, text, len(text) // 2 + 4096, int(arrow))
    except OracleError as e:                     # many tiny entries / huge well gaps: the image outgrows the text
        if e.code != -2:
            raise
        return _run(lib().orc_dexta, text, 64 * len(text) + (1 << 24), int(arrow))


def undexta(data: bytes, arrow: bool = False, width: int = 80, upper: bool = False) -> bytes:
    cap = len(data) * 9 + 65536
    return _run(lib().orc_undexta, data, cap, int(arrow), width, int(upper))


def dexqv(text: bytes, lossy: bool = False) -> bytes:
    return _run(lib().orc_dexqv, text, len(text) * 3 + 65536, int(lossy))


def undexqv(data: bytes, upper: bool = False) -> bytes:
    return _run(lib().orc_undexqv, data, len(data) * 40 + 65536, int(upper))


def qv_scan(text: bytes) -> Stats:
    st = Stats()
    r = lib().orc_qv_scan(text, len(text), C.byref(st))
    if r < 0:
        raise OracleError(r)
    return st


def qv_create(st: Stats, lossy: bool = False) -> Coding:
    c = Coding()
    lib().orc_qv_create(C.byref(st), int(lossy), C.byref(c))
    return c


def huffman(hist, inscheme: Scheme | None = None) -> Scheme:
    h = (C.c_uint64 * 256)(*[int(x) for x in hist])
    out = Scheme()
    lib().orc_huffman(h, C.byref(inscheme) if inscheme is not None else None, C.byref(out))
    return out


def write_coding(c: Coding, prefix: bytes) -> bytes:
    out = C.create_string_buffer(16384 + len(prefix))
    r = lib().orc_write_coding(C.byref(c), prefix, len(prefix), out, len(out))
    if r < 0:
        raise OracleError(r)
    return out.raw[:r]


def encode_stream(sym: Scheme, run: Scheme | None, rchar: int, s: bytes) -> bytes:
    cap = len(s) * 5 + 64
    out = C.create_string_buffer(cap)
    r = lib().orc_encode_stream(C.byref(sym), C.byref(run) if run is not None else None,
                                rchar, s, len(s), out, cap)
    if r < 0:
        raise OracleError(r)
    return out.raw[:r]


def dexqv_offsets(data: bytes, max_entries: int) -> np.ndarray:
    offs = (C.c_int64 * (max_entries + 2))()
    r = lib().orc_dexqv_offsets(data, len(data), offs, max_entries)
    if r < 0:
        raise OracleError(r)
    return np.frombuffer(offs, dtype=np.int64)[: r + 1].copy()


# ---------------------------------------------------------------------------------------------
#  The reference tools themselves (oracle/_ref, built by oracle/Makefile from /root/reference)
# ---------------------------------------------------------------------------------------------

def have_ref() -> bool:
    return all(os.path.exists(os.path.join(REF_DIR, t))
               for t in ("dexta", "undexta", "dexar", "undexar", "dexqv", "undexqv"))


_EXT = {"dexta": (".fasta", ".dexta"), "undexta": (".dexta", ".fasta"),
        "dexar": (".arrow", ".dexar"), "undexar": (".dexar", ".arrow"),
        "dexqv": (".quiva", ".dexqv"), "undexqv": (".dexqv", ".quiva")}


def ref_tool(tool: str, data: bytes, *flags: str, tmpdir: str | None = None,
             taskset: int | None = None):
    """Run one reference tool on `data`; returns (output bytes, wall seconds)."""
    import time
    src, dst = _EXT[tool]
    base = tmpdir or ("/dev/shm" if os.path.isdir("/dev/shm") else None)
    d = tempfile.mkdtemp(prefix="dxref_", dir=base)
    try:
        with open(os.path.join(d, "x" + src), "wb") as f:
            f.write(data)
        cmd = [os.path.join(REF_DIR, tool), "-k", *flags, os.path.join(d, "x" + src)]
        if taskset is not None and shutil.which("taskset"):
            cmd = ["taskset", "-c", str(taskset)] + cmd
        t0 = time.perf_counter()
        p = subprocess.run(cmd, capture_output=True)
        dt = time.perf_counter() - t0
        if p.returncode != 0:
            raise RuntimeError(f"{tool} failed rc={p.returncode}: {p.stderr.decode()[:500]}")
        with open(os.path.join(d, "x" + dst), "rb") as f:
            return f.read(), dt
    finally:
        shutil.rmtree(d, ignore_errors=True)


def ref_tool_parallel(tool: str, data: bytes, nproc: int, *flags: str):
    """`nproc` independent runs of one reference tool at the same time, each on its own copy of
    `data` (the reference is single threaded; independent files are how it uses more cores).
    Returns (output of the first run, wall seconds until the last one finished)."""
    import time
    src, dst = _EXT[tool]
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    d = tempfile.mkdtemp(prefix="dxrefp_", dir=base)
    try:
        for k in range(nproc):
            with open(os.path.join(d, f"x{k}" + src), "wb") as f:
                f.write(data)
        t0 = time.perf_counter()
        procs = [subprocess.Popen([os.path.join(REF_DIR, tool), "-k", *flags,
                                   os.path.join(d, f"x{k}" + src)],
                                  stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
                 for k in range(nproc)]
        for p in procs:
            _, err = p.communicate()
            if p.returncode != 0:
                raise RuntimeError(f"{tool} failed rc={p.returncode}: {err.decode()[:500]}")
        dt = time.perf_counter() - t0
        with open(os.path.join(d, "x0" + dst), "rb") as f:
            return f.read(), dt
    finally:
        shutil.rmtree(d, ignore_errors=True)
