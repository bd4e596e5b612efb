"""The six C command-line tools (tools/bin) on a GPU box: same files in, same files out as the
reference tools, including cross-decoding (our encoder -> reference decoder and the reverse)."""
import os
import shutil
import subprocess
import tempfile

import pytest

from tests import cases

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tools", "bin")


def run_tool(tool, path, *flags):
    p = subprocess.run([os.path.join(BIN, tool), *flags, path], capture_output=True)
    assert p.returncode == 0, p.stderr.decode()
    return p


@pytest.fixture()
def tmp():
    d = tempfile.mkdtemp(prefix="dxcli_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    yield d
    shutil.rmtree(d, ignore_errors=True)


def _write(path, data):
    with open(path, "wb") as f:
        f.write(data)


def _read(path):
    with open(path, "rb") as f:
        return f.read()


@pytest.mark.parametrize("kind,enc,dec,src,dst", [
    ("fasta", "dexta", "undexta", ".fasta", ".dexta"),
    ("arrow", "dexar", "undexar", ".arrow", ".dexar"),
    ("quiva", "dexqv", "undexqv", ".quiva", ".dexqv"),
])
def test_tools_match_reference(orc, tmp, kind, enc, dec, src, dst):
    text = cases.all_cases()[kind]["lognormal_40"]
    base = os.path.join(tmp, "x")
    _write(base + src, text)
    run_tool(enc, base + src, "-v")                   # without -k the source must disappear
    assert not os.path.exists(base + src)
    got = _read(base + dst)
    want = orc.dexqv(text) if kind == "quiva" else orc.dexta(text, arrow=(kind == "arrow"))
    assert got == want
    run_tool(dec, base, "-k")                          # name given without extension
    back = _read(base + src)
    exp = orc.undexqv(want) if kind == "quiva" else orc.undexta(want, arrow=(kind == "arrow"))
    assert back == exp
    assert os.path.exists(base + dst)
    if orc.have_ref():                                 # cross-decode with the reference binaries
        ref_back, _ = orc.ref_tool(dec, got)
        assert ref_back == exp
        ref_enc, _ = orc.ref_tool(enc, text)
        _write(base + "2" + dst, ref_enc)
        run_tool(dec, base + "2" + dst)
        assert _read(base + "2" + src) == exp


def test_flags_and_pipe_mode(orc, tmp):
    text = cases.all_cases()["fasta"]["upper_and_n"]
    p = subprocess.run([os.path.join(BIN, "dexta"), "-i"], input=text, capture_output=True)
    assert p.returncode == 0 and p.stdout == orc.dexta(text)
    q = subprocess.run([os.path.join(BIN, "undexta"), "-i", "-U", "-w37"], input=p.stdout,
                       capture_output=True)
    assert q.returncode == 0 and q.stdout == orc.undexta(p.stdout, width=37, upper=True)
    qtext = cases.all_cases()["quiva"]["short_file"]
    base = os.path.join(tmp, "q")
    _write(base + ".quiva", qtext)
    run_tool("dexqv", base + ".quiva", "-kl")
    assert _read(base + ".dexqv") == orc.dexqv(qtext, lossy=True)
    assert os.path.exists(base + ".quiva")
    run_tool("undexqv", base + ".dexqv", "-U", "-k")
    assert _read(base + ".quiva") == orc.undexqv(orc.dexqv(qtext, lossy=True), upper=True)


def test_errors_are_reported_like_the_reference(tmp):
    base = os.path.join(tmp, "bad")
    _write(base + ".quiva", b"@m/1/0_4 RQ=0.8\nabcd\nacgt\nabcd\nabc\nabcd\n")
    p = subprocess.run([os.path.join(BIN, "dexqv"), base], capture_output=True)
    assert p.returncode == 1 and b"not the same length" in p.stderr
    assert os.path.exists(base + ".quiva")            # the source survives a failure
    _write(base + ".dexta", b"\x12\x34rubbish")
    p = subprocess.run([os.path.join(BIN, "undexta"), base], capture_output=True)
    assert p.returncode == 1 and b"endian key invalid" in p.stderr


@pytest.mark.parametrize("tool,src,dst,arrow", [("dexta", ".fasta", ".dexta", False), ("dexar", ".arrow", ".dexar", True)])
def test_many_tiny_entries_outgrow_the_first_buffer(orc, tmp, tool, src, dst, arrow):
    """An image is 13 (17) header bytes per entry plus a quarter of the bases: 150 000 one-base entries
    with well gaps give an image LARGER than the text.  The tool's first buffer (n/3 + 200 000) is
    too small; it must ask the library what the call needs (dx_needed_bytes) and go again -- the
    reference compresses such a file without complaint (dexta.c:139-205)."""
    parts = []
    well = 0
    for i in range(150000):
        well += 1 + (i % 7) * 300                        # up to seven 0xff delta bytes
        if arrow:
            parts.append(b">m/%d/0_1 SN=5.00,6.00,7.00,8.00\n%c\n" % (well, b"1234"[i % 4]))
        else:
            parts.append(b">m/%d/0_1 RQ=0.850\n%c\n" % (well, b"acgt"[i % 4]))
    text = b"".join(parts)
    want = orc.dexta(text, arrow=arrow)
    assert len(want) > len(text) // 3 + 200000
    base = os.path.join(tmp, "tiny")
    _write(base + src, text)
    run_tool(tool, base + src, "-k")
    assert _read(base + dst) == want
    import dextractor_b200 as dx                      # and the same through the host-buffer binding
    ctx = dx.Context(0)
    try:
        assert ctx.dexta(text, kind=dx.ARROW if arrow else dx.FASTA) == want
    finally:
        ctx.close()
