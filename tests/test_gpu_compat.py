"""Drop-in check of the QV.h / DB.h boundary: the reference's OWN mains (dexqv.c, undexqv.c, dexta.c,
undexta.c, dexar.c, undexar.c -- compiled from /root/reference with the reference's headers by
`make -C oracle compat`) linked against libdexcompat.so, i.e. against this repository's GPU
implementation of Compress_Next_QVentry, Uncompress_Next_QVentry, Compress_Read, ... instead of
DB.c + QV.c.  Their output files must equal those of the real reference tools byte for byte."""
import os
import shutil
import subprocess
import tempfile

import pytest

from tests import cases

pytestmark = pytest.mark.gpu

REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
EXT = {"dexta": (".fasta", ".dexta"), "undexta": (".dexta", ".fasta"),
       "dexar": (".arrow", ".dexar"), "undexar": (".dexar", ".arrow"),
       "dexqv": (".quiva", ".dexqv"), "undexqv": (".dexqv", ".quiva")}


def run(binary: str, tool: str, data: bytes, *flags: str) -> bytes:
    src, dst = EXT[tool]
    d = tempfile.mkdtemp(prefix="dxcompat_")
    try:
        with open(os.path.join(d, "x" + src), "wb") as f:
            f.write(data)
        p = subprocess.run([os.path.join(REF, binary), "-k", *flags, os.path.join(d, "x" + src)],
                           capture_output=True, timeout=600)
        assert p.returncode == 0, (binary, p.stderr.decode()[:500])
        with open(os.path.join(d, "x" + dst), "rb") as f:
            return f.read()
    finally:
        shutil.rmtree(d, ignore_errors=True)


@pytest.fixture(scope="module")
def have():
    names = ["compat_" + t for t in EXT] + list(EXT)
    if not all(os.path.exists(os.path.join(REF, n)) for n in names):
        pytest.skip("oracle/_ref/compat_* not built (make -C oracle ref compat, needs /root/reference)")


@pytest.mark.parametrize("name", ["lognormal_40", "edge_lengths", "short_file", "long_runs",
                                  "big_well_gaps", "rare_symbols", "no_n_tags"])
def test_reference_dexqv_mains_over_the_gpu_library(have, name):
    text = dict(cases.quiva_cases())[name]
    want = run("dexqv", "dexqv", text)
    got = run("compat_dexqv", "dexqv", text)
    assert got == want
    assert run("compat_undexqv", "undexqv", want) == run("undexqv", "undexqv", want)
    assert run("compat_undexqv", "undexqv", want, "-U") == run("undexqv", "undexqv", want, "-U")


def test_reference_dexqv_main_lossy(have):
    text = dict(cases.quiva_cases())["lognormal_40"]
    assert run("compat_dexqv", "dexqv", text, "-l") == run("dexqv", "dexqv", text, "-l")


@pytest.mark.parametrize("name", ["edge_lengths", "width_60", "upper_and_n", "ragged", "big_well_gaps"])
def test_reference_dexta_mains_over_the_gpu_library(have, name):
    text = dict(cases.fasta_cases())[name]
    want = run("dexta", "dexta", text)
    assert run("compat_dexta", "dexta", text) == want
    assert run("compat_undexta", "undexta", want) == run("undexta", "undexta", want)
    assert run("compat_undexta", "undexta", want, "-U", "-w60") == run("undexta", "undexta", want, "-U", "-w60")


@pytest.mark.parametrize("name", ["edge_lengths", "odd_symbols"])
def test_reference_dexar_mains_over_the_gpu_library(have, name):
    text = dict(cases.arrow_cases())[name]
    want = run("dexar", "dexar", text)
    assert run("compat_dexar", "dexar", text) == want
    assert run("compat_undexar", "undexar", want) == run("undexar", "undexar", want)
