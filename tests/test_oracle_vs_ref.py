"""Pins the oracle (oracle/dx_oracle.c) against the reference tools themselves (oracle/_ref,
compiled from the mounted reference sources by oracle/Makefile).  CPU only."""
import pytest

from tests import cases

FASTA = dict(cases.fasta_cases())
ARROW = dict(cases.arrow_cases())
QUIVA = dict(cases.quiva_cases())


@pytest.mark.parametrize("name", sorted(FASTA))
def test_dexta_undexta(ref, name):
    text = FASTA[name]
    want, _ = ref.ref_tool("dexta", text)
    assert ref.dexta(text) == want
    if name.startswith("enc_only"):
        return
    back, _ = ref.ref_tool("undexta", want)
    assert ref.undexta(want) == back
    back_u, _ = ref.ref_tool("undexta", want, "-U", "-w37")
    assert ref.undexta(want, width=37, upper=True) == back_u


@pytest.mark.parametrize("name", sorted(ARROW))
def test_dexar_undexar(ref, name):
    text = ARROW[name]
    want, _ = ref.ref_tool("dexar", text)
    assert ref.dexta(text, arrow=True) == want
    back, _ = ref.ref_tool("undexar", want)
    assert ref.undexta(want, arrow=True) == back
    back_w, _ = ref.ref_tool("undexar", want, "-w100")
    assert ref.undexta(want, arrow=True, width=100) == back_w


@pytest.mark.parametrize("lossy", [False, True])
@pytest.mark.parametrize("name", sorted(QUIVA))
def test_dexqv_undexqv(ref, name, lossy):
    text = QUIVA[name]
    flags = ("-l",) if lossy else ()
    want, _ = ref.ref_tool("dexqv", text, *flags)
    got = ref.dexqv(text, lossy=lossy)
    assert got == want
    back, _ = ref.ref_tool("undexqv", want)
    assert ref.undexqv(want) == back
    if not lossy and name not in cases.QUIVA_NOT_IDENTITY:
        assert back == text
    back_u, _ = ref.ref_tool("undexqv", want, "-U")
    assert ref.undexqv(want, upper=True) == back_u


def test_entry_offsets_walk(ref):
    text = QUIVA["edge_lengths"]
    data = ref.dexqv(text)
    n = text.count(b"\n") // 6
    offs = ref.dexqv_offsets(data, n)
    assert len(offs) == n + 1 and offs[-1] == len(data)
    assert all(offs[i] < offs[i + 1] for i in range(n))
