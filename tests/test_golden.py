"""Oracle against the committed reference outputs in tests/golden/ (made by
tests/golden/make_golden.py with the reference tools).  CPU only; needs no reference mount."""
import hashlib
import json
import os

import pytest

from tests import cases

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MANIFEST = json.load(open(os.path.join(GOLD, "manifest.json")))


def _sha(b):
    return hashlib.sha256(b).hexdigest()


def load_entry(ent):
    """(input text, reference-encoded bytes) for one manifest entry; checks the fixture hashes."""
    if "input" in ent:
        text = open(os.path.join(GOLD, ent["input"]), "rb").read()
    else:
        text = cases.all_cases()[ent["kind"]][ent["case"]]
    assert _sha(text) == ent["input_sha256"], "seeded generator drifted from the golden input"
    enc = open(os.path.join(GOLD, ent["encoded"]), "rb").read()
    assert _sha(enc) == ent["encoded_sha256"]
    return text, enc


@pytest.mark.parametrize("ent", MANIFEST, ids=lambda e: e["encoded"])
def test_oracle_matches_golden(orc, ent):
    text, enc = load_entry(ent)
    kind = ent["kind"]
    if kind == "quiva":
        assert orc.dexqv(text, lossy=bool(ent["flags"])) == enc
        back = orc.undexqv(enc)
    else:
        assert orc.dexta(text, arrow=(kind == "arrow")) == enc
        back = orc.undexta(enc, arrow=(kind == "arrow"))
    assert len(back) == ent["decoded_len"] and _sha(back) == ent["decoded_sha256"]
    assert (back == text) == ent["decoded_equals_input"]
