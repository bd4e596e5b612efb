#!/usr/bin/env python
"""Regenerates tests/golden/ from the REFERENCE TOOLS (oracle/_ref, compiled by oracle/Makefile
from the mounted reference sources).  Run in the build container only:

    python tests/golden/make_golden.py

For every chosen case of tests/cases.py it stores the reference encoder's output bytes and, in
manifest.json, the sha256 of the (seed-regenerated) input and of the reference decoder's output.
Tiny hand-written inputs are stored verbatim as well.
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import orc          # noqa: E402
from tests import cases         # noqa: E402

CHOSEN = {
    "fasta": ["ragged", "edge_lengths", "width_60", "upper_and_n", "big_well_gaps", "no_rq"],
    "arrow": ["odd_symbols", "edge_lengths", "big_well_gaps"],
    "quiva": ["short_file", "edge_lengths", "mid_file", "lognormal_40", "long_runs",
              "symbol_255"],
}
TOOLS = {"fasta": ("dexta", "undexta", ".dexta"), "arrow": ("dexar", "undexar", ".dexar"),
         "quiva": ("dexqv", "undexqv", ".dexqv")}
STORE_INPUT = {("fasta", "ragged"), ("arrow", "odd_symbols")}


def sha(b):
    return hashlib.sha256(b).hexdigest()


def main():
    assert orc.have_ref(), "build oracle/_ref first (make -C oracle ref)"
    allc = cases.all_cases()
    manifest = []
    for kind, names in CHOSEN.items():
        enc, dec, ext = TOOLS[kind]
        for name in names:
            text = allc[kind][name]
            variants = [()]
            if kind == "quiva" and name in ("lognormal_40", "short_file"):
                variants.append(("-l",))
            for flags in variants:
                out, _ = orc.ref_tool(enc, text, *flags)
                back, _ = orc.ref_tool(dec, out)
                tag = f"{kind}_{name}" + ("_lossy" if flags else "")
                fn = tag + ext
                with open(os.path.join(HERE, fn), "wb") as f:
                    f.write(out)
                ent = {"kind": kind, "case": name, "flags": list(flags), "encoded": fn,
                       "input_len": len(text), "input_sha256": sha(text),
                       "encoded_sha256": sha(out), "decoded_len": len(back),
                       "decoded_sha256": sha(back), "decoded_equals_input": back == text}
                if (kind, name) in STORE_INPUT:
                    inp = tag + ".input"
                    with open(os.path.join(HERE, inp), "wb") as f:
                        f.write(text)
                    ent["input"] = inp
                manifest.append(ent)
                print(f"{tag:32s} in={len(text):8d} out={len(out):8d} identity={back == text}")
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)


if __name__ == "__main__":
    main()
