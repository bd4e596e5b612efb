#!/usr/bin/env python
"""Foreign-endian and old-layout .dexqv fixtures (the reference reads both: undexqv.c:103-110,
135-180, QV.c:553-568 GETFLIP, 1226-1255), made in the build container and checked against the
REFERENCE undexqv before they are stored:

    python tests/golden/make_legacy.py

A .dexqv is rebuilt from a text with the oracle's table construction and stream encoder
(build(text, flip=False, old=False) must reproduce the reference encoder's file byte for byte -- that
pins this builder), then written again
  * with every 16/32-bit quantity of the writer in the OTHER byte order: the 0x55aa and 0x33cc keys,
    delChar/subChar, the prefix length, the code bits of the schemes, beg/end/qv, and every 32-bit
    word of the four Huffman streams (the packed tags are bytes and stay);
  * in the OLD layout: no 0x55aa key in front of the coding header, beg/end/qv as three uint16;
  * both.
The reference undexqv must decode each variant to the same text as the native file.
"""
import hashlib
import json
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from dextractor_b200 import synth      # noqa: E402
from oracle import orc                 # noqa: E402


def _swap_words(b: bytes) -> bytes:
    assert len(b) % 4 == 0
    return np.frombuffer(b, dtype="<u4").byteswap().tobytes()


def _flip_header(h: bytes, ntab: int) -> bytes:
    """native coding header (0x33cc key first) -> the same header written by a foreign-endian host"""
    out = bytearray()
    key, dc, sc, plen = struct.unpack_from("<HHHi", h, 0)
    out += struct.pack(">HHHi", key, dc, sc, plen)
    p = 10
    out += h[p:p + plen]
    p += plen
    for _ in range(ntab):
        out.append(h[p]); p += 1
        for _i in range(256):
            ln = h[p]; out.append(ln); p += 1
            if ln > 0:
                out += struct.pack(">I", struct.unpack_from("<I", h, p)[0]); p += 4
    assert p == len(h)
    return bytes(out)


def _pack_tags(tags: np.ndarray) -> bytes:
    code = np.zeros(256, dtype=np.uint8)
    for ch, v in ((b"c", 1), (b"g", 2), (b"t", 3), (b"C", 1), (b"G", 2), (b"T", 3)):
        code[ch[0]] = v
    c = code[tags]
    pad = (-len(c)) % 4
    c = np.concatenate([c, np.zeros(pad, dtype=np.uint8)]).reshape(-1, 4)
    return ((c[:, 0] << 6) | (c[:, 1] << 4) | (c[:, 2] << 2) | c[:, 3]).astype(np.uint8).tobytes()


def build(text: bytes, flip: bool, old: bool) -> bytes:
    st = orc.qv_scan(text)
    cd = orc.qv_create(st)
    lines = text.split(b"\n")
    prefix = lines[0][: lines[0].index(b"/")]
    hdr = orc.write_coding(cd, prefix)
    ntab = 4 + (cd.delchar >= 0) + (cd.subchar >= 0)
    out = bytearray()
    if not old:
        out += struct.pack(">H" if flip else "<H", 0x55aa)
    out += _flip_header(hdr, ntab) if flip else hdr
    well = 0
    end = ">" if flip else "<"
    for e in range(len(lines) // 6):
        h = lines[6 * e].decode()
        f = h.split("/")
        w = int(f[1]); beg, rest = f[2].split("_"); en, rq = rest.split(" RQ=0.")
        beg, en, qv = int(beg), int(en), int(rq)
        d = w - well
        out += b"\xff" * (d // 255) + bytes([d % 255])
        well = w
        out += struct.pack(end + ("HHH" if old else "iii"), beg, en, qv)
        dl, tg, ins, mrg, sub = (np.frombuffer(lines[6 * e + k], dtype=np.uint8) for k in range(1, 6))

        def stream(sym, run, rc, s):
            b = orc.encode_stream(cd.tab[sym], cd.tab[run] if rc >= 0 else None, rc, s.tobytes())
            return _swap_words(b) if flip else b

        out += stream(0, 1, cd.delchar, dl)
        kept = tg[dl != cd.delchar] if cd.delchar >= 0 else tg
        out += _pack_tags(kept)
        out += stream(2, 0, -1, ins)
        out += stream(3, 0, -1, mrg)
        out += stream(4, 5, cd.subchar, sub)
    return bytes(out)


def sha(b):
    return hashlib.sha256(b).hexdigest()


def main():
    assert orc.have_ref(), "build oracle/_ref first (make -C oracle ref)"
    rng = np.random.default_rng(77)
    # beg/end < 65536 so that the old 16-bit layout can hold them
    lens = [int(x) for x in rng.integers(1, 9000, size=60)] + [1, 2, 3, 4, 5, 31, 32, 33, 255, 256, 257]
    text = synth.make_quiva(77, lens)
    native = orc.ref_tool("dexqv", text)[0]
    mine = build(text, False, False)
    assert mine == native, "the builder does not reproduce the reference encoder"
    want = orc.ref_tool("undexqv", native)[0]
    manifest = {"input_sha256": sha(text), "decoded_sha256": sha(want), "decoded_len": len(want),
                "decoded_equals_input": want == text, "files": {}}
    # (seeded numpy draws are not bit-reproducible across CPUs -- SIMD log/exp -- so the native file is
    #  stored too: the expected text of every variant is what the oracle decodes from it)
    with open(os.path.join(HERE, "legacy_native.dexqv"), "wb") as f:
        f.write(native)
    manifest["native_sha256"] = sha(native)
    for flip, old in ((True, False), (False, True), (True, True)):
        b = build(text, flip, old)
        back = orc.ref_tool("undexqv", b)[0]
        assert back == want, (flip, old)
        name = "legacy_" + "_".join([x for x, on in (("foreign", flip), ("old", old)) if on]) + ".dexqv"
        with open(os.path.join(HERE, name), "wb") as f:
            f.write(b)
        manifest["files"][name] = {"flip": flip, "old": old, "sha256": sha(b)}
        print(f"{name:32s} {len(b)} bytes, reference undexqv decodes it to the native text")
    assert want == text            # the tests regenerate the expected text from the seed
    manifest["seed"] = 77
    manifest["lengths"] = lens
    with open(os.path.join(HERE, "legacy_manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)


if __name__ == "__main__":
    main()
