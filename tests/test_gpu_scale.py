"""Parity at BASELINE.json sizes, against the reference tools themselves (oracle/_ref, compiled from
the reference sources; the oracle port when they are absent): 1 GB .quiva / .fasta / .arrow generated
on the device, the GPU image compared byte for byte (sha256) with the reference tool's output for the
same text, and the reference's image decoded on the GPU -- with discovered and with known entry
offsets -- back to the text.  Then one 8 GB shard (the 8-GPU share of configs[3]'s 64 GB file): offsets
beyond 2^31 in text and image, round trip through both decoders.
"""
import hashlib

import numpy as np
import pytest

import dextractor_b200 as dx
from dextractor_b200 import lib as dxl
from dextractor_b200 import synth_torch

pytestmark = pytest.mark.gpu

GB = 10 ** 9


@pytest.fixture(scope="module")
def ctx():
    c = dx.Context(0)
    yield c
    c.close()


def sha(b) -> str:
    return hashlib.sha256(b).hexdigest()


def _ref(orc, tool, data, **kw):
    if orc.have_ref():
        return orc.ref_tool(tool, data)[0]
    port = {"dexqv": orc.dexqv, "undexqv": orc.undexqv, "dexta": orc.dexta, "undexta": orc.undexta,
            "dexar": lambda d: orc.dexta(d, arrow=True), "undexar": lambda d: orc.undexta(d, arrow=True)}
    return port[tool](data)


def test_quiva_1gb_against_the_reference_tools(ctx, orc):
    import torch
    dev = torch.device("cuda", 0)
    text_t, nent, npos = synth_torch.make_quiva_device(3, 1 * GB, dev)
    torch.cuda.synchronize()
    U = text_t.numel()
    text = text_t.cpu().numpy().tobytes()
    want = _ref(orc, "dexqv", text)                              # the reference's .dexqv for this text

    # GPU encode, device resident, with the entry index
    enc = torch.empty(U // 2 + (1 << 20), dtype=torch.uint8, device=dev)
    n = ctx.dexqv_dev(text_t.data_ptr(), U, False, enc.data_ptr(), enc.numel())
    got = enc[:n].cpu().numpy().tobytes()
    assert n == len(want) and sha(got) == sha(want), "GPU .dexqv differs from the reference's at 1 GB"

    # the REFERENCE's image decoded on the GPU: entries discovered, then entries known
    img = torch.from_numpy(np.frombuffer(want, dtype=np.uint8).copy()).to(dev)
    back = torch.zeros(U + 4096, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()            # torch fills on its own stream, the library runs on another
    m = ctx.undexqv_dev(img.data_ptr(), len(want), False, back.data_ptr(), back.numel())
    assert m == U and bool(torch.equal(back[:U], text_t)), "GPU undexqv (discovered) differs at 1 GB"
    ctx.keep_index(True)
    ctx.undexqv_dev(img.data_ptr(), len(want), False, back.data_ptr(), back.numel())
    rows = ctx.last_index()
    ctx.keep_index(False)
    assert len(rows) == nent
    # entry starts = first well-delta byte: walk back from the stream offsets over beg/end/qv and the delta bytes
    hdr_len = 2 + dxl.read_coding(want[2:2 + 200000])[2]
    offs = np.empty(nent + 1, dtype=np.int64)
    offs[0] = hdr_len
    for i, r in enumerate(rows):
        offs[i + 1] = r[1]
    back.zero_()
    torch.cuda.synchronize()
    m = ctx.undexqv_dev(img.data_ptr(), len(want), False, back.data_ptr(), back.numel(), entry_off=offs)
    assert m == U and bool(torch.equal(back[:U], text_t)), "GPU undexqv (offsets known) differs at 1 GB"
    # and the reference's own decoder on the GPU's image gives the text back
    assert sha(_ref(orc, "undexqv", got)) == sha(text)


@pytest.mark.parametrize("arrow", [False, True], ids=["fasta", "arrow"])
def test_fasta_arrow_1gb_against_the_reference_tools(ctx, orc, arrow):
    import torch
    dev = torch.device("cuda", 0)
    kind = dx.ARROW if arrow else dx.FASTA
    enc_tool, dec_tool = ("dexar", "undexar") if arrow else ("dexta", "undexta")
    fa, nfa = synth_torch.make_fasta_device(11 + int(arrow), 1 * GB, dev, arrow=arrow)
    torch.cuda.synchronize()
    U = fa.numel()
    text = fa.cpu().numpy().tobytes()
    want = _ref(orc, enc_tool, text)
    pk = torch.empty(U // 3 + (1 << 20), dtype=torch.uint8, device=dev)
    n = ctx.dexta_dev(kind, fa.data_ptr(), U, pk.data_ptr(), pk.numel())
    got = pk[:n].cpu().numpy().tobytes()
    assert n == len(want) and sha(got) == sha(want), f"GPU {enc_tool} differs from the reference's at 1 GB"
    ref_text = _ref(orc, dec_tool, want)                         # (arrow: SN= digits pass through a float)
    img = torch.from_numpy(np.frombuffer(want, dtype=np.uint8).copy()).to(dev)
    un = torch.zeros(U + 4096, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    k = ctx.undexta_dev(kind, img.data_ptr(), len(want), 80, False, un.data_ptr(), un.numel())
    assert k == len(ref_text) and sha(un[:k].cpu().numpy().tobytes()) == sha(ref_text), \
        f"GPU {dec_tool} differs from the reference's at 1 GB"


def test_quiva_8gb_shard_round_trip(ctx):
    """cfg4's per-GPU share at 8 GPUs: byte offsets beyond 2^31 (text) and 2^31 (image), ~157 k entries."""
    import torch
    dev = torch.device("cuda", 0)
    free, _ = torch.cuda.mem_get_info()
    if free < 60 * GB:
        pytest.skip("needs ~45 GB of device memory")
    text_t, nent, npos = synth_torch.make_quiva_device(5, 8 * GB, dev)
    torch.cuda.synchronize()
    U = text_t.numel()
    assert U > 2 ** 32
    enc = torch.empty(U // 2 + (1 << 20), dtype=torch.uint8, device=dev)
    st = ctx.qv_scan_dev(text_t.data_ptr(), U, None)
    assert int(st.nentries) == nent and int(st.totchar) == npos
    cd = dxl.make_coding(st, False)
    prefix = bytes(text_t[:200].cpu().numpy().tobytes())
    prefix = prefix[: prefix.index(b"/", 1)]
    hdr = b"\xaa\x55" + dxl.write_coding(cd, prefix)
    ctx.h2d(enc.data_ptr(), hdr)
    body, lastw, offs = ctx.qv_encode_dev(text_t.data_ptr(), U, cd, False, 0, enc.data_ptr() + len(hdr),
                                          enc.numel() - len(hdr), want_offsets=nent)
    n = len(hdr) + body
    assert n > 2 ** 31 and offs[-1] == body
    back = torch.zeros(U + 4096, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()            # torch fills on its own stream, the library runs on another
    m = ctx.undexqv_dev(enc.data_ptr(), n, False, back.data_ptr(), back.numel(), entry_off=offs + len(hdr))
    assert m == U and bool(torch.equal(back[:U], text_t)), "8 GB round trip (offsets known) differs"
    back.zero_()
    torch.cuda.synchronize()
    m = ctx.undexqv_dev(enc.data_ptr(), n, False, back.data_ptr(), back.numel())
    assert m == U and bool(torch.equal(back[:U], text_t)), "8 GB round trip (offsets discovered) differs"
    # a checksum of the image that a sharded run must reproduce: first and last MB + length
    head = enc[: 1 << 20].cpu().numpy().tobytes()
    tail = enc[n - (1 << 20): n].cpu().numpy().tobytes()
    assert len(sha(head + tail)) == 64
