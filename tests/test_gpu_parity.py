"""Parity tests proper: the CUDA path, called through the C ABI (libdexb200.so), against the
oracle on the same seeded inputs and against the committed golden reference outputs.
Bit-exact: this is byte/integer work, there is no tolerance anywhere."""
import hashlib
import json
import os

import numpy as np
import pytest

import dextractor_b200 as dx
from tests import cases

pytestmark = pytest.mark.gpu

FASTA = dict(cases.fasta_cases())
ARROW = dict(cases.arrow_cases())
QUIVA = dict(cases.quiva_cases())
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MANIFEST = json.load(open(os.path.join(GOLD, "manifest.json")))


@pytest.fixture(scope="module")
def ctx():
    c = dx.Context(0)
    yield c
    c.close()


def first_diff(a: bytes, b: bytes) -> str:
    if a == b:
        return "equal"
    n = min(len(a), len(b))
    x = np.frombuffer(a[:n], dtype=np.uint8) != np.frombuffer(b[:n], dtype=np.uint8)
    i = int(np.argmax(x)) if x.any() else n
    return f"len {len(a)} vs {len(b)}, first difference at byte {i}: {a[i:i+8]!r} vs {b[i:i+8]!r}"


@pytest.mark.parametrize("name", sorted(FASTA))
def test_dexta_undexta(ctx, orc, name):
    text = FASTA[name]
    want = orc.dexta(text)
    got = ctx.dexta(text)
    assert got == want, first_diff(got, want)
    if name.startswith("enc_only"):
        return
    for width, upper in ((80, False), (37, True), (1, False), (100000, False)):
        back = ctx.undexta(want, width=width, upper=upper)
        exp = orc.undexta(want, width=width, upper=upper)
        assert back == exp, (width, upper, first_diff(back, exp))


@pytest.mark.parametrize("name", sorted(ARROW))
def test_dexar_undexar(ctx, orc, name):
    text = ARROW[name]
    want = orc.dexta(text, arrow=True)
    got = ctx.dexta(text, kind=dx.ARROW)
    assert got == want, first_diff(got, want)
    for width in (80, 100):
        back = ctx.undexta(want, kind=dx.ARROW, width=width)
        exp = orc.undexta(want, arrow=True, width=width)
        assert back == exp, first_diff(back, exp)


@pytest.mark.parametrize("lossy", [False, True])
@pytest.mark.parametrize("name", sorted(QUIVA))
def test_dexqv(ctx, orc, name, lossy):
    text = QUIVA[name]
    want = orc.dexqv(text, lossy=lossy)
    got = ctx.dexqv(text, lossy=lossy)
    assert got == want, first_diff(got, want)


@pytest.mark.parametrize("name", sorted(QUIVA))
def test_undexqv(ctx, orc, name):
    text = QUIVA[name]
    enc = orc.dexqv(text)
    want = orc.undexqv(enc)
    got = ctx.undexqv(enc)
    assert got == want, first_diff(got, want)
    gotu = ctx.undexqv(enc, upper=True)
    assert gotu == orc.undexqv(enc, upper=True)


@pytest.mark.parametrize("name", ["lognormal_40", "late_n", "short_file", "no_n_tags"])
def test_scan_statistics(ctx, orc, name):
    """QVcoding_Scan: histograms, run histograms and the order-dependent run characters."""
    import torch
    text = QUIVA[name]
    t = torch.frombuffer(bytearray(text), dtype=torch.uint8).cuda()
    st = ctx.qv_scan_dev(t.data_ptr(), len(text))
    ref = orc.qv_scan(text)
    assert (st.delchar, st.subchar, st.totchar, st.nentries) == \
           (ref.delchar, ref.subchar, ref.totchar, ref.nentries)
    for k, nm in enumerate(["del_", "ins", "mrg", "sub"]):
        assert list(st.hist[k]) == list(getattr(ref, nm)), nm
    assert [x + 1 for x in st.hist[4]] == list(ref.delrun)
    assert [x + 1 for x in st.hist[5]] == list(ref.subrun)


def test_device_pointer_round_trip_with_index(ctx, orc):
    """dx_dexqv_dev / dx_undexqv_dev on torch device buffers; decode with and without the
    encoder's entry index gives the same text."""
    import torch
    text = QUIVA["lognormal_40"]
    t = torch.frombuffer(bytearray(text), dtype=torch.uint8).cuda()
    out = torch.empty(len(text), dtype=torch.uint8, device="cuda")
    st = ctx.qv_scan_dev(t.data_ptr(), len(text))
    cd = dx.lib.make_coding(st, False)
    hdr = b"\xaa\x55" + dx.lib.write_coding(cd, text[: text.index(b"/", 1)])
    nent = st.nentries
    body, lastw, offs = ctx.qv_encode_dev(t.data_ptr(), len(text), cd, False, 0,
                                          out.data_ptr() + 0, out.numel(), want_offsets=nent)
    want = orc.dexqv(text)
    got = hdr + out[:body].cpu().numpy().tobytes()
    assert got == want, first_diff(got, want)
    # decode the full image on the device, index given (offsets are relative to the image)
    img = torch.frombuffer(bytearray(want), dtype=torch.uint8).cuda()
    back = torch.empty(len(text) + 64, dtype=torch.uint8, device="cuda")
    m = ctx.undexqv_dev(img.data_ptr(), len(want), False, back.data_ptr(), back.numel(),
                        entry_off=offs + len(hdr))
    assert back[:m].cpu().numpy().tobytes() == text
    m2 = ctx.undexqv_dev(img.data_ptr(), len(want), False, back.data_ptr(), back.numel())
    assert back[:m2].cpu().numpy().tobytes() == text
    assert ctx.undexqv_size_dev(img.data_ptr(), len(want)) == len(text)


def test_two_shard_encode_equals_whole_file(ctx, orc):
    """The multi-GPU decomposition on one GPU: scan two shards with the carry, sum the
    statistics, encode each shard with the previous shard's last well; concatenation equals
    the reference output of the whole file."""
    import torch
    text = QUIVA["lognormal_40"]
    cut = text.index(b"\n@", len(text) // 2) + 1
    parts = [text[:cut], text[cut:]]
    dev = [torch.frombuffer(bytearray(p), dtype=torch.uint8).cuda() for p in parts]
    s0 = ctx.qv_scan_dev(dev[0].data_ptr(), len(parts[0]))
    carry = dx.Carry()
    carry.delchar, carry.subchar, carry.totchar = s0.delchar, s0.subchar, s0.totchar
    for i in range(256):
        carry.sub[i] = s0.sub_prefix[i]
    s1 = ctx.qv_scan_dev(dev[1].data_ptr(), len(parts[1]), carry)
    tot = dx.Stats()
    for k in range(6):
        for i in range(256):
            tot.hist[k][i] = s0.hist[k][i] + s1.hist[k][i]
    tot.totchar = s0.totchar + s1.totchar
    tot.nentries = s0.nentries + s1.nentries
    tot.delchar, tot.subchar = s1.delchar, s1.subchar
    cd = dx.lib.make_coding(tot, False)
    hdr = b"\xaa\x55" + dx.lib.write_coding(cd, text[: text.index(b"/", 1)])
    out = torch.empty(len(text), dtype=torch.uint8, device="cuda")
    b0, w0, _ = ctx.qv_encode_dev(dev[0].data_ptr(), len(parts[0]), cd, False, 0,
                                  out.data_ptr(), out.numel())
    piece0 = out[:b0].cpu().numpy().tobytes()
    b1, w1, _ = ctx.qv_encode_dev(dev[1].data_ptr(), len(parts[1]), cd, False, w0,
                                  out.data_ptr(), out.numel())
    piece1 = out[:b1].cpu().numpy().tobytes()
    want = orc.dexqv(text)
    got = hdr + piece0 + piece1
    assert got == want, first_diff(got, want)


@pytest.mark.parametrize("ent", MANIFEST, ids=lambda e: e["encoded"])
def test_golden(ctx, ent):
    """CUDA path against the committed reference outputs (no oracle involved)."""
    if "input" in ent:
        text = open(os.path.join(GOLD, ent["input"]), "rb").read()
    else:
        text = cases.all_cases()[ent["kind"]][ent["case"]]
    enc = open(os.path.join(GOLD, ent["encoded"]), "rb").read()
    kind = ent["kind"]
    if kind == "quiva":
        got = ctx.dexqv(text, lossy=bool(ent["flags"]))
        back = ctx.undexqv(enc)
    else:
        k = dx.ARROW if kind == "arrow" else dx.FASTA
        got = ctx.dexta(text, kind=k)
        back = ctx.undexta(enc, kind=k)
    assert got == enc, first_diff(got, enc)
    assert hashlib.sha256(back).hexdigest() == ent["decoded_sha256"]


def test_batched_reads(ctx, orc):
    """dx_compress_reads_dev / dx_uncompress_reads_dev (the DB loader form)."""
    import torch
    rng = np.random.default_rng(3)
    lens = np.array([1, 2, 3, 4, 5, 15, 16, 17, 63, 64, 65, 1000, 4097, 20000], dtype=np.int32)
    reads = [rng.choice(np.frombuffer(b"acgtACGTn", dtype=np.uint8), size=int(n)).tobytes()
             for n in lens]
    src = b"".join(reads)
    src_off = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    clen = (lens + 3) // 4
    dst_off = np.concatenate([[0], np.cumsum(clen)[:-1]]).astype(np.int64)
    d_src = torch.frombuffer(bytearray(src), dtype=torch.uint8).cuda()
    d_so, d_len = torch.from_numpy(src_off).cuda(), torch.from_numpy(lens).cuda()
    d_do = torch.from_numpy(dst_off).cuda()
    d_dst = torch.zeros(int(clen.sum()), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()        # torch fills on its own stream, the library runs on another
    ctx.compress_reads_dev(dx.FASTA, d_src.data_ptr(), d_so.data_ptr(), d_len.data_ptr(),
                           len(lens), d_dst.data_ptr(), d_do.data_ptr())
    ctx.sync()
    packed = d_dst.cpu().numpy().tobytes()
    import ctypes as C
    for i, r in enumerate(reads):
        buf = C.create_string_buffer(r, len(r) + 8)
        orc.lib().orc_number_read(buf)
        orc.lib().orc_compress_read(len(r), buf)
        assert packed[dst_off[i]: dst_off[i] + clen[i]] == buf.raw[: clen[i]], i
    d_back = torch.zeros(len(src), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.uncompress_reads_dev(dx.FASTA, False, d_dst.data_ptr(), d_do.data_ptr(), d_len.data_ptr(),
                             len(lens), d_back.data_ptr(), d_so.data_ptr())
    ctx.sync()
    back = d_back.cpu().numpy().tobytes()
    assert back == src.lower().replace(b"n", b"a")


def test_batched_arrows(ctx, orc):
    """The same batch form for pulse widths, as Load_All_Arrows uses the codec (DB.c:1556-1614):
    Number_Arrow + Compress_Read one way, Uncompress_Read + Letter_Arrow the other."""
    import ctypes as C
    import torch
    rng = np.random.default_rng(4)
    lens = np.array([1, 2, 3, 4, 5, 7, 8, 9, 31, 32, 33, 255, 256, 257, 1000, 4097, 30001], dtype=np.int32)
    reads = [rng.choice(np.frombuffer(b"1234", dtype=np.uint8), size=int(n)).tobytes() for n in lens]
    src = b"".join(reads)
    src_off = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    clen = (lens + 3) // 4
    dst_off = np.concatenate([[0], np.cumsum(clen)[:-1]]).astype(np.int64)
    d_src = torch.frombuffer(bytearray(src), dtype=torch.uint8).cuda()
    d_so, d_len = torch.from_numpy(src_off).cuda(), torch.from_numpy(lens).cuda()
    d_do = torch.from_numpy(dst_off).cuda()
    d_dst = torch.zeros(int(clen.sum()), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.compress_reads_dev(dx.ARROW, d_src.data_ptr(), d_so.data_ptr(), d_len.data_ptr(),
                           len(lens), d_dst.data_ptr(), d_do.data_ptr())
    ctx.sync()
    packed = d_dst.cpu().numpy().tobytes()
    for i, r in enumerate(reads):
        buf = C.create_string_buffer(r, len(r) + 8)
        orc.lib().orc_number_arrow(buf)
        orc.lib().orc_compress_read(len(r), buf)
        assert packed[dst_off[i]: dst_off[i] + clen[i]] == buf.raw[: clen[i]], i
    d_back = torch.zeros(len(src), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.uncompress_reads_dev(dx.ARROW, False, d_dst.data_ptr(), d_do.data_ptr(), d_len.data_ptr(),
                             len(lens), d_back.data_ptr(), d_so.data_ptr())
    ctx.sync()
    assert d_back.cpu().numpy().tobytes() == src
