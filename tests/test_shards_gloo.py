"""The N > 1 path on CPU: two processes over gloo run the host side of the sharded dexqv
(dextractor_b200/shards.py, the code bench.py runs under torchrun) on two shards of one file.

Each rank computes the statistics of its own shard -- rank 0 with the oracle's scan (it owns the
prefix that fixes the run characters), rank 1 with a plain numpy count that uses rank 0's run
characters from its first entry on, which is what dx_qv_scan_dev does with a carry -- the rows are
all-gathered, and every rank must end up with the whole file's statistics, the whole file's
coding header, and the right well to encode its first entry against."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dextractor_b200 import lib as dxl
from dextractor_b200 import shards, synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _entries(text: bytes):
    """-> list of (start offset, well, [del, tag, ins, mrg, sub] lines)"""
    lines = text.split(b"\n")[:-1]
    out, off = [], 0
    for k in range(0, len(lines), 6):
        well = int(lines[k].split(b"/")[1])
        out.append((off, well, lines[k + 1:k + 6]))
        off += sum(len(x) + 1 for x in lines[k:k + 6])
    return out


def _count_shard(ents, delchar, subchar):
    """Statistics of a shard that is NOT the first: histograms of the four QV lines and run-length
    histograms with the given run characters from the first entry on (QV.c:702-724; the '+1' start of
    the run histograms belongs to the sum, so it is not added here)."""
    st = dxl.Stats()
    h = np.ctypeslib.as_array(st.hist)
    tot = 0
    for _, _, (dele, _tag, ins, mrg, sub) in ents:
        for k, line in ((0, dele), (1, ins), (2, mrg), (3, sub)):
            h[k] += np.bincount(np.frombuffer(line, dtype=np.uint8), minlength=256).astype(np.uint64)
        for k, line, rc in ((4, dele, delchar), (5, sub, subchar)):
            if rc < 0:
                continue
            a = np.frombuffer(line, dtype=np.uint8)
            pos = np.flatnonzero(a != rc)
            runs = np.diff(np.concatenate([[-1], pos])) - 1
            if len(a) and (len(pos) == 0 or pos[-1] != len(a) - 1):
                runs = np.append(runs, len(a) - 1 - (pos[-1] if len(pos) else -1))
            h[k] += np.bincount(np.minimum(runs, 255), minlength=256).astype(np.uint64)
        tot += len(dele)
    st.totchar, st.nentries = tot, len(ents)
    return st


def _worker(rank, world, port, text, q):
    from oracle import orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ents = _entries(text)
        cuts = shards.split_entries([e[0] for e in ents], len(text), world)
        a, b, lo, hi = cuts[rank]
        mine = ents[a:b]
        # rank 0 resolves the run characters in its own prefix and hands them on
        rc = torch.zeros(2, dtype=torch.int64)
        if rank == 0:
            s0 = orc.Stats()
            assert orc.lib().orc_qv_scan(text[lo:hi], hi - lo, s0) == 0
            rc[:] = torch.tensor([s0.delchar, s0.subchar])
        dist.broadcast(rc, 0)
        delchar, subchar = int(rc[0]), int(rc[1])
        if rank == 0:
            st = dxl.Stats()
            h = np.ctypeslib.as_array(st.hist)
            for k, f in enumerate(("del_", "ins", "mrg", "sub", "delrun", "subrun")):
                h[k] = np.ctypeslib.as_array(getattr(s0, f))
            # the oracle's run histograms start at 1 (QV.c:934-935): that belongs to the sum, once
            h[4] -= 1; h[5] -= 1
            st.totchar, st.nentries = s0.totchar, s0.nentries
        else:
            st = _count_shard(mine, delchar, subchar)
        last_well = mine[-1][1] if mine else 0
        row = torch.from_numpy(shards.pack_stats(st, last_well))
        rows = shards.exchange(row, world).numpy()
        tot, lwell_in = shards.merge_stats(rows, rank, (delchar, subchar))
        cd = dxl.make_coding(tot, False)                             # ... by dx_qv_make_coding
        prefix = text[: text.index(b"/")]
        hdr = dxl.write_coding(cd, prefix)
        q.put((rank, bytes(np.ctypeslib.as_array(tot.hist).tobytes()), int(tot.totchar), int(tot.nentries),
               hdr, lwell_in, a))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_ranks_over_gloo_agree_with_the_whole_file(orc):
    rng = np.random.default_rng(21)
    lengths = [int(x) for x in rng.integers(3000, 9000, size=60)]     # > 100 000 positions in shard 0
    text = synth.make_quiva(21, lengths)
    whole = orc.Stats()
    assert orc.lib().orc_qv_scan(text, len(text), whole) == 0
    enc = orc.dexqv(text)
    ents = _entries(text)

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, text, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=150) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    want_hist = np.stack([np.ctypeslib.as_array(getattr(whole, f))
                          for f in ("del_", "ins", "mrg", "sub", "delrun", "subrun")]).astype(np.uint64)
    want_hist[4:6] -= 1                        # the oracle's run buckets start at 1 (QV.c:934-935)
    for rank, hist, totchar, nentries, hdr, lwell_in, first in got:
        assert np.array_equal(np.frombuffer(hist, dtype=np.uint64).reshape(6, 256), want_hist), rank
        assert totchar == whole.totchar and nentries == whole.nentries == len(ents)
        assert enc[2:2 + len(hdr)] == hdr, "coding header differs from the whole file's"
        assert lwell_in == (ents[first - 1][1] if first > 0 else 0)
    # the hand-off value is what the reference encodes the first entry of shard 1 against
    first1 = got[1][6]
    offs = (orc.C.c_int64 * (len(ents) + 1))()
    assert orc.lib().orc_dexqv_offsets(enc, len(enc), offs, len(ents) + 1) == len(ents)
    p = offs[first1]
    delta = 0
    while enc[p] == 255:
        delta += 255; p += 1
    delta += enc[p]
    assert delta == ents[first1][1] - got[1][5]


def test_split_entries_balances_bytes_and_handles_more_shards_than_entries():
    starts = [0, 100, 250, 900, 1000]
    cuts = shards.split_entries(starts, 1200, 2)
    assert cuts == [(0, 3, 0, 900), (3, 5, 900, 1200)]
    cuts = shards.split_entries(starts, 1200, 8)
    assert cuts[0][0] == 0 and cuts[-1][1] == 5 and cuts[-1][3] == 1200
    assert all(c[1] == n[0] and c[3] == n[2] for c, n in zip(cuts, cuts[1:]))
    assert sum(b - a for a, b, _, _ in cuts) == 5
    rows = np.zeros((3, shards.STAT_WORDS), dtype=np.int64)
    rows[0, -2:] = (4, 77); rows[1, -2:] = (0, 0); rows[2, -2:] = (2, 99)
    assert shards.merge_stats(rows, 2, (50, 63))[1] == 77          # skips the empty shard
    assert shards.merge_stats(rows, 0, (50, 63))[1] == 0
    assert list(shards.shard_offsets([10, 0, 5])) == [0, 10, 10]


def test_split_entries_properties():
    """For any entry layout and shard count: the shards are contiguous, in order, cover every entry
    and every byte exactly once, and a cut never falls inside an entry."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=200, deadline=None)
    @given(st.lists(st.integers(min_value=1, max_value=5000), min_size=0, max_size=60),
           st.integers(min_value=1, max_value=9))
    def check(sizes, nshards):
        starts = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64) if sizes else np.zeros(0, np.int64)
        total = int(sum(sizes))
        cuts = shards.split_entries(starts, total, nshards)
        assert len(cuts) == nshards
        assert cuts[0][0] == 0 and cuts[-1][1] == len(sizes) and cuts[-1][3] == total
        assert cuts[0][2] == (0 if sizes else total)
        for (a, b, lo, hi), nxt in zip(cuts, cuts[1:] + [None]):
            assert a <= b and lo <= hi
            assert lo == (int(starts[a]) if a < len(sizes) else total)      # cuts sit on entry starts
            assert hi - lo == int(sum(sizes[a:b]))
            if nxt is not None:
                assert nxt[0] == b and nxt[2] == hi
        assert list(shards.shard_offsets([c[3] - c[2] for c in cuts])) == [c[2] for c in cuts]

    check()


# ---- cutting one file by counting lines (no entry index) ----------------------------------------------------

def _line_count_cuts(text: bytes, world: int):
    """what every rank would compute, here in one process: raw cuts at line starts, newline counts,
    skipped lines -> entry-aligned shard starts (the last element is len(text))"""
    a = np.frombuffer(text, dtype=np.uint8)
    nom = shards.nominal_cuts(len(text), world)
    raw = [shards.raw_line_cut(a, p) for p in nom]
    nls = [np.flatnonzero(a[raw[r]:raw[r + 1]] == 10) + raw[r] for r in range(world)]
    counts = [len(x) for x in nls]
    starts = []
    for r in range(world):
        k = shards.lines_to_skip(counts, r)
        starts.append(shards.entry_aligned_start(nls[r], raw[r], k, raw[r + 1]))
    return shards.resolve_starts(starts, len(text))


def _at_sign_hook(i, streams):
    # quality lines that begin with '@' (QV 31) and look like headers to a pattern matcher
    for k in (0, 2, 3, 4):
        streams[k][0] = ord("@")
    if len(streams[2]) > 12:
        streams[2][:12] = np.frombuffer(b"@m/1/0_9 RQ=", dtype=np.uint8)


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_line_count_cuts_land_on_entry_starts(world):
    rng = np.random.default_rng(world)
    for seed, lengths in ((1, synth.draw_lengths(rng, 40)), (2, [1, 2, 3, 1, 1, 5000, 1, 1]),
                          (3, [7] * 5), (4, [30000])):
        text = synth.make_quiva(seed, lengths, stream_hook=_at_sign_hook)
        true_starts = {e[0] for e in _entries(text)} | {len(text)}
        cuts = _line_count_cuts(text, world)
        assert cuts[0] == 0 and cuts[-1] == len(text)
        assert all(c in true_starts for c in cuts), (world, seed)
        assert all(cuts[i] <= cuts[i + 1] for i in range(world)), (world, seed)
        # balanced up to one entry: no shard start is further from its nominal cut than one entry
        longest = 6 * (max(lengths) + 1) + 120
        nom = shards.nominal_cuts(len(text), world)
        assert all(0 <= cuts[r] - nom[r] <= longest for r in range(world) if cuts[r] < cuts[r + 1]), (world, seed)


def _cut_worker(rank, world, port, text, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        a = np.frombuffer(text, dtype=np.uint8)
        nom = shards.nominal_cuts(len(text), world)
        lo, hi = shards.raw_line_cut(a, nom[rank]), shards.raw_line_cut(a, nom[rank + 1])
        nl = np.flatnonzero(a[lo:hi] == 10) + lo                # on the GPU: the newline index
        mine = torch.tensor([len(nl)], dtype=torch.int64)
        counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(counts, mine)
        counts = [int(c) for c in counts]
        start = shards.entry_aligned_start(nl, lo, shards.lines_to_skip(counts, rank), hi)
        starts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(starts, torch.tensor([start], dtype=torch.int64))
        q.put((rank, shards.resolve_starts([int(x) for x in starts], len(text))[rank]))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_three_ranks_over_gloo_cut_one_file_at_entry_starts():
    rng = np.random.default_rng(9)
    text = synth.make_quiva(5, synth.draw_lengths(rng, 25), stream_hook=_at_sign_hook)
    world, port = 3, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_cut_worker, args=(r, world, port, text, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert [got[r] for r in range(world)] == _line_count_cuts(text, world)[:world]
    true_starts = {e[0] for e in _entries(text)}
    assert all(got[r] in true_starts for r in range(world))
