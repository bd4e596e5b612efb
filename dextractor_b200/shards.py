"""Host-side logic of the sharded dexqv path (one process per GPU).

A .quiva file is cut at entry boundaries into contiguous shards, one per rank.  Entries are
independent once the coding scheme is fixed, so the only things that cross rank boundaries are
  * the six 256-bin histograms of QVcoding_Scan (QV.c:922-1023), which are additive -- the run
    characters and the first ~100 000 positions that fix them (QV.c:993-1015) belong to rank 0, which
    hands the two characters to the other ranks before they count run lengths;
  * the last well number of a shard: the first entry of the next shard stores its well as a delta
    against it (dexqv.c:128-135).
Both travel in ONE all-gather of 6*256+3 int64 per rank; every rank then forms the same sums and
builds the same code tables on its host (Create_QVcoding is deterministic, QV.c:1029-1169).

Nothing here touches a GPU: the functions take and return host arrays and work with any
torch.distributed backend (nccl on the B200s, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np

from .lib import Stats

STAT_WORDS = 6 * 256 + 3          # histograms, totchar, nentries, last well of the shard


def split_entries(entry_starts, total_bytes: int, nshards: int):
    """Cut [0, total_bytes) at entry starts into nshards contiguous ranges balanced by BYTES (not by
    entry count).  -> list of (first_entry, end_entry, byte_start, byte_end); a shard may be empty
    when there are fewer entries than shards."""
    starts = np.asarray(entry_starts, dtype=np.int64)
    n = len(starts)
    cuts = [0]
    for r in range(1, nshards):
        target = total_bytes * r // nshards
        k = int(np.searchsorted(starts, target, side="left"))     # first entry at or after the target
        cuts.append(max(cuts[-1], min(k, n)))
    cuts.append(n)
    out = []
    for r in range(nshards):
        a, b = cuts[r], cuts[r + 1]
        lo = int(starts[a]) if a < n else total_bytes
        hi = int(starts[b]) if b < n else total_bytes
        out.append((a, b, lo, hi))
    return out


def pack_stats(st: Stats, last_well: int) -> np.ndarray:
    """One rank's contribution to the all-gather."""
    row = np.empty(STAT_WORDS, dtype=np.int64)
    row[: 6 * 256] = np.ctypeslib.as_array(st.hist).reshape(-1).astype(np.int64)
    row[6 * 256:] = (int(st.totchar), int(st.nentries), int(last_well))
    return row


def merge_stats(rows: np.ndarray, rank: int, run_chars) -> tuple[Stats, int]:
    """rows: [world, STAT_WORDS] as gathered.  -> (statistics of the whole file, the well number the
    first entry of this rank's shard is a delta against: the last well of the nearest earlier
    non-empty shard, 0 for the first)."""
    rows = np.asarray(rows, dtype=np.int64).reshape(-1, STAT_WORDS)
    tot = Stats()
    np.ctypeslib.as_array(tot.hist)[:] = rows[:, : 6 * 256].sum(axis=0).reshape(6, 256).astype(np.uint64)
    tot.totchar = int(rows[:, 6 * 256].sum())
    tot.nentries = int(rows[:, 6 * 256 + 1].sum())
    tot.delchar, tot.subchar = int(run_chars[0]), int(run_chars[1])
    lwell = 0
    for r in range(rank - 1, -1, -1):
        if rows[r, 6 * 256 + 1] > 0:               # an empty shard has no well to hand on
            lwell = int(rows[r, 6 * 256 + 2])
            break
    return tot, lwell


def shard_offsets(sizes) -> np.ndarray:
    """Exclusive scan of the shards' compressed sizes: where each rank's bytes go behind the file
    header (the implicit file position of the reference's fwrite calls)."""
    s = np.asarray(sizes, dtype=np.int64)
    return np.concatenate([[0], np.cumsum(s)[:-1]])


def exchange(row, world: int, device=None):
    """All-gather one row per rank (torch.distributed must be initialised when world > 1).
    row: torch int64 tensor of STAT_WORDS on `device`.  -> [world, STAT_WORDS] tensor on `device`."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return row.view(1, -1)
    out = torch.empty(world * row.numel(), dtype=torch.int64, device=row.device if device is None else device)
    dist.all_gather_into_tensor(out, row)
    return out.view(world, -1)
