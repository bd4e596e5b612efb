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


# ---- cutting ONE file into shards without knowing its entry starts --------------------------------------
#
# A .quiva entry is 6 lines (header + 5 streams, QV.c:751-798); a line that begins with '@' proves
# nothing ('@' is QV 31), so entry starts in the middle of a file are found by COUNTING lines:
#   1. rank r takes the byte range between two nominal cuts, each moved forward to the next line start
#      (raw_line_cut on a small window around the nominal position);
#   2. it counts the newlines of its range (on the GPU: the newline index of k_pred_slots);
#   3. the counts are exclusive-scanned over the ranks (one word per rank in the all-gather): the
#      global number of a rank's first line.  The rank skips lines_to_skip() lines to reach the next
#      line whose number is a multiple of 6 -- its first entry start -- and those skipped lines belong
#      to its predecessor's last entry;
#   4. the starts are gathered (one more word): a rank whose range lies inside one long entry has no
#      entry start and takes an empty shard (resolve_starts).

def nominal_cuts(total_bytes: int, nshards: int):
    """Equal byte targets (before they are moved to line starts)."""
    return [total_bytes * r // nshards for r in range(nshards + 1)]


def raw_line_cut(buf, pos: int) -> int:
    """First line start at or after byte `pos` of buf (bytes / uint8 array): pos itself if it is 0 or
    follows a newline, else one past the next newline; len(buf) if there is none."""
    a = np.frombuffer(buf, dtype=np.uint8) if isinstance(buf, (bytes, bytearray, memoryview)) else np.asarray(buf)
    n = len(a)
    if pos <= 0:
        return 0
    if pos >= n:
        return n
    if a[pos - 1] == 10:
        return pos
    nl = np.flatnonzero(a[pos:] == 10)
    return n if len(nl) == 0 else pos + int(nl[0]) + 1


def lines_to_skip(line_counts, rank: int, lines_per_entry: int = 6) -> int:
    """line_counts[q] = newlines in rank q's raw range.  -> how many whole lines at the start of rank's
    range still belong to the previous rank's last entry."""
    first_line = int(np.asarray(line_counts[:rank], dtype=np.int64).sum())
    return (-first_line) % lines_per_entry


def entry_aligned_start(newline_pos, raw_start: int, skip: int, raw_end: int) -> int:
    """newline_pos: ascending offsets of the newlines inside [raw_start, raw_end).  -> offset of the
    rank's first entry start, or -1 when no entry starts inside the range (it lies inside one long
    entry, or is empty): such a rank gets an empty shard, see resolve_starts."""
    if raw_start >= raw_end:
        return -1
    if skip == 0:
        return raw_start
    if len(newline_pos) < skip:
        return -1
    p = int(newline_pos[skip - 1]) + 1
    return p if p < raw_end else -1


def resolve_starts(starts, total_bytes: int):
    """starts[r] as gathered from entry_aligned_start (one word per rank).  A rank without an entry
    start takes its successor's, so shard r = [out[r], out[r+1]) is empty for it and the shards stay
    contiguous and entry-aligned.  -> world + 1 offsets, out[0] == 0, out[-1] == total_bytes."""
    out = [int(x) for x in starts] + [int(total_bytes)]
    for r in range(len(starts) - 1, -1, -1):
        if out[r] < 0:
            out[r] = out[r + 1]
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
