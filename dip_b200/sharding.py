"""Host-side logic of the block-sharded multi-GPU round: which blocks a rank
prices, and how rank 0 merges the shards' packed results.  Pure host code (numpy + the host-only
C-ABI function dipgpu_unpack_result); the NCCL plumbing lives in bench.py."""
from __future__ import annotations

import numpy as np

from .pricer import RoundResult, unpack_result


def block_range(m: int, world: int, rank: int, work: np.ndarray | None = None) -> tuple[int, int]:
    """Contiguous block range [first, first+count) of `rank`.  With `work` (per-block cost, e.g.
    n_b * (cap_b + 1)) the cut points balance the prefix sums; otherwise block counts."""
    if world <= 1:
        return 0, m
    if work is None:
        per = (m + world - 1) // world
        first = min(rank * per, m)
        return first, max(0, min(per, m - first))
    w = np.asarray(work, np.float64)
    pre = np.concatenate([[0.0], np.cumsum(w)])
    tot = pre[-1]
    cuts = [int(np.searchsorted(pre, tot * r / world, side="left")) for r in range(world + 1)]
    cuts[0], cuts[-1] = 0, m
    for r in range(1, world + 1):
        cuts[r] = max(cuts[r], cuts[r - 1])
    return cuts[rank], cuts[rank + 1] - cuts[rank]


def merge_shards(packed: list[np.ndarray], m: int, cap_ind: int) -> RoundResult:
    """Decode every shard (block ids are global inside the packed buffers) and concatenate the kept
    columns in block order; mostNegReducedCost is re-added in block order on the host, which is the
    reference's summation order (DecompAlgo.cpp:5229-5232)."""
    total = RoundResult(m, m, cap_ind)
    parts = []
    for buf in packed:
        part = RoundResult(m, m, cap_ind)
        # share the per-block arrays so each shard fills its own global slots
        for name in ("blockRedCost", "blockMostNegRC", "blockOrigCost", "blockObj", "blockNnz"):
            setattr(part, name, getattr(total, name))
            setattr(part.c, name, getattr(total, name).ctypes.data)
        unpack_result(buf, m, cap_ind, result=part)
        parts.append(part)
    k = 0
    pos = 0
    cells = 0
    for part in sorted(parts, key=lambda p: int(p.colBlock[0]) if p.nCols else 1 << 30):
        n = part.nCols
        ni = int(part.colStart[n]) if n else 0
        total.colBlock[k:k + n] = part.colBlock[:n]
        total.colRedCost[k:k + n] = part.colRedCost[:n]
        total.colOrigCost[k:k + n] = part.colOrigCost[:n]
        total.colStart[k:k + n] = part.colStart[:n] + pos
        total.colInd[pos:pos + ni] = part.colInd[:ni]
        k += n
        pos += ni
        cells += part.cells
    total.colStart[k] = pos
    total.c.nCols = k
    total.c.nInd = pos
    total.c.cells = cells
    s = 0.0
    for b in range(m):
        s += float(total.blockMostNegRC[b])
    total.c.mostNegReducedCost = s
    return total
