"""Test helper: build a shard's packed result buffer (layout of dip_b200/csrc/common.cuh:
ResultHeader + ResultLayout) from oracle outputs, so the host-side shard/merge logic can be
exercised without a GPU."""
import struct

import numpy as np

MAGIC = 0x52504944


def align_up(x, a):
    return (x + a - 1) // a * a


def pack_shard(first, red_cost, orig_cost, obj, nnz, keep, cols, cells, ind_cap):
    m = len(red_cost)
    kept = [lb for lb in range(m) if keep[lb]]
    starts = [0]
    ind = []
    for lb in kept:
        ind.extend(int(v) for v in cols[lb])
        starts.append(len(ind))
    hdr = struct.pack("<IiiiqqqQQQ", MAGIC, m, first, len(kept), len(ind), int(cells), int(ind_cap), 0, 0, 0)
    assert len(hdr) == 64
    kept_start = np.zeros(m + 1, np.int64)
    kept_start[:len(starts)] = starts
    kept_block = np.zeros(m, np.int32)
    kept_block[:len(kept)] = kept
    body = b"".join([
        np.asarray(red_cost, np.float64).tobytes(),
        np.minimum(0.0, np.asarray(red_cost, np.float64)).tobytes(),
        np.asarray(orig_cost, np.float64).tobytes(),
        np.asarray(obj, np.float64).tobytes(),
        kept_start.tobytes(),
        np.asarray(nnz, np.int32).tobytes(),
        kept_block.tobytes(),
    ])
    buf = hdr + body
    buf += b"\0" * (align_up(len(buf), 16) - len(buf))
    indarr = np.zeros(ind_cap, np.int32)
    indarr[:len(ind)] = ind
    buf += indarr.tobytes()
    buf += b"\0" * (align_up(len(buf), 16) - len(buf))
    return np.frombuffer(buf, np.uint8).copy()
