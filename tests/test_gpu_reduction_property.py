"""Property test of the fused kernel's reduction pre-pass (knapsack.cu) on > 10^5 blocks, at the margins
its tolerances are about: the pre-pass may drop an item only if NO optimal solution contains it, so optimum (bit for
bit), column and sums must equal the oracle's plain DP (gap_func.py:160-181) -- with the pre-pass on and off.

  near_tie      profits = e0 * w * (1 + d), |d| from 1e-12 to 1e-6: efficiencies a hair apart, far closer than a
                bucket's width (2^-5 relative) and around the 1e-9 / 2e-6 margins of the survivor test
  bucket_edge   efficiencies exactly ON the edges of the efficiency buckets (2^k * (1 + j/32)) and one ulp to either
                side: the float division that bins an item may round across the edge
  pisinger      reduced costs with 0..4 decimals: GAP_KnapPisinger's scale factors 1..10^4, integer profits, many ties
  pass1         blocks whose optimum contains an item OUTSIDE the most efficient ones pass 0 runs over (a big item
                that beats the best fill of equal small ones): the survivor test must keep it and pass 1 must run
  pass1_hair    the same with the big item's profit within 1e-12 .. 1e-6 (relative, either sign) of the best fill
  fallback      more than 224 items tie at the top efficiency: the pre-pass gives up and the plain DP runs
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PER_LAUNCH = 250          # fused shapes: <= 256 blocks per shard
THREADS = os.cpu_count() or 1


def _gap_weights(rng, size):
    return np.floor(1 - 10 * np.log(rng.uniform(1e-9, 1, size=size))).astype(np.int32)


def _family(kind, rng, m):
    """m blocks of one family: (bcs, w, cap, rc, big_mask or None)."""
    n = int(rng.integers(1040, 1300))   # (the pre-pass needs its histograms to fit the item arrays: >= 1031 columns per block)
    N = m * n
    bcs = (np.arange(m + 1) * n).astype(np.int32)
    big = None
    if kind == "near_tie":
        w = _gap_weights(rng, N)
        cap = rng.integers(40, 220, size=m).astype(np.int32)
        e0 = np.repeat(rng.uniform(0.5, 200.0, size=m), n)
        d = 10.0 ** rng.uniform(-12, -6, size=N) * rng.choice([-1.0, 1.0], size=N)
        p = e0 * w * (1.0 + d)
        spread = rng.random(N) < 0.3
        p = np.where(spread, rng.uniform(0, 2, size=N) * e0 * w, p)
    elif kind == "bucket_edge":
        w = rng.integers(1, 17, size=N).astype(np.int32)
        cap = rng.integers(30, 160, size=m).astype(np.int32)
        k = np.repeat(rng.integers(-3, 9, size=m), n) + rng.integers(0, 2, size=N)
        j = rng.integers(0, 32, size=N)
        eff = np.ldexp(1.0 + j / 32.0, k)                       # exactly a bucket edge
        ulp = rng.integers(-1, 2, size=N)
        eff = np.where(ulp > 0, np.nextafter(eff, np.inf), np.where(ulp < 0, np.nextafter(eff, -np.inf), eff))
        p = eff * w
    elif kind == "pisinger":
        w = _gap_weights(rng, N)
        cap = rng.integers(40, 200, size=m).astype(np.int32)
        dec = np.repeat(rng.integers(0, 5, size=m), n)
        p = np.round(rng.uniform(0, 60, size=N) / np.maximum(w, 1) ** 0.5, 4)
        p = np.round(p * 10.0 ** dec) / 10.0 ** dec               # at most `dec` decimals -> scale 10^dec
        p = np.maximum(p, 0.0)
    elif kind in ("pass1", "pass1_hair"):
        w = np.empty(N, np.int32)
        p = np.empty(N)
        cap = rng.integers(60, 200, size=m).astype(np.int32)
        big = np.zeros(N, bool)
        for b in range(m):
            C = int(cap[b])
            ws = int(rng.choice([q for q in (5, 6, 7, 8, 9, 11) if C % q >= 2]))
            es = float(rng.integers(4, 40))                        # the small items' efficiency (exact products)
            s = slice(b * n, (b + 1) * n)
            wb = np.full(n, ws, np.int32)
            pb = np.full(n, es * ws)
            lo = np.ones(n, bool)                                  # low-efficiency noise the pre-pass drops, except ...
            lo[rng.choice(n, size=int(rng.integers(60, 160)), replace=False)] = False   # ... 60-160 equal small items (<= 224: no fallback)
            wb[lo] = rng.integers(1, 30, size=int(lo.sum()))
            pb[lo] = rng.uniform(0.05, 0.6, size=int(lo.sum())) * es * wb[lo]
            r = C % ws
            kbig = rng.choice(n, size=3, replace=False)
            for q, idx in enumerate(kbig):
                wb[idx] = C - q                                     # fills the capacity (almost) alone
                if kind == "pass1":
                    pb[idx] = es * (C - r) + es * r * rng.uniform(0.2, 0.9)   # above the best fill of small items, efficiency below es
                else:  # the best fill of small items, es * (C - r), a hair more or a hair less: the survivor test's knife edge
                    pb[idx] = es * (C - r) * (1.0 + 10.0 ** rng.uniform(-12, -6) * rng.choice([-1.0, 1.0]))
            w[s], p[s] = wb, pb
            big[b * n + kbig] = True
    else:                                                          # "fallback"
        w = np.where(rng.random(N) < 0.8, 1, rng.integers(1, 6, size=N)).astype(np.int32)
        cap = rng.integers(120, 200, size=m).astype(np.int32)
        e0 = np.repeat(rng.uniform(1.0, 50.0, size=m), n)
        p = e0 * w * (1.0 + np.where(rng.random(N) < 0.5, 0.0, rng.integers(-2, 3, size=N) * 2.0 ** -50))
    rc = -p
    rc = np.where(rng.random(N) < 0.15, rng.uniform(0, 3, size=N), rc)   # columns that are no items
    if big is not None:
        rc[big] = -p[big]
    return bcs, w, cap, rc, big


def _run(pricer, oracle, bcs, w, cap, rc, oc, alpha, mode):
    pricer.setKnapsackBlocks(bcs, w, cap, mode)   # (first: the block structure bounds the objective's length from below)
    pricer.setObjective(oc)
    out = []
    for no_prune in (0, 1):
        pricer.set_option("dp_no_prune", no_prune)
        res = pricer.solveBlocks(rc, alpha, redCostEps=-np.inf)
        out.append((res.blockObj.copy(), res.blockNnz.copy(), res.blockRedCost.copy(), res.colStart[:res.nCols + 1].copy(),
                    res.colInd[:res.c.nInd].copy(), res.cells, res.cellsEvaluated))
    pricer.set_option("dp_no_prune", 0)
    ref = oracle.solve_blocks(bcs, w, cap, rc, oc, alpha, red_cost_eps=-np.inf, profit_mode=mode, n_threads=THREADS)
    return out, ref


@pytest.mark.parametrize("kind,launches", [("near_tie", 130), ("bucket_edge", 100), ("pisinger", 100), ("pass1", 40),
                                           ("pass1_hair", 40), ("fallback", 40)])
def test_reduction_property(oracle, pricer, kind, launches):
    from dip_b200.pricer import PROFIT_FP64, PROFIT_PISINGER_INT
    mode = PROFIT_PISINGER_INT if kind == "pisinger" else PROFIT_FP64
    rng = np.random.default_rng({"near_tie": 11, "bucket_edge": 12, "pisinger": 13, "pass1": 14, "fallback": 15, "pass1_hair": 16}[kind])
    cells = evaluated = blocks = with_big = 0
    for it in range(launches):
        m = PER_LAUNCH
        bcs, w, cap, rc, big = _family(kind, rng, m)
        oc = rng.integers(1, 50, size=len(w)).astype(np.float64)
        alpha = rng.uniform(-5, 5, size=m)
        (on, off), ref = _run(pricer, oracle, bcs, w, cap, rc, oc, alpha, mode)
        assert pricer.dpGeometry()["fused_tail"] == 1
        ref_start = np.concatenate([[0], np.cumsum(ref.col_nnz)])
        ref_ind = np.concatenate([np.asarray(c, np.int32) for c in ref.col_ind]) if m else np.zeros(0, np.int32)
        for tag, (z, nnz, rcb, cs, ci, cl, ev) in (("pre-pass on", on), ("pre-pass off", off)):
            np.testing.assert_array_equal(z, ref.z, err_msg=f"{kind} launch {it} {tag}: optimum")
            np.testing.assert_array_equal(nnz, ref.col_nnz, err_msg=f"{kind} launch {it} {tag}: column sizes")
            np.testing.assert_array_equal(cs, ref_start, err_msg=f"{kind} launch {it} {tag}")
            np.testing.assert_array_equal(ci, ref_ind, err_msg=f"{kind} launch {it} {tag}: columns")
            np.testing.assert_array_equal(rcb, ref.col_red_cost, err_msg=f"{kind} launch {it} {tag}: reduced costs")
            assert cl == ref.cells
        assert off[6] == off[5]                       # switched off: every reference cell is evaluated
        if kind != "fallback":
            assert on[6] < on[5], f"{kind} launch {it}: the pre-pass did not engage"
        cells += on[5]
        evaluated += on[6]
        blocks += m
        if big is not None:
            with_big += sum(bool(big[np.asarray(c, np.int64)].any()) for c in ref.col_ind)
    assert blocks == launches * PER_LAUNCH
    if kind == "pass1":
        assert with_big > 0.9 * blocks                # the optimum needs an item pass 0 never saw ...
        assert evaluated < 0.9 * cells                # ... and the pre-pass engaged (no fallback to the plain DP)
    if kind == "fallback":
        assert evaluated == cells                     # > 224 items at the top efficiency: the plain DP ran
    print(f"{kind}: {blocks} blocks, {evaluated / max(cells, 1):.3f} of the reference's cells evaluated")


def test_reduction_property_block_count():
    """The six families together: 112 500 blocks."""
    assert (130 + 100 + 100 + 40 + 40 + 40) * PER_LAUNCH >= 100_000
