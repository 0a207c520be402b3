"""The multi-device context (dipgpu_create_multi): every partitioned entry point returns, field by field, what
the one-GPU context returns on the same inputs.  Runs on a one-GPU box with the same ordinal listed several
times (the shards then share the GPU: same code paths, same kernels); with two or more GPUs the real device
list is tested as well, including the NVLink (peer) exchange of the device-resident entry point."""
import ctypes as C

import numpy as np
import pytest

from dip_b200 import capi
from dip_b200 import instances as I

pytestmark = pytest.mark.gpu


def _device_lists():
    import torch
    n = torch.cuda.device_count()
    lists = [[0, 0], [0, 0, 0]]
    if n >= 2:
        lists.append(list(range(min(n, 8))))
    return lists


def _same_round(a, b, m):
    assert a.nCols == b.nCols
    np.testing.assert_array_equal(a.blockObj[:m], b.blockObj[:m])
    np.testing.assert_array_equal(a.blockRedCost[:m], b.blockRedCost[:m])
    np.testing.assert_array_equal(a.blockMostNegRC[:m], b.blockMostNegRC[:m])
    np.testing.assert_array_equal(a.blockOrigCost[:m], b.blockOrigCost[:m])
    np.testing.assert_array_equal(a.blockNnz[:m], b.blockNnz[:m])
    n = a.nCols
    np.testing.assert_array_equal(a.colBlock[:n], b.colBlock[:n])
    np.testing.assert_array_equal(a.colRedCost[:n], b.colRedCost[:n])
    np.testing.assert_array_equal(a.colOrigCost[:n], b.colOrigCost[:n])
    np.testing.assert_array_equal(a.colStart[:n + 1], b.colStart[:n + 1])
    np.testing.assert_array_equal(a.colInd[:a.c.nInd], b.colInd[:b.c.nInd])
    assert a.mostNegReducedCost == b.mostNegReducedCost
    # (cellsEvaluated is a diagnostic: whether the reduction pre-pass engages depends on the shard's widest block)
    assert a.cells == b.cells and 0 <= b.cellsEvaluated <= b.cells


def _gap(m=12, n=300, seed=5, cls="d"):
    g = (I.gap_class_d if cls == "d" else I.gap_class_e)(m=m, n=n, seed=seed)
    model = I.gap_pricing_model(g)
    rng = np.random.default_rng(seed)
    br = I.gap_branch_rows(model, rng)
    duals = [I.gap_duals(model, rng, branch_rows=br) for _ in range(6)]
    model.dual_support = I.gap_dual_support(model, br)
    return model, duals


@pytest.fixture(scope="module")
def single():
    from dip_b200.pricer import GpuPricer
    model, duals = _gap()
    p = GpuPricer(0)
    p.setModel(model)
    yield p, model, duals
    p.close()


@pytest.mark.parametrize("devs", _device_lists())
@pytest.mark.parametrize("support", [False, True])
@pytest.mark.parametrize("pinned", [False, True])
@pytest.mark.parametrize("force", [True, False])
def test_sharded_round_equals_single(single, devs, support, pinned, force):
    """force: option shard_rounds = number of devices (the blocks ARE partitioned); else the default policy -- this
    model's blocks are co-resident on one GPU, so the single round stays on the primary device."""
    from dip_b200.pricer import GpuPricer, RoundResult
    p1, model, duals = single
    p1.setDualSupport(model.dual_support if support else None)
    with GpuPricer(devs) as pm:
        if force:
            pm.set_option("shard_rounds", len(devs))
        pm.setModel(model)
        assert pm.deviceCount()[0] == len(devs)
        rng = pm.shardRanges()
        assert rng[0][0] == 0 and sum(c for _, c in rng) == model.m
        if force:
            assert all(rng[d][0] + rng[d][1] == rng[d + 1][0] for d in range(len(devs) - 1))
            assert len(devs) == 1 or rng[0][1] < model.m
        else:
            assert rng[0][1] == model.m
        if support:
            pm.setDualSupport(model.dual_support)
        pin = capi.PinnedArray((model.u_len,), np.float64) if pinned else None
        for u in duals[:3]:
            for phase in (capi.PHASE_PRICE2, capi.PHASE_PRICE1):
                a, rca = p1.priceRound(u, phase=phase, want_redcost=True, result=RoundResult(model.m, model.m, model.N))
                rb = RoundResult(model.m, model.m, model.N)
                if pinned:
                    pin.array[:] = u
                    rcb = np.empty(model.N)
                    pm._check(pm._lib.dipgpu_price_round(pm._ctx, pin.array.ctypes.data, model.u_len, phase, 1e-4, C.byref(rb.c), capi.ptr(rcb)))
                else:
                    rb, rcb = pm.priceRound(u, phase=phase, want_redcost=True, result=rb)
                _same_round(a, rb, model.m)
                np.testing.assert_array_equal(rca, rcb)
                # stage (a) alone, column-partitioned
                np.testing.assert_array_equal(pm.calcRedCost(u, phase), rca)
        # the expansion of a partitioned round: every shard expands the columns it kept
        a = p1.priceRound(duals[0], result=RoundResult(model.m, model.m, model.N))
        ea = p1.expandColumns()
        pm.priceRound(duals[0], result=RoundResult(model.m, model.m, model.N))
        eb = pm.expandColumns()
        for x, y in zip(ea, eb):
            np.testing.assert_array_equal(x, y)
    p1.setDualSupport(None)


@pytest.mark.parametrize("devs", _device_lists())
@pytest.mark.parametrize("profit_mode", [0, 1])
def test_batches_and_ladder_equal_single(devs, profit_mode):
    from dip_b200.pricer import GpuPricer
    model, duals = _gap(m=9, n=260, seed=8)
    U = np.stack(duals[:5])
    with GpuPricer(0) as p1, GpuPricer(devs) as pm:
        for p in (p1, pm):
            p.setModel(model, profit_mode)
            p.setDualSupport(model.dual_support)
        ra, rb = p1.priceRoundsBatch(U), pm.priceRoundsBatch(U)
        for a, b in zip(ra, rb):
            _same_round(a, b, model.m)
        # fewer vectors than devices, and a single one
        for k in (1, 2):
            for a, b in zip(p1.priceRoundsBatch(U[:k]), pm.priceRoundsBatch(U[:k])):
                _same_round(a, b, model.m)
        # the smoothing ladder: centre, steps, accept, again
        for p in (p1, pm):
            p.stabSetCenter(duals[0])
        for it in range(3):
            (la, ala), (lb, alb) = p1.priceRoundStabLadder(duals[1 + it], 0.3, 5), pm.priceRoundStabLadder(duals[1 + it], 0.3, 5)
            np.testing.assert_array_equal(ala, alb)
            for a, b in zip(la, lb):
                _same_round(a, b, model.m)
            for step in (0, 4, 2):
                np.testing.assert_array_equal(p1.smoothedDuals(step), pm.smoothedDuals(step))
            p1.stabAccept(1 + it)
            pm.stabAccept(1 + it)
        a, sa = p1.priceRoundStab(duals[5], 0.25)
        b, sb = pm.priceRoundStab(duals[5], 0.25)
        _same_round(a, b, model.m)
        np.testing.assert_array_equal(sa, sb)
        # the selected step's columns expand alike
        ea, eb = p1.expandColumns(), pm.expandColumns()
        for x, y in zip(ea, eb):
            np.testing.assert_array_equal(x, y)


@pytest.mark.parametrize("devs", _device_lists())
def test_shard_policy_switches_after_the_model_is_set(single, devs):
    """shard_rounds set AFTER the blocks: the partition follows, rounds stay equal to the one-GPU round."""
    from dip_b200.pricer import GpuPricer, RoundResult
    p1, model, duals = single
    p1.setDualSupport(model.dual_support)
    with GpuPricer(devs) as pm:
        pm.setModel(model)
        pm.setDualSupport(model.dual_support)
        want = p1.priceRound(duals[1], result=RoundResult(model.m, model.m, model.N))
        for k in (len(devs), -1, 1, len(devs)):
            pm.set_option("shard_rounds", k)
            rng = pm.shardRanges()
            assert sum(c for _, c in rng) == model.m
            assert (rng[0][1] < model.m) == (k == len(devs) and len(devs) > 1)
            _same_round(want, pm.priceRound(duals[1], result=RoundResult(model.m, model.m, model.N)), model.m)
    p1.setDualSupport(None)


@pytest.mark.parametrize("devs", _device_lists())
def test_solve_blocks_partitioned(devs):
    """Stages (b)+(c) on caller-supplied reduced costs: fused shards (few blocks) and large shards (the
    three-kernel path), ragged and empty blocks, node bounds."""
    from dip_b200.pricer import GpuPricer
    for m, nhi, caphi, seed in ((40, 90, 300, 3), (900, 40, 150, 4), (5, 1500, 900, 6)):
        b = I.random_batch(m, 0, nhi, 0, caphi, seed)
        with GpuPricer(0) as p1, GpuPricer(devs) as pm:
            pm.set_option("shard_rounds", len(devs))
            for p in (p1, pm):
                p.setObjective(b.orig_cost)
                p.setKnapsackBlocks(b.block_col_start, b.weight, b.capacity)
            _same_round(p1.solveBlocks(b.red_cost, b.alpha), pm.solveBlocks(b.red_cost, b.alpha), m)
            rng = np.random.default_rng(seed)
            lb = (rng.random(len(b.weight)) < 0.02).astype(np.float64)
            ub = 1.0 - (rng.random(len(b.weight)) < 0.05) * (1 - lb)
            for p in (p1, pm):
                p.setColBounds(lb, ub)
            _same_round(p1.solveBlocks(b.red_cost, b.alpha), pm.solveBlocks(b.red_cost, b.alpha), m)


@pytest.mark.parametrize("devs", _device_lists())
@pytest.mark.parametrize("support", [False, True])
def test_device_resident_round_over_peer_memory(single, devs, support):
    """dipgpu_price_round_dev on a multi-device context: the duals in the primary device's memory, the shards'
    packed results mirrored into the primary's gather buffer; decoded == the one-GPU round."""
    import torch
    from dip_b200.pricer import GpuPricer, RoundResult, unpack_results
    p1, model, duals = single
    p1.setDualSupport(None)
    with GpuPricer(devs) as pm:
        pm.set_option("shard_rounds", len(devs))
        pm.setModel(model)
        if support:
            pm.setDualSupport(model.dual_support)
        pm.set_option("stage_timing", 1)
        dev0 = torch.device("cuda", devs[0])
        for u in duals[:3]:
            ud = torch.from_numpy(u).to(dev0)
            torch.cuda.synchronize(dev0)
            pm.priceRoundDev(ud.data_ptr(), model.u_len)
            pm.sync()
            ms, mx = pm.lastDeviceMs()
            assert mx > 0 and len(ms) == len(devs)
            ptr, nbytes = pm.resultDev()
            off, siz = pm.resultShards()

            class _DevBytes:
                def __init__(self, p, n):
                    self.__cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (p, False), "version": 2}
            host = torch.as_tensor(_DevBytes(ptr, nbytes), device=dev0).cpu().numpy()
            got = unpack_results(host, off, siz, model.m, model.N)
            want = p1.priceRound(u, result=RoundResult(model.m, model.m, model.N))
            _same_round(want, got, model.m)


def test_errors_and_forwarded_calls():
    from dip_b200.pricer import GpuPricer, DipGpuError, RoundResult
    model, duals = _gap(m=6, n=120, seed=2)
    with GpuPricer([0, 0]) as pm, GpuPricer(0) as p1:
        with pytest.raises(DipGpuError) as e:
            pm.priceRound(duals[0])
        assert e.value.code == capi.ERR_STATE
        for p in (pm, p1):
            p.setModel(model)
        with pytest.raises(DipGpuError) as e:
            pm.setBlockRange(0, 3)
        assert e.value.code == capi.ERR_UNSUPPORTED
        with pytest.raises(DipGpuError) as e:
            pm.priceRound(duals[0][:-1])
        assert e.value.code == capi.ERR_INVALID and "uLen" in str(e.value)
        # pool re-pricing lives on the primary; the round that follows it with u == NULL too
        r = p1.priceRound(duals[0], result=RoundResult(model.m, model.m, model.N))
        cs, ri, va, hs = p1.expandColumns()
        for p in (pm, p1):
            p.poolAppend(cs, ri, va, r.colOrigCost[:r.nCols])
        a, fa = p1.poolSetReducedCosts(duals[1])
        b, fb = pm.poolSetReducedCosts(duals[1])
        np.testing.assert_array_equal(a, b)
        assert fa == fb
        _same_round(p1.priceRoundResident(result=RoundResult(model.m, model.m, model.N)),
                    pm.priceRoundResident(result=RoundResult(model.m, model.m, model.N)), model.m)
        # a subset of the blocks (not partitioned)
        ra, alla = p1.priceRoundWhich(duals[2], [1, 4])
        rb, allb = pm.priceRoundWhich(duals[2], [1, 4])
        _same_round(ra, rb, model.m)
        assert alla == allb
        k = pm.counters()
        assert k["kernel_launches"] > 0 and k["h2d_bytes"] > 0


@pytest.mark.parametrize("n_vec", [8, 13, 37])
def test_pipelined_batch_of_pinned_vectors(single, n_vec):
    """A batch of page-locked dual vectors with a dual support is priced in chunks (the gathers of u[support] on a
    second stream, chunk k's fused launch behind its own gather): equal, vector by vector, to the unpipelined batch
    and to the single rounds; on one GPU and partitioned over a multi-device context."""
    from dip_b200.pricer import GpuPricer
    p1, model, duals = single
    p1.setDualSupport(model.dual_support)
    pin = capi.PinnedArray((n_vec, model.u_len), np.float64)
    for k in range(n_vec):
        pin.array[k] = duals[k % len(duals)]
        pin.array[k][:model.meta["n"]] *= 1.0 + 0.01 * k   # distinct vectors (the assignment rows are in the support)
    ref = [_snap(p1.priceRound(pin.array[k].copy())) for k in range(n_vec)]
    for pipe in (1, 0, 1):
        p1.set_option("batch_pipeline", pipe)
        res, arr = p1._batch_results(n_vec)
        p1.priceRoundsBatchRaw(n_vec, pin.array.ctypes.data, model.u_len, capi.PHASE_PRICE2, 1e-4, arr)
        got = p1._batch_finish(res, arr)
        assert [_snap(g) for g in got] == ref, pipe
    p1.set_option("batch_pipeline", 0)
    with GpuPricer([0, 0]) as pm:
        pm.set_option("batch_pipeline", 1)
        pm.setModel(model)
        pm.setDualSupport(model.dual_support)
        res, arr = pm._batch_results(n_vec)
        pm.priceRoundsBatchRaw(n_vec, pin.array.ctypes.data, model.u_len, capi.PHASE_PRICE2, 1e-4, arr)
        assert [_snap(g) for g in pm._batch_finish(res, arr)] == ref
    p1.setDualSupport(None)
    pin.free()


@pytest.mark.parametrize("n_vec,coop", [(7, 0), (40, 0), (150, 1), (333, 1)])
def test_ticketed_batch_pulls_its_duals_itself(n_vec, coop):
    """Batches too large to be co-resident (or any batch with dp_coop = 0) draw tickets; with a dual support their
    CTAs fetch u[support] from host memory themselves (option batch_pull, from batch_pull_min vectors on) -- CTA
    (vector, block) its slice, the CTAs of a vector meet at a counter -- instead of a gather kernel in front of
    the launch: equal, vector by vector, to the gathered batch and to the single rounds; pinned and pageable
    vectors; the counters are left clean for the next launch (three launches in a row)."""
    from dip_b200.pricer import GpuPricer
    model, duals = _gap(m=40, n=260, seed=33, cls="e")
    with GpuPricer(0) as p:
        p.set_option("dp_coop", coop)
        p.set_option("batch_pull_min", 0)
        p.setModel(model)
        p.setDualSupport(model.dual_support)
        pin = capi.PinnedArray((n_vec, model.u_len), np.float64)
        for k in range(n_vec):
            pin.array[k] = duals[k % len(duals)]
            pin.array[k][:model.meta["n"]] *= 1.0 + 0.003 * k
        ref = [_snap(p.priceRound(pin.array[k].copy())) for k in range(min(n_vec, 24))]
        outs = {}
        for pull in (1, 0, 1, 1):
            p.set_option("batch_pull", pull)
            l0 = p.counters()["kernel_launches"]
            res, arr = p._batch_results(n_vec)
            p.priceRoundsBatchRaw(n_vec, pin.array.ctypes.data, model.u_len, capi.PHASE_PRICE2, 1e-4, arr)
            got = [_snap(g) for g in p._batch_finish(res, arr)]
            assert got[:len(ref)] == ref, pull
            outs.setdefault(pull, got)
            assert got == outs[pull]
            if pull:
                assert p.counters()["kernel_launches"] - l0 == 1      # no gather kernel in front
        assert outs[0] == outs[1]
        # pageable vectors: gathered on the host into mapped staging, pulled from there
        p.set_option("batch_pull", 1)
        got = [_snap(g) for g in p.priceRoundsBatch(np.array(pin.array[:min(n_vec, 30)]))]
        assert got == outs[1][:len(got)]
        pin.free()


def test_mixed_batch_sizes_over_many_launches(single):
    """a publication word that a launch does not rewrite (vector >= the launch's nVec) must never be
    taken for a current one, however many launches lie in between: batch(4), 260 x batch(1), batch(4)."""
    p1, model, duals = single
    p1.setDualSupport(None)
    U = np.stack(duals[:4])
    ref = [_snap(p1.priceRound(u)) for u in U]
    for rnd in range(2):
        got = p1.priceRoundsBatch(U)
        for g, r in zip(got, ref):
            assert _snap(g) == r
        for k in range(260):
            g = p1.priceRoundsBatch(U[(k % 4):(k % 4) + 1])
            if k % 37 == 0:
                assert _snap(g[0]) == ref[k % 4]


def _snap(r):
    n = r.nCols
    return (n, r.blockObj.tobytes(), r.blockRedCost.tobytes(), r.blockNnz.tobytes(), r.colBlock[:n].tobytes(),
            r.colStart[:n + 1].tobytes(), r.colInd[:r.c.nInd].tobytes(), r.mostNegReducedCost, r.cells)


@pytest.mark.parametrize("support", [False, True])
def test_one_kernel_round_equals_three_kernel_round(support):
    """The fused shapes form the reduced costs inside the DP kernel (option fuse_redcost, default on): same bits as
    the stream SpMV kernel in front of it -- single rounds, phase 1, resident duals, batches, the ladder, node bounds,
    cut rows appended behind the convexity rows."""
    from dip_b200.pricer import GpuPricer, RoundResult
    model, duals = _gap(m=24, n=1100, seed=21, cls="e")   # blocks wide enough for the reduction pre-pass
    with GpuPricer(0) as pa, GpuPricer(0) as pb:
        pb.set_option("fuse_redcost", 0)
        for p in (pa, pb):
            p.setModel(model)
            if support:
                p.setDualSupport(model.dual_support)
        la0 = pa.counters()["kernel_launches"]
        for u in duals[:3]:
            for phase in (capi.PHASE_PRICE2, capi.PHASE_PRICE1):
                _same_round(pa.priceRound(u, phase=phase, result=RoundResult(model.m, model.m, model.N)),
                            pb.priceRound(u, phase=phase, result=RoundResult(model.m, model.m, model.N)), model.m)
        if support:   # (without a declared support a block's 3 x 1100 stored entries exceed the kernel's staging areas:
            #  the stream SpMV stays in front)
            assert pa.counters()["kernel_launches"] - la0 == 6          # ONE kernel per round
        _same_round(pa.priceRoundResident(result=RoundResult(model.m, model.m, model.N)),
                    pb.priceRoundResident(result=RoundResult(model.m, model.m, model.N)), model.m)
        ra, rca = pa.priceRound(duals[3], want_redcost=True, result=RoundResult(model.m, model.m, model.N))
        rb, rcb = pb.priceRound(duals[3], want_redcost=True, result=RoundResult(model.m, model.m, model.N))
        _same_round(ra, rb, model.m)
        np.testing.assert_array_equal(rca, rcb)
        for a, b in zip(pa.priceRoundsBatch(np.stack(duals[:5])), pb.priceRoundsBatch(np.stack(duals[:5]))):
            _same_round(a, b, model.m)
        for p in (pa, pb):
            p.stabSetCenter(duals[0])
        (la, _), (lb, _) = pa.priceRoundStabLadder(duals[1], 0.3, 4), pb.priceRoundStabLadder(duals[1], 0.3, 4)
        for a, b in zip(la, lb):
            _same_round(a, b, model.m)
        # the pool re-pricing after a one-kernel round reads the dense vector (formed on the device from the compact one)
        r = pa.priceRound(duals[2], result=RoundResult(model.m, model.m, model.N))
        cs, ri, va, hs = pa.expandColumns()
        for p in (pa, pb):
            p.priceRound(duals[2], result=RoundResult(model.m, model.m, model.N))
            p.poolClear()
            p.poolAppend(cs, ri, va, r.colOrigCost[:r.nCols])
        np.testing.assert_array_equal(pa.poolSetReducedCosts(None)[0], pb.poolSetReducedCosts(None)[0])
        # node bounds
        rng = np.random.default_rng(3)
        lb_ = (rng.random(model.N) < 0.01).astype(np.float64)
        ub_ = 1.0 - (rng.random(model.N) < 0.03) * (1 - lb_)
        for p in (pa, pb):
            p.setColBounds(lb_, ub_)
        _same_round(pa.priceRound(duals[4], result=RoundResult(model.m, model.m, model.N)),
                    pb.priceRound(duals[4], result=RoundResult(model.m, model.m, model.N)), model.m)
        for p in (pa, pb):
            p.setColBounds(None, None)
        # cut rows behind the convexity rows
        n_cut = 7
        rs = np.zeros(n_cut + 1, np.int64)
        ci, va2 = [], []
        for k in range(n_cut):
            cols = np.sort(rng.choice(model.N, size=25, replace=False))
            ci.append(cols)
            va2.append(rng.integers(1, 5, size=25).astype(np.float64))
            rs[k + 1] = rs[k] + 25
        for p in (pa, pb):
            p.appendCoreRows(rs, np.concatenate(ci).astype(np.int32), np.concatenate(va2))
        u2 = np.concatenate([duals[0], rng.uniform(-2, 2, size=n_cut)])
        if support:
            sup2 = np.concatenate([model.dual_support, np.arange(model.u_len, model.u_len + n_cut)]).astype(np.int32)
            for p in (pa, pb):
                p.setDualSupport(sup2)
        _same_round(pa.priceRound(u2, result=RoundResult(model.m, model.m, model.N)),
                    pb.priceRound(u2, result=RoundResult(model.m, model.m, model.N)), model.m)


def test_cut_rows_appended_as_a_delta_on_a_multi_device_context():
    """dipgpu_append_core_rows reaches every device; each keeps the new rows as a delta (append_delta_min_nnz = 0
    forces the path on this small model) and sweeps its own column range: forced block partition, whole rounds and a
    batch equal to the one-GPU context that rebuilds everything."""
    from dip_b200.pricer import GpuPricer
    model, duals = _gap(m=16, n=280, seed=41, cls="e")
    rng = np.random.default_rng(4)
    with GpuPricer(0) as p1, GpuPricer([0, 0, 0]) as pm:
        pm.set_option("append_delta_min_nnz", 0)
        pm.set_option("shard_rounds", 3)
        for p in (p1, pm):
            p.setModel(model)
            p.calcRedCost(duals[0])          # tiles built on every device
        R = model.R
        for k, nn in ((6, 300), (3, 40)):
            r2 = np.sort(rng.integers(0, k, size=nn))
            rs2 = np.zeros(k + 1, np.int64)
            np.cumsum(np.bincount(r2, minlength=k), out=rs2[1:])
            ci2 = rng.integers(0, model.N, size=nn).astype(np.int32)
            v2 = np.round(rng.uniform(-3, 3, size=nn), 1)
            for p in (p1, pm):
                p.appendCoreRows(rs2, ci2, v2)
            R += k
            U = [np.concatenate([u, rng.standard_normal(R - model.R)]) for u in duals[:4]]
            np.testing.assert_array_equal(pm.calcRedCost(U[0]), p1.calcRedCost(U[0]))
            np.testing.assert_array_equal(pm.calcRedCost(U[1], capi.PHASE_PRICE1), p1.calcRedCost(U[1], capi.PHASE_PRICE1))
            for u in U[:2]:
                assert _snap(pm.priceRound(u)) == _snap(p1.priceRound(u))
            assert [_snap(g) for g in pm.priceRoundsBatch(np.stack(U))] == [_snap(g) for g in p1.priceRoundsBatch(np.stack(U))]
