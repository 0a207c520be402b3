"""The N > 1 path on CPU: world_size-2 gloo.  Each rank prices its contiguous block range of one
GAP round with the CPU oracle (standing in for the CUDA kernels, which need a GPU), packs the shard
in the library's wire format, the shards are all-gathered like bench.py does over NCCL, and rank 0
decodes them through the C ABI (dipgpu_unpack_result, host-only) and merges.  The merged round must
equal the unsharded one."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dip_b200 import instances as I
        from dip_b200.sharding import block_range, merge_shards
        from oracle import dip_oracle as orc
        from packing import pack_shard

        model = I.gap_pricing_model(I.gap_class_d(9, 60, seed=4))
        # rank 0 owns the master duals and broadcasts them (ncclBroadcast(u) on the GPUs)
        u = torch.zeros(model.u_len, dtype=torch.float64)
        if rank == 0:
            u.copy_(torch.from_numpy(I.gap_duals(model, np.random.default_rng(3))))
        dist.broadcast(u, src=0)
        un = u.numpy()
        work = np.diff(model.block_col_start) * (model.capacity.astype(np.int64) + 1)
        first, count = block_range(model.m, world, rank, work)
        # shard-local pricing
        rcx = orc.calc_redcost(model.R, model.N, model.row_start, model.col_ind, model.val,
                               orc.adjust_duals(un, model.n_base_rows, model.m), model.objective)
        bcs = model.block_col_start[first:first + count + 1]
        alpha = un[model.n_base_rows + first:model.n_base_rows + first + count]
        sh = orc.solve_blocks(bcs, model.weight, model.capacity[first:first + count], rcx, model.objective, alpha)
        ind_cap = int(bcs[-1] - bcs[0]) if count else 0
        buf = pack_shard(first, sh.col_red_cost, sh.col_orig_cost, sh.z, sh.col_nnz, sh.col_keep, sh.col_ind,
                         sh.cells, ind_cap)
        # fixed-size all-gather of the packed shards
        size = torch.tensor([len(buf)], dtype=torch.int64)
        dist.all_reduce(size, op=dist.ReduceOp.MAX)
        send = torch.zeros(int(size.item()), dtype=torch.uint8)
        send[:len(buf)] = torch.from_numpy(buf)
        recv = [torch.zeros_like(send) for _ in range(world)]
        dist.all_gather(recv, send)
        if rank == 0:
            merged = merge_shards([r.numpy() for r in recv], model.m, model.N)
            full = orc.price_round(model.R, model.N, model.row_start, model.col_ind, model.val, model.objective, un,
                                   model.n_base_rows, model.m, model.block_col_start, model.weight, model.capacity)
            ok = (merged.nCols == full.n_kept
                  and np.array_equal(merged.blockRedCost, full.col_red_cost)
                  and np.array_equal(merged.blockObj, full.z)
                  and np.array_equal(merged.blockMostNegRC, full.most_neg_rc)
                  and merged.mostNegReducedCost == full.most_neg_reduced_cost
                  and merged.cells == full.cells)
            kept = [b for b in range(model.m) if full.col_keep[b]]
            ok = ok and list(merged.colBlock[:merged.nCols]) == kept
            for k, b in enumerate(kept):
                ok = ok and list(merged.column(k)) == list(full.col_ind[b])
            q.put(("ok" if ok else "mismatch", first, count))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_sharded_round_world2_gloo(oracle, gpu_lib):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(150)
        assert p.exitcode == 0
    assert q.get(timeout=5)[0] == "ok"


def test_block_range_partitions():
    from dip_b200.sharding import block_range
    for m in (0, 1, 7, 80, 81):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                f, c = block_range(m, world, r)
                seen.extend(range(f, f + c))
            assert seen == list(range(m))
    work = np.array([1, 1, 1, 1, 100, 1, 1, 1], np.float64)
    parts = [block_range(8, 2, r, work) for r in range(2)]
    assert parts[0][0] == 0 and parts[0][1] + parts[1][1] == 8 and parts[1][0] == parts[0][1]
