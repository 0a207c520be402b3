import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import bench
from dip_b200.pricer import GpuPricer
torch.cuda.set_device(0)
fb = torch.empty(256 << 20, dtype=torch.uint8, device='cuda'); fr = torch.ones(64 << 20, dtype=torch.float32, device='cuda')
def flush():
    fb.fill_(1); fr.sum()
import sys
for k in (0, 2, 1):
    r = bench.bench_pool(torch, 0, 6542.4, 'm', flush, 0, pool_kernel=k)
    print(k, r['avg_launch_ms'], r['frac'], r.get('without_dual_support'))
for opts in ({'pool_compact': 0}, {'pool_chunk': 512}, {'pool_warps': 16}, {'pool_warps': 20}, {'pool_warps': 24}):
    r = bench.bench_pool(torch, 0, 6542.4, 'm', flush, 0, pool_kernel=1, options=opts)
    print(opts, r['avg_launch_ms'], r['frac'])
