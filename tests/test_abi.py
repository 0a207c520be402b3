"""The C-ABI library loads and exports every symbol include/dipgpu.h declares.  No compute calls
without a GPU: on a CPU-only box dipgpu_create must FAIL (there is no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "dipgpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\bint\s+(dipgpu_\w+)\s*\(", txt)))


def test_header_symbols_are_exported(gpu_lib):
    from dip_b200 import capi
    names = _declared()
    assert len(names) >= 20
    assert set(names) == set(capi.SYMBOLS)
    for n in names:
        assert hasattr(gpu_lib, n), n
    assert gpu_lib.dipgpu_abi_version() == 4


def test_only_the_abi_is_exported():
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", os.path.join(ROOT, "dip_b200", "libdipgpu.so")],
                         capture_output=True, text=True, check=True).stdout
    syms = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert syms and all(s.startswith("dipgpu_") for s in syms), syms


def test_no_cpu_fallback(gpu_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ctx = C.c_void_p()
    assert gpu_lib.dipgpu_create(C.byref(ctx), 0) == 2  # DIPGPU_ERR_CUDA
    assert not ctx.value
    from dip_b200.pricer import GpuPricer
    from dip_b200.capi import DipGpuError
    with pytest.raises(DipGpuError):
        GpuPricer(0)


def test_product_does_not_touch_the_oracle():
    """Nothing under dip_b200/ or include/ may import, link or call oracle/."""
    bad = []
    for base in ("dip_b200", "include"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            for f in fs:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    s = open(os.path.join(dp, f), errors="replace").read()
                    if re.search(r"dip_oracle|from oracle|import oracle|oracle/", s):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_unpack_rejects_garbage(gpu_lib):
    from dip_b200 import capi
    cols = capi.Cols()
    buf = (C.c_char * 256)()
    assert gpu_lib.dipgpu_unpack_result(buf, 256, C.byref(cols)) == capi.ERR_INVALID
    assert gpu_lib.dipgpu_unpack_result(None, 0, C.byref(cols)) == capi.ERR_INVALID


def test_dp_geometry_planner_host_only():
    """Host logic of the DP shape choice (no device): covers the row, odd cells/thread, smem limit."""
    from dip_b200 import capi
    for nb, cap, cols, mw in [(80, 176, 1600, 60), (10_000, 10_000, 500, 1000), (20, 404, 200, 100),
                              (1, 0, 1, 0), (5000, 12_000, 300, 12_000), (100_000, 30, 20, 30)]:
        g = capi.plan_dp_geometry(nb, cap, cols, mw)
        assert g["threads"] % 32 == 0 and 32 <= g["threads"] <= 1024
        assert g["cells_per_thread"] % 2 == 1
        assert g["threads"] * g["cells_per_thread"] >= cap + 1
        assert g["smem_bytes"] <= 227 * 1024
        assert g["guard"] == 0 or g["guard"] >= mw * g["items_per_step"]
    # latency regime (few small knapsacks): one cell per thread, fused backtrack + filter
    g = capi.plan_dp_geometry(80, 176, 1600, 60)
    assert g["cells_per_thread"] == 1 and g["fused_tail"] == 1
    # throughput regime (many large knapsacks): many cells per thread
    g = capi.plan_dp_geometry(10_000, 10_000, 500, 1000)
    assert g["cells_per_thread"] >= 11 and g["fused_tail"] == 0
    # a row beyond one CTA's registers / shared memory (> 12 800 cells): rows in global memory, any capacity
    g = capi.plan_dp_geometry(10, 1_000_000, 10, 10)
    assert g["cells_per_thread"] == 0 and g["fused_tail"] == 0
