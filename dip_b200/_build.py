"""In-tree build of libdipgpu.so for sm_100a (nvcc cross-compiles without a GPU).

    python -m dip_b200._build [--force] [--verbose]

The built library lives at dip_b200/libdipgpu.so (git-ignored, shipped to the
GPU box by gpurun).  cudart is linked statically so the .so has no CUDA toolkit
run-time dependency beyond the driver.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libdipgpu.so")
SOURCES = ["dipgpu_api.cu", "redcost.cu", "knapsack.cu", "filter.cu", "microbench.cu", "pisinger.cu", "expand.cu", "pool.cu", "mckp.cu", "multi.cu", "bigdp.cu", "intknap.cu"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "filter.cuh"), os.path.join(os.path.dirname(HERE), "include", "dipgpu.h")]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# -fmad=false: no fp contraction anywhere -- every product and sum is rounded on its own, as the reference's
# C++ / Python arithmetic does it (e.g. floor(-rc*scale + 0.5) rounds the product first, GAP_KnapPisinger.h:174-175)
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-fmad=false", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
         "--expt-relaxed-constexpr"]


def nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libdipgpu cannot be built (there is no CPU fallback)")


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src: str, verbose: bool) -> str:
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    path = os.path.join(CSRC, src)
    if _stale(obj, [path] + HEADERS):
        cmd = [nvcc(), *ARCH, *FLAGS, "-c", path, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
        if os.path.exists(LIB):
            os.remove(LIB)
    with cf.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), SOURCES))
    if _stale(LIB, objs):
        cmd = [nvcc(), *ARCH, "-shared", "-o", LIB, *objs, "-cudart", "static"]
        env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
        subprocess.run(cmd, check=True, env=env)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
