"""The header-only C++ host mirror (include/dipgpu_decomp.hpp): compiles with g++ against the C ABI,
refuses to run without a GPU (no CPU fallback), and on a B200 reproduces the oracle exactly."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_adapter")


def _build(oracle):
    src = os.path.join(ROOT, "tests", "cpp", "test_adapter.cpp")
    deps = [src, os.path.join(ROOT, "include", "dipgpu_decomp.hpp"), os.path.join(ROOT, "include", "dipgpu.h")]
    if os.path.exists(EXE) and all(os.path.getmtime(EXE) > os.path.getmtime(d) for d in deps):
        return
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-o", EXE, src,
                    "-L" + os.path.join(ROOT, "dip_b200"), "-ldipgpu",
                    "-L" + os.path.join(ROOT, "oracle"), "-ldip_oracle",
                    "-Wl,-rpath," + os.path.join(ROOT, "dip_b200"), "-Wl,-rpath," + os.path.join(ROOT, "oracle")],
                   check=True, env=env)


def _run():
    return subprocess.run([EXE], capture_output=True, text=True, timeout=300)


def test_cpp_adapter_compiles_and_refuses_cpu(oracle, gpu_lib):
    import torch
    _build(oracle)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = _run()
    assert r.returncode == 0 and r.stdout.strip() == "NO_GPU", r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_adapter_parity(oracle, gpu_lib):
    _build(oracle)
    r = _run()
    assert r.returncode == 0 and r.stdout.startswith("PARITY_OK"), r.stdout + r.stderr
