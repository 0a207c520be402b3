"""pytest configuration: the `gpu` marker and shared fixtures.

`-m "not gpu"` runs here (no GPU): oracle vs golden vectors, host logic, ABI symbol checks.
`-m gpu` runs on a B200: parity of the CUDA path (through the C ABI) against the oracle.
"""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _unhex(a):
    return np.array([float.fromhex(v) for v in a], np.float64)


@pytest.fixture(scope="session")
def golden_knapsack():
    d = json.load(open(os.path.join(GOLDEN, "knapsack01_cases.json")))
    cases = []
    for c in d["cases"]:
        cases.append(dict(tag=c["tag"], obj=_unhex(c["obj"]), w=np.array(c["w"], np.int32), cap=int(c["cap"]),
                          z=float.fromhex(c["z"]), solution=list(c["solution"]),
                          hs_z=float.fromhex(c["hs_z"]) if "hs_z" in c else None,
                          hs_status=c.get("hs_status")))
    return cases


@pytest.fixture(scope="session")
def golden_intknap():
    d = json.load(open(os.path.join(GOLDEN, "intknap_kp_cases.json")))
    return [dict(tag=c["tag"], obj=_unhex(c["obj"]), w=np.array(c["w"], np.int32), cap=int(c["cap"]),
                 z=float.fromhex(c["z"]), counts=list(c["counts"]), plain=bool(c["plain"])) for c in d["cases"]]


@pytest.fixture(scope="session")
def golden_gap0515():
    d = json.load(open(os.path.join(GOLDEN, "gap0515_2_rounds.json")))
    for r in d["rounds"]:
        r["u"] = _unhex(r["u"])
        r["red_cost"] = _unhex(r["red_cost"])
        for b in r["blocks"]:
            b["z"] = float.fromhex(b["z"])
    return d


@pytest.fixture(scope="session")
def oracle():
    from oracle import dip_oracle
    dip_oracle.build()
    return dip_oracle


@pytest.fixture(scope="session")
def gpu_lib():
    """The built CUDA library; a GPU test never falls back to anything else."""
    from dip_b200 import capi
    return capi.lib()


@pytest.fixture()
def pricer(gpu_lib):
    from dip_b200.pricer import GpuPricer
    p = GpuPricer(0)
    yield p
    p.close()
