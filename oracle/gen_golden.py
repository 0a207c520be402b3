#!/usr/bin/env python3
"""Generate tests/golden/*.json from the REFERENCE ITSELF (run in the build container).

TEST INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box, so the
vectors made here are committed; this script is the recipe that made them.

What is executed unchanged from the reference:
  * ``knapsack01`` -- AST-extracted from
    /root/reference/Dip/src/dippy/examples/gap/gap_func.py:132-183 (the module
    itself cannot be imported: it needs PuLP at gap_func.py:12) and exec'd as is.
  * ``KnapsackOptimizeHS`` -- /root/reference/Dip/src/UtilKnapsack.cpp:246-502
    compiled into oracle/_ref/libdip_ref_knap.so by oracle/Makefile.
  * ``kp`` -- the integer knapsack of the cutting-stock pricing problem, AST-extracted from
    /root/reference/Dip/src/dippy/tests/test_cutting_stock.py:154-179 (nested in
    create_cutting_stock_problem; the module imports PuLP and coinor.dippy) and exec'd as is.  The
    function is an exponential recursion: small cases run it as it stands; larger ones run the same
    function body with its recursive calls memoised on the capacity (the global name ``kp`` is bound
    to a wrapper that returns a COPY of the cached (z, counts), because the body mutates the list it
    gets back); both ways are checked equal on the small cases.
  * the instance /root/reference/Dip/src/dippy/examples/gap/gap0515-2.dat
    (format: Dip/examples/GAP/GAP_Instance.cpp:22-74), known optimum 269
    (Dip/data/GAP/gap.opt:9).

Usage:  python oracle/gen_golden.py   (writes tests/golden/)
"""
from __future__ import annotations

import ast
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF = "/root/reference"
GAP_FUNC = f"{REF}/Dip/src/dippy/examples/gap/gap_func.py"
GAP_DAT = f"{REF}/Dip/src/dippy/examples/gap/gap0515-2.dat"


def extract_reference_knapsack01():
    src = open(GAP_FUNC).read()
    tree = ast.parse(src)
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "knapsack01"]
    assert len(fn) == 1
    mod = ast.Module(body=fn, type_ignores=[])
    ns: dict = {}
    exec(compile(mod, GAP_FUNC, "exec"), ns)  # noqa: S102 - the reference's own function
    return ns["knapsack01"]


CSP_TEST = f"{REF}/Dip/src/dippy/tests/test_cutting_stock.py"


def extract_reference_kp():
    """Returns (kp_plain, kp_memo): the reference's kp unchanged, and the same body with memoised recursion."""
    tree = ast.parse(open(CSP_TEST).read())
    outer = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "create_cutting_stock_problem"][0]
    fn = [n for n in outer.body if isinstance(n, ast.FunctionDef) and n.name == "kp"]
    assert len(fn) == 1
    code = compile(ast.Module(body=fn, type_ignores=[]), CSP_TEST, "exec")
    ns_plain: dict = {"range": range}
    exec(code, ns_plain)  # noqa: S102 - the reference's own function
    ns_memo: dict = {"range": range}
    exec(code, ns_memo)  # noqa: S102
    body = ns_memo["kp"]
    cache: dict = {}

    def wrapper(obj, weights, capacity):
        if capacity not in cache:
            z, sol = body(obj, weights, capacity)
            cache[capacity] = (z, list(sol))
        z, sol = cache[capacity]
        return z, list(sol)

    ns_memo["kp"] = wrapper

    def kp_memo(obj, weights, capacity):
        cache.clear()
        return wrapper(list(obj), list(weights), capacity)

    return ns_plain["kp"], kp_memo


def intknap_cases(rng):
    cases = []

    def add(tag, obj, w, cap, plain):
        cases.append(dict(tag=tag, obj=[float(v) for v in obj], w=[int(v) for v in w], cap=int(cap), plain=plain))

    # the pricing problems of the reference's own test instance (lengths 9, 7, 5; roll length 20)
    add("csp_test_unit", [1.0, 1.0, 1.0], [9, 7, 5], 20, True)
    add("csp_test_duals", [0.5, 0.375, 0.25], [9, 7, 5], 20, True)
    add("csp_test_two_items", [0.45, 0.25], [9, 5], 20, True)
    add("empty", [], [], 12, True)
    add("cap0", [1.0, 2.0], [3, 4], 0, True)
    add("none_fits", [1.0, 2.0], [30, 40], 20, True)
    # small enough for the plain recursion
    for k in range(30):
        n = int(rng.integers(1, 5))
        cap = int(rng.integers(1, 22))
        w = rng.integers(3, 12, size=n)
        obj = rng.uniform(0.05, 3.0, size=n)
        add(f"small_fp64_{k}", obj, w, cap, True)
    for k in range(20):
        n = int(rng.integers(2, 5))
        cap = int(rng.integers(4, 20))
        w = rng.integers(2, 8, size=n)
        obj = rng.integers(1, 5, size=n).astype(np.float64)
        add(f"small_ties_{k}", obj, w, cap, True)
    # memoised recursion: realistic sizes
    for k in range(40):
        n = int(rng.integers(1, 41))
        cap = int(rng.integers(1, 400))
        w = rng.integers(1, max(2, cap // 2 + 2), size=n)
        obj = rng.uniform(0.001, 2.0, size=n)
        add(f"rand_fp64_{k}", obj, w, cap, False)
    for k in range(30):
        n = int(rng.integers(2, 30))
        cap = int(rng.integers(5, 200))
        w = rng.integers(1, 9, size=n)
        obj = rng.integers(1, 6, size=n).astype(np.float64)
        add(f"ties_int_{k}", obj, w, cap, False)
    for k in range(20):
        n = int(rng.integers(2, 30))
        cap = int(rng.integers(20, 300))
        w = rng.integers(2, 40, size=n)
        add(f"proportional_{k}", w * 0.25, w, cap, False)          # profit per unit of weight equal: every fill ties
    for k in range(20):
        n = int(rng.integers(3, 30))
        cap = int(rng.integers(30, 300))
        w = rng.integers(1, 25, size=n)
        base = rng.uniform(0.1, 0.3)
        obj = base * w * (1.0 + rng.integers(-3, 4, size=n) * 2.0 ** -50)  # sums differ in the last bits
        add(f"ulp_{k}", obj, w, cap, False)
    for k, cap in enumerate([31, 32, 33, 63, 64, 65, 255, 256, 257, 1000, 1023, 1024, 1025]):
        n = int(rng.integers(5, 60))
        w = rng.integers(1 + k % 3, max(3, cap // 4), size=n)
        obj = rng.uniform(0.01, 1.0, size=n) * w ** 0.9
        add(f"wide_{k}_cap{cap}", obj, w, cap, False)
    return cases


def gen_intknap():
    """tests/golden/intknap_kp_cases.json: (z, counts) of the reference's kp."""
    from oracle import dip_oracle as orc

    sys.setrecursionlimit(20000)
    kp_plain, kp_memo = extract_reference_kp()
    rng = np.random.default_rng(20261020)
    out = []
    n_plain = 0
    for cs in intknap_cases(rng):
        z, sol = kp_memo(cs["obj"], cs["w"], cs["cap"])
        if cs["plain"]:
            z0, sol0 = kp_plain(list(cs["obj"]), list(cs["w"]), cs["cap"])
            assert float(z0) == float(z) and list(sol0) == list(sol), cs["tag"]
            n_plain += 1
        zo, so = orc.intknap_kp(cs["obj"], cs["w"], cs["cap"])
        assert float(z) == zo and list(sol) == list(so), (cs["tag"], z, zo, sol, so)
        out.append(dict(tag=cs["tag"], obj=hx(cs["obj"]), w=cs["w"], cap=cs["cap"], plain=cs["plain"],
                        z=float(z).hex(), counts=[int(v) for v in sol]))
    json.dump(dict(source="test_cutting_stock.py:154-179 kp, AST-extracted and exec'd (plain: unchanged recursion; "
                          "others: the same body, recursive calls memoised on the capacity)", cases=out),
              open(f"{ROOT}/tests/golden/intknap_kp_cases.json", "w"), indent=0)
    print(f"intknap kp: {len(out)} cases ({n_plain} through the plain recursion)")


def hx(a):
    return [float(v).hex() for v in a]


def knap_cases(rng):
    cases = []

    def add(tag, obj, w, cap):
        cases.append(dict(tag=tag, obj=[float(v) for v in obj], w=[int(v) for v in w], cap=int(cap)))

    # random fp64 profits
    for k in range(60):
        n = int(rng.integers(2, 41))
        cap = int(rng.integers(0, 121))
        w = rng.integers(1, max(2, cap // 2 + 3), size=n)
        obj = rng.uniform(0.01, 50.0, size=n)
        add(f"rand_fp64_{k}", obj, w, cap)
    # tie-heavy: small integer profits, small weights -> many equal-value sets
    for k in range(40):
        n = int(rng.integers(2, 25))
        cap = int(rng.integers(1, 40))
        w = rng.integers(1, 6, size=n)
        obj = rng.integers(1, 5, size=n).astype(np.float64)
        add(f"ties_int_{k}", obj, w, cap)
    # profit proportional to weight (subset-sum like; every fill of equal weight ties)
    for k in range(20):
        n = int(rng.integers(2, 30))
        cap = int(rng.integers(5, 80))
        w = rng.integers(1, 20, size=n)
        add(f"subset_sum_{k}", w.astype(np.float64) * 0.5, w, cap)
    # items heavier than the capacity, zero-weight items, capacity 0
    for k in range(20):
        n = int(rng.integers(2, 20))
        cap = int(rng.integers(0, 30))
        w = rng.integers(0, 2 * cap + 3, size=n)
        obj = rng.uniform(0.5, 9.5, size=n)
        add(f"edge_w_{k}", obj, w, cap)
    # fp64 rounding sensitivity: profits that differ in the last bits
    for k in range(20):
        n = int(rng.integers(3, 30))
        cap = int(rng.integers(10, 60))
        w = rng.integers(1, 12, size=n)
        base = rng.uniform(1, 3)
        obj = base * w * (1.0 + rng.integers(-3, 4, size=n) * 2.0 ** -50)
        add(f"ulp_{k}", obj, w, cap)
    # wider rows (several 32-cell groups; crosses the warp-width boundaries)
    for k, cap in enumerate([31, 32, 33, 63, 64, 65, 127, 128, 129, 255, 256, 257, 500, 1023, 1024, 1025]):
        n = int(rng.integers(20, 60))
        w = rng.integers(1, max(2, cap // 3), size=n)
        obj = rng.uniform(0.1, 100.0, size=n)
        add(f"wide_{k}_cap{cap}", obj, w, cap)
    return cases


def main():
    from oracle import dip_oracle as orc

    if "--intknap" in sys.argv:
        return gen_intknap()

    ref_knap = extract_reference_knapsack01()
    rng = np.random.default_rng(20261019)
    os.makedirs(f"{ROOT}/tests/golden", exist_ok=True)

    # ---------------------------------------------------------- knapsack01 golden
    out = []
    hs_checked = 0
    for cs in knap_cases(rng):
        z, sol = ref_knap(list(cs["obj"]), list(cs["w"]), cs["cap"])
        rec = dict(tag=cs["tag"], obj=hx(cs["obj"]), w=cs["w"], cap=cs["cap"],
                   z=float(z).hex(), solution=[int(i) for i in sol])
        # reference HS B&B on integer-profit cases with n >= 3 and positive weights
        ints = all(float(p).is_integer() for p in cs["obj"]) and min(cs["w"]) >= 1
        if ints and len(cs["obj"]) >= 3 and orc.have_ref():
            st, zhs, _ = orc.ref_knapsack_hs(cs["obj"], cs["w"], cs["cap"])
            rec["hs_status"] = int(st)
            rec["hs_z"] = float(zhs).hex()
            hs_checked += 1
        out.append(rec)
        # the oracle must agree with the reference function right here
        zo, so = orc.knapsack01(cs["obj"], cs["w"], cs["cap"])
        assert float(z) == zo and list(sol) == list(so), cs["tag"]
    json.dump(dict(source="gap_func.py:132-183 knapsack01, AST-extracted and exec'd unchanged",
                   cases=out), open(f"{ROOT}/tests/golden/knapsack01_cases.json", "w"), indent=0)
    print(f"knapsack01: {len(out)} cases ({hs_checked} with reference HS values)")

    # ------------------------------------------------------------- gap0515-2 round
    tok = open(GAP_DAT).read().split()
    m, n = int(tok[0]), int(tok[1])
    vals = [int(t) for t in tok[2:]]
    cost = np.array(vals[: m * n], np.int64).reshape(m, n)
    wgt = np.array(vals[m * n: 2 * m * n], np.int64).reshape(m, n)
    cap = np.array(vals[2 * m * n: 2 * m * n + m], np.int64)
    N = m * n
    rounds = []
    for r in range(6):
        # master dual layout (DecompAlgo.cpp:826-828): n AP rows | 2N branch rows | m convex
        u_ap = rng.uniform(0.0, 2.0 * cost.mean(), size=n)
        u_br = np.where(rng.uniform(size=2 * N) < 0.02, rng.uniform(-3, 3, size=2 * N), 0.0)
        alpha = rng.uniform(-10.0, 0.0, size=m)
        if r == 0:
            u_br[:] = 0.0
        u = np.concatenate([u_ap, u_br, alpha])
        # c - A''^T u, straight from the model definition (GAP_DecompApp.cpp:69-135 + the
        # identity rows of DecompAlgo.cpp:1431-1455)
        rc = cost.astype(np.float64).reshape(-1).copy()
        for i in range(m):
            for j in range(n):
                k = i * n + j
                acc = 0.0
                # descending row order, as the oracle's scatter
                if u_br[N + k] != 0.0:
                    acc += u_br[N + k] * 1.0
                if u_br[k] != 0.0:
                    acc += u_br[k] * 1.0
                if u_ap[j] != 0.0:
                    acc += u_ap[j] * 1.0
                rc[k] = cost[i, j] - acc
        blocks = []
        for i in range(m):
            rcb = rc[i * n:(i + 1) * n]
            idx = [t for t in range(n) if rcb[t] < 0]      # gap_func.py:102
            obj = [-rcb[t] for t in idx]                   # gap_func.py:104
            ws = [int(wgt[i, t]) for t in idx]             # gap_func.py:105
            z, sol = ref_knap(obj, ws, int(cap[i]))        # gap_func.py:107
            if len(idx) == 1:
                # documented divergence: the Python aliases row -1 onto row 0 when n == 1
                z = obj[0] if ws[0] <= cap[i] else 0.0
                sol = [0] if ws[0] <= cap[i] else []
            chosen = sorted(i * n + idx[s] for s in sol)
            blocks.append(dict(z=float(z).hex(), ind=chosen))
        rounds.append(dict(u=hx(u), red_cost=hx(rc), blocks=blocks))
    json.dump(dict(source="gap0515-2.dat + knapsack01 (reference), duals seeded 20261019",
                   m=m, n=n, cost=cost.reshape(-1).tolist(), weight=wgt.reshape(-1).tolist(),
                   capacity=cap.tolist(), known_optimum_min=269, rounds=rounds),
              open(f"{ROOT}/tests/golden/gap0515_2_rounds.json", "w"), indent=0)
    print(f"gap0515-2: {len(rounds)} rounds")
    gen_intknap()


if __name__ == "__main__":
    main()
