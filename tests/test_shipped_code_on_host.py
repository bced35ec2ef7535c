"""Shipped CUDA source, compiled for the HOST with g++ and checked on the CPU.

The library's float64 leaf evaluation (csrc/mpcb_exact.cuh: the numbers `best_cost`, `best_traj` and `first_control`
are made of) and its invariant-divisor index decoding (csrc/mpcb_types.cuh) are plain host/device functions.  These
tests build them into a small shared object (tests/native/exact_host.cpp) and compare with the reference's golden
outputs (tests/golden/*.json, produced by the reference's own unmodified functions) and with the float64 oracle --
so the CPU suite, too, exercises code that ships, not only its restatement.  Nothing here is a product path: the
library itself still needs a GPU (tests/test_abi.py::test_no_cpu_fallback_without_a_device)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import closed_form as C

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
L, DT = C.CONFIG["L"], C.CONFIG["delta_t"]


@pytest.fixture(scope="module")
def host():
    out_dir = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "exact_host.so")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-std=c++17",
                           "-I", os.path.join(ROOT, "diplomjourney_b200", "csrc"), "-I", cuda_inc,
                           os.path.join(ROOT, "tests", "native", "exact_host.cpp"), "-o", so])
    lib = ctypes.CDLL(so)
    dp, u64p, u32p = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint32)
    lib.mpcb_test_fastdiv64.argtypes = [ctypes.c_uint64, ctypes.c_longlong, u64p, u64p]
    lib.mpcb_test_fastdiv32.argtypes = [ctypes.c_uint32, ctypes.c_longlong, u32p, u32p]
    lib.mpcb_test_exact_cost.argtypes = [dp, ctypes.c_int, dp, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_int,
                                         ctypes.c_int, ctypes.c_int, dp, dp, dp, ctypes.c_longlong, dp,
                                         ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.c_double]
    lib.mpcb_test_exact_cost.restype = ctypes.c_double
    lib.mpcb_test_apply_event.argtypes = [ctypes.c_int] + [ctypes.c_double] * 6 + [ctypes.c_int, dp]
    lib.mpcb_test_exact_from_node.argtypes = [dp, ctypes.c_int, dp, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_int,
                                              ctypes.c_int, dp, dp, dp, ctypes.c_int, ctypes.c_double, ctypes.c_int,
                                              ctypes.c_longlong, ctypes.POINTER(ctypes.c_longlong), dp, dp, u32p, u32p]
    return lib


def _exact(lib, V, B, mode, cost, H, state, target, origin, j, slow=False):
    v = np.ascontiguousarray(V, np.float64); b = np.ascontiguousarray(B, np.float64)
    st = np.ascontiguousarray(np.asarray(state, np.float64)[:3]); tg = np.ascontiguousarray(target, np.float64)
    og = np.ascontiguousarray(origin, np.float64)
    traj = np.empty((H, 3)); fc = ctypes.c_int(-1)
    dp = ctypes.POINTER(ctypes.c_double)
    p = lambda a: a.ctypes.data_as(dp)
    J = lib.mpcb_test_exact_cost(p(v), v.size, p(b), b.size, L, DT, mode, 0 if cost == C.COST_MM else 1, H,
                                 p(st), p(tg), p(og), int(j), p(traj), ctypes.byref(fc), 1 if slow else 0, C.CONFIG["v_min"])
    return J, traj, fc.value


def test_invariant_divisor_decoding(host):
    """FastDiv64 / FastDiv32 (leaf index -> control digits) against exact integer division: powers of grid sizes up to
    2^62, awkward divisors, numerators up to the largest leaf index."""
    rng = np.random.default_rng(0)
    divisors = [1, 2, 3, 7, 12, 255, 256, 257, 451, 1024, 24321, 2 ** 31 - 1, 2 ** 32, 2 ** 32 + 1, 451 ** 2, 451 ** 3,
                24321 ** 3, 1024 ** 5, 256 ** 7, 2 ** 62, 2 ** 62 + 12345]
    u64p = ctypes.POINTER(ctypes.c_uint64)
    for d in divisors:
        n = np.concatenate([rng.integers(0, 2 ** 63 - 1, 4000, dtype=np.uint64),
                            np.array([0, 1, d - 1, d, d + 1, 2 * d - 1, 2 * d, 2 ** 63 - 1], dtype=np.uint64),
                            (rng.integers(0, max(2, (2 ** 63 - 1) // d), 2000, dtype=np.uint64) * np.uint64(d))])
        n = n[n <= np.uint64(2 ** 63 - 1)]
        q = np.empty_like(n)
        host.mpcb_test_fastdiv64(d, n.size, n.ctypes.data_as(u64p), q.ctypes.data_as(u64p))
        np.testing.assert_array_equal(q, n // np.uint64(d), err_msg=f"d={d}")
    u32p = ctypes.POINTER(ctypes.c_uint32)
    for d in [1, 2, 3, 12, 255, 256, 451, 1024, 24321, 451 ** 2, 65536, 2 ** 31 - 1]:
        n = np.concatenate([rng.integers(0, 2 ** 32 - 1, 4000, dtype=np.uint32),
                            np.array([0, 1, d - 1, d, min(d + 1, 2 ** 32 - 1), 2 ** 32 - 1], dtype=np.uint32)])
        q = np.empty_like(n)
        host.mpcb_test_fastdiv32(d, n.size, n.ctypes.data_as(u32p), q.ctypes.data_as(u32p))
        np.testing.assert_array_equal(q, n // np.uint32(d), err_msg=f"d={d}")


def test_exact_evaluation_reproduces_the_reference_full_ticks(host, golden):
    """Every accepted tick of the reference's FULL closed loops (both scripts): the shipped float64 evaluation of the
    winning leaf gives the reference's criterion, trajectory and returned control."""
    g = golden("full_h3")
    checked = 0
    for case in g["cases"]:
        sc = case["scenario"]
        target, origin = (sc["x_t"], sc["y_t"]), (sc["x_0"], sc["y_0"])
        V, B = case["vector_v"], case["vector_beta"]
        for tick in case["ticks"]:
            r = C.solve_full(tick["state"], target, origin, V, B, 3, C.COST_MM, threshold=tick["threshold"])
            if not r["accepted"]:
                continue
            J, traj, fc = _exact(host, V, B, 0, C.COST_MM, 3, tick["state"], target, origin, r["index"])
            assert J == pytest.approx(tick["criterion_after"], rel=1e-12)
            np.testing.assert_allclose(traj, np.array(tick["traj"]), rtol=0, atol=1e-12)
            S = len(V) * len(B)
            assert fc == r["index"] // S ** 2
            vv, bb, _ = C.control_tables(V, B, L, DT)
            assert [vv[fc], bb[fc]] == pytest.approx(tick["ret"][3:5], abs=1e-12)
            checked += 1
    assert checked >= 40


def test_exact_evaluation_reproduces_the_reference_held_solves(host, golden):
    """The online controller's solves (HELD tree, tree-script cost), with and without the slow-down override
    (every velocity := max(min(V), v_min), math_model_tree.py:312-316)."""
    g = golden("held_single")
    checked = slowed = 0
    for case in g["cases"]:
        slow = bool(case["slow"])
        slowed += slow
        o = C.solve_held(case["state"], case["target"], case["origin"], case["vector_v"], case["vector_beta"], 3, C.COST_TREE,
                         slow=slow)
        J, traj, fc = _exact(host, case["vector_v"], case["vector_beta"], 1, C.COST_TREE, 3, case["state"], case["target"],
                             case["origin"], o["index"], slow=slow)
        assert J == pytest.approx(o["cost"], rel=1e-12)
        np.testing.assert_allclose(traj, np.array(case["traj"]), rtol=0, atol=1e-12)
        assert fc == o["index"]
        checked += 1
    assert checked >= 8 and slowed >= 1


@pytest.mark.parametrize("cost", [C.COST_MM, C.COST_TREE])
def test_exact_evaluation_matches_the_oracle_on_random_leaves(host, cost):
    """Random leaves of deeper trees (H = 4, 5: index decoding over more digits), incl. the line-origin special case."""
    rng = np.random.default_rng(5)
    V, B = np.linspace(0.0, 1.0, 6), np.linspace(-1.0, 1.0, 7)
    S = 42
    for H in (4, 5):
        for x in C.random_scenarios(6, 40 + H):
            for j in np.r_[0, S ** H - 1, rng.integers(0, S ** H, 20)]:
                J, traj, fc = _exact(host, V, B, 0, cost, H, x[:3], x[3:5], x[:2], int(j))
                Jo = C.exact_leaf_cost(x[:3], x[3:5], x[:2], V, B, int(j), H, cost)
                assert J == pytest.approx(Jo, rel=1e-12), (H, j)
                assert fc == int(j) // S ** (H - 1)


def test_operator_events_reproduce_the_reference(host, golden):
    """apply_event of csrc/mpcb_exact.cuh -- what the device-resident closed loop executes between two ticks -- against
    the reference's own new_target / turn_left / turn_right (+ slow_down) in every heading quadrant (bit-identical on
    the host, where sin / cos are the libm the reference used)."""
    g = golden("operator_events")
    kinds = {"new_target": 1, "turn_left": 2, "turn_right": 3}
    for c in g["cases"]:
        line = np.full(4, np.nan)
        slow = host.mpcb_test_apply_event(kinds[c["kind"]], c["a"], c["b"], g["radius_u_turn"], *c["pose"], c["slow_before"],
                                          line.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
        np.testing.assert_allclose(line, c["out"], rtol=0, atol=2e-15, err_msg=str(c))
        assert slow == c["steps_for_slowing"], c


@pytest.mark.parametrize("node32", [1, 0])
@pytest.mark.parametrize("H", [2, 3, 5])
def test_leaf_from_its_nodes_pose_equals_the_whole_walk(host, node32, H):
    """The refinement scan evaluates an in-window leaf from the float64 pose of its depth-(H-1) node (computed once per
    node) with ONE more step; that must be the very number exact_cost gets from walking all H steps -- bit for bit --
    and the node's controls (and those of its parent, which the tile test walks) must be the base-S digits of its index,
    with the 32-bit and the 64-bit dividers alike."""
    rng = np.random.default_rng(40 + H)
    V, B = np.linspace(0.0, 1.0, 7), np.linspace(-1.0, 1.0, 9)
    S = V.size * B.size
    dp = ctypes.POINTER(ctypes.c_double)
    p = lambda a: a.ctypes.data_as(dp)
    for cost in (0, 1):
        for slow in (0, 1):
            st, tg, og = rng.uniform(-3, 3, 3), rng.uniform(-3, 3, 2), rng.uniform(-1, 1, 2)
            j = rng.integers(0, S ** H, 400).astype(np.int64)
            j[:3] = (0, S ** H - 1, S ** (H - 1))
            out, whole = np.empty(j.size), np.empty(j.size)
            digits = np.zeros(j.size * (H - 1), np.uint32)
            pdigits = np.zeros(max(1, j.size * (H - 2)), np.uint32)
            v, b = np.ascontiguousarray(V), np.ascontiguousarray(B)
            host.mpcb_test_exact_from_node(p(v), v.size, p(b), b.size, L, DT, cost, H, p(st), p(tg), p(og), slow,
                                           C.CONFIG["v_min"], node32, j.size,
                                           j.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)), p(out), p(whole),
                                           digits.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)),
                                           pdigits.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)))
            assert np.array_equal(out.view(np.uint64), whole.view(np.uint64))
            want = np.array([[(int(q) // S) // S ** (H - 2 - k) % S for k in range(H - 1)] for q in j], np.uint32)
            np.testing.assert_array_equal(digits.reshape(j.size, H - 1), want)
            if H > 2:       # the tile test walks the node's parent: the same digits without the last
                np.testing.assert_array_equal(pdigits.reshape(j.size, H - 2), want[:, :-1])
