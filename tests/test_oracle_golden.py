"""The oracle (numpy restatement + C restatement) against the committed fixtures that
were produced by the reference's own unmodified functions (oracle/make_golden.py)."""
import math

import numpy as np
import pytest

from oracle import c_oracle as K
from oracle import closed_form as C


def test_step_matches_reference_quad(golden):
    g = golden("pieces")
    worst = 0.0
    for s in g["steps"]:
        out = C.step(s["state"], s["v"], s["beta"])
        ref = np.array(s["out"])
        worst = max(worst, np.max(np.abs(np.array(out) - ref) / np.maximum(1.0, np.abs(ref))))
    # quad integrates a constant over [t, t+dt]; (t+dt)-t != dt in floating point -> ~1e-14
    assert worst < 1e-13


def test_costs_match_reference(golden):
    g = golden("pieces")
    for c in g["costs"]:
        x, y, p = c["state"]
        assert float(C.leaf_cost(x, y, p, c["target"], c["origin"], C.COST_MM)) == pytest.approx(c["mm"], rel=1e-14)
        assert float(C.leaf_cost(x, y, p, c["target"], c["origin"], C.COST_TREE)) == pytest.approx(c["tree"], rel=1e-14)
    # origin special case: distance 1000 (MM) / 1000**2 (TREE)
    last = g["costs"][-1]
    assert last["mm"] > 1e8 and last["tree"] > 1e10


def test_grid_generators(golden):
    g = golden("pieces")
    for e in g["grids"]:
        fn = C.vector_of_velocities if e["kind"] == "v" else C.vector_of_beta_angles
        assert fn(e["arg"]) == e["out"]
    v, b = C.grid_full_default()
    d = g["full_default_grid"]
    assert (v.size, b.size) == (d["nv"], d["nb"]) == (201, 121)
    assert list(v[:3]) == d["v_head"] and list(b[-3:]) == d["b_tail"]


@pytest.mark.parametrize("impl", ["numpy", "c"])
def test_full_h3_against_reference(golden, impl):
    g = golden("full_h3")
    solve = C.solve_full if impl == "numpy" else K.solve_full
    n_stall = 0
    for case in g["cases"]:
        sc = case["scenario"]
        target, origin = (sc["x_t"], sc["y_t"]), (sc["x_0"], sc["y_0"])
        prev_ret = None
        for tick in case["ticks"]:
            r = solve(tick["state"], target, origin, case["vector_v"], case["vector_beta"], 3, C.COST_MM,
                      threshold=tick["threshold"])
            if r["accepted"]:
                assert r["cost"] == pytest.approx(tick["criterion_after"], rel=1e-13)
                np.testing.assert_allclose(r["traj"], np.array(tick["traj"]), rtol=0, atol=1e-12)
                ret = list(r["traj"][0]) + list(r["first_control"])
                np.testing.assert_allclose(ret, tick["ret"], rtol=0, atol=1e-12)
            else:  # stall: the reference returns the previously accepted path again
                n_stall += 1
                assert tick["ret"] == prev_ret
                assert tick["criterion_after"] == tick["threshold"]
            prev_ret = tick["ret"]
    assert n_stall >= 5  # the fixtures do exercise the carried threshold


@pytest.mark.parametrize("impl", ["numpy", "c"])
def test_held_single_against_reference(golden, impl):
    g = golden("held_single")
    solve = C.solve_held if impl == "numpy" else K.solve_held
    assert any(c["slow"] for c in g["cases"])
    for c in g["cases"]:
        r = solve(c["state"], c["target"], c["origin"], c["vector_v"], c["vector_beta"], 3, C.COST_TREE,
                  slow=c["slow"])
        np.testing.assert_allclose(r["traj"], np.array(c["traj"]), rtol=0, atol=1e-12)
        ret = list(r["traj"][0]) + list(r["first_control"])
        np.testing.assert_allclose(ret, c["ret"], rtol=0, atol=1e-12)


def test_c_and_numpy_oracles_agree_on_all_leaves():
    sc = C.random_scenarios(4, 11)
    V, B = [0.0, 0.3, 0.6, 1.0], np.linspace(-1, 1, 7)
    for s in sc:
        for cost in (C.COST_MM, C.COST_TREE):
            for H in (1, 2, 3, 4):
                a = K.full_leaf_costs(s[:3], s[3:], s[:2], V, B, H, cost)
                b = C.full_leaf_costs(s[:3], s[3:], s[:2], V, B, H, cost)
                np.testing.assert_allclose(a, b, rtol=1e-14)
                r1 = K.solve_full(s[:3], s[3:], s[:2], V, B, H, cost)
                r2 = C.solve_full(s[:3], s[3:], s[:2], V, B, H, cost)
                assert r1["index"] == r2["index"] == int(np.argmin(b))


def test_split_tree_ranges_recombine():
    """One rank's share = a contiguous range of first controls; lexicographic (cost,index)
    min over the shares equals the whole-tree answer."""
    s = C.random_scenarios(1, 5)[0]
    V, B = [0.0, 0.5, 1.0], np.linspace(-1, 1, 5)
    whole = K.solve_full(s[:3], s[3:], s[:2], V, B, 3)
    parts = [K.solve_full(s[:3], s[3:], s[:2], V, B, 3, i0_range=(a, b)) for a, b in ((0, 4), (4, 9), (9, 15))]
    best = min(parts, key=lambda r: (r["cost"], r["raw_index"]))
    assert (best["cost"], best["raw_index"]) == (whole["cost"], whole["raw_index"])


def test_zero_velocity_ties_pick_first_leaf():
    # v=0 makes beta irrelevant: exact ties, first leaf in enumeration order wins (math_model.py:195)
    V, B = [0.0], np.linspace(-1, 1, 5)
    r = C.solve_full([1.0, 2.0, 0.3], (3.0, 4.0), (0.0, 0.0), V, B, 3)
    assert r["index"] == 0
    r = C.solve_held([1.0, 2.0, 0.3], (3.0, 4.0), (0.0, 0.0), V, B, 3)
    assert r["index"] == 0
