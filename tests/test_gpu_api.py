"""GPU: the reference-API modules end to end (real CUDA backend) against the reference's logs."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_math_mpc_closed_loop_matches_reference_log(golden):
    mt = importlib.reload(importlib.import_module("diplomjourney_b200.math_model_tree"))
    mt._backend = None
    mt.math_mpc([0, 0, 0, 0, 0], [2, 3], False)
    log = golden("held_closed_loop")["log"]
    for key in ("result_trajectory_x", "result_trajectory_y", "result_trajectory_phi", "result_trajectory_v",
                "result_trajectory_beta", "predicted_trajectory_x_anim2", "predicted_trajectory_y_anim1"):
        np.testing.assert_allclose(np.array(getattr(mt, key), dtype=float), log[key], rtol=0, atol=1e-9, err_msg=key)
    assert (mt.p, mt.m, mt.steps_for_slowing) == (log["final"]["p"], log["final"]["m"], log["final"]["steps_for_slowing"])


def test_full_module_matches_reference_ticks(golden):
    rm = importlib.import_module("diplomjourney_b200.run_math_model")
    g = golden("full_h3")
    for case in [c for c in g["cases"] if c["script"] == "run_math_model.py"]:
        sc = case["scenario"]
        rm._backend = None
        rm._grid_key = None
        rm.vector_v, rm.vector_beta = np.array(case["vector_v"]), np.array(case["vector_beta"])
        rm.reset_scenario(sc["x_0"], sc["y_0"], sc["phi_0"], sc["x_t"], sc["y_t"])
        for tick in case["ticks"]:
            rm.optimal_criterion = tick["threshold"]
            r = rm.predictive_control(*tick["state"], 0, sc["x_t"], sc["y_t"])
            np.testing.assert_allclose(r, tick["ret"], rtol=0, atol=1e-12)
            assert rm.optimal_criterion == pytest.approx(tick["criterion_after"], rel=1e-12)


def test_actual_mode_runs_with_seeded_noise():
    mt = importlib.reload(importlib.import_module("diplomjourney_b200.math_model_tree"))
    np.random.seed(3)
    mt.math_mpc([0, 0, 0, 0, 0], [2, 3], True)
    assert len(mt.actual_result_trajectory_x) > 20
    assert mt.is_on_target(mt.actual_result_trajectory_x[-1], mt.actual_result_trajectory_y[-1], mt.x_t, mt.y_t)[0] \
        or mt.recursive


def _window_params(max_ticks=200):
    from diplomjourney_b200 import _native as nat, config
    return nat, nat.LoopParams.from_config(config, nat.COST_TREE, 3, max_ticks)


def test_device_closed_loop_matches_reference_runs(golden):
    """mpcb_held_closed_loop: whole event-free math_mpc runs on the device vs the reference's own logs
    (incl. one run that ends in the 'Recursive error' stall)."""
    nat, params = _window_params()
    g = golden("held_short_loops")
    s = nat.Solver(0)
    cases = g["cases"]
    r = s.held_closed_loop(params, [c["init"] for c in cases], [c["target"] for c in cases],
                           [c["origin"] for c in cases], first_threshold=[c["first_threshold"] for c in cases])
    for i, c in enumerate(cases):
        ref = np.array(c["log"]).T                      # [ticks][5]
        assert r["ticks"][i] == ref.shape[0], (i, r["ticks"][i], ref.shape)
        np.testing.assert_allclose(r["log"][i, :ref.shape[0]], ref, rtol=0, atol=1e-9)
        assert r["status"][i] == (nat.LOOP_STALLED if c["recursive"] else nat.LOOP_ON_TARGET)
    s.close()


def test_device_closed_loop_batch_equals_per_tick_host_loop():
    """A batch of random robots: the device-resident loop and the per-tick host loop (math_mpc with the
    scripted events switched off, every tick one GPU HELD solve) produce the same trajectories."""
    mt = importlib.reload(importlib.import_module("diplomjourney_b200.math_model_tree"))
    rng = np.random.default_rng(11)
    n = 6
    init = np.zeros((n, 5))
    init[:, 2] = rng.uniform(-1.0, 1.0, n)
    init[:, 3] = rng.choice([0.0, 0.3, 0.6], n)
    ang = init[:, 2] + rng.uniform(-0.6, 0.6, n)
    dist_ = rng.uniform(0.8, 2.5, n)
    tgt = np.stack([dist_ * np.cos(ang), dist_ * np.sin(ang)], 1)
    batch = mt.math_mpc_batch(init, tgt, max_ticks=300)
    assert set(batch["status"]) <= {0, 1}
    mt.scripted_events = False
    for i in range(n):
        mt.reset_state()
        mt.x_t, mt.y_t = tgt[i]
        mt.optimal_criterion = mt.control_criterion([mt.x_0, mt.y_0, mt.phi_0])
        mt.math_mpc(list(init[i]), list(tgt[i]), False)
        host = np.array([mt.result_trajectory_x[1:], mt.result_trajectory_y[1:], mt.result_trajectory_phi[1:],
                         mt.result_trajectory_v[1:], mt.result_trajectory_beta[1:]], dtype=float).T
        k = batch["ticks"][i]
        assert k == host.shape[0], (i, k, host.shape)
        np.testing.assert_allclose(batch["log"][i, :k], host, rtol=0, atol=1e-9)


def test_actual_mode_matches_reference_seeded_runs(golden):
    mt = importlib.reload(importlib.import_module("diplomjourney_b200.math_model_tree"))
    for c in golden("held_actual")["cases"]:
        mt.x_0, mt.y_0, mt.phi_0 = 0, 0, 0
        mt.reset_state()
        np.random.seed(c["seed"])
        mt.math_mpc([0, 0, 0, 0, 0], [2, 3], True)
        for key, ref in c["log"].items():
            np.testing.assert_allclose(np.array(getattr(mt, key), dtype=float), ref, rtol=0, atol=1e-9, err_msg=key)
        assert (mt.p, mt.m) == (c["p"], c["m"])


def test_full_module_leaf_cloud():
    from oracle import closed_form as C
    rm = importlib.import_module("diplomjourney_b200.run_math_model")
    rm._backend = None
    rm._grid_key = None
    rm.vector_v, rm.vector_beta = np.array([0.0, 0.4, 0.7, 1.0]), np.round(np.radians([-60, -30, 0, 30, 60]), 3)
    rm.reset_scenario(0.5, -1.0, 0.3, 3.0, 2.0)
    x, y, J = rm.leaf_cloud(0.5, -1.0, 0.3)
    Jo, xo, yo, _ = C.full_leaf_costs([0.5, -1.0, 0.3], (3.0, 2.0), (0.5, -1.0), rm.vector_v, rm.vector_beta, 3,
                                      C.COST_MM, return_xy=True)
    assert x.shape == (20 ** 3,)
    np.testing.assert_allclose(x, xo, atol=2e-6)
    np.testing.assert_allclose(y, yo, atol=2e-6)
    ok = Jo < 1e7
    np.testing.assert_allclose(J[ok], Jo[ok], atol=5e-3)


def test_full_closed_loop_batch_matches_reference_ticks(golden):
    """mpcb_full_closed_loop (run_batch): all scenarios of a grid in one batched loop vs the reference's own
    tick-by-tick returns, incl. carried thresholds and the repeated-position stop."""
    rm = importlib.import_module("diplomjourney_b200.run_math_model")
    g = golden("full_h3")
    for grid in ("g4x5", "g5x7", "g3x9"):
        cases = [c for c in g["cases"] if c["script"] == "run_math_model.py" and c["grid"] == grid]
        rm._backend = None
        rm._grid_key = None
        rm.vector_v, rm.vector_beta = np.array(cases[0]["vector_v"]), np.array(cases[0]["vector_beta"])
        sc = np.array([[c["scenario"][k] for k in ("x_0", "y_0", "phi_0", "x_t", "y_t")] for c in cases], dtype=float)
        nt = max(len(c["ticks"]) for c in cases)
        r = rm.run_batch(sc, max_ticks=nt)
        for i, c in enumerate(cases):
            k_rep, knife_seen = 0, False
            for t, tick in enumerate(c["ticks"]):
                if t >= r["ticks"][i]:
                    # our loop stopped (the fixture ran a fixed number of ticks): it must have had the reference's reason
                    last = r["log"][i, t - 1]
                    on_target = (c["scenario"]["x_t"] - last[0]) ** 2 + (c["scenario"]["y_t"] - last[1]) ** 2 <= 0.001
                    assert (r["status"][i] == 0 and on_target) or (r["status"][i] == 1 and k_rep >= 2), (grid, i, t)
                    break
                knife_edge = (abs(tick["criterion_after"] - tick["threshold"]) < 1e-9 * tick["threshold"]
                              and tick["criterion_after"] != tick["threshold"])
                knife_seen = knife_seen or knife_edge               # ... and the stalls after it repeat that choice
                cols = slice(0, 3) if knife_seen else slice(0, 5)   # accepted by 3e-13: controls may differ, pose not
                np.testing.assert_allclose(r["log"][i, t, cols], tick["ret"][cols], rtol=0, atol=1e-12,
                                           err_msg=f"{grid} scenario {i} tick {t}")
                if t > 0 and tick["ret"][:2] == c["ticks"][t - 1]["ret"][:2]:
                    k_rep += 1
                elif t == 0 and tick["ret"][:2] == tick["state"][:2]:
                    k_rep += 1


def test_integration_md_stub_runs_verbatim(golden):
    """The ctypes stub printed in INTEGRATION.md section 2, executed as written (only the library path is filled in),
    reproduces a reference HELD solve."""
    import os
    import re
    from diplomjourney_b200 import _native, config
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    md = open(os.path.join(root, "INTEGRATION.md")).read()
    code = re.search(r"```python\n(import ctypes as C.*?)```", md, re.S).group(1)
    code = code.replace("/path/to/diplomjourney_b200/lib/libmpcb200.so", _native.library_path())
    c = golden("held_single")["cases"][1]
    ns = {k: getattr(config, k) for k in config.__all__}
    ns.update(x_t=c["target"][0], y_t=c["target"][1], x_0=c["origin"][0], y_0=c["origin"][1])
    exec(compile(code, "INTEGRATION.md", "exec"), ns)
    cost, k, traj, ctl = ns["_gpu_held_solve"](c["state"], c["vector_v"], c["vector_beta"], c["slow"], 1e300)
    assert k >= 0
    np.testing.assert_allclose(list(traj[0]) + list(ctl), c["ret"], rtol=0, atol=1e-12)


def test_held_windows_batch_equals_per_robot_solves():
    """mpcb_solve_held_windows: one launch for a batch of robots that each have their own acceleration window
    (math_model_tree.py:239-256) == the reference flow robot by robot (window lists built on the host, set_grid +
    HELD solve), bit for bit, and == the float64 oracle.  Covers clipped windows (v near 0 and v_max, beta at the
    limit), the slow-down override, thresholds that reject, skipped entries and an empty velocity window."""
    from oracle import closed_form as C
    nat, params = _window_params()
    mt = importlib.import_module("diplomjourney_b200.math_model_tree")
    s = nat.Solver(0)
    rng = np.random.default_rng(4)
    n = 96
    sc = C.random_scenarios(n, 31)
    vb = np.stack([rng.uniform(0.0, 0.995, n), rng.uniform(-1.05, 1.05, n)], 1)
    vb[0] = (0.0, 0.0); vb[1] = (0.999, 1.04); vb[2] = (0.01, -1.047); vb[3] = (1.2, 0.0)      # [3]: empty velocity window
    flags = np.zeros(n, np.uint8); flags[5:20:3] = nat.FLAG_SLOW; flags[7] = nat.FLAG_SKIP
    thr = np.full(n, np.inf); thr[10:14] = 1.0
    r = s.solve_held_windows(params, sc[:, :3], vb, sc[:, 3:5], sc[:, :2], threshold=thr, flags=flags)
    for i in range(n):
        V, B = mt.vector_of_velocities(vb[i, 0]), mt.vector_of_beta_angles(vb[i, 1])
        assert tuple(r["shape"][i]) == ((0, 0) if flags[i] & nat.FLAG_SKIP else (len(V), len(B))), i
        if not V or not B or flags[i] & nat.FLAG_SKIP:
            assert r["index"][i] == -1 and np.isnan(r["cost"][i])
            continue
        s.set_grid(V, B, params.L, params.delta_t, params.v_min)
        one = s.solve(nat.MODE_HELD, nat.COST_TREE, 3, sc[i, :3], sc[i, 3:5], sc[i, :2], threshold=thr[i],
                      flags=int(flags[i]))
        assert r["index"][i] == one["index"][0] and r["cost"][i] == one["cost"][0], i
        np.testing.assert_array_equal(r["traj"][i], one["traj"][0])
        np.testing.assert_array_equal(r["first_control"][i], one["first_control"][0])
        o = C.solve_held(sc[i, :3], sc[i, 3:5], sc[i, :2], V, B, 3, C.COST_TREE, threshold=thr[i],
                         slow=bool(flags[i] & nat.FLAG_SLOW), v_min=params.v_min)
        assert r["index"][i] == o["index"] and r["cost"][i] == pytest.approx(o["cost"], rel=1e-12)
    assert 3 in [i for i in range(n) if r["shape"][i][0] == 0]
    s.close()


def test_actual_mode_batch_matches_reference_seeded_runs(golden):
    """Row f2/f3 on the device: N seeded actual-mode robots (actuator noise, operator events) in ONE batch -- one
    mpcb_solve_held_windows launch per tick -- equal the reference's own seeded runs (held_actual.json) and a
    sequential math_mpc(..., True) run of the module."""
    from test_host_api import _actual_batch_equals_fixture_and_sequential
    mt = importlib.reload(importlib.import_module("diplomjourney_b200.math_model_tree"))
    mt._backend = None
    _actual_batch_equals_fixture_and_sequential(mt, golden)


def test_device_closed_loop_with_operator_events(golden):
    """mpcb_held_closed_loop_events: the reference's programmed run math_mpc([0,0,0,0,0], [2,3], False) -- 150 ticks with
    turn_right at 60, turn_left at 90, new_target at 110 and the slow-down override after each -- executed in ONE
    launch, events applied on the device, against the reference's own log; and a batch of robots that start elsewhere
    against the per-tick path on which the module's own new_target / turn_* functions apply the events."""
    nat, _ = _window_params()
    mt = importlib.reload(importlib.import_module("diplomjourney_b200.math_model_tree"))
    mt._backend = None
    log = golden("held_closed_loop")["log"]
    rng = np.random.default_rng(5)
    n = 24
    init = np.zeros((n, 5)); init[1:, :2] = rng.uniform(-1, 1, (n - 1, 2)); init[1:, 2] = rng.uniform(-0.5, 7.0, n - 1)
    tgt = np.tile([2.0, 3.0], (n, 1)); tgt[1:] += rng.uniform(-1, 1, (n - 1, 2))
    r = mt.math_mpc_batch(init, tgt, max_ticks=200, events=True)
    ref = np.array([log[k][1:] for k in ("result_trajectory_x", "result_trajectory_y", "result_trajectory_phi",
                                         "result_trajectory_v", "result_trajectory_beta")], dtype=float).T
    fin = log["final"]
    assert r["ticks"][0] == ref.shape[0] == fin["p"] - 1 and r["status"][0] == nat.LOOP_ON_TARGET
    np.testing.assert_allclose(r["log"][0, :ref.shape[0]], ref, rtol=0, atol=1e-9)
    np.testing.assert_allclose(r["final"][0], [2.0, 3.0, fin["x_0"], fin["y_0"], fin["steps_for_slowing"], fin["m"]],
                               rtol=0, atol=1e-9)
    host = mt.math_mpc_batch(init, tgt, max_ticks=200, events=True, host_loop=True)
    np.testing.assert_array_equal(r["ticks"], host["ticks"])
    np.testing.assert_array_equal(r["status"], host["status"])
    np.testing.assert_allclose(r["log"], host["log"], rtol=0, atol=1e-9, equal_nan=True)
    for i, c in enumerate(host["robots"]):
        np.testing.assert_allclose(r["final"][i], [c["x_t"], c["y_t"], c["x_0"], c["y_0"], c["steps_for_slowing"], c["m"]],
                                   rtol=0, atol=1e-9)
    assert (r["ticks"] > 110).sum() >= 4 and len({int(q) for q in r["ticks"]}) > 3      # events really fired, runs differ
    # a custom script: one new target after tick 5 == two event-free loops chained by hand
    a = mt.math_mpc_batch(init[:1], tgt[:1], max_ticks=200, events=[(5, nat.EVENT_NEW_TARGET, -1.0, 1.5)])
    assert a["final"][0, 0] == -1.0 and a["final"][0, 1] == 1.5 and a["ticks"][0] > 5
    np.testing.assert_array_equal(a["final"][0, 2:4], a["log"][0, 4, :2])               # the line restarts at the pose of tick 5
    np.testing.assert_array_equal(a["log"][0, :5], r["log"][0, :5])


def test_device_closed_loop_many_robots_per_cta():
    """More robots than resident CTAs: every CTA runs several robots back to back (the stop flags of one robot's loop
    must not leak into the next).  4,096 robots == the same robots in batches of 64."""
    nat, params = _window_params(max_ticks=256)
    rng = np.random.default_rng(0)
    n = 4096
    init = np.zeros((n, 5)); init[:, 2] = rng.uniform(-1, 1, n)
    ang = init[:, 2] + rng.uniform(-0.5, 0.5, n); d = rng.uniform(0.05, 1.0, n)
    tgt = np.stack([d * np.cos(ang), d * np.sin(ang)], 1)
    s = nat.Solver(0)
    big = s.held_closed_loop(params, init, tgt, [[0.0, 0.0]], first_threshold=1e10)
    for lo in range(0, n, 1024):
        part = s.held_closed_loop(params, init[lo:lo + 64], tgt[lo:lo + 64], [[0.0, 0.0]], first_threshold=1e10)
        np.testing.assert_array_equal(big["ticks"][lo:lo + 64], part["ticks"])
        np.testing.assert_array_equal(big["status"][lo:lo + 64], part["status"])
        np.testing.assert_array_equal(big["log"][lo:lo + 64], part["log"])      # rows past a robot's last tick are NaN
    assert nat.LOOP_ON_TARGET in set(np.unique(big["status"]))
    assert np.isnan(big["log"][0, big["ticks"][0]:]).all() and not np.isnan(big["log"][0, :big["ticks"][0]]).any()
    s.close()


def test_held_tick_one_call_equals_the_two_call_path():
    """mpcb_held_tick_host (window lists + HELD solve in one call; inputs and results in mapped pinned host memory) ==
    set_grid + solve, with the staged-copy flavour and with the slow-down override, on ticks of the reference run."""
    from diplomjourney_b200 import config
    nat, _ = _window_params()
    mt = importlib.import_module("diplomjourney_b200.math_model_tree")
    s = nat.Solver(0)
    rng = np.random.default_rng(11)
    for k in range(24):
        v, beta = rng.uniform(0.0, 0.99), rng.uniform(-1.0, 1.0)
        V, B = mt.vector_of_velocities(v), mt.vector_of_beta_angles(beta)
        st, tg, og = rng.uniform(-3, 3, 3), rng.uniform(-3, 3, 2), rng.uniform(-1, 1, 2)
        thr = float("inf") if k % 5 else 1.0
        flags = nat.FLAG_SLOW if k % 3 == 0 else 0
        s.set_grid(V, B, config.L, config.delta_t, config.v_min)
        ref = s.solve(nat.MODE_HELD, nat.COST_TREE, 3, st, tg, og, threshold=thr, flags=flags)
        for zc in (1, 0):
            s.set_option("zero_copy", zc)
            cost, idx, traj, ctl = s.held_tick(V, B, config.L, config.delta_t, config.v_min, nat.COST_TREE, 3, st, tg, og, thr, flags)
            assert (cost, idx) == (ref["cost"][0], ref["index"][0])
            np.testing.assert_array_equal(traj, ref["traj"][0])
            assert ctl == tuple(ref["first_control"][0])
    s.set_option("zero_copy", 1)
    s.close()
