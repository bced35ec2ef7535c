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
