"""Container-only: the restatement against the reference's own functions executed live
(oracle/ref_exec.py).  Skipped where /root/reference is absent (the GPU box)."""
import numpy as np
import pytest

from oracle import closed_form as C
from oracle import ref_exec as R

pytestmark = pytest.mark.skipif(not R.reference_available(), reason="needs /root/reference")


def test_full_live_small_grid():
    V, B = [0.0, 0.6, 1.0], np.round(np.radians([-50, 0, 50]), 3)
    sc = dict(x_0=1.5, y_0=-2.0, phi_0=0.7, x_t=4.0, y_t=1.0)
    m = R.load_full("run_math_model.py", vector_v=V, vector_beta=B, overrides=sc)
    thr = float(m.optimal_criterion)
    st = [sc["x_0"], sc["y_0"], sc["phi_0"]]
    for _ in range(2):
        r = m.predictive_control(st[0], st[1], st[2], 0, sc["x_t"], sc["y_t"])
        o = C.solve_full(st, (sc["x_t"], sc["y_t"]), (sc["x_0"], sc["y_0"]), V, B, 3, C.COST_MM, threshold=thr)
        assert o["accepted"]
        np.testing.assert_allclose(list(o["traj"][0]) + list(o["first_control"]), r, atol=1e-12, rtol=0)
        assert o["cost"] == pytest.approx(float(m.optimal_criterion), rel=1e-13)
        thr, st = o["cost"], r[:3]


def test_held_live_default_window():
    T = R.load_tree()
    T.x_t, T.y_t = 2, 3
    V, B = T.vector_of_velocities(0.5), T.vector_of_beta_angles(0.0)
    assert (len(V), len(B)) == (11, 41)
    T.optimal_criterion = 1e300
    r = T.predictive_control(0.1, 0.2, 0.3, 2, 3, V, B, False)
    o = C.solve_held([0.1, 0.2, 0.3], (2, 3), (0, 0), V, B, 3, C.COST_TREE)
    np.testing.assert_allclose(list(o["traj"][0]) + list(o["first_control"]), r, atol=1e-12, rtol=0)
