"""TEST INFRASTRUCTURE ONLY -- regenerates tests/golden/*.json from the reference itself.

Runs ONLY in the build container (needs /root/reference).  Every number written
here is an output of the reference's own, unmodified function bodies executed
through ``oracle/ref_exec.py``; nothing comes from our restatement.

    python -m oracle.make_golden            # ~2 min, rewrites tests/golden/

Fixtures:
  full_h3.json        FULL predictive_control (run_math_model.py:133-228 and
                      math_model.py:136-231) on reduced grids: per tick the inputs
                      (state, threshold) and outputs (returned 5-list, accepted
                      cost, the 3 predicted poses), incl. threshold carry + stalls.
  held_closed_loop.json  math_mpc([0,0,0,0,0],[2,3],False) (math_model_tree.py:515-579):
                      the whole programmed closed loop, S<=451, with the scripted
                      operator events at ticks 60/90/110 and the slow-down override.
  held_single.json    single HELD predictive_control calls incl. slow flag.
  pieces.json         iteration_of_predict / control_criterion / grid generators.
  operator_events.json  new_target / turn_left / turn_right / slow_down in every heading quadrant.
"""
from __future__ import annotations

import json
import math
import os
import sys

import numpy as np
import scipy

from . import ref_exec as R

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _f(x):
    if isinstance(x, (list, tuple, np.ndarray)):
        return [_f(v) for v in x]
    return float(x)


def _meta():
    return dict(generator="oracle/make_golden.py", numpy=np.__version__, scipy=scipy.__version__,
                python=sys.version.split()[0], source="unmodified reference functions via AST extraction")


def gen_full():
    cases = []
    grids = {
        "g4x5": ([0.0, 0.4, 0.7, 1.0], np.round(np.radians([-60, -30, 0, 30, 60]), 3)),
        "g5x7": ([0, 0.25, 0.5, 0.75, 1.0], np.round(np.radians([-60, -40, -20, 0, 20, 40, 60]), 3)),
        "g3x9": ([0.0, 0.5, 1.0], np.round(np.arange(-math.radians(60), math.radians(60) + 0.26, math.radians(15)), 3)),
    }
    rng = np.random.default_rng(20261018)
    scen = [dict(x_0=0, y_0=0, phi_0=0, x_t=1, y_t=5)]
    for _ in range(5):
        x0, y0 = rng.uniform(-10, 10, 2)
        scen.append(dict(x_0=float(x0), y_0=float(y0), phi_0=float(rng.uniform(-math.pi, math.pi)),
                         x_t=float(rng.uniform(x0 - 10, x0 + 10)), y_t=float(rng.uniform(y0 - 10, y0 + 10))))
    # a near-target scenario so that the carried threshold produces stalls
    scen.append(dict(x_0=0.0, y_0=0.0, phi_0=1.2, x_t=0.06, y_t=0.17))
    for script in ("run_math_model.py", "math_model.py"):
        for gname, (V, B) in grids.items():
            for si, sc in enumerate(scen):
                if script == "math_model.py" and (gname != "g4x5" or si > 2):
                    continue
                nticks = 6 if si == len(scen) - 1 else 3
                m = R.load_full(script, vector_v=V, vector_beta=B, overrides=sc)
                x, y, phi, v = sc["x_0"], sc["y_0"], sc["phi_0"], 0
                ticks = []
                for _ in range(nticks):
                    thr_in = float(m.optimal_criterion)
                    try:
                        r = m.predictive_control(x, y, phi, v, sc["x_t"], sc["y_t"])
                    except (TypeError, IndexError) as e:  # first-ever solve with no improving leaf
                        ticks.append(dict(state=_f([x, y, phi]), threshold=thr_in, error=type(e).__name__))
                        break
                    traj = [_f(list(p)[:3]) for p in m.optimal_trajectory[0]]
                    ticks.append(dict(state=_f([x, y, phi]), threshold=thr_in, ret=_f(r),
                                      criterion_after=float(m.optimal_criterion), traj=traj))
                    x, y, phi, v = r[0], r[1], r[2], r[3]
                cases.append(dict(script=script, grid=gname, vector_v=_f(V), vector_beta=_f(B),
                                  scenario=sc, ticks=ticks))
    return dict(meta=_meta(), cost="mm", H=3, cases=cases)


def gen_held_closed_loop():
    T = R.load_tree()
    T.math_mpc([0, 0, 0, 0, 0], [2, 3], False)
    keys = ["result_trajectory_x", "result_trajectory_y", "result_trajectory_phi", "result_trajectory_v",
            "result_trajectory_beta", "result_trajectory_angle_speed", "time_arr_for_plotting",
            "result_x_velocity", "result_y_velocity", "result_x_acceleration", "result_y_acceleration"]
    for comp in ("x", "y", "phi"):
        for k in range(3):
            keys.append(f"predicted_trajectory_{comp}_anim{k}")
    out = {k: _f(T.ns[k]) for k in keys}
    out["final"] = dict(x_t=float(T.x_t), y_t=float(T.y_t), x_0=float(T.x_0), y_0=float(T.y_0),
                        phi_0=float(T.phi_0), p=int(T.p), m=int(T.m), steps_for_slowing=int(T.steps_for_slowing),
                        recursive=bool(T.recursive))
    return dict(meta=_meta(), call="math_mpc([0,0,0,0,0],[2,3],False)", cost="tree", H=3, log=out)


def gen_held_short_loops():
    """Event-free closed loops: targets close enough that math_mpc finishes before its first
    scripted operator event at tick 60 (math_model_tree.py:564)."""
    cases = []
    for init, tgt in (([0, 0, 0, 0, 0], [1.0, 0.8]), ([0, 0, 1.0, 0, 0], [0.3, 1.1]),
                      ([0.2, -0.1, -0.7, 0.3, 0.1], [1.2, -0.9]), ([0, 0, 3.0, 0, 0], [0.9, 0.2]),
                      ([0, 0, 0.4, 0.8, -0.3], [1.5, 1.0])):
        T = R.load_tree()
        T.math_mpc(list(init), list(tgt), False)
        assert T.p < 60, T.p
        cases.append(dict(init=_f(init), target=_f(tgt), origin=[float(T.x_0), float(T.y_0)],
                          first_threshold=float(T.control_criterion([T.x_0, T.y_0, T.phi_0])),
                          ticks=int(T.p) - 1, m=int(T.m), recursive=bool(T.recursive),
                          steps_for_slowing=int(T.steps_for_slowing),
                          log=[_f(T.ns["result_trajectory_" + k][1:]) for k in ("x", "y", "phi", "v", "beta")]))
    return dict(meta=_meta(), cost="tree", H=3, cases=cases)


def gen_held_actual():
    """math_mpc(..., isActual=True) (math_model_tree.py:580-629): actuator noise from numpy.random
    (what ``random`` is after the reference's ``from scipy import *``), seeded, plus the operator
    events at ticks 1/60/90/110."""
    cases = []
    for seed in (5, 11):
        T = R.load_tree()
        np.random.seed(seed)
        T.math_mpc([0, 0, 0, 0, 0], [2, 3], True)
        cases.append(dict(seed=seed, p=int(T.p), m=int(T.m), recursive=bool(T.recursive),
                          log={k: _f(T.ns[k]) for k in ("actual_result_trajectory_x", "actual_result_trajectory_y",
                                                        "actual_result_trajectory_phi", "actual_result_trajectory_v",
                                                        "actual_result_trajectory_beta")}))
    return dict(meta=_meta(), call="np.random.seed(seed); math_mpc([0,0,0,0,0],[2,3],True)", cases=cases)


def gen_held_single():
    cases = []
    rng = np.random.default_rng(7)
    for i in range(12):
        T = R.load_tree()
        x0, y0 = rng.uniform(-3, 3, 2)
        phi = float(rng.uniform(-math.pi, math.pi))
        xt, yt = float(rng.uniform(x0 - 4, x0 + 4)), float(rng.uniform(y0 - 4, y0 + 4))
        v_now = float(rng.choice([0.0, 0.02, 0.3, 0.5, 0.97, 0.995]))
        b_now = float(rng.choice([0.0, 0.3, -1.0, 1.04, -1.047]))
        T.x_t, T.y_t = xt, yt
        # line origin: half of the cases use the start (origin special case reachable when v grid has 0)
        if i % 2 == 0:
            T.x_0, T.y_0 = float(x0), float(y0)
        else:
            T.x_0, T.y_0 = float(x0 - 0.3), float(y0 + 0.2)
        slow = 5 if i % 3 == 0 else 0
        T.steps_for_slowing = slow
        T.optimal_criterion = sys.maxsize
        V, B = T.vector_of_velocities(v_now), T.vector_of_beta_angles(b_now)
        r = T.predictive_control(float(x0), float(y0), phi, xt, yt, V, B, False)
        traj = [_f(list(p)[:3]) for p in T.optimal_trajectory[0]]
        cases.append(dict(state=_f([x0, y0, phi]), target=[xt, yt], origin=[float(T.x_0), float(T.y_0)],
                          v_now=v_now, beta_now=b_now, vector_v=_f(V), vector_beta=_f(B),
                          slow=bool(slow), ret=_f(r), traj=traj,
                          steps_for_slowing_after=int(T.steps_for_slowing)))
    return dict(meta=_meta(), cost="tree", H=3, cases=cases)


def gen_pieces():
    M = R.load_full("math_model.py", vector_v=[0, 1], vector_beta=[0, 0.1])
    T = R.load_tree()
    rng = np.random.default_rng(3)
    steps = []
    for _ in range(40):
        st = _f(rng.uniform(-5, 5, 3))
        v = float(rng.uniform(0, 1))
        b = float(rng.uniform(-1.05, 1.05))
        M.t = float(rng.integers(0, 200)) * 0.05  # quad window moves with t (math_model.py:98-107)
        steps.append(dict(state=st, v=v, beta=b, t=float(M.t), out=_f(M.iteration_of_predict(st, v, b))))
    costs = []
    for _ in range(40):
        st = _f(rng.uniform(-5, 5, 3))
        costs.append(dict(state=st, target=[1, 5], origin=[0, 0],
                          mm=float(M.control_criterion(st)), tree=float(T.control_criterion(st))))
    costs.append(dict(state=[0, 0, 0.3], target=[1, 5], origin=[0, 0],
                      mm=float(M.control_criterion([0, 0, 0.3])), tree=float(T.control_criterion([0, 0, 0.3]))))
    grids = []
    for v_now in (0.0, 0.01, 0.02, 0.5, 0.975, 0.98, 0.999):
        grids.append(dict(kind="v", arg=v_now, out=_f(T.vector_of_velocities(v_now))))
    for b_now in (0.0, 0.5, -0.9, 1.0471975511965976, -1.0471975511965976, 1.03):
        grids.append(dict(kind="beta", arg=b_now, out=_f(T.vector_of_beta_angles(b_now))))
    Vd = np.round(np.arange(0, 1 + 0.005, 0.005), 3)
    Bd = np.round(np.arange(-math.radians(60), math.radians(60) + math.radians(1), math.radians(1)), 3)
    return dict(meta=_meta(), steps=steps, costs=costs, grids=grids,
                full_default_grid=dict(nv=int(Vd.size), nb=int(Bd.size), v_head=_f(Vd[:3]), v_tail=_f(Vd[-3:]),
                                       b_head=_f(Bd[:3]), b_tail=_f(Bd[-3:])))


def gen_operator_events():
    """new_target / turn_left / turn_right (+ slow_down) of math_model_tree.py:118-226 on poses in every heading
    quadrant the functions distinguish (incl. phi < pi/2, phi > 2 pi and the quadrant borders)."""
    T = R.load_tree()
    rng = np.random.default_rng(17)
    phis = list(rng.uniform(-1.0, 8.0, 44)) + [math.pi / 2, math.pi, 3 * math.pi / 2, 2 * math.pi]
    cases = []
    for i, phi in enumerate(phis):
        x, y = (float(q) for q in rng.uniform(-4, 4, 2))
        d = float(rng.uniform(0.5, 3.0))
        for kind in ("turn_left", "turn_right", "new_target"):
            T.steps_for_slowing = int(rng.integers(0, 30))
            before = int(T.steps_for_slowing)
            if kind == "new_target":
                a, b = (float(q) for q in rng.uniform(-4, 4, 2))
                T.new_target(x, y, float(phi), a, b, 0.4)
            else:
                a, b = d, 0.0
                getattr(T, kind)(x, y, float(phi), d, 0.4)
            cases.append(dict(kind=kind, pose=[x, y, float(phi)], a=a, b=b, slow_before=before,
                              out=_f([T.x_t, T.y_t, T.x_0, T.y_0]), steps_for_slowing=int(T.steps_for_slowing)))
    slow = [dict(deg=float(dg), before=7, after=None) for dg in (0, 5, 9.99, 10, 30, 45, 45.01, 90, 90.01, 180, -30, -95)]
    for c in slow:
        T.steps_for_slowing = c["before"]
        T.slow_down(math.radians(c["deg"]))
        c["after"] = int(T.steps_for_slowing)
    return dict(meta=_meta(), radius_u_turn=float(T.radius_u_turn), cases=cases, slow_down=slow)


def main():
    if not R.reference_available():
        raise SystemExit("needs /root/reference (build container only)")
    os.makedirs(OUT, exist_ok=True)
    only = sys.argv[1:]
    for name, fn in (("pieces", gen_pieces), ("held_single", gen_held_single), ("held_short_loops", gen_held_short_loops),
                     ("held_actual", gen_held_actual), ("operator_events", gen_operator_events),
                     ("full_h3", gen_full), ("held_closed_loop", gen_held_closed_loop)):
        if only and name not in only:
            continue
        data = fn()
        with open(os.path.join(OUT, name + ".json"), "w") as f:
            json.dump(data, f, indent=None, separators=(",", ":"))
        print("wrote", name, os.path.getsize(os.path.join(OUT, name + ".json")), "bytes")


if __name__ == "__main__":
    main()
