"""TEST INFRASTRUCTURE ONLY -- executes the UNMODIFIED reference function bodies.

The reference scripts (/root/reference/math_model.py, run_math_model.py,
math_model_tree.py) cannot be imported: they run the whole experiment at module
level, star-import names modern SciPy no longer exports, and call
``np.set_printoptions(threshold=np.nan)``.  This loader parses a script with
``ast``, keeps only its top-level ``def`` statements, and ``exec``s those --
byte for byte, no edits -- into a namespace that supplies what the dead
star-imports and the module-level statements used to supply.

It only works where /root/reference exists (the build container).  It is used
by ``oracle/make_golden.py`` to generate the fixtures under ``tests/golden/``
and by the container-only tests that pin ``oracle/closed_form.py`` to the
reference.  Nothing in the product (``diplomjourney_b200/``) may import it.

Parity status: the reference has no tests or golden vectors of its own
(SURVEY.md section 4), so the pin is "outputs of the reference itself run here".
"""
from __future__ import annotations

import ast
import math
import os
import sys
import time
import types

import numpy as np
import scipy.integrate as sp

REFERENCE_DIR = os.environ.get("MPCB_REFERENCE_DIR", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "math_model.py"))


class _NoOp:
    """Stands in for matplotlib.pyplot: every attribute is a callable no-op."""

    def __getattr__(self, _name):
        return _NoOp()

    def __call__(self, *a, **k):
        return _NoOp()


def _load_config() -> dict:
    src = open(os.path.join(REFERENCE_DIR, "config.py")).read()
    ns: dict = {}
    exec(compile(src, "config.py", "exec"), ns)
    return {k: v for k, v in ns.items() if not k.startswith("__") and k != "math"}


def _function_defs(script: str):
    path = os.path.join(REFERENCE_DIR, script)
    tree = ast.parse(open(path).read(), filename=path)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef)]
    return ast.Module(body=keep, type_ignores=[]), path


def _base_namespace(quiet: bool = True) -> dict:
    ns = {
        "np": np, "sp": sp, "math": math, "time": time, "sys": sys,
        # what ``from scipy import *`` exported on the author's SciPy
        "cos": np.cos, "sin": np.sin, "tan": np.tan, "arctan": np.arctan,
        "size": np.size, "random": np.random,
        "plt": _NoOp(), "Polygon": _NoOp(), "Circle": _NoOp(), "FuncAnimation": _NoOp(),
    }
    if quiet:
        ns["print"] = lambda *a, **k: None
    ns.update(_load_config())
    return ns


def load_full(script: str = "math_model.py", vector_v=None, vector_beta=None,
              overrides: dict | None = None) -> types.SimpleNamespace:
    """Namespace holding the FULL-tree reference functions (math_model.py:40-231
    or run_math_model.py:42-228) with the given control grids as the module
    globals ``vector_v`` / ``vector_beta``.  Default grids are the reference's
    own (math_model.py:23-30) -- note S^3 rows are allocated, keep S small."""
    ns = _base_namespace()
    if overrides:
        ns.update(overrides)
    v_max, delta_v = ns["v_max"], ns["delta_v"]
    beta_max, delta_beta = ns["beta_max"], ns["delta_beta"]
    if vector_v is None:
        vector_v = np.round(np.arange(0, v_max + delta_v, delta_v), 3)
    if vector_beta is None:
        vector_beta = np.round(np.arange(-beta_max, beta_max + delta_beta, delta_beta), 3)
    ns["vector_v"] = np.asarray(vector_v, dtype=float)
    ns["vector_beta"] = np.asarray(vector_beta, dtype=float)
    ns["prediction_horizon"] = 3
    ns["t"] = 0
    s1 = np.size(ns["vector_beta"]) * np.size(ns["vector_v"])
    ns["size_max_1"], ns["size_max_2"], ns["size_max_3"] = s1, pow(s1, 2), pow(s1, 3)
    mod, path = _function_defs(script)
    exec(compile(mod, path, "exec"), ns)
    # module-level state the scripts set up after the defs (math_model.py:131-133,
    # run_math_model.py:128-130)
    ns["optimal_trajectory"] = [[[0]]] if script == "math_model.py" else [0]
    ns["optimal_criterion"] = ns["control_criterion"]([ns["x_0"], ns["y_0"], ns["phi_0"]])
    return _View(ns)


def load_tree(overrides: dict | None = None, quiet: bool = True) -> "_View":
    """Namespace holding the HELD-tree reference functions
    (math_model_tree.py:48-635) plus the module-level state of :638-717."""
    sys.path.insert(0, REFERENCE_DIR)
    try:
        import importlib
        ct = importlib.import_module("CoordinateTree")
    finally:
        sys.path.pop(0)
    ns = _base_namespace(quiet)
    if overrides:
        ns.update(overrides)
    ns["CoordinateTree"] = ct.CoordinateTree
    ns["prediction_horizon"] = 3
    ns["radius_u_turn"] = ns["L"] / np.sin(ns["beta_max"])
    for k in ("beta", "v"):
        ns[k] = 0
    ns["phi"], ns["x"], ns["y"] = ns["phi_0"], ns["x_0"], ns["y_0"]
    mod, path = _function_defs("math_model_tree.py")
    exec(compile(mod, path, "exec"), ns)
    x_0, y_0, phi_0 = ns["x_0"], ns["y_0"], ns["phi_0"]
    ns.update(
        t=0, dt=ns["delta_t"], time_arr_for_plotting=[0], actual_time_arr_for_plotting=[0],
        optimal_trajectory=[[[0]]],
        result_trajectory_phi=[phi_0], actual_result_trajectory_phi=[phi_0],
        result_trajectory_x=[x_0], actual_result_trajectory_x=[x_0],
        result_x_velocity=[0], actual_result_x_velocity=[0],
        result_x_acceleration=[0], actual_result_x_acceleration=[0],
        result_trajectory_y=[y_0], actual_result_trajectory_y=[y_0],
        result_y_velocity=[0], actual_result_y_velocity=[0],
        result_y_acceleration=[0], actual_result_y_acceleration=[0],
        result_trajectory_v=[0], actual_result_trajectory_v=[0],
        result_trajectory_beta=[0], actual_result_trajectory_beta=[0],
        result_trajectory_angle_speed=[0], actual_result_trajectory_angle_speed=[0],
        result_v=0, result_beta=0, m=0, steps_for_slowing=0, need_scatter=False,
    )
    for pre in ("", "actual_"):
        for comp in ("x", "y", "phi"):
            for k in range(3):
                ns[f"{pre}predicted_trajectory_{comp}_anim{k}"] = []
    ns["optimal_criterion"] = ns["control_criterion"]([x_0, y_0, phi_0])
    return _View(ns)


class _View:
    """Attribute access to the exec namespace (functions read/write it as their
    module globals, so state such as ``optimal_criterion`` is visible here)."""

    def __init__(self, ns: dict):
        object.__setattr__(self, "ns", ns)

    def __getattr__(self, k):
        try:
            return self.ns[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self.ns[k] = v
