"""TEST INFRASTRUCTURE ONLY -- ctypes loader for oracle/mpc_oracle.c (the C float64
restatement).  Same semantics as oracle/closed_form.py, faster on big trees."""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

from . import closed_form as cf

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmpc_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "mpc_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        dp, i64p = C.POINTER(C.c_double), C.POINTER(C.c_int64)
        common = [dp, dp, C.c_int, C.c_int, C.c_int, C.c_double, dp, dp, dp, C.c_double]
        _lib.mpco_solve_full.argtypes = common + [C.c_int64, C.c_int64, dp, i64p]
        _lib.mpco_solve_held.argtypes = common + [dp, i64p, dp]
        _lib.mpco_full_leaf_costs.argtypes = common + [dp]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _args(state, target, origin, vector_v, vector_beta, H, cost, L, delta_t, slow=False, v_min=None):
    vv, bb, dphi = cf.control_tables(vector_v, vector_beta, L, delta_t, slow, v_min)
    vv = np.ascontiguousarray(vv)
    dphi = np.ascontiguousarray(dphi)
    st = np.asarray(state, dtype=np.float64)[:3].copy()
    tg = np.asarray(target, dtype=np.float64).copy()
    og = np.asarray(origin, dtype=np.float64).copy()
    theta = cf.heading_reference(tg[0], tg[1])
    keep = (vv, dphi, st, tg, og)
    return keep, [_p(vv), _p(dphi), int(vv.size), int(H), 0 if cost == cf.COST_MM else 1, float(delta_t),
                  _p(st), _p(tg), _p(og), theta], bb


def solve_full(state, target, origin, vector_v, vector_beta, H=3, cost=cf.COST_MM, threshold=math.inf,
               L=cf.CONFIG["L"], delta_t=cf.CONFIG["delta_t"], i0_range=None):
    keep, a, bb = _args(state, target, origin, vector_v, vector_beta, H, cost, L, delta_t)
    S = a[2]
    lo, hi = (0, S) if i0_range is None else i0_range
    c, i = C.c_double(), C.c_int64()
    lib().mpco_solve_full(*a, lo, hi, C.byref(c), C.byref(i))
    return cf._finish(state, target, origin, keep[0], bb, keep[1], delta_t, S, H, c.value, i.value,
                      threshold, cost, held=False)


def solve_held(state, target, origin, vector_v, vector_beta, H=3, cost=cf.COST_TREE, threshold=math.inf,
               slow=False, v_min=cf.CONFIG["v_min"], L=cf.CONFIG["L"], delta_t=cf.CONFIG["delta_t"]):
    keep, a, bb = _args(state, target, origin, vector_v, vector_beta, H, cost, L, delta_t, slow, v_min)
    S = a[2]
    c, i = C.c_double(), C.c_int64()
    allc = np.empty(S)
    lib().mpco_solve_held(*a, C.byref(c), C.byref(i), _p(allc))
    res = cf._finish(state, target, origin, keep[0], bb, keep[1], delta_t, S, H, c.value, i.value,
                     threshold, cost, held=True)
    res["leaf_costs"] = allc
    return res


def full_leaf_costs(state, target, origin, vector_v, vector_beta, H=3, cost=cf.COST_MM,
                    L=cf.CONFIG["L"], delta_t=cf.CONFIG["delta_t"]):
    keep, a, _ = _args(state, target, origin, vector_v, vector_beta, H, cost, L, delta_t)
    out = np.empty(a[2] ** H)
    lib().mpco_full_leaf_costs(*a, _p(out))
    return out
