"""ctypes binding of libmpcb200.so (include/mpcb200.h).

This is the only way the package computes an MPC solve: there is NO CPU fallback.
If the shared library is missing it is built with nvcc (diplomjourney_b200.build);
if that is impossible, or no CUDA device is present, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np

from . import build as _build

MODE_FULL, MODE_HELD = 0, 1
COST_MM, COST_TREE = 0, 1
ALGO_AUTO, ALGO_LEAFWALK, ALGO_PREFIX = 0, 1, 2
FLAG_SLOW = 1
FLAG_SKIP = 2
MAX_H = 8

_ERR = {-1: "invalid argument", -2: "CUDA error", -3: "no grid set", -4: "empty control grid",
        -5: "tree too large", -6: "no CUDA device", -7: "NCCL error"}


class MpcbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libmpcb200: {_ERR.get(code, code)}: {msg}")
        self.code = code


class Stats(C.Structure):
    _fields_ = [("units", C.c_int64), ("leaves_per_solve", C.c_int64), ("segments", C.c_int64),
                ("refine_segments", C.c_int64), ("refine_candidates", C.c_int64), ("pruned_units", C.c_int64),
                ("algo", C.c_int32), ("kernel_launches", C.c_int32)]


class LoopParams(C.Structure):
    """mpcb_loop_params (include/mpcb200.h)."""
    _fields_ = [("L", C.c_double), ("delta_t", C.c_double), ("delta_v", C.c_double), ("delta_beta", C.c_double),
                ("v_max", C.c_double), ("v_min", C.c_double), ("beta_limit", C.c_double), ("half_v", C.c_double),
                ("half_beta", C.c_double), ("eps", C.c_double), ("n_v", C.c_int32), ("n_beta", C.c_int32),
                ("cost_kind", C.c_int32), ("H", C.c_int32), ("max_ticks", C.c_int32), ("reserved", C.c_int32)]

    @classmethod
    def from_config(cls, cfg, cost_kind=COST_TREE, H=3, max_ticks=512):
        """Window constants computed exactly as math_model_tree.py:239-256 computes them."""
        half_v = (cfg.v_acc_max * cfg.delta_t) / cfg.delta_v
        half_b = (math.degrees(cfg.beta_acc_max) * cfg.delta_t) / math.degrees(cfg.delta_beta)
        return cls(L=cfg.L, delta_t=cfg.delta_t, delta_v=cfg.delta_v, delta_beta=cfg.delta_beta, v_max=cfg.v_max,
                   v_min=cfg.v_min, beta_limit=cfg.beta_max + math.radians(cfg.eps_beta), half_v=half_v,
                   half_beta=half_b, eps=cfg.eps, n_v=1 + 2 * int(half_v), n_beta=1 + 2 * int(half_b),
                   cost_kind=cost_kind, H=H, max_ticks=max_ticks, reserved=0)


LOOP_ON_TARGET, LOOP_STALLED, LOOP_MAX_TICKS, LOOP_NO_LEAF = 0, 1, 2, 3


class LoopEvent(C.Structure):
    """mpcb_loop_event (include/mpcb200.h): one operator event of a closed-loop script."""
    _fields_ = [("tick", C.c_int32), ("kind", C.c_int32), ("a", C.c_double), ("b", C.c_double)]


EVENT_NEW_TARGET, EVENT_TURN_LEFT, EVENT_TURN_RIGHT = 1, 2, 3

_lib = None
_dp, _i64p, _u8p, _fp = C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_uint8), C.POINTER(C.c_float)


def library_path() -> str:
    return _build.LIB


def load():
    """Load (building first if needed) libmpcb200.so. Raises if it cannot be had."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.build_library()       # returns at once unless a source is newer than the library (or MPCB_LIB is set)
    lib = C.CDLL(path)
    vp = C.c_void_p
    lib.mpcb_version.restype = C.c_int
    lib.mpcb_device_count.restype = C.c_int
    lib.mpcb_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.mpcb_destroy.argtypes = [vp]
    lib.mpcb_last_error.argtypes = [vp]
    lib.mpcb_last_error.restype = C.c_char_p
    lib.mpcb_stream.argtypes = [vp]
    lib.mpcb_stream.restype = vp
    lib.mpcb_sync.argtypes = [vp]
    lib.mpcb_set_grid.argtypes = [vp, _dp, C.c_int, _dp, C.c_int, C.c_double, C.c_double, C.c_double]
    lib.mpcb_set_option.argtypes = [vp, C.c_char_p, C.c_double]
    solve = [vp, C.c_int, C.c_int, C.c_int, C.c_int64, vp, vp, vp, vp, vp, C.c_int64, C.c_int64, vp, vp, vp, vp]
    lib.mpcb_solve_batch_host.argtypes = solve
    lib.mpcb_solve_batch_device.argtypes = solve
    lib.mpcb_dump_leaves_host.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, C.c_uint8,
                                          C.c_int64, C.c_int64, _fp, _dp]
    lib.mpcb_get_stats.argtypes = [vp, C.POINTER(Stats)]
    loop = [vp, C.POINTER(LoopParams), C.c_int64, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.mpcb_held_closed_loop_host.argtypes = loop
    lib.mpcb_held_closed_loop_device.argtypes = loop
    lib.mpcb_held_closed_loop_events_host.argtypes = loop[:8] + [vp, C.c_int32, C.c_double] + loop[8:] + [vp]
    lib.mpcb_held_tick_host.argtypes = [vp, vp, C.c_int, vp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                                        vp, vp, vp, C.c_double, C.c_int, vp, vp, vp, vp]
    win = [vp, C.POINTER(LoopParams), C.c_int64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.mpcb_solve_held_windows_host.argtypes = win
    lib.mpcb_solve_held_windows_device.argtypes = win
    lib.mpcb_full_closed_loop_host.argtypes = [vp, C.c_int, C.c_int, C.c_int64, vp, vp, vp, vp, C.c_double, C.c_int, vp, vp, vp]
    split = [vp, vp, C.c_int, C.c_int, C.c_int64, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.mpcb_solve_tree_split_host.argtypes = split
    lib.mpcb_solve_tree_split_device.argtypes = split
    lib.mpcb_allreduce_min.argtypes = [vp, vp, vp, vp]
    lib.mpcb_nccl_unique_id.argtypes = [vp]
    lib.mpcb_nccl_comm_create.argtypes = [vp, C.c_int, C.c_int, vp, C.POINTER(vp)]
    lib.mpcb_nccl_comm_destroy.argtypes = [vp]
    _lib = lib
    return lib


def _arr(a, dtype, shape=None):
    a = np.ascontiguousarray(a, dtype=dtype)
    if shape is not None:
        a = a.reshape(shape)
    return a


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Solver:
    """One device, one stream, one control grid at a time (mpcb_handle)."""

    def __init__(self, device: int = 0):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.mpcb_create(int(device), C.byref(h))
        if rc != 0:
            raise MpcbError(rc, "mpcb_create failed (a CUDA device is required; there is no CPU fallback)")
        self.h = h
        self.device = int(device)
        self.S = 0
        self.grid = None
        self._owner = None
        self._tick = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.mpcb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise MpcbError(rc, (self.lib.mpcb_last_error(self.h) or b"").decode())

    # ---- configuration
    def set_grid(self, vector_v, vector_beta, L, delta_t, v_min=0.0):
        v = _arr(vector_v, np.float64).ravel()
        b = _arr(vector_beta, np.float64).ravel()
        self._ck(self.lib.mpcb_set_grid(self.h, v.ctypes.data_as(_dp), v.size, b.ctypes.data_as(_dp), b.size,
                                        float(L), float(delta_t), float(v_min)))
        self.S = v.size * b.size
        self.grid = (v, b, float(L), float(delta_t), float(v_min))
        self._owner = None      # whoever cached "my grid is set on this solver" must set it again

    def set_option(self, name: str, value: float):
        self._ck(self.lib.mpcb_set_option(self.h, name.encode(), float(value)))

    @property
    def stream(self) -> int:
        return int(self.lib.mpcb_stream(self.h) or 0)

    def sync(self):
        self._ck(self.lib.mpcb_sync(self.h))

    def stats(self) -> dict:
        s = Stats()
        self._ck(self.lib.mpcb_get_stats(self.h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in Stats._fields_}

    # ---- solves
    def solve(self, mode, cost, H, state, target, origin, threshold=None, flags=None, i0_range=None):
        """Host-buffer batch solve (mpcb_solve_batch_host). Arrays: state[N,3], target[N,2], origin[N,2]."""
        st = _arr(state, np.float64)
        st = st.reshape(-1, st.shape[-1])[:, :3].copy() if st.ndim > 1 else st[:3].reshape(1, 3).copy()
        N = st.shape[0]
        tg = _arr(target, np.float64, (-1, 2))
        og = _arr(origin, np.float64, (-1, 2))
        if tg.shape[0] == 1 and N > 1:
            tg = np.repeat(tg, N, 0)
        if og.shape[0] == 1 and N > 1:
            og = np.repeat(og, N, 0)
        if tg.shape[0] != N or og.shape[0] != N:
            raise ValueError("state/target/origin batch sizes differ")
        thr = None if threshold is None else _arr(np.broadcast_to(np.asarray(threshold, np.float64), (N,)), np.float64)
        fl = None if flags is None else _arr(np.broadcast_to(np.asarray(flags, np.uint8), (N,)), np.uint8)
        lo, hi = (0, -1) if i0_range is None else (int(i0_range[0]), int(i0_range[1]))
        cost_o = np.empty(N, np.float64)
        idx_o = np.empty(N, np.int64)
        traj_o = np.empty((N, H, 3), np.float64)
        ctl_o = np.empty((N, 2), np.float64)
        self._ck(self.lib.mpcb_solve_batch_host(self.h, mode, cost, H, N, _ptr(st), _ptr(tg), _ptr(og), _ptr(thr),
                                                _ptr(fl), lo, hi, _ptr(cost_o), _ptr(idx_o), _ptr(traj_o), _ptr(ctl_o)))
        return dict(cost=cost_o, index=idx_o, traj=traj_o, first_control=ctl_o)

    def solve_device(self, mode, cost, H, N, state, target, origin, threshold, flags, out_cost, out_index,
                     out_traj, out_ctl, i0_range=None):
        """Device-pointer batch solve (mpcb_solve_batch_device); arguments are raw device addresses
        (e.g. torch.Tensor.data_ptr()) or 0/None. Enqueues on self.stream, does not synchronise."""
        lo, hi = (0, -1) if i0_range is None else (int(i0_range[0]), int(i0_range[1]))
        vp = lambda x: C.c_void_p(int(x)) if x else None
        self._ck(self.lib.mpcb_solve_batch_device(self.h, mode, cost, H, int(N), vp(state), vp(target), vp(origin),
                                                  vp(threshold), vp(flags), lo, hi, vp(out_cost), vp(out_index),
                                                  vp(out_traj), vp(out_ctl)))

    def solve_tree_split(self, comm: "NcclComm", cost, H, state, target, origin, threshold=None):
        """ONE FULL tree per solve shared by the ranks of ``comm`` (mpcb_solve_tree_split_host): collective call,
        every rank passes the same arguments and gets the whole tree's result (global leaf indices)."""
        st = _arr(state, np.float64)
        st = st.reshape(-1, st.shape[-1])[:, :3].copy() if st.ndim > 1 else st[:3].reshape(1, 3).copy()
        N = st.shape[0]
        tg = _arr(np.broadcast_to(_arr(target, np.float64, (-1, 2)), (N, 2)), np.float64)
        og = _arr(np.broadcast_to(_arr(origin, np.float64, (-1, 2)), (N, 2)), np.float64)
        thr = None if threshold is None else _arr(np.broadcast_to(np.asarray(threshold, np.float64), (N,)), np.float64)
        cost_o, idx_o = np.empty(N, np.float64), np.empty(N, np.int64)
        traj_o, ctl_o = np.empty((N, H, 3), np.float64), np.empty((N, 2), np.float64)
        self._ck(self.lib.mpcb_solve_tree_split_host(self.h, comm.comm, cost, H, N, _ptr(st), _ptr(tg), _ptr(og),
                                                     _ptr(thr), _ptr(cost_o), _ptr(idx_o), _ptr(traj_o), _ptr(ctl_o)))
        return dict(cost=cost_o, index=idx_o, traj=traj_o, first_control=ctl_o)

    def solve_tree_split_device(self, comm: "NcclComm", cost, H, N, state, target, origin, threshold, out_cost,
                                out_index, out_traj, out_ctl):
        """Device-pointer flavour (mpcb_solve_tree_split_device): enqueues the solve of this rank's share, the
        16-byte all-gather and the finalisation on self.stream; does not synchronise."""
        vp = lambda x: C.c_void_p(int(x)) if x else None
        self._ck(self.lib.mpcb_solve_tree_split_device(self.h, comm.comm, cost, H, int(N), vp(state), vp(target),
                                                       vp(origin), vp(threshold), vp(out_cost), vp(out_index),
                                                       vp(out_traj), vp(out_ctl)))

    def held_tick(self, vector_v, vector_beta, L, delta_t, v_min, cost, H, state, target, origin, threshold=math.inf,
                  flags=0):
        """One online tick in one call (mpcb_held_tick_host): the control window lists of this tick and ONE HELD
        solve.  The lean path of math_model_tree.predictive_control: preallocated buffers, no per-call numpy work
        beyond turning the two window lists into arrays.  Returns (cost, index, traj[H,3], (v, beta))."""
        buf = self._tick
        if buf is None or buf["H"] != H:
            buf = self._tick = dict(H=H, inp=np.empty(7), cost=np.empty(1), index=np.empty(1, np.int64),
                                    traj=np.empty((H, 3)), ctl=np.empty(2))
            buf["ptr"] = tuple(C.c_void_p(buf[k].ctypes.data) for k in ("cost", "index", "traj", "ctl"))
            base = buf["inp"].ctypes.data
            buf["in_ptr"] = (C.c_void_p(base), C.c_void_p(base + 24), C.c_void_p(base + 40))
        v = np.asarray(vector_v, dtype=np.float64)
        b = np.asarray(vector_beta, dtype=np.float64)
        inp = buf["inp"]
        inp[0], inp[1], inp[2] = state[0], state[1], state[2]
        inp[3], inp[4] = target[0], target[1]
        inp[5], inp[6] = origin[0], origin[1]
        self._ck(self.lib.mpcb_held_tick_host(self.h, C.c_void_p(v.ctypes.data), v.size, C.c_void_p(b.ctypes.data), b.size,
                                              L, delta_t, v_min, cost, H, *buf["in_ptr"], float(threshold), int(flags),
                                              *buf["ptr"]))
        self.S = v.size * b.size
        self.grid = (v, b, float(L), float(delta_t), float(v_min))
        self._owner = None
        return float(buf["cost"][0]), int(buf["index"][0]), buf["traj"], (float(buf["ctl"][0]), float(buf["ctl"][1]))

    def solve_held_windows(self, params: "LoopParams", state, v_beta, target, origin, threshold=None, flags=None):
        """One online tick of N robots, each with its OWN acceleration window built on the device around its current
        (v, beta) (mpcb_solve_held_windows_host).  state[N,3], v_beta[N,2], target[N,2], origin[N,2].
        Returns dict(cost, index, traj, first_control, shape[N,2] = (nV, nB))."""
        st = _arr(state, np.float64)
        st = st.reshape(-1, st.shape[-1])[:, :3].copy()
        N = st.shape[0]
        vb = _arr(np.broadcast_to(_arr(v_beta, np.float64, (-1, 2)), (N, 2)), np.float64)
        tg = _arr(np.broadcast_to(_arr(target, np.float64, (-1, 2)), (N, 2)), np.float64)
        og = _arr(np.broadcast_to(_arr(origin, np.float64, (-1, 2)), (N, 2)), np.float64)
        thr = None if threshold is None else _arr(np.broadcast_to(np.asarray(threshold, np.float64), (N,)), np.float64)
        fl = None if flags is None else _arr(np.broadcast_to(np.asarray(flags, np.uint8), (N,)), np.uint8)
        H = params.H
        cost_o, idx_o = np.empty(N, np.float64), np.empty(N, np.int64)
        traj_o, ctl_o = np.empty((N, H, 3), np.float64), np.empty((N, 2), np.float64)
        shape_o = np.empty((N, 2), np.int32)
        self._ck(self.lib.mpcb_solve_held_windows_host(self.h, C.byref(params), N, _ptr(st), _ptr(vb), _ptr(tg), _ptr(og),
                                                       _ptr(thr), _ptr(fl), _ptr(cost_o), _ptr(idx_o), _ptr(traj_o),
                                                       _ptr(ctl_o), _ptr(shape_o)))
        return dict(cost=cost_o, index=idx_o, traj=traj_o, first_control=ctl_o, shape=shape_o)

    def held_closed_loop(self, params: "LoopParams", init, target, origin, first_threshold=None, slow_steps=None,
                         events=None, radius_u_turn=0.0):
        """Whole closed loops of the online controller for a batch of robots on the device
        (mpcb_held_closed_loop_host / _events_host). ``events``: an operator script [(tick, kind, a, b), ...] with kind
        EVENT_NEW_TARGET (a, b = the new target) or EVENT_TURN_LEFT / _RIGHT (a = distance), applied to every robot
        after the tick it names; ``radius_u_turn`` = L / sin(beta_max) for the turns.
        Returns dict(log[N,max_ticks,5], ticks[N], status[N], final[N,6] = x_t, y_t, x_0, y_0, steps_for_slowing, m)."""
        ini = _arr(init, np.float64, (-1, 5))
        N = ini.shape[0]
        tg = _arr(np.broadcast_to(np.asarray(target, np.float64).reshape(-1, 2), (N, 2)), np.float64)
        og = _arr(np.broadcast_to(np.asarray(origin, np.float64).reshape(-1, 2), (N, 2)), np.float64)
        thr = None if first_threshold is None else _arr(np.broadcast_to(np.asarray(first_threshold, np.float64), (N,)), np.float64)
        sl = None if slow_steps is None else _arr(np.broadcast_to(np.asarray(slow_steps, np.int32), (N,)), np.int32)
        log = np.full((N, params.max_ticks, 5), np.nan)
        ticks = np.empty(N, np.int32)
        status = np.empty(N, np.int32)
        ev = (LoopEvent * max(1, len(events or ())))()
        for i, (tick, kind, a, b) in enumerate(events or ()):
            ev[i] = LoopEvent(int(tick), int(kind), float(a), float(b))
        final = np.empty((N, 6))
        self._ck(self.lib.mpcb_held_closed_loop_events_host(self.h, C.byref(params), N, _ptr(ini), _ptr(tg), _ptr(og),
                                                            _ptr(thr), _ptr(sl), ev if events else None,
                                                            len(events or ()), float(radius_u_turn), _ptr(log),
                                                            _ptr(ticks), _ptr(status), _ptr(final)))
        return dict(log=log, ticks=ticks, status=status, final=final)

    def full_closed_loop(self, cost, H, init_state, target, origin, first_threshold=None, eps=0.001, max_ticks=256):
        """Closed loop of the FULL-tree scripts for a batch of robots (mpcb_full_closed_loop_host): one batched
        FULL solve per tick with the carried threshold; stops per robot on target / after two repeated
        positions / at max_ticks. Returns dict(log[N,max_ticks,5], ticks[N], status[N])."""
        ini = _arr(init_state, np.float64, (-1, 3))
        N = ini.shape[0]
        tg = _arr(np.broadcast_to(np.asarray(target, np.float64).reshape(-1, 2), (N, 2)), np.float64)
        og = _arr(np.broadcast_to(np.asarray(origin, np.float64).reshape(-1, 2), (N, 2)), np.float64)
        thr = None if first_threshold is None else _arr(np.broadcast_to(np.asarray(first_threshold, np.float64), (N,)), np.float64)
        log = np.empty((N, max_ticks, 5))
        ticks = np.empty(N, np.int32)
        status = np.empty(N, np.int32)
        self._ck(self.lib.mpcb_full_closed_loop_host(self.h, cost, H, N, _ptr(ini), _ptr(tg), _ptr(og), _ptr(thr),
                                                     float(eps), int(max_ticks), _ptr(log), _ptr(ticks), _ptr(status)))
        return dict(log=log, ticks=ticks, status=status)

    def dump_leaves(self, mode, cost, H, state, target, origin, flags=0, algo=ALGO_LEAFWALK, leaf_begin=0, count=None):
        st = _arr(state, np.float64).ravel()[:3].copy()
        tg = _arr(target, np.float64).ravel()[:2].copy()
        og = _arr(origin, np.float64).ravel()[:2].copy()
        if count is None:
            count = (self.S if mode == MODE_HELD else self.S ** H) - leaf_begin
        xy = np.empty((count, 2), np.float32)
        cs = np.empty(count, np.float64)
        self._ck(self.lib.mpcb_dump_leaves_host(self.h, mode, cost, H, algo, st.ctypes.data_as(_dp),
                                                tg.ctypes.data_as(_dp), og.ctypes.data_as(_dp), int(flags),
                                                int(leaf_begin), int(count), xy.ctypes.data_as(_fp),
                                                cs.ctypes.data_as(_dp)))
        return xy, cs


def nccl_unique_id() -> bytes:
    """128-byte NCCL id (rank 0 creates it, the caller distributes it)."""
    buf = C.create_string_buffer(128)
    rc = load().mpcb_nccl_unique_id(buf)
    if rc != 0:
        raise MpcbError(rc, "ncclGetUniqueId failed (is libnccl.so.2 loadable?)")
    return buf.raw


class NcclComm:
    """Native NCCL communicator bound to a Solver's device (mpcb_nccl_comm_create)."""

    def __init__(self, solver: "Solver", nranks: int, rank: int, unique_id: bytes):
        self.solver = solver
        self.comm = C.c_void_p()
        rc = solver.lib.mpcb_nccl_comm_create(solver.h, nranks, rank, C.c_char_p(unique_id), C.byref(self.comm))
        if rc != 0:
            raise MpcbError(rc, "ncclCommInitRank failed")

    def allreduce_min(self, cost_ptr: int, index_ptr: int):
        """Lexicographic (cost, index) min over ranks, in place on 1-element device buffers."""
        rc = self.solver.lib.mpcb_allreduce_min(self.solver.h, self.comm, C.c_void_p(cost_ptr), C.c_void_p(index_ptr))
        if rc != 0:
            raise MpcbError(rc, "mpcb_allreduce_min failed")

    def close(self):
        if self.comm:
            self.solver.lib.mpcb_nccl_comm_destroy(self.comm)
            self.comm = None


_default = {}


def default_solver(device: int = 0) -> Solver:
    """Process-wide solver per device used by the reference-API modules."""
    s = _default.get(device)
    if s is None:
        s = _default[device] = Solver(device)
    return s
