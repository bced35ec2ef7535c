"""Test-only stand-in for diplomjourney_b200._native.Solver that answers from the float64
oracle.  It lets the CPU suite exercise the HOST logic of the reference-API modules (closed loop,
operator events, finishing heuristic, logs) without a GPU.  Never used by the product."""
import numpy as np

from oracle import closed_form as C


class OracleBackend:
    def __init__(self):
        self.grid = None
        self.calls = 0

    def set_grid(self, vector_v, vector_beta, L, delta_t, v_min=0.0):
        self.grid = (list(vector_v), list(vector_beta), L, delta_t, v_min)
        self.S = len(self.grid[0]) * len(self.grid[1])

    def solve(self, mode, cost, H, state, target, origin, threshold=None, flags=None, i0_range=None):
        V, B, L, dt, v_min = self.grid
        self.calls += 1
        st = np.asarray(state, float)
        if st.ndim == 2:      # batch: loop
            tg = np.asarray(target, float).reshape(-1, 2); og = np.asarray(origin, float).reshape(-1, 2)
            n = st.shape[0]
            rs = [self.solve(mode, cost, H, st[i], tg[i if tg.shape[0] == n else 0], og[i if og.shape[0] == n else 0],
                             None if threshold is None else np.broadcast_to(threshold, (n,))[i],
                             None if flags is None else np.broadcast_to(flags, (n,))[i], i0_range) for i in range(n)]
            return {k: np.concatenate([r[k] for r in rs], axis=0) for k in rs[0]}
        thr = np.inf if threshold is None else float(threshold)
        ck = C.COST_MM if cost == 0 else C.COST_TREE
        if mode == 1:
            r = C.solve_held(state, target, origin, V, B, H, ck, threshold=thr, slow=bool(flags), v_min=v_min,
                             L=L, delta_t=dt)
        else:
            r = C.solve_full(state, target, origin, V, B, H, ck, threshold=thr, L=L, delta_t=dt, i0_range=i0_range)
        return dict(cost=np.array([r["cost"]]), index=np.array([r["index"]]), traj=r["traj"][None],
                    first_control=np.array([r["first_control"]]))

    def solve_held_windows(self, params, state, v_beta, target, origin, threshold=None, flags=None):
        """One online tick per robot with the robot's own acceleration window (what mpcb_solve_held_windows does on
        the device), answered by the oracle: windows from closed_form.vector_of_*, HELD solve per robot."""
        st = np.asarray(state, float).reshape(-1, 3)
        n = st.shape[0]
        vb = np.broadcast_to(np.asarray(v_beta, float).reshape(-1, 2), (n, 2))
        tg = np.broadcast_to(np.asarray(target, float).reshape(-1, 2), (n, 2))
        og = np.broadcast_to(np.asarray(origin, float).reshape(-1, 2), (n, 2))
        thr = np.full(n, np.inf) if threshold is None else np.broadcast_to(np.asarray(threshold, float), (n,))
        fl = np.zeros(n, np.uint8) if flags is None else np.broadcast_to(np.asarray(flags, np.uint8), (n,))
        H = params.H
        out = dict(cost=np.full(n, np.nan), index=np.full(n, -1, np.int64), traj=np.full((n, H, 3), np.nan),
                   first_control=np.full((n, 2), np.nan), shape=np.zeros((n, 2), np.int32))
        ck = C.COST_MM if params.cost_kind == 0 else C.COST_TREE
        self.calls += 1
        for i in range(n):
            V, B = C.vector_of_velocities(vb[i, 0]), C.vector_of_beta_angles(vb[i, 1])
            out["shape"][i] = (len(V), len(B))
            if not V or not B or (fl[i] & 2):
                continue
            r = C.solve_held(st[i], tg[i], og[i], V, B, H, ck, threshold=float(thr[i]), slow=bool(fl[i] & 1),
                             v_min=params.v_min, L=params.L, delta_t=params.delta_t)
            out["cost"][i], out["index"][i] = r["cost"], r["index"]
            out["traj"][i], out["first_control"][i] = r["traj"], r["first_control"]
        return out
