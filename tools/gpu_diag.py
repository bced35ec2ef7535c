"""Developer diagnostics on a GPU box: fp32-stage error vs model, refinement statistics, first timings."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diplomjourney_b200 import _native as nat
from oracle import closed_form as C, c_oracle as K

L, DT, VMIN = 0.5, 0.05, 0.4
s = nat.Solver(0)

def errs():
    V, B = [0.0, 0.3, 0.6, 1.0], np.linspace(-1, 1, 9)
    s.set_grid(V, B, L, DT, VMIN)
    sc = C.random_scenarios(20, 5)
    near = sc[:6].copy(); near[:, 3] = near[:, 0] + 0.07; near[:, 4] = near[:, 1] - 0.05
    for algo in (nat.ALGO_LEAFWALK, nat.ALGO_PREFIX):
        for cost, ck in ((C.COST_MM, nat.COST_MM), (C.COST_TREE, nat.COST_TREE)):
            for name, arr, off in (("far", sc, 0.0), ("near", near, 0.0), ("offline", sc, 0.5)):
                w = 0
                for x in arr:
                    og = x[:2] + off
                    xy, J = s.dump_leaves(nat.MODE_FULL, ck, 3, x[:3], x[3:5], og, algo=algo)
                    Jo = K.full_leaf_costs(x[:3], x[3:5], og, V, B, 3, cost)
                    ok = Jo < 1e7
                    w = max(w, np.abs(J - Jo)[ok].max())
                print(f"algo={algo} cost={cost} {name}: max |J32-J64| = {w:.3e}")

def timing():
    import torch
    V, B = C.vector_of_velocities(0.5), C.vector_of_beta_angles(0.0)
    s.set_grid(V, B, L, DT, VMIN)
    for N in (16, 256, 1024):
        sc = C.random_scenarios(N, 0)
        st = torch.tensor(sc[:, :3].copy(), device="cuda"); tg = torch.tensor(sc[:, 3:5].copy(), device="cuda"); og = torch.tensor(sc[:, :2].copy(), device="cuda")
        oc = torch.empty(N, dtype=torch.float64, device="cuda"); oi = torch.empty(N, dtype=torch.int64, device="cuda")
        ot = torch.empty(N, 3, 3, dtype=torch.float64, device="cuda"); ou = torch.empty(N, 2, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        for algo in (nat.ALGO_PREFIX, nat.ALGO_LEAFWALK):
            if algo == nat.ALGO_LEAFWALK and N > 16: continue
            s.set_option("algo", algo)
            for rep in range(3):
                t = time.perf_counter()
                s.solve_device(nat.MODE_FULL, nat.COST_MM, 3, N, st.data_ptr(), tg.data_ptr(), og.data_ptr(), 0, 0, oc.data_ptr(), oi.data_ptr(), ot.data_ptr(), ou.data_ptr())
                s.sync(); dt = time.perf_counter() - t
            leaves = N * len(V) ** 3 * len(B) ** 3
            print(f"FULL S=451 H=3 N={N} algo={algo}: {dt*1e3:.2f} ms  {leaves/dt:.3e} rollouts/s  stats={s.stats()}")
        s.set_option("algo", 0)
    # HELD batch on the 201x121 grid
    V, B = C.grid_full_default(); s.set_grid(V, B, L, DT, VMIN)
    for N in (1, 1024):
        sc = C.random_scenarios(N, 1)
        for rep in range(3):
            t = time.perf_counter(); r = s.solve(nat.MODE_HELD, nat.COST_TREE, 3, sc[:, :3], sc[:, 3:5], sc[:, :2]); dt = time.perf_counter() - t
        print(f"HELD S=24321 H=3 N={N} (host API): {dt*1e3:.3f} ms  {N*24321/dt:.3e} rollouts/s stats={s.stats()}")

def latency():
    import importlib
    from diplomjourney_b200 import config
    V, B = C.vector_of_velocities(0.5), C.vector_of_beta_angles(0.0)
    x = C.random_scenarios(1, 3)[0]
    for name, fn in (("set_grid S=451", lambda: s.set_grid(V, B, L, DT, VMIN)),
                     ("HELD solve N=1 S=451 (host API)", lambda: s.solve(nat.MODE_HELD, nat.COST_TREE, 3, x[:3], x[3:5], x[:2]))):
        fn(); fn()
        t = time.perf_counter()
        for _ in range(200): fn()
        print(f"{name}: {(time.perf_counter()-t)/200*1e6:.1f} us")
    params = nat.LoopParams.from_config(config, nat.COST_TREE, 3, 256)
    rng = np.random.default_rng(0)
    for N in (1, 1024, 16384):
        init = np.zeros((N, 5)); init[:, 2] = rng.uniform(-1, 1, N)
        ang = init[:, 2] + rng.uniform(-0.5, 0.5, N); d = rng.uniform(1.0, 4.0, N)
        tgt = np.stack([d*np.cos(ang), d*np.sin(ang)], 1)
        for rep in range(2):
            t = time.perf_counter(); r = s.held_closed_loop(params, init, tgt, [[0.0, 0.0]], first_threshold=1e10); dt = time.perf_counter() - t
        ticks = int(r["ticks"].sum())
        print(f"device closed loop N={N}: {dt*1e3:.2f} ms, {ticks} ticks total ({ticks/N:.1f}/robot), {ticks/dt:.3e} MPC solves/s, status counts {np.bincount(r['status'], minlength=4).tolist()}")

if len(sys.argv) > 1 and sys.argv[1] == "latency":
    latency()
else:
    errs()
    timing()
    latency()
