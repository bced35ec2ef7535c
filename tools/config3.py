"""BASELINE configs[2] on one GPU: single robot (config.py scenario), horizon 5, 32x32 (speed x steer) grid, S=1024.
  * HELD (the tree the CoordinateTree rules actually build: S leaves, 5 S node expansions): latency per solve;
  * FULL (S^5 = 1.126e15 leaves): a stated range of first controls i0 (each i0 = one S^4-leaf subtree), exhaustive and
    with the exact branch-and-bound, and the whole-tree time extrapolated from it (the split path of
    diplomjourney_b200/distributed.py runs exactly these per-rank ranges);
  * parity: HELD against the float64 oracle; FULL pruned == FULL exhaustive on the same range (index, cost, trajectory).
usage: python tools/config3.py [n_i0=8] [whole_pruned: 0|1 = 0]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from diplomjourney_b200 import _native as nat, config as cfg
from oracle import closed_form as C

n_i0 = int(sys.argv[1]) if len(sys.argv) > 1 else 8
whole_pruned = len(sys.argv) > 2 and sys.argv[2] == "1"
H = 5
V = np.linspace(0.0, cfg.v_max, 32)
B = np.linspace(-cfg.beta_max, cfg.beta_max, 32)
state, target, origin = (cfg.x_0, cfg.y_0, cfg.phi_0), (cfg.x_t, cfg.y_t), (cfg.x_0, cfg.y_0)
s = nat.default_solver()
s.set_grid(V, B, cfg.L, cfg.delta_t, cfg.v_min)
S = s.S

# ---- HELD
r = s.solve(nat.MODE_HELD, nat.COST_MM, H, state, target, origin)          # warm-up: lazy module load
o = C.solve_held(np.array(state, float), np.array(target, float), np.array(origin, float), V, B, H, C.COST_MM)
held_ok = int(r["index"][0]) == int(o["index"]) and abs(r["cost"][0] - o["cost"]) <= 1e-12 * abs(o["cost"])
reps = 2000
t = time.perf_counter()
for _ in range(reps):
    s.solve(nat.MODE_HELD, nat.COST_MM, H, state, target, origin)
dt = (time.perf_counter() - t) / reps
print(f"config3 HELD H={H} S={S}: {dt*1e6:.1f} us/solve through the host API ({1/dt:.0f} solves/s, {S/dt:.3e} rollouts/s, "
      f"{H*S/dt:.3e} node expansions/s) leaf={int(r['index'][0])} cost={r['cost'][0]:.6f} oracle_parity={'OK' if held_ok else 'FAIL'}",
      flush=True)

# ---- FULL on a range of first controls around the winner of the HELD solve (where pruning is least effective)
mid = int(r["index"][0])
lo = max(0, min(S - n_i0, mid - n_i0 // 2))
rng = (lo, lo + n_i0)
leaves = n_i0 * S ** (H - 1)
res = {}
for prune in (1, 0):
    s.set_option("prune", prune)
    t = time.perf_counter()
    res[prune] = s.solve(nat.MODE_FULL, nat.COST_MM, H, state, target, origin, i0_range=rng)
    dt = time.perf_counter() - t
    st = s.stats()
    print(f"config3 FULL H={H} S={S} i0 in [{rng[0]},{rng[1]}) prune={prune}: leaves={leaves:.4e} time={dt:.2f}s "
          f"rate={leaves/dt:.3e} rollouts/s -> whole tree (x{S/n_i0:.0f}) = {dt*S/n_i0:.0f} s on 1 GPU, {dt*S/n_i0/8:.0f} s on 8 "
          f"leaf={int(res[prune]['index'][0])} cost={res[prune]['cost'][0]:.6f} pruned_nodes={st['pruned_units']}/{st['units']} "
          f"refine(seg={st['refine_segments']},cand={st['refine_candidates']})", flush=True)
same = (int(res[0]["index"][0]) == int(res[1]["index"][0]) and res[0]["cost"][0] == res[1]["cost"][0]
        and np.array_equal(res[0]["traj"], res[1]["traj"]))
print(f"config3 FULL pruned == exhaustive on the range: {'OK' if same else 'FAIL'}", flush=True)

if whole_pruned:
    s.set_option("prune", 1)
    t = time.perf_counter()
    w = s.solve(nat.MODE_FULL, nat.COST_MM, H, state, target, origin)
    dt = time.perf_counter() - t
    st = s.stats()
    print(f"config3 FULL WHOLE tree prune=1: leaves={S**H:.4e} time={dt:.2f}s effective rate={S**H/dt:.3e} rollouts/s "
          f"leaf={int(w['index'][0])} cost={w['cost'][0]:.6f} first_control={w['first_control'][0].tolist()} "
          f"pruned_nodes={st['pruned_units']}/{st['units']}", flush=True)
sys.exit(0 if (held_ok and same) else 1)
