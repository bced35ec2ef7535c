"""BASELINE configs[3] at FULL size on one GPU: 65,536 scenarios (default_rng(1), distribution of
run_math_model.py:235-239), horizon 4, 16x16 linspace grid (S=256), FULL tree = 4.295e9 leaves per scenario,
2.815e14 in total.  Solved in batches through the host C ABI with the exact branch-and-bound (option prune=1,
bit-identical to evaluating every leaf: tests/test_gpu_parity.py::test_branch_and_bound_is_exact), then
  * `n_check` scenarios are re-solved with prune=0 (every leaf evaluated) and compared bit for bit,
  * `n_oracle` of those are solved by the float64 C oracle (oracle/mpc_oracle.c, all host cores) and compared.
usage: python tools/config4_full.py [n_scenarios=65536] [batch=4096] [n_check=32] [n_oracle=2]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from diplomjourney_b200 import _native as nat
from oracle import closed_form as C, c_oracle as K

arg = lambda i, d: int(sys.argv[i]) if len(sys.argv) > i else d
n_total, batch, n_check, n_oracle = arg(1, 65536), arg(2, 4096), arg(3, 32), arg(4, 2)
H = 4
V = np.linspace(0.0, 1.0, 16)
B = np.linspace(-np.radians(60), np.radians(60), 16)
sc = C.random_scenarios(n_total, 1)
s = nat.Solver(0)
s.set_grid(V, B, 0.5, 0.05, 0.4)
leaves = s.S ** H
s.set_option("prune", 1)
s.solve(nat.MODE_FULL, nat.COST_MM, H, sc[:8, :3], sc[:8, 3:5], sc[:8, :2])       # warm-up (lazy loading, tables)

cost = np.empty(n_total); index = np.empty(n_total, np.int64); ctl = np.empty((n_total, 2))
units = pruned = 0
t0 = time.perf_counter()
for b0 in range(0, n_total, batch):
    x = sc[b0:b0 + batch]
    r = s.solve(nat.MODE_FULL, nat.COST_MM, H, x[:, :3], x[:, 3:5], x[:, :2])
    cost[b0:b0 + batch], index[b0:b0 + batch], ctl[b0:b0 + batch] = r["cost"], r["index"], r["first_control"]
    st = s.stats()
    units += st["units"] * len(x); pruned += st["pruned_units"]
dt = time.perf_counter() - t0
print(f"config4 FULL-SIZE: {n_total} scenarios x {leaves:.4e} leaves = {n_total*leaves:.4e} rollouts, H={H}, S={s.S}, "
      f"prune=1 (exact), batches of {batch}: {dt:.2f} s host wall incl. copies -> {n_total/dt:.1f} solves/s, "
      f"effective {n_total*leaves/dt:.3e} rollouts/s; depth-3 nodes evaluated {1 - pruned/max(units,1):.4f} of all; "
      f"checksum index_sum={int(index.sum())} cost_sum={cost.sum():.6f}", flush=True)

rng = np.random.default_rng(7)
pick = np.sort(rng.choice(n_total, size=min(n_check, n_total), replace=False))
s.set_option("prune", 0)
t0 = time.perf_counter()
x = sc[pick]
r = s.solve(nat.MODE_FULL, nat.COST_MM, H, x[:, :3], x[:, 3:5], x[:, :2])
dte = time.perf_counter() - t0
same = bool(np.array_equal(r["index"], index[pick]) and np.array_equal(r["cost"], cost[pick])
            and np.array_equal(r["first_control"], ctl[pick]))
print(f"config4 check: {len(pick)} scenarios re-solved with prune=0 in {dte:.2f} s ({len(pick)*leaves/dte:.3e} rollouts/s): "
      f"identical index/cost/control = {same}", flush=True)
ok = same
for k in pick[:n_oracle]:
    t0 = time.perf_counter()
    o = K.solve_full(sc[k, :3], sc[k, 3:5], sc[k, :2], V, B, H, C.COST_MM)
    dto = time.perf_counter() - t0
    good = int(o["index"]) == int(index[k]) and abs(o["cost"] - cost[k]) <= 1e-12 * abs(o["cost"])
    ok = ok and good
    print(f"config4 oracle: scenario {k}: C oracle leaf={int(o['index'])} cost={o['cost']:.9f} in {dto:.1f} s "
          f"({leaves/dto:.3e} rollouts/s on {os.cpu_count()} host threads) GPU leaf={int(index[k])} cost={cost[k]:.9f} "
          f"{'OK' if good else 'MISMATCH'}", flush=True)
s.close()
sys.exit(0 if ok else 1)
