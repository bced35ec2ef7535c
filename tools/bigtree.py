"""One oversized FULL tree on ONE GPU with the exact branch-and-bound: BASELINE configs[2] (H=5, 32x32 grid,
1.126e15 leaves) and configs[4] (H=6, 16x16 grid, 2.815e14 leaves), whole tree, with every mode of the subtree cut;
every run must return the same leaf, cost and trajectory.
usage: python tools/bigtree.py <H> <n> [first-control range for the subtree_cut=0 comparison = S/8] [start heading = config.phi_0]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from diplomjourney_b200 import _native as nat, config as cfg

H, n = int(sys.argv[1]), int(sys.argv[2])
V = np.linspace(0.0, cfg.v_max, n)
B = np.linspace(-cfg.beta_max, cfg.beta_max, n)
phi0 = float(sys.argv[4]) if len(sys.argv) > 4 else cfg.phi_0
state, target, origin = (cfg.x_0, cfg.y_0, phi0), (cfg.x_t, cfg.y_t), (cfg.x_0, cfg.y_0)
s = nat.default_solver()
s.set_grid(V, B, cfg.L, cfg.delta_t, cfg.v_min)
S = s.S
part = int(sys.argv[3]) if len(sys.argv) > 3 else max(1, S // 8)
s.solve(nat.MODE_FULL, nat.COST_MM, 2, state, target, origin)          # warm-up: module load, tables


def run(label, subtree, i0_range=None):
    s.set_option("prune", 1)
    s.set_option("subtree_cut", subtree)
    t = time.perf_counter()
    r = s.solve(nat.MODE_FULL, nat.COST_MM, H, state, target, origin, i0_range=i0_range)
    dt = time.perf_counter() - t
    st = s.stats()
    leaves = st["leaves_per_solve"]
    print(f"bigtree H={H} S={S} phi0={phi0:g} {label}: leaves={leaves:.4e} time={dt*1e3:.3f} ms effective {leaves/dt:.3e} rollouts/s "
          f"leaf={int(r['index'][0])} cost={r['cost'][0]:.6f} first_control={r['first_control'][0].tolist()} "
          f"pruned_nodes={st['pruned_units']}/{st['units']} launches={st['kernel_launches']} "
          f"refine(seg={st['refine_segments']},cand={st['refine_candidates']})", flush=True)
    return r


whole = run("whole tree, subtree_cut=2 (frontier descent)", 2)
again = run("whole tree, subtree_cut=2 (second call)", 2)
tiles = run("whole tree, subtree_cut=1 (tile bound)", 1)
i0w = int(whole["index"][0]) // S ** (H - 1)                            # first control of the winner
rng = (max(0, min(S - part, i0w - part // 2)), max(0, min(S - part, i0w - part // 2)) + part)
a = run(f"i0 in [{rng[0]},{rng[1]}), subtree_cut=2", 2, rng)
b = run(f"i0 in [{rng[0]},{rng[1]}), subtree_cut=0 (node-level cut only)", 0, rng)
same = all(int(x["index"][0]) == int(y["index"][0]) and x["cost"][0] == y["cost"][0] and np.array_equal(x["traj"], y["traj"])
           for x, y in ((whole, again), (whole, tiles), (a, b)))
inside = rng[0] * S ** (H - 1) <= int(whole["index"][0]) < rng[1] * S ** (H - 1)
if inside:
    same = same and int(a["index"][0]) == int(whole["index"][0]) and a["cost"][0] == whole["cost"][0]
print(f"bigtree H={H} S={S}: identical results = {same} (winner inside the compared range: {inside})", flush=True)
s.set_option("subtree_cut", 2)
sys.exit(0 if same else 1)
