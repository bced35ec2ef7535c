"""Randomised exactness sweep of everything that skips work on the GPU: thousands of random scenarios (incl. robots facing
away from / next to / far from their targets, targets off the tracked line) x grids x horizons x both costs, each batch
solved with prune=0, screen=0 (every leaf, one sqrt each: no bound of any kind) and then with the screened pass 1
(prune=0, screen=1) and with the branch-and-bound (prune=1) in both subtree-cut modes, with and without the fp32
pre-filter; index, cost, trajectory and first control must be bit-identical.
usage: python tools/prune_sweep.py [scenarios per case = 512]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from diplomjourney_b200 import _native as nat
from oracle import closed_form as C

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
s = nat.Solver(0)
s.set_option("algo", nat.ALGO_PREFIX)
rad = np.radians
cases = [("window 11x41", C.vector_of_velocities(0.5), C.vector_of_beta_angles(0.0), 3),
         ("window at v_max", C.vector_of_velocities(1.0), C.vector_of_beta_angles(rad(55)), 3),
         ("16x16 from rest", np.linspace(0, 1, 16), np.linspace(-rad(60), rad(60), 16), 3),
         ("16x16 from rest", np.linspace(0, 1, 16), np.linspace(-rad(60), rad(60), 16), 4),
         ("7x9 wide steer", np.linspace(0.1, 1, 7), np.linspace(-1.4, 1.4, 9), 4),
         ("5x5 with reverse", np.linspace(-0.5, 1, 5), np.linspace(-1, 1, 5), 5),
         ("3x4", np.linspace(0, 1, 3), np.linspace(-1, 1, 4), 6)]
bad = total = 0
t0 = time.perf_counter()
for ci, (name, V, B, H) in enumerate(cases):
    s.set_grid(V, B, 0.5, 0.05, 0.4)
    sc = C.random_scenarios(n, 1000 + ci)
    k = n // 8
    sc[:k, 3:5] = sc[:k, :2] + np.random.default_rng(ci).uniform(-0.1, 0.1, (k, 2))        # target within a few steps
    sc[k:2 * k, 3:5] = sc[k:2 * k, :2] + np.random.default_rng(ci + 50).uniform(-300, 300, (k, 2))   # far away
    origin = sc[:, :2].copy()
    origin[2 * k:3 * k] += np.random.default_rng(ci + 99).uniform(-5, 5, (k, 2))           # robot off the tracked line
    for cost in (nat.COST_MM, nat.COST_TREE):
        args = (nat.MODE_FULL, cost, H, sc[:, :3], sc[:, 3:5], origin)
        s.set_option("prune", 0); s.set_option("screen", 0)
        ref = s.solve(*args)
        for label, opts in (("screen=1", dict(prune=0, screen=1)),
                            ("subtree_cut=1", dict(prune=1, subtree_cut=1, prefilter=1)),
                            ("subtree_cut=3", dict(prune=1, subtree_cut=3, prefilter=1)),
                            ("subtree_cut=1 prefilter=0", dict(prune=1, subtree_cut=1, prefilter=0))):
            for k_, v_ in opts.items():
                s.set_option(k_, v_)
            r = s.solve(*args)
            st = s.stats()
            same = all(np.array_equal(r[key], ref[key], equal_nan=True) for key in ("index", "cost", "traj", "first_control"))
            total += n; bad += 0 if same else int(np.sum(r["index"] != ref["index"]) or 1)
            print(f"{name:18s} H={H} S={s.S:4d} cost={'mm' if cost == nat.COST_MM else 'tree'} {label:26s}: "
                  f"{n} scenarios identical={same} nodes evaluated {1 - st['pruned_units'] / (st['units'] * n):.5f}", flush=True)
s.set_option("subtree_cut", 2); s.set_option("screen", 1); s.set_option("prefilter", 1)
print(f"prune_sweep: {total} screened / pruned solves compared with exhaustive ones, {bad} differ, {time.perf_counter() - t0:.1f} s")
sys.exit(1 if bad else 0)
