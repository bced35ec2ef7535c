"""Safety margin of the refinement window: worst measured |J32 - J64| / eps over many small trees."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from diplomjourney_b200 import _native as nat
from oracle import closed_form as C
from test_gpu_parity import _eps_model, COSTS, L, DT, VMIN
s = nat.Solver(0)
for grid_name, V, B in (("4x9 v<=1", [0.0, 0.3, 0.6, 1.0], np.linspace(-1, 1, 9)),
                        ("window 5x9", np.array(C.vector_of_velocities(0.5))[::2][:5], np.array(C.vector_of_beta_angles(0.0))[::5])):
    s.set_grid(V, B, L, DT, VMIN)
    vv, bb, dphi = C.control_tables(V, B, L, DT)
    smax, dphimax = float(np.max(vv) * DT), float(np.max(np.abs(dphi)))
    sc = C.random_scenarios(40, 123)
    for algo_name in ("leafwalk", "leafwalk_direct", "prefix", "prefix_direct"):
        algo = nat.ALGO_LEAFWALK if algo_name.startswith("leafwalk") else nat.ALGO_PREFIX
        direct = algo_name.endswith("_direct")         # the forms the pass-1 kernels rank with, against tol1/2
        s.set_option("dump_direct", 1 if direct else 0)
        for cost in (C.COST_MM, C.COST_TREE):
            worst = {}
            for x in sc:
                for name, st, og in (("far", x, x[:2]), ("offline", x, x[:2] + np.array([5.0, -4.0])),
                                     ("near", np.r_[x[:3], x[0] + 0.05, x[1] + 0.03], x[:2]),
                                     ("far-away", np.r_[x[:3], x[0] + 300.0, x[1] - 200.0], x[:2])):
                    _, J = s.dump_leaves(nat.MODE_FULL, COSTS[cost], 3, st[:3], st[3:5], og, algo=algo)
                    Jo = C.full_leaf_costs(st[:3], st[3:5], og, V, B, 3, cost)
                    ok = Jo < 1e7
                    eps = _eps_model(st, og, 3, cost, algo == nat.ALGO_PREFIX, smax, dphimax, direct)
                    worst[name] = max(worst.get(name, 0.0), float(np.abs(J - Jo)[ok].max() / eps))
            print(f"{grid_name:12s} algo={algo_name:15s} cost={cost:4s} worst |err|/eps: " +
                  "  ".join(f"{k}={v:.3f}" for k, v in worst.items()))
