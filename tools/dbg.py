import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from diplomjourney_b200 import _native as nat
from oracle import closed_form as C
s = nat.Solver(0)
V = [0.0, 0.4, 0.7, 1.0]; B = np.round(np.radians([-60, -30, 0, 30, 60]), 3)
s.set_grid(V, B, 0.5, 0.05, 0.0)
o = C.solve_full([0,0,0], (1,5), (0,0), V, B, 3, 'mm', threshold=100050990.584786)
print("oracle", o['index'], o['cost'])
for algo in (1, 2):
    for refine in (1, 0):
        s.set_option("algo", algo); s.set_option("refine", refine)
        r = s.solve(nat.MODE_FULL, nat.COST_MM, 3, [0,0,0], (1,5), (0,0), threshold=100050990.584786)
        print(algo, refine, r['index'], r['cost'], r['first_control'], s.stats())
xy, J = s.dump_leaves(nat.MODE_FULL, nat.COST_MM, 3, [0,0,0], (1,5), (0,0), algo=1)
Jo = C.full_leaf_costs([0,0,0], (1,5), (0,0), V, B, 3, 'mm')
print(J[:6], Jo[:6], np.argmin(J), np.argmin(Jo), np.abs(J-Jo)[Jo<1e7].max())
