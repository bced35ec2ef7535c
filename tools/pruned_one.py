"""A few pruned steps (for ncu captures of the pruned chain): python tools/pruned_one.py [cfg2|cfg4]"""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from diplomjourney_b200 import _native as nat
from oracle import closed_form as C
s = nat.Solver(0)
if len(sys.argv) > 1 and sys.argv[1] == "cfg4":      # BASELINE configs[3]: 16 x 16 grid from rest, H = 4 (a 4,096-scenario slice)
    s.set_grid(np.linspace(0, 1, 16), np.linspace(-math.radians(60), math.radians(60), 16), 0.5, 0.05, 0.4)
    sc, H = C.random_scenarios(4096, 1), 4
else:
    s.set_grid(C.vector_of_velocities(0.5), C.vector_of_beta_angles(0.0), 0.5, 0.05, 0.4)
    sc, H = C.random_scenarios(1024, 0), 3
for _ in range(3):
    r = s.solve(nat.MODE_FULL, nat.COST_MM, H, sc[:, :3], sc[:, 3:5], sc[:, :2])
print(int(r["index"].sum()), s.stats())
