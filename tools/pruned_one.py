"""A few pruned cfg2 steps (for ncu captures of the pruned chain): python tools/pruned_one.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from diplomjourney_b200 import _native as nat
from oracle import closed_form as C
s = nat.Solver(0)
s.set_grid(C.vector_of_velocities(0.5), C.vector_of_beta_angles(0.0), 0.5, 0.05, 0.4)
sc = C.random_scenarios(1024, 0)
for _ in range(3):
    r = s.solve(nat.MODE_FULL, nat.COST_MM, 3, sc[:, :3], sc[:, 3:5], sc[:, :2])
print(int(r["index"].sum()), s.stats())
