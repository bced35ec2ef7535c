"""BASELINE configs[0] on the GPU: one tick of run_math_model.py's default scenario (S=24,321, H=3, 1.44e13 leaves)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diplomjourney_b200 import run_math_model as rm
t = time.perf_counter()
r = rm.predictive_control(rm.x, rm.y, rm.phi, rm.v, rm.x_t, rm.y_t)
dt = time.perf_counter() - t
from diplomjourney_b200 import _native
st = _native.default_solver().stats()
print(f"config0 tick 1: S={rm.size_max_1} leaves={rm.size_max_3} time={dt:.2f}s rate={rm.size_max_3/dt:.3e} rollouts/s "
      f"result={[float(x) for x in r]} leaf={rm.last_leaf_index} criterion={rm.optimal_criterion:.6f} stats={st}")
