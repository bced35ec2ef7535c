"""BASELINE configs[0] on the GPU: ticks of run_math_model.py's default scenario (S=24,321, H=3, 1.44e13 leaves).
usage: python tools/config0.py [prune: 0|1]   (default 1)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diplomjourney_b200 import run_math_model as rm
from diplomjourney_b200 import _native

prune = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
_native.default_solver().set_option("prune", prune)
state = (rm.x, rm.y, rm.phi, rm.v)
for tick in (1, 2):     # tick 1 includes the one-off table build/upload and lazy kernel loading
    t = time.perf_counter()
    r = rm.predictive_control(*state, rm.x_t, rm.y_t)
    dt = time.perf_counter() - t
    st = _native.default_solver().stats()
    print(f"config0 prune={int(prune)} tick {tick}: S={rm.size_max_1} leaves={rm.size_max_3} time={dt*1e3:.1f} ms "
          f"rate={rm.size_max_3/dt:.3e} rollouts/s result={[float(x) for x in r]} leaf={rm.last_leaf_index} "
          f"criterion={rm.optimal_criterion:.6f} stats={st}")
    state = (r[0], r[1], r[2], r[3])
