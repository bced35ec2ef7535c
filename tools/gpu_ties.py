import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from diplomjourney_b200 import _native as nat
from oracle import closed_form as C
s = nat.Solver(0)
V = np.linspace(0.0, 1.0, 16); B = np.linspace(-np.radians(60), np.radians(60), 16)
s.set_grid(V, B, 0.5, 0.05, 0.4)
sc = C.random_scenarios(6, 77)
for H in (3, 4):
    for i, x in enumerate(sc):
        for rep in range(2):
            t = time.perf_counter(); r = s.solve(nat.MODE_FULL, nat.COST_MM, H, x[:3], x[3:5], x[:2]); dt = time.perf_counter() - t
        st = s.stats()
        print(f"H={H} scen={i} idx={int(r['index'][0])} cost={r['cost'][0]:.3f} time={dt*1e3:.2f} ms rate={256**H/dt:.3e} refine_segments={st['refine_segments']} candidates={st['refine_candidates']} segs={st['segments']}")
