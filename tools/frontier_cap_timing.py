"""Pruned 4,096-scenario slice of BASELINE configs[3] (16 x 16 grid from rest, H = 4) under different frontier capacities:
python tools/frontier_cap_timing.py"""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from diplomjourney_b200 import _native as nat
from oracle import closed_form as C
dev = torch.device("cuda", 0)
s = nat.Solver(0)
ext = torch.cuda.ExternalStream(s.stream, device=dev)
s.set_grid(np.linspace(0, 1, 16), np.linspace(-math.radians(60), math.radians(60), 16), 0.5, 0.05, 0.4)
n, H = 4096, 4
sc = C.random_scenarios(n, 1)
st, tg, og = (torch.from_numpy(np.ascontiguousarray(sc[:, a:b])).to(dev) for a, b in ((0, 3), (3, 5), (0, 2)))
oc = torch.empty(n, dtype=torch.float64, device=dev); oi = torch.empty(n, dtype=torch.int64, device=dev)
ref = None
for cap in (1 << 22, 1 << 24, 1 << 25, 1 << 26):
    s.set_option("frontier_cap", cap)
    def step():
        s.solve_device(nat.MODE_FULL, nat.COST_MM, H, n, st.data_ptr(), tg.data_ptr(), og.data_ptr(), 0, 0, oc.data_ptr(), oi.data_ptr(), 0, 0)
    with torch.cuda.stream(ext):
        step(); step()
    s.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(ext):
        e0.record()
        for _ in range(5):
            step()
        e1.record()
    s.sync(); torch.cuda.synchronize()
    cur = (oi.cpu().numpy().copy(), oc.cpu().numpy().copy())
    same = ref is None or (np.array_equal(cur[0], ref[0]) and np.array_equal(cur[1], ref[1]))
    ref = ref or cur
    stt = s.stats()
    print(f"frontier_cap 2^{int(math.log2(cap))}: {e0.elapsed_time(e1) / 5:.3f} ms per {n} solves, launches {stt['kernel_launches']}, pruned {stt['pruned_units']}, same={same}", flush=True)
