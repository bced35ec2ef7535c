"""Pruned cfg2 / cfg4 step (device API, CUDA events) under option combinations: python tools/pruned_timing.py"""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from diplomjourney_b200 import _native as nat
from oracle import closed_form as C
dev = torch.device("cuda", 0)
s = nat.Solver(0)
ext = torch.cuda.ExternalStream(s.stream, device=dev)
for name, V, B, H, n, seed in (("cfg2", np.array(C.vector_of_velocities(0.5)), np.array(C.vector_of_beta_angles(0.0)), 3, 1024, 0),
                               ("cfg4", np.linspace(0, 1, 16), np.linspace(-math.radians(60), math.radians(60), 16), 4, 4096, 1)):
    s.set_grid(V, B, 0.5, 0.05, 0.4)
    sc = C.random_scenarios(n, seed)
    st, tg, og = (torch.from_numpy(np.ascontiguousarray(sc[:, a:b])).to(dev) for a, b in ((0, 3), (3, 5), (0, 2)))
    oc = torch.empty(n, dtype=torch.float64, device=dev); oi = torch.empty(n, dtype=torch.int64, device=dev)
    ref = None
    for opts in ({"prefilter": 0}, {"node_list": 0}, {}, {"candidate_list": 1 << 20}):
        s.set_option("prune", 1); s.set_option("prefilter", 1); s.set_option("candidate_list", -1); s.set_option("node_list", 1 << 15)
        for k, v in opts.items():
            s.set_option(k, v)
        def step():
            s.solve_device(nat.MODE_FULL, nat.COST_MM, H, n, st.data_ptr(), tg.data_ptr(), og.data_ptr(), 0, 0, oc.data_ptr(), oi.data_ptr(), 0, 0)
        with torch.cuda.stream(ext):
            step(); step()
        s.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(ext):
            e0.record()
            for _ in range(10):
                step()
            e1.record()
        s.sync(); torch.cuda.synchronize()
        cur = (oi.cpu().numpy().copy(), oc.cpu().numpy().copy())
        same = ref is None or (np.array_equal(cur[0], ref[0]) and np.array_equal(cur[1], ref[1]))
        ref = ref or cur
        stt = s.stats()
        print(f"{name} {opts}: {e0.elapsed_time(e1) / 10:.3f} ms per {n} solves, pruned {stt['pruned_units']}/{stt['units'] * n}, launches {stt['kernel_launches']}, refine segments {stt['refine_segments']} candidates {stt['refine_candidates']}, same={same}", flush=True)
