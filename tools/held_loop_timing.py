"""Where the time of the device-resident HELD closed loop goes: kernel alone (device pointers,
mpcb_held_closed_loop_device) vs the host call with its copies and the 40 B/tick log.

    gpurun -- 'python tools/held_loop_timing.py'
"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diplomjourney_b200 import _native as nat, config  # noqa: E402


def main():
    s = nat.default_solver(0)
    rng = np.random.default_rng(0)
    for n, max_ticks in ((4096, 256), (16384, 128), (1184, 256), (148, 256)):
        params = nat.LoopParams.from_config(config, nat.COST_TREE, 3, max_ticks)
        init = np.zeros((n, 5)); init[:, 2] = rng.uniform(-1, 1, n)
        ang = init[:, 2] + rng.uniform(-0.5, 0.5, n); d = rng.uniform(1.0, 4.0, n)
        tgt = np.stack([d * np.cos(ang), d * np.sin(ang)], 1)
        org = np.zeros((n, 2))
        thr = np.full(n, 1e10)
        dev = torch.device("cuda:0")
        t_init, t_tgt, t_org, t_thr = (torch.from_numpy(a).to(dev) for a in (init, tgt, org, thr))
        t_log = torch.empty((n, max_ticks, 5), dtype=torch.float64, device=dev)
        t_ticks = torch.empty(n, dtype=torch.int32, device=dev)
        t_status = torch.empty(n, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        vp = lambda t: C.c_void_p(t.data_ptr())
        best = 1e9
        for _ in range(5):
            t0 = time.perf_counter()
            rc = s.lib.mpcb_held_closed_loop_device(s.h, C.byref(params), n, vp(t_init), vp(t_tgt), vp(t_org), vp(t_thr),
                                                    None, vp(t_log), vp(t_ticks), vp(t_status))
            assert rc == 0
            s.sync()
            best = min(best, time.perf_counter() - t0)
        ticks = int(t_ticks.sum().item())
        t0 = time.perf_counter()
        r = s.held_closed_loop(params, init, tgt, org, first_threshold=1e10)
        host = time.perf_counter() - t0
        assert int(r["ticks"].sum()) == ticks
        print(f"robots {n:6d} max_ticks {max_ticks}: ticks {ticks}  device {best * 1e3:8.3f} ms "
              f"({ticks / best:.3e} solves/s)  host call {host * 1e3:8.3f} ms ({ticks / host:.3e} solves/s)")


if __name__ == "__main__":
    main()
