"""Timing of the device-resident closed loop with and without the operator-event script, next to the per-tick path on
which the host applies the events (one batched launch per tick).  usage: python tools/events_timing.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time, numpy as np, importlib
from diplomjourney_b200 import _native as nat, config
mt = importlib.import_module("diplomjourney_b200.math_model_tree")
rng = np.random.default_rng(1)
for n in (1, 1024, 16384):
    init = np.zeros((n, 5)); init[:, :2] = rng.uniform(-1, 1, (n, 2)); init[:, 2] = rng.uniform(-0.5, 7, n)
    tgt = np.tile([2.0, 3.0], (n, 1)) + rng.uniform(-1, 1, (n, 2))
    for ev in (False, True):
        mt.math_mpc_batch(init, tgt, max_ticks=200, events=ev)
        t0 = time.perf_counter(); r = mt.math_mpc_batch(init, tgt, max_ticks=200, events=ev); dt = time.perf_counter() - t0
        print(f"device loop N={n} events={ev}: {dt*1e3:.2f} ms, {r['ticks'].sum()} ticks, {r['ticks'].sum()/dt:.3e} ticks/s, status {np.bincount(r['status'], minlength=4)}", flush=True)
    if n == 1024:
        t0 = time.perf_counter(); h = mt.math_mpc_batch(init, tgt, max_ticks=200, events=True, host_loop=True); dt = time.perf_counter() - t0
        print(f"per-tick host loop N={n} events=True: {dt*1e3:.1f} ms, {h['ticks'].sum()/dt:.3e} ticks/s")
