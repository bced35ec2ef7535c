#!/usr/bin/env python
"""bench.py -- MPC inner-loop benchmark (BASELINE.json metric: candidate rollouts/s and MPC solves/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cfg2|cfg4]

Workload at every N (weak scaling: fixed work per GPU, robots are independent -- no collective):
  cfg2 = BASELINE.json configs[1]: 1,024 independent robots per GPU, horizon 3, FULL tree
  (math_model.py:159-200) with the MM tracking cost (math_model.py:82-86) on the default
  acceleration-window control grid 11 x 41 (math_model_tree.py:239-256 at v=0.5, beta=0; S=451):
  91,733,851 leaf rollouts per robot, 9.39e10 per step.  Scenarios: numpy default_rng(0) over the
  distribution of run_math_model.py:235-239.  A "step" is one batched solve of all robots.

One JSON line on stdout (rank 0).  `value` = rollouts/s with inputs resident in HBM (device
API, CUDA events on the library's stream, max over ranks); `e2e` = the same through the
host-buffer C-ABI call (pinned host inputs copied H2D and results copied D2H inside the timed
region).  `--impl reference` times the CPU port of the reference path (oracle/ref_port.py:
per-node scipy.integrate.quad, Python loops) on all host cores, on a bounded slice of the
same workload.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

H = 3
N_ROBOTS = 1024
SM_COUNT = 148
MUFU_PER_CLK_SM = 16       # XU lanes per SM per clock
FP32_PER_CLK_SM = 128


def workload(name):
    from oracle import closed_form as C
    if name == "cfg4":   # configs[3] shape, trimmed batch: 16x16 grid, H=4 (4.29e9 leaves per scenario)
        V = np.linspace(0.0, 1.0, 16)
        B = np.linspace(-math.radians(60), math.radians(60), 16)
        return dict(name="cfg4: FULL tree, H=4, 16x16 linspace grid (S=256), MM cost", V=V, B=B, H=4, n=32, seed=1)
    V, B = C.vector_of_velocities(0.5), C.vector_of_beta_angles(0.0)
    return dict(name="cfg2: 1024 robots/GPU, FULL tree, H=3, acceleration-window grid 11x41 (S=451), MM cost",
                V=np.array(V), B=np.array(B), H=3, n=N_ROBOTS, seed=0)


def peaks():
    p = dict(sm_max_mhz=1965.0, hbm_gbs=6650.0, src="fallback")
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            m = json.load(f)
        p.update(sm_max_mhz=float(m["sm_max_mhz"]), hbm_gbs=float(m["hbm_gbs"]), src="measured")
    except Exception:
        pass
    return p


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
                getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
            }
            while not self._stop_evt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
                time.sleep(0.02)
        except Exception as e:  # NVML missing: report it rather than fail the bench
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        busy = s[len(s) // 4:] if s else []
        return dict(sm_mhz=(statistics.median(busy) if busy else None), sm_max_mhz=self.max_mhz,
                    reasons=sorted(self.reasons), samples=len(s))


def cpu_sample(wl, seconds_target=12.0, procs=None):
    """Bounded sample of the same workload on the host cores with the reference-style port."""
    from oracle import closed_form as C
    from oracle import ref_port as P
    procs = procs or os.cpu_count() or 1
    scen = C.random_scenarios(max(procs, 8), wl["seed"])
    S = len(wl["V"]) * len(wl["B"])
    if wl["H"] != 3:
        raise SystemExit("the reference CPU path is hard-coded to H=3 (math_model.py:160-186)")
    # ~33 us per leaf per core (measured on the B200 box's hosts) -> second-level subtrees per process for ~12 s
    n_i1 = max(1, min(S, int(seconds_target / (33e-6 * S))))
    leaves, wall, per = P.timed_sample_full(list(wl["V"]), list(wl["B"]), scen, "mm", n_i1, procs)
    return dict(value=leaves / wall, unit="rollouts/s", cores=procs, kind="port",
                sample=f"{procs} processes x {n_i1} second-level subtrees ({n_i1 * S} leaves each) of the FULL H=3 "
                       f"S={S} tree, per-node scipy.integrate.quad as in math_model.py:90-114; {wall:.1f}s wall",
                rollouts_per_s_per_core=leaves / sum(per))


def closed_form_cpu(wl):
    """The same solve with the float64 closed-form C port on all cores (context, not the baseline)."""
    from oracle import c_oracle as K
    from oracle import closed_form as C
    s = C.random_scenarios(2, wl["seed"])
    K.solve_full(s[0][:3], s[0][3:5], s[0][:2], wl["V"], wl["B"], wl["H"])   # warm (builds the .so)
    t = time.perf_counter()
    K.solve_full(s[1][:3], s[1][3:5], s[1][:2], wl["V"], wl["B"], wl["H"])
    dt = time.perf_counter() - t
    S = len(wl["V"]) * len(wl["B"])
    return dict(value=S ** wl["H"] / dt, unit="rollouts/s", cores=os.cpu_count(),
                note="oracle/mpc_oracle.c: float64 closed form with prefix sharing, pthreads, 1 solve")


def held_metrics(solver, nat, C, device_index):
    """MPC solves/s of the online (HELD, CoordinateTree-pruned) controller -- the other half of the
    BASELINE metric: (a) whole closed loops resident on the device for a batch of robots,
    (b) one tick through the reference-API call on host buffers, (c) a 1,024-robot batch on the
    FULL scripts' 201x121 grid (S=24,321 candidates per solve)."""
    from diplomjourney_b200 import config
    out = {}
    rng = np.random.default_rng(0)
    n = 4096
    params = nat.LoopParams.from_config(config, nat.COST_TREE, 3, 256)
    init = np.zeros((n, 5)); init[:, 2] = rng.uniform(-1, 1, n)
    ang = init[:, 2] + rng.uniform(-0.5, 0.5, n); d = rng.uniform(1.0, 4.0, n)
    tgt = np.stack([d * np.cos(ang), d * np.sin(ang)], 1)
    for _ in range(3):
        t = time.perf_counter()
        r = solver.held_closed_loop(params, init, tgt, [[0.0, 0.0]], first_threshold=1e10)
        dt = time.perf_counter() - t
    out["closed_loop_batch"] = dict(robots=n, ticks=int(r["ticks"].sum()), seconds=dt,
                                    solves_per_s=float(r["ticks"].sum() / dt), S_max=451,
                                    note="mpcb_held_closed_loop_host: whole math_mpc loops on the device, host wall incl. copies")
    V, B = C.vector_of_velocities(0.5), C.vector_of_beta_angles(0.0)
    x = C.random_scenarios(1, 3)[0]
    reps = 300
    solver.set_grid(V, B, C.CONFIG["L"], C.CONFIG["delta_t"], C.CONFIG["v_min"])
    solver.solve(nat.MODE_HELD, nat.COST_TREE, 3, x[:3], x[3:5], x[:2])
    t = time.perf_counter()
    for _ in range(reps):
        solver.set_grid(V, B, C.CONFIG["L"], C.CONFIG["delta_t"], C.CONFIG["v_min"])
        solver.solve(nat.MODE_HELD, nat.COST_TREE, 3, x[:3], x[3:5], x[:2])
    dt = (time.perf_counter() - t) / reps
    out["single_tick_host_api"] = dict(us_per_solve=dt * 1e6, solves_per_s=1.0 / dt, S=len(V) * len(B),
                                       note="set_grid + solve per tick, host buffers, 1 H2D + 1 kernel + 1 D2H")
    Vf, Bf = C.grid_full_default()
    solver.set_grid(Vf, Bf, C.CONFIG["L"], C.CONFIG["delta_t"], C.CONFIG["v_min"])
    sc = C.random_scenarios(1024, 1)
    for _ in range(3):
        t = time.perf_counter()
        solver.solve(nat.MODE_HELD, nat.COST_TREE, 3, sc[:, :3], sc[:, 3:5], sc[:, :2])
        dt = time.perf_counter() - t
    out["batch_201x121"] = dict(robots=1024, S=int(Vf.size * Bf.size), seconds=dt, solves_per_s=1024 / dt,
                                rollouts_per_s=1024 * Vf.size * Bf.size / dt)
    return out


def bigtree_metrics(solver, nat):
    """BASELINE configs[2] and configs[4]: ONE oversized FULL tree (config.py scenario) on this GPU through the host API
    with the exact branch-and-bound -- the only way such a tree is solved in finite time; `exhaustive_s` is the same
    tree at the measured every-leaf rate of the headline kernel (profiles/r1i_config3.txt, r1g_multigpu.txt)."""
    from diplomjourney_b200 import config as cfg
    out = {}
    solver.set_option("prune", 1)
    for name, H, n, exhaustive_s in (("configs[2] H=5 32x32", 5, 32, 364.0), ("configs[4] H=6 16x16", 6, 16, 19.2 * 8)):
        V = np.linspace(0.0, cfg.v_max, n)
        B = np.linspace(-cfg.beta_max, cfg.beta_max, n)
        solver.set_grid(V, B, cfg.L, cfg.delta_t, cfg.v_min)
        st, tg, og = (cfg.x_0, cfg.y_0, cfg.phi_0), (cfg.x_t, cfg.y_t), (cfg.x_0, cfg.y_0)
        best = None
        for _ in range(3):
            t = time.perf_counter()
            r = solver.solve(nat.MODE_FULL, nat.COST_MM, H, st, tg, og)
            dt = time.perf_counter() - t
            best = dt if best is None else min(best, dt)
        stt = solver.stats()
        out[name] = dict(leaves=int(stt["leaves_per_solve"]), seconds=best, solves_per_s=1.0 / best,
                         leaf=int(r["index"][0]), cost=float(r["cost"][0]),
                         nodes_set_up=int(stt["units"] - stt["pruned_units"]), nodes=int(stt["units"]),
                         exhaustive_seconds_one_gpu=exhaustive_s)
    solver.set_option("prune", 0)
    return out


def run_reference(args):
    wl = workload(args.workload)
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    for i in range(args.warmup + args.steps):
        r = cpu_sample(wl, seconds_target=float(os.environ.get("MPCB_CPU_SAMPLE_SECONDS", "6.0")))
        if i >= args.warmup:
            vals.append(r)
    v = statistics.mean(x["value"] for x in vals)
    last = vals[-1]
    S = len(wl["V"]) * len(wl["B"])
    line = dict(metric="candidate rollouts/s (MPC inner loop)", value=v, unit="rollouts/s", impl="reference",
                n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * (S ** wl["H"] * wl["n"]) / v, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload=wl["name"], robots_per_gpu=wl["n"], H=wl["H"], S=S),
                solves_per_s=v / S ** wl["H"],
                cpu_baseline=dict(value=v, unit="rollouts/s", cores=last["cores"], kind="port", sample=last["sample"]),
                e2e=dict(value=v, unit="rollouts/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                note="ms_per_step is extrapolated from the timed slice: the reference cannot finish one step")
    print(json.dumps(line), flush=True)


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from diplomjourney_b200 import _native as nat
    from oracle import closed_form as C   # scenario generator + parity gate only

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)

    wl = workload(args.workload)
    Hh, n = wl["H"], wl["n"]
    S = len(wl["V"]) * len(wl["B"])
    leaves_per_solve = S ** Hh
    solver = nat.Solver(local)
    solver.set_grid(wl["V"], wl["B"], C.CONFIG["L"], C.CONFIG["delta_t"], C.CONFIG["v_min"])
    solver.set_option("algo", {"auto": nat.ALGO_AUTO, "prefix": nat.ALGO_PREFIX, "leafwalk": nat.ALGO_LEAFWALK}[args.algo])
    # the headline evaluates EVERY leaf (as the reference does); the exact branch-and-bound is reported separately
    solver.set_option("prune", 0)
    if args.nodes_per_thread:
        solver.set_option("nodes_per_thread", args.nodes_per_thread)

    # each rank owns its own robots (contiguous ranges of the global batch): no data-path collective
    scen = C.random_scenarios(n * world, wl["seed"])[rank * n:(rank + 1) * n]
    st_h = torch.from_numpy(np.ascontiguousarray(scen[:, :3])).pin_memory()
    tg_h = torch.from_numpy(np.ascontiguousarray(scen[:, 3:5])).pin_memory()
    og_h = torch.from_numpy(np.ascontiguousarray(scen[:, :2])).pin_memory()
    st, tg, og = st_h.to(dev), tg_h.to(dev), og_h.to(dev)
    oc = torch.empty(n, dtype=torch.float64, device=dev)
    oi = torch.empty(n, dtype=torch.int64, device=dev)
    ot = torch.empty(n, Hh, 3, dtype=torch.float64, device=dev)
    ou = torch.empty(n, 2, dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2
    ext = torch.cuda.ExternalStream(solver.stream, device=dev)

    def step_device():
        solver.solve_device(nat.MODE_FULL, nat.COST_MM, Hh, n, st.data_ptr(), tg.data_ptr(), og.data_ptr(), 0, 0,
                            oc.data_ptr(), oi.data_ptr(), ot.data_ptr(), ou.data_ptr())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing (value + roofline)
    with torch.cuda.stream(ext):
        for _ in range(args.warmup):
            flush.fill_(1)
            step_device()
    stats = solver.stats()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(ext):
        t0.record()
        for k in range(args.steps):
            flush.fill_(k & 0xFF)                 # L2 flush between timed iterations (inside the region)
            evs[k][0].record()
            step_device()
            evs[k][1].record()
        t1.record()
    barrier()
    clocks = sampler.stop()
    total_ms = t0.elapsed_time(t1)
    step_ms = [a.elapsed_time(b) for a, b in evs]
    tt = torch.tensor([total_ms, statistics.mean(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms, kern_ms = float(tt[0]), float(tt[1])
    rollouts_per_step = leaves_per_solve * n * world
    value = rollouts_per_step * args.steps / (total_ms * 1e-3)

    # ---------------- end-to-end through the host-buffer C-ABI call (pinned inputs, H2D + D2H inside)
    out_np = None
    for _ in range(max(1, args.warmup)):
        out_np = solver.solve(nat.MODE_FULL, nat.COST_MM, Hh, st_h.numpy(), tg_h.numpy(), og_h.numpy())
    barrier()
    e0 = time.perf_counter()
    for _ in range(args.steps):
        out_np = solver.solve(nat.MODE_FULL, nat.COST_MM, Hh, st_h.numpy(), tg_h.numpy(), og_h.numpy())
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - e0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = rollouts_per_step * args.steps / float(te[0])
    h2d = n * (3 + 2 + 2) * 8
    d2h = n * (8 + 8 + 3 * Hh * 8 + 16)

    # ---------------- parity gate on the benchmark's own output (a few robots vs the float64 C oracle)
    parity = None
    if rank == 0:
        from oracle import c_oracle as K
        ok = 0
        chk = min(3, n)
        for i in range(chk):
            o = K.solve_full(scen[i, :3], scen[i, 3:5], scen[i, :2], wl["V"], wl["B"], Hh, C.COST_MM)
            ok += int(out_np["index"][i] == o["index"] and abs(out_np["cost"][i] - o["cost"]) <= 1e-12 * o["cost"]
                      and int(oi[i]) == o["index"])
        parity = f"{ok}/{chk} robots identical to the float64 oracle (index, cost rtol 1e-12)"
        if ok != chk:
            raise SystemExit("PARITY FAILURE in bench: " + parity)

    # ---------------- same step with the exact branch-and-bound on (identical results, fewer leaves evaluated)
    pruned = None
    if args.algo != "leafwalk":
        solver.set_option("prune", 1)
        with torch.cuda.stream(ext):
            step_device()
        barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(ext):
            p0.record()
            for _ in range(args.steps):
                step_device()
            p1.record()
        barrier()
        pst = solver.stats()
        same = bool(torch.equal(oi.cpu(), torch.from_numpy(out_np["index"])))
        tp = torch.tensor([p0.elapsed_time(p1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        parents = pst["units"] * n
        pruned = dict(ms_per_step=float(tp[0]) / args.steps, solves_per_s=n * world * args.steps / (float(tp[0]) * 1e-3),
                      evaluated_fraction=1.0 - pst["pruned_units"] / max(parents, 1), same_leaves_as_unpruned=same,
                      note="option prune=1 (the library default): exact branch-and-bound -- nodes and subtrees whose leaves "
                           "provably cannot reach the refinement window are skipped, identical records; `value` above "
                           "is measured with prune=0 (every leaf evaluated)")
        if not same:
            raise SystemExit("PARITY FAILURE in bench: pruned and unpruned solves disagree")
        solver.set_option("prune", 0)
    solver.set_option("algo", nat.ALGO_AUTO)
    held = held_metrics(solver, nat, C, local) if rank == 0 and args.algo == "auto" else None
    bigtree = bigtree_metrics(solver, nat) if rank == 0 and args.algo == "auto" else None
    if rank == 0:
        pk = peaks()
        clk_hz = pk["sm_max_mhz"] * 1e6
        mufu_peak = SM_COUNT * MUFU_PER_CLK_SM * clk_hz
        fp32_peak = SM_COUNT * FP32_PER_CLK_SM * clk_hz
        per_gpu_rate = leaves_per_solve * n / (kern_ms * 1e-3)           # rollouts/s of one GPU, per step
        mufu_a = 2 * Hh + 1                                              # accounting A
        issue_cyc = {1: 10.5, 4: 9.75}.get(args.nodes_per_thread, 10.0)         # issue cycles per rollout of the pair loop
        line = dict(
            metric="candidate rollouts/s (MPC inner loop)", value=value, unit="rollouts/s", n_gpus=world,
            steps=args.steps, warmup=args.warmup, ms_per_step=total_ms / args.steps, higher_is_better=True,
            scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
            config=dict(workload=wl["name"], robots_per_gpu=n, H=Hh, S=S, leaves_per_solve=leaves_per_solve,
                        l2="flushed between timed steps (256 MiB fill)", selection="float64-refined argmin"),
            solves_per_s=n * world * args.steps / (total_ms * 1e-3),
            roofline=dict(
                bound="mufu", unit="Tops/s",
                achieved=per_gpu_rate * mufu_a / 1e12, peak=mufu_peak / 1e12,
                frac=per_gpu_rate * mufu_a / mufu_peak,
                traffic=6840064,   # dram__bytes_read+write per launch, ncu --set full (profiles/r1l_cfg2_full.txt)
                accounting="A: (2H+1) MUFU per rollout, one-thread-per-leaf design (SURVEY 8d); the prefix kernel "
                           "shares prefixes and executes 1 MUFU + 8.5 FP32 lane-ops per rollout (SASS of "
                           "prefix_min_loop_far2x2: per leaf pair and node 7 FFMA2 + 1 FADD2 at 2 issue cycles each, "
                           "2 MUFU.SQRT, 1 FMNMX3, 2 LDS.128 shared by the thread's nodes), so frac>1 under A is expected",
                executed=dict(mufu_per_rollout=1.0, fp32_ops_per_rollout=8.5, issue_cycles_per_rollout=issue_cyc,
                              mufu_frac=per_gpu_rate * 1.0 / mufu_peak, fp32_frac=per_gpu_rate * 8.5 / fp32_peak,
                              issue_frac=per_gpu_rate * issue_cyc / fp32_peak),
                hbm_gbs=(h2d + d2h) / (kern_ms * 1e-3) / 1e9, hbm_peak_gbs=pk["hbm_gbs"],
                peak_src=f"{pk['src']} sm_max_mhz={pk['sm_max_mhz']:.0f} x {SM_COUNT} SMs x {MUFU_PER_CLK_SM} MUFU/clk/SM",
                kernel=("leafwalk_kernel<1,true,1>" if stats["algo"] == nat.ALGO_LEAFWALK else "prefix_kernel<1,true>" if args.nodes_per_thread == 1 else f"prefixn_kernel<true,{args.nodes_per_thread or 2}>") + " (pass 1)",
                kernel_ms_per_step=kern_ms),
            e2e=dict(value=e2e_value, unit="rollouts/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                     solves_per_s=n * world * args.steps / float(te[0])),
            gpu_launches=stats["kernel_launches"] * args.steps,
            clocks=clocks, parity=parity, held=held, pruned=pruned, bigtree=bigtree,
            refine=dict(segments=stats["refine_segments"], candidates=stats["refine_candidates"]),
        )
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_sample(wl)
            line["cpu_closed_form"] = closed_form_cpu(wl)
        print(json.dumps(line), flush=True)
    solver.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg4"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--nodes-per-thread", type=int, default=0, choices=[0, 1, 2, 4],
                    help="prefix pass 1: depth-(H-1) nodes per thread (0 = library default)")
    ap.add_argument("--algo", default="auto", choices=["auto", "prefix", "leafwalk"],
                    help="expansion kernel: prefix (default for this workload) or the one-thread-per-leaf design")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
