"""Independent pin of the big-tree answers (task 2): BASELINE configs[2] (H=5, 32x32 grid, 1.126e15 leaves)
and configs[4] (H=6, 16x16 grid, 2.815e14 leaves) on the config.py scenario with EVERY leaf evaluated -- prune=0 and
screen=0: one MUFU.SQRT per leaf, no bound of any kind involved -- split over the GPUs of one box by first control
(mpcb_solve_tree_split_device).  Rank 0 then solves the same trees on its own GPU with the exact branch-and-bound and
compares, and writes the records to gpurun_out/r2_bigtree_exhaustive.json (committed as profiles/r2_bigtree_exhaustive.json,
which bench.py's bigtree leg checks against on every run).

   python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 tools/pin_bigtree.py
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from diplomjourney_b200 import _native as nat, config as cfg

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
s = nat.Solver(local)
ids = [nat.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
comm = nat.NcclComm(s, world, rank, ids[0])
ext = torch.cuda.ExternalStream(s.stream)
x = np.array([cfg.x_0, cfg.y_0, cfg.phi_0, cfg.x_t, cfg.y_t], dtype=np.float64)
st = torch.tensor(x[:3].copy(), device="cuda"); tg = torch.tensor(x[3:5].copy(), device="cuda"); og = torch.tensor(x[:2].copy(), device="cuda")
oc = torch.empty(1, dtype=torch.float64, device="cuda"); oi = torch.empty(1, dtype=torch.int64, device="cuda")
ot = torch.empty(8 * 3, dtype=torch.float64, device="cuda"); ou = torch.empty(2, dtype=torch.float64, device="cuda")
trees = [("configs[4] H=6 16x16", 6, 16), ("configs[2] H=5 32x32", 5, 32)]
only = os.environ.get("MPCB_PIN_ONLY")
out, ok = {}, True
for name, H, n in trees:
    if only and only not in name:
        continue
    V = np.linspace(0.0, cfg.v_max, n); B = np.linspace(-cfg.beta_max, cfg.beta_max, n)
    s.set_grid(V, B, cfg.L, cfg.delta_t, cfg.v_min)
    s.set_option("prune", 0); s.set_option("screen", 0)
    # warm-up on a small tree of the same grid (tables, NCCL connections)
    s.solve_tree_split_device(comm, nat.COST_MM, 3, 1, st.data_ptr(), tg.data_ptr(), og.data_ptr(), 0, oc.data_ptr(), oi.data_ptr(), ot.data_ptr(), ou.data_ptr())
    s.sync(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(ext):
        e0.record()
        s.solve_tree_split_device(comm, nat.COST_MM, H, 1, st.data_ptr(), tg.data_ptr(), og.data_ptr(), 0, oc.data_ptr(), oi.data_ptr(), ot.data_ptr(), ou.data_ptr())
        e1.record()
    s.sync(); torch.cuda.synchronize()
    tt = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    recs = [None] * world
    dist.all_gather_object(recs, (float(oc[0]), int(oi[0]), ot[:3 * H].cpu().tolist(), ou.cpu().tolist()))
    agree = all(r == recs[0] for r in recs)
    leaves = (n * n) ** H
    rec = dict(H=H, grid=f"{n}x{n}", leaves=leaves, scenario="config.py: start (0,0,0), target (1,5)", mode="prune=0, screen=0: every leaf evaluated, one MUFU.SQRT each",
               gpus=world, seconds=float(tt[0]), rollouts_per_s=leaves / float(tt[0]), cost=recs[0][0], leaf=recs[0][1],
               traj=recs[0][2], first_control=recs[0][3], ranks_agree=agree)
    if rank == 0:
        s.set_option("prune", 1); s.set_option("screen", 1)
        t = time.perf_counter()
        p = s.solve(nat.MODE_FULL, nat.COST_MM, H, x[:3], x[3:5], x[:2])
        dt = time.perf_counter() - t
        same = int(p["index"][0]) == rec["leaf"] and float(p["cost"][0]) == rec["cost"] and p["traj"][0].reshape(-1).tolist() == rec["traj"] \
            and p["first_control"][0].tolist() == rec["first_control"]
        rec["pruned_one_gpu"] = dict(seconds=dt, identical=bool(same))
        ok = ok and same and agree
        print(json.dumps({name: rec}), flush=True)
    out[name] = rec
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/r2_bigtree_exhaustive.json", "w") as f:
        json.dump(out, f, indent=1)
    print("pin", "OK" if ok else "FAIL", flush=True)
comm.close(); s.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
