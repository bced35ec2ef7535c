"""Run under torchrun on a multi-GPU box:
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py
Checks the split-tree path (torch.distributed NCCL rounds AND the library's native NCCL mpcb_allreduce_min)
against a whole-tree solve and the float64 oracle, and times a config-5-shaped tree."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from diplomjourney_b200 import _native as nat, distributed as D
from oracle import closed_form as C, c_oracle as K

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
s = nat.Solver(local)
V = np.linspace(0.0, 1.0, 16); B = np.linspace(-np.radians(60), np.radians(60), 16)
s.set_grid(V, B, 0.5, 0.05, 0.4)
sc = C.random_scenarios(3, 77)
ok = True
for H in (3, 4):
    for x in sc:
        r = D.solve_tree_split(s, nat.COST_MM, H, x[:3], x[3:5], x[:2])
        whole = s.solve(nat.MODE_FULL, nat.COST_MM, H, x[:3], x[3:5], x[:2])
        same = int(r["index"][0]) == int(whole["index"][0]) and r["cost"][0] == whole["cost"][0] and np.array_equal(r["traj"], whole["traj"])
        if H == 3:
            o = K.solve_full(x[:3], x[3:5], x[:2], V, B, H, C.COST_MM)
            same = same and int(r["index"][0]) == o["index"]
        ok = ok and same
# native NCCL reconciliation
ids = [nat.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
comm = nat.NcclComm(s, world, rank, ids[0])
x = sc[0]
lo, hi = D.shard_range(s.S, world, rank)
oc = torch.empty(1, dtype=torch.float64, device="cuda"); oi = torch.empty(1, dtype=torch.int64, device="cuda")
st = torch.tensor(x[:3].copy(), device="cuda"); tg = torch.tensor(x[3:5].copy(), device="cuda"); og = torch.tensor(x[:2].copy(), device="cuda")
torch.cuda.synchronize()
ext = torch.cuda.ExternalStream(s.stream)
runs = [(int(h), 0) for h in os.environ.get("MPCB_SPLIT_H", "4,5").split(",")] + \
       [(int(h), 1) for h in os.environ.get("MPCB_SPLIT_H_PRUNED", "").split(",") if h]
for H, prune in runs:
    s.set_option("prune", prune)
    # 256^5 = 1.1e12 leaves; 256^6 = 2.8e14 leaves is BASELINE config 5
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(); torch.cuda.synchronize()
    with torch.cuda.stream(ext):
        ev0.record()
        s.solve_device(nat.MODE_FULL, nat.COST_MM, H, 1, st.data_ptr(), tg.data_ptr(), og.data_ptr(), 0, 0, oc.data_ptr(), oi.data_ptr(), 0, 0, i0_range=(lo, hi))
        comm.allreduce_min(oc.data_ptr(), oi.data_ptr())
        ev1.record()
    s.sync(); torch.cuda.synchronize()
    tt = torch.tensor([ev0.elapsed_time(ev1) * 1e-3], device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    allrec = [None] * world
    dist.all_gather_object(allrec, (float(oc[0]), int(oi[0])))
    ok = ok and all(a == allrec[0] for a in allrec)
    st_ = s.stats()
    if rank == 0:
        leaves = s.S ** H
        print(f"world={world} split-tree H={H} prune={prune} leaves={leaves:.3e} record={allrec[0]} time={float(tt[0]):.4f}s "
              f"rate={leaves/float(tt[0]):.3e} rollouts/s refine(seg={st_['refine_segments']},cand={st_['refine_candidates']}) pruned_nodes={st_['pruned_units']}/{st_['units']} "
              f"[first H includes NCCL connection setup]", flush=True)
if rank == 0:
    print(f"world={world} parity={'OK' if ok else 'FAIL'}", flush=True)
comm.close(); s.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
