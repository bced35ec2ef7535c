"""Run under torchrun on a multi-GPU box:
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py
Checks the split-tree path -- the library's own NCCL path (mpcb_solve_tree_split: one 16-byte all-gather per solve,
local lexicographic minimum, local re-roll of the winner), the stand-alone mpcb_allreduce_min, and the
torch.distributed flavour -- against a whole-tree solve and the float64 oracle, and times config-5-shaped trees
(MPCB_SPLIT_H = horizons solved exhaustively, MPCB_SPLIT_H_PRUNED = with the exact branch-and-bound)."""
import os, sys
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
ids = [nat.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
comm = nat.NcclComm(s, world, rank, ids[0])
sc = C.random_scenarios(3, 77)
ok = True


def same(a, b):
    return int(a["index"][0]) == int(b["index"][0]) and a["cost"][0] == b["cost"][0] and np.array_equal(a["traj"], b["traj"]) \
        and np.array_equal(a["first_control"], b["first_control"])


for H in (3, 4):
    for prune in (0, 1):
        s.set_option("prune", prune)
        # a batch of three trees in one collective call, and one with a threshold nothing beats
        nat_b = s.solve_tree_split(comm, nat.COST_MM, H, sc[:, :3], sc[:, 3:5], sc[:, :2])
        none = s.solve_tree_split(comm, nat.COST_MM, H, sc[0, :3], sc[0, 3:5], sc[0, :2], threshold=1.0)
        ok = ok and int(none["index"][0]) == -1
        for i, x in enumerate(sc):
            whole = s.solve(nat.MODE_FULL, nat.COST_MM, H, x[:3], x[3:5], x[:2])
            one = {k: v[i:i + 1] for k, v in nat_b.items()}
            t = D.solve_tree_split(s, nat.COST_MM, H, x[:3], x[3:5], x[:2])          # torch.distributed flavour
            good = same(one, whole) and same(t, whole)
            if H == 3:
                o = K.solve_full(x[:3], x[3:5], x[:2], V, B, H, C.COST_MM)
                good = good and int(whole["index"][0]) == o["index"]
            ok = ok and good
# the reconciliation step on its own (in place on 1-element device buffers): ranks without a leaf, NaN costs
oc = torch.tensor([5.0 + rank], dtype=torch.float64, device="cuda"); oi = torch.tensor([100 - rank], dtype=torch.int64, device="cuda")
comm.allreduce_min(oc.data_ptr(), oi.data_ptr()); s.sync()
ok = ok and (float(oc[0]), int(oi[0])) == (5.0, 100)
oc = torch.tensor([float("nan") if rank == 0 else 7.0], dtype=torch.float64, device="cuda")
oi = torch.tensor([3 if rank == 0 else 50 + rank], dtype=torch.int64, device="cuda")
comm.allreduce_min(oc.data_ptr(), oi.data_ptr()); s.sync()
ok = ok and (float(oc[0]), int(oi[0])) == (7.0, 51)
oc = torch.tensor([float("nan")], dtype=torch.float64, device="cuda"); oi = torch.tensor([-1], dtype=torch.int64, device="cuda")
comm.allreduce_min(oc.data_ptr(), oi.data_ptr()); s.sync()
ok = ok and np.isnan(float(oc[0])) and int(oi[0]) == -1

# a share that holds no leaf near the optimum must not build a refinement window of its own: on the config.py tree the share
# of the v = 0 first controls holds 16^(H-1) exactly tied unmoved nodes (once 2.3e7 float64 candidates on that rank)
from diplomjourney_b200 import config as cfg
s.set_option("prune", 0)
r = s.solve_tree_split(comm, nat.COST_MM, 4, [cfg.x_0, cfg.y_0, cfg.phi_0], [cfg.x_t, cfg.y_t], [cfg.x_0, cfg.y_0])
cands = [None] * world
dist.all_gather_object(cands, s.stats()["refine_candidates"])
ok = ok and int(r["index"][0]) == s.S ** 4 - 1 and sum(cands) <= 64
if rank == 0:
    print(f"world={world} config.py tree H=4, every leaf evaluated: float64 candidates per rank {cands}", flush=True)

x = sc[0]
oc = torch.empty(1, dtype=torch.float64, device="cuda"); oi = torch.empty(1, dtype=torch.int64, device="cuda")
st = torch.tensor(x[:3].copy(), device="cuda"); tg = torch.tensor(x[3:5].copy(), device="cuda"); og = torch.tensor(x[:2].copy(), device="cuda")
torch.cuda.synchronize()
ext = torch.cuda.ExternalStream(s.stream)
runs = [(int(h), 0) for h in os.environ.get("MPCB_SPLIT_H", "4,5").split(",") if h] + \
       [(int(h), 1) for h in os.environ.get("MPCB_SPLIT_H_PRUNED", "").split(",") if h]
for H, prune in runs:
    s.set_option("prune", prune)
    # 256^5 = 1.1e12 leaves; 256^6 = 2.8e14 leaves is BASELINE config 5
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(); torch.cuda.synchronize()
    with torch.cuda.stream(ext):
        ev0.record()
        s.solve_tree_split_device(comm, nat.COST_MM, H, 1, st.data_ptr(), tg.data_ptr(), og.data_ptr(), 0, oc.data_ptr(), oi.data_ptr(), 0, 0)
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
              f"[the first run includes NCCL connection setup]", flush=True)
if rank == 0:
    print(f"world={world} parity={'OK' if ok else 'FAIL'}", flush=True)
comm.close(); s.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
