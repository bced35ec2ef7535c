"""The N>1 host logic on CPU: world_size-2 gloo processes, the oracle standing in for the GPU solver.
Covers the split-tree reconciliation (two all-reduce(min) rounds, lowest index on exact ties,
owner lookup, trajectory broadcast) and the sharded batch."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import closed_form as C

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, q):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from diplomjourney_b200 import distributed as D
    from oracle_backend import OracleBackend
    out = {}
    # 1. combine_min: exact ties -> lowest index; NaN and "no leaf" records never win
    c = torch.tensor([5.0, 7.0, float("nan"), 3.0] if rank == 0 else [5.0, 6.0, 2.0, float("inf")], dtype=torch.float64)
    i = torch.tensor([40, 1, 9, 8] if rank == 0 else [12, 2, 7, -1], dtype=torch.int64)
    D.combine_min(c, i)
    out["combine"] = (c.tolist(), i.tolist())
    # 2. one tree split by first control
    V, B = [0.0, 0.5, 1.0], list(np.linspace(-1, 1, 5))
    be = OracleBackend()
    be.set_grid(V, B, 0.5, 0.05, 0.4)
    s = C.random_scenarios(1, 5)[0]
    r = D.solve_tree_split(be, 0, 3, s[:3], s[3:5], s[:2], device="cpu")
    out["split"] = (float(r["cost"][0]), int(r["index"][0]), r["traj"][0].tolist(), r["first_control"][0].tolist())
    r = D.solve_tree_split(be, 0, 3, s[:3], s[3:5], s[:2], threshold=float(r["cost"][0]), device="cpu")
    out["split_rejected"] = int(r["index"][0])
    # 3. sharded batch of 5 robots over 2 ranks (3 + 2)
    sc = C.random_scenarios(5, 9)
    res, rng = D.solve_batch_sharded(be, 1, 1, 3, sc[:, :3], sc[:, 3:5], sc[:, :2])
    out["batch"] = (res["index"].tolist(), res["cost"].tolist(), rng)
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_helpers():
    from diplomjourney_b200 import distributed as D
    for n, w in ((256, 8), (451, 8), (5, 2), (3, 4), (24321, 7)):
        edges = [D.shard_range(n, w, r) for r in range(w)]
        assert edges[0][0] == 0 and edges[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
        assert max(hi - lo for lo, hi in edges) - min(hi - lo for lo, hi in edges) <= 1
        for i0 in range(n):
            r = D.owner_of_first_control(i0, n, w)
            assert edges[r][0] <= i0 < edges[r][1]


def test_two_rank_gloo_split_tree_and_sharded_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0]["batch"][2] == (0, 3) and got[1]["batch"][2] == (3, 5)
    for r in got.values():
        r["batch"] = r["batch"][:2]
    assert got[0] == got[1]                                   # every rank ends with the same answer
    cost, idx = got[0]["combine"]
    assert cost == [5.0, 6.0, 2.0, 3.0] and idx == [12, 2, 7, 8]
    V, B = [0.0, 0.5, 1.0], list(np.linspace(-1, 1, 5))
    s = C.random_scenarios(1, 5)[0]
    whole = C.solve_full(s[:3], s[3:5], s[:2], V, B, 3, C.COST_MM)
    c, i, traj, ctl = got[0]["split"]
    assert i == whole["index"] and c == whole["cost"]
    np.testing.assert_allclose(traj, whole["traj"], atol=0)
    assert tuple(ctl) == whole["first_control"]
    assert got[0]["split_rejected"] == -1
    sc = C.random_scenarios(5, 9)
    Vh, Bh = V, B
    ref = [C.solve_held(x[:3], x[3:5], x[:2], Vh, Bh, 3, C.COST_TREE) for x in sc]
    assert got[0]["batch"][0] == [r["index"] for r in ref]
    np.testing.assert_allclose(got[0]["batch"][1], [r["cost"] for r in ref], rtol=0)
