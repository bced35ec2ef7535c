"""Multi-GPU plumbing (one process per GPU, torch.distributed): SURVEY section 8e.

* Batches of independent robots/scenarios shard by contiguous ranges -- no data-path collective.
* ONE oversized tree is split by contiguous ranges of the FIRST control, so that rank order equals
  leaf-index order; each rank reduces its share to a (cost, index) record and the records are
  reconciled by a lexicographic minimum: two 8-byte all-reduce(min) rounds (float64 cost, then int64
  index among the ranks that hold that cost).  A single packed 64-bit word cannot hold a 32-bit cost
  and the up-to-2^50 leaf index, hence two rounds.  With the NCCL backend the tensors live on the GPU
  and the rounds run over NVLink; the same code runs on gloo/CPU tensors in the tests.
  ``Solver``-level alternative without torch: ``mpcb_allreduce_min`` of the C ABI (native NCCL).
"""
from __future__ import annotations

import math

import numpy as np

from . import _native

_I64_MAX = np.iinfo(np.int64).max


def shard_range(n: int, world: int, rank: int):
    """Contiguous, balanced [lo, hi) of n items for ``rank`` (first n % world ranks get one more)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def owner_of_first_control(i0: int, S: int, world: int) -> int:
    """Rank whose shard_range(S, world, rank) contains first control i0."""
    base, extra = divmod(S, world)
    edge = extra * (base + 1)
    return i0 // (base + 1) if i0 < edge else extra + (i0 - edge) // max(base, 1)


def combine_min(cost, index, group=None):
    """In-place lexicographic (cost, index) minimum over the ranks of ``group``.
    ``cost``: float64 tensor [n], ``index``: int64 tensor [n] (-1 = no leaf); any device the backend
    supports.  NaN costs are treated as +inf."""
    import torch
    import torch.distributed as dist

    c = torch.where(torch.isnan(cost), torch.full_like(cost, math.inf), cost)
    gmin = c.clone()
    dist.all_reduce(gmin, op=dist.ReduceOp.MIN, group=group)
    contrib = torch.where((c == gmin) & (index >= 0), index, torch.full_like(index, _I64_MAX))
    dist.all_reduce(contrib, op=dist.ReduceOp.MIN, group=group)
    cost.copy_(gmin)
    index.copy_(torch.where(contrib == _I64_MAX, torch.full_like(contrib, -1), contrib))
    return cost, index


def solve_tree_split(solver: "_native.Solver", cost_kind, H, state, target, origin, threshold=None, group=None,
                     device=None):
    """One FULL tree (a single robot) split over the ranks of ``group`` by first control.
    Every rank returns the same dict(cost, index, traj, first_control) as ``Solver.solve``."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    S = solver.S
    lo, hi = shard_range(S, world, rank)
    part = solver.solve(_native.MODE_FULL, cost_kind, H, state, target, origin, threshold=None, i0_range=(lo, hi))
    dev = device if device is not None else ("cuda" if dist.get_backend(group) == "nccl" else "cpu")
    c = torch.tensor(part["cost"], dtype=torch.float64, device=dev)
    i = torch.tensor(part["index"], dtype=torch.int64, device=dev)
    combine_min(c, i, group)
    win = int(i[0])
    payload = torch.zeros(3 * H + 2, dtype=torch.float64, device=dev)
    owner = 0
    if win >= 0:
        owner = owner_of_first_control(win // (S ** (H - 1)), S, world)
        if rank == owner:
            payload[:3 * H] = torch.from_numpy(part["traj"][0].reshape(-1)).to(dev)
            payload[3 * H:] = torch.from_numpy(part["first_control"][0]).to(dev)
    dist.broadcast(payload, src=dist.get_global_rank(group, owner) if group is not None else owner, group=group)
    best = float(c[0])
    thr = math.inf if threshold is None else float(threshold)
    accepted = win >= 0 and best < thr
    out = payload.cpu().numpy()
    return dict(cost=np.array([best]), index=np.array([win if accepted else -1], dtype=np.int64),
                traj=out[:3 * H].reshape(1, H, 3), first_control=out[3 * H:].reshape(1, 2))


def solve_batch_sharded(solver: "_native.Solver", mode, cost_kind, H, state, target, origin, threshold=None,
                        flags=None, group=None, gather=True):
    """N independent solves sharded by contiguous ranges over the ranks; optionally all-gathered so
    that every rank holds all N results (object gather: control-plane, not on the solve's critical path)."""
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    st = np.asarray(state, np.float64).reshape(-1, np.asarray(state).shape[-1])
    n = st.shape[0]
    lo, hi = shard_range(n, world, rank)
    sl = slice(lo, hi)
    pick = lambda a: None if a is None else (np.asarray(a)[sl] if np.ndim(a) > 0 and np.shape(a)[0] == n else a)
    tg = np.asarray(target, np.float64).reshape(-1, 2)
    og = np.asarray(origin, np.float64).reshape(-1, 2)
    mine = solver.solve(mode, cost_kind, H, st[sl], tg[sl] if tg.shape[0] == n else tg,
                        og[sl] if og.shape[0] == n else og, threshold=pick(threshold), flags=pick(flags)) \
        if hi > lo else dict(cost=np.empty(0), index=np.empty(0, np.int64), traj=np.empty((0, H, 3)),
                             first_control=np.empty((0, 2)))
    if not gather:
        return mine, (lo, hi)
    parts = [None] * world
    dist.all_gather_object(parts, mine, group=group)
    return {k: np.concatenate([p[k] for p in parts], axis=0) for k in mine}, (lo, hi)
