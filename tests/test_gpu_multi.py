"""Multi-GPU paths on a real box (skipped with fewer than 2 visible GPUs): the split tree reconciled through
torch.distributed NCCL rounds and through the library's own NCCL path, and bench.py at N=2."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    try:
        from diplomjourney_b200 import _native
        return _native.load().mpcb_device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ngpus() < 2, reason="needs 2 GPUs")
def test_split_tree_two_ranks_nccl():
    env = dict(os.environ, MPCB_SPLIT_H="4", MPCB_SPLIT_H_PRUNED="4")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "multigpu_check.py")],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "parity=OK" in r.stdout


@pytest.mark.skipif(_ngpus() < 2, reason="needs 2 GPUs")
def test_bench_two_ranks():
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29534", os.path.join(ROOT, "bench.py"),
                        "--gpus", "2", "--steps", "2", "--warmup", "3", "--cfg4-scenarios", "4096"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(line) == 1
    d = json.loads(line[0])
    assert d["n_gpus"] == 2 and d["scaling"] == "weak" and d["value"] > 1e12
    # the two multi-GPU configurations of the baseline ride on the same line: configs[3] strong-scaled (no collective) and
    # configs[4], ONE tree split over the ranks and reconciled by the library's own NCCL all-gather
    assert d["cfg4_strong"]["scaling"] == "strong" and d["cfg4_strong"]["exhaustive_check"]["identical"]
    runs = d["split_tree"]["runs"]
    assert {(r["H"], r["scenario"]) for r in runs} == {(6, "config.py"), (6, "tie-heavy"), (5, "config.py"), (5, "tie-heavy")}
    assert all(r["ranks_agree"] for r in runs) and all(r["equals_one_gpu_whole_tree"] for r in runs if r["H"] == 6)


@pytest.mark.skipif(_ngpus() < 2, reason="needs 2 GPUs")
def test_two_handles_on_two_devices_in_one_process():
    """Two handles on two devices driven from ONE process: plain solves alternate from a single host thread (every
    entry point must select its handle's device), and the split-tree collective -- whose scratch lives in the handle,
    on the handle's device -- runs with one driver thread per handle."""
    import threading

    import numpy as np
    from diplomjourney_b200 import _native as nat
    from oracle import closed_form as C

    s = [nat.Solver(0), nat.Solver(1)]
    V = np.linspace(0.0, 1.0, 9); B = np.linspace(-1.0, 1.0, 11)
    sc = C.random_scenarios(4, 5)
    for k in (0, 1):
        s[k].set_grid(V, B, 0.5, 0.05, 0.4)
    whole = [None, None]
    for rep in range(2):                                   # one host thread, devices interleaved
        for k in (1, 0):
            whole[k] = s[k].solve(nat.MODE_FULL, nat.COST_MM, 3, sc[:, :3], sc[:, 3:5], sc[:, :2])
    for key in ("index", "cost", "traj"):
        np.testing.assert_array_equal(whole[0][key], whole[1][key])
    uid = nat.nccl_unique_id()
    comms, out, err = [None, None], [None, None], []

    def drive(k):
        try:
            comms[k] = nat.NcclComm(s[k], 2, k, uid)
            for _ in range(2):
                out[k] = s[k].solve_tree_split(comms[k], nat.COST_MM, 3, sc[:, :3], sc[:, 3:5], sc[:, :2])
        except Exception as e:                             # surfaced below
            err.append(e)

    th = [threading.Thread(target=drive, args=(k,)) for k in (0, 1)]
    [t.start() for t in th]
    [t.join(timeout=300) for t in th]
    assert not err and out[0] is not None and out[1] is not None, err
    for k in (0, 1):
        for key in ("index", "cost", "traj", "first_control"):
            np.testing.assert_array_equal(out[k][key], whole[0][key])
    for k in (0, 1):
        comms[k].close()
        s[k].close()
