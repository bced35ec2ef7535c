"""The C-ABI library: builds with nvcc for sm_100a, loads, and exports every symbol that
include/mpcb200.h declares.  No compute calls here (no GPU in the CPU suite)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "mpcb200.h")).read()
    return re.findall(r"MPCB_API\s+[\w\s\*]+?\b(mpcb_\w+)\s*\(", src)


def test_header_is_plain_c():
    # the boundary is C: the header must compile as C11 with no CUDA/torch headers
    r = subprocess.run(["gcc", "-std=c11", "-fsyntax-only", "-x", "c", os.path.join(ROOT, "include", "mpcb200.h")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_library_builds_and_exports_every_declared_symbol():
    from diplomjourney_b200 import _native, build
    path = build.build_library()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    names = declared_symbols()
    assert len(names) >= 17 and "mpcb_solve_batch_host" in names and "mpcb_allreduce_min" in names
    for n in names:
        assert hasattr(lib, n), n
    assert _native.load().mpcb_version() >= 100


def test_sm100a_code_is_in_the_library():
    from diplomjourney_b200 import build
    out = subprocess.run(["cuobjdump", "-lelf", build.build_library()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback_without_a_device():
    """Without a CUDA device the product path must fail loudly, not answer from the CPU."""
    from diplomjourney_b200 import _native
    if _native.load().mpcb_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(_native.MpcbError):
        _native.Solver(0)
    from diplomjourney_b200 import math_model_tree as mt
    mt._backend = None
    with pytest.raises(_native.MpcbError):
        mt.predictive_control(0.0, 0.0, 0.0, 2, 3, [0.5], [0.0], False)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "diplomjourney_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+(oracle|tests)\b", txt, re.M), f


def test_every_option_the_library_accepts_is_documented_in_the_header():
    api = open(os.path.join(ROOT, "diplomjourney_b200", "csrc", "mpcb_api.cu")).read()
    body = api[api.index("int mpcb_set_option("):]
    body = body[:body.index("\n}\n")]
    accepted = set(re.findall(r'strcmp\(name, "(\w+)"\)', body))
    assert {"prune", "subtree_cut", "algo", "refine"} <= accepted
    header = open(os.path.join(ROOT, "include", "mpcb200.h")).read()
    for name in accepted:
        assert f'"{name}"' in header, f"option {name} is not described in include/mpcb200.h"


def test_bench_line_helpers():
    """bench.py without a GPU: both arms describe the workload with the same `config` object, the per-rollout operation
    counts cover every kernel the line can name, and the roofline's ncu figures come from a committed capture."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    wl = bench.workload("cfg2")
    cfg = bench.config_of(wl)
    assert cfg["S"] == 451 and cfg["H"] == 3 and cfg["robots_per_gpu"] == 1024 and cfg["leaves_per_solve"] == 451 ** 3
    assert set(bench.EXECUTED) == {"prefix_screen", "prefix_full", "leafwalk"}
    for key in bench.EXECUTED:
        rec = bench.ncu_record(key)
        assert rec is not None and os.path.exists(os.path.join(ROOT, rec["file"])) and rec["traffic"] > 0, key
        assert 0 < rec["fma_pipe_pct"] < 100 and 0 < rec["xu_pipe_pct"] < 100
    pins = bench.pinned_bigtree()
    assert {name for name, _, _ in bench.BIGTREES} == set(pins) and all(p["ranks_agree"] for p in pins.values())
