#!/bin/bash
# One GPU box: the GPU test-suite, the cfg2/cfg4 bench lines and the big-tree / config-0 tools; outputs under gpurun_out/<tag>_*.
# usage (from the repo root): gpurun --timeout 700 -- 'bash tools/gpu/check_pruning.sh r1m'
tag=${1:-run}
mkdir -p gpurun_out
(timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/${tag}_pytest.log; cat gpurun_out/${tag}_pytest.log
timeout 100 python bench.py --no-cpu --steps 6 > gpurun_out/${tag}_bench_nocpu.json 2> gpurun_out/${tag}_bench.err
timeout 100 python bench.py --no-cpu --steps 6 --workload cfg4 > gpurun_out/${tag}_bench_cfg4.json 2>> gpurun_out/${tag}_bench.err
timeout 120 python tools/bigtree.py 5 32 > gpurun_out/${tag}_bigtree.txt 2>&1
timeout 150 python tools/bigtree.py 6 16 >> gpurun_out/${tag}_bigtree.txt 2>&1
timeout 150 python tools/bigtree.py 5 32 64 2.5 >> gpurun_out/${tag}_bigtree.txt 2>&1
timeout 60 python tools/config0.py 1 > gpurun_out/${tag}_config0.txt 2>&1
for f in gpurun_out/${tag}_bench_nocpu.json gpurun_out/${tag}_bench_cfg4.json; do
  python -c "import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['pruned']['ms_per_step'], d['pruned']['evaluated_fraction'])" $f
done
cat gpurun_out/${tag}_bigtree.txt gpurun_out/${tag}_config0.txt; tail -3 gpurun_out/${tag}_bench.err
