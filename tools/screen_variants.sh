#!/bin/bash
# GPU box: the exhaustive pass-1 variants of the prefix kernel on the cfg2 bench (rollouts/s), bit-identical answers checked by bench.py's parity gate
mkdir -p gpurun_out
for scr in 0 1 2; do for npt in 2 4; do
  echo "== screen=$scr npt=$npt" >> gpurun_out/r2a_variants.txt
  python bench.py --no-cpu --steps 6 --screen $scr --nodes-per-thread $npt 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('value %.4e  ms/step %.3f  kernel_ms %.3f  e2e %.4e  parity %s  refine %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['e2e']['value'], d['parity'], d['refine']))
    elif 'Error' in l or 'FAIL' in l: print(l.strip())
" >> gpurun_out/r2a_variants.txt
done; done
cat gpurun_out/r2a_variants.txt
