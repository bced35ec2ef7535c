#!/bin/bash
# usage: run_variants.sh OUT "screen levels" variant...   (on the GPU box)
out=$1; shift; levels=$1; shift
for v in "$@"; do for scr in $levels; do
  echo "== $v screen=$scr" >> $out
  MPCB_LIB=$PWD/build/variants/lib_$v.so python bench.py --no-cpu --no-legs --steps 6 --screen $scr 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('value %.4e  ms/step %.3f  kernel_ms %.3f  parity %s  refine %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['parity'][:3], d['refine']))
    elif 'Error' in l or 'FAIL' in l: print(l.strip())
" >> $out
done; done
cat $out
