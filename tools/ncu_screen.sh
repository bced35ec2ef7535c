#!/bin/bash
# ncu --set full of the exhaustive pass-1 kernel for each screen level (one B200)
set -x
for scr in 1 2; do
  python bench.py --no-cpu --steps 2 --screen $scr > gpurun_out/r2b_plain_s$scr.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:prefixn_kernel -s 3 -c 1 -f -o gpurun_out/r2b_screen$scr \
      python bench.py --no-cpu --steps 2 --screen $scr > gpurun_out/r2b_ncu_s$scr.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
