#!/bin/bash
# ncu evidence of the shipped binary (one B200): launch list of the bench command, --set full of the dominant pass-1
# kernel (prefixn_kernel<.,2,screen>) and of the one-thread-per-leaf kernel.  Outputs under gpurun_out/$1_*.
p=gpurun_out/$1
set -x
python bench.py --no-cpu --no-legs --steps 2 > ${p}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${p}_launches.csv \
    python bench.py --no-cpu --no-legs --steps 2 > ${p}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:prefixn_kernel -s 3 -c 1 -f -o ${p}_prefixn \
    python bench.py --no-cpu --no-legs --steps 2 > ${p}_ncu_prefixn.log 2>&1
python bench.py --no-cpu --no-legs --steps 2 --algo leafwalk > ${p}_plain_leafwalk.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:leafwalk_kernel -s 6 -c 2 -f -o ${p}_leafwalk \
    python bench.py --no-cpu --no-legs --steps 2 --algo leafwalk > ${p}_ncu_leafwalk.log 2>&1
ls -la gpurun_out/*.ncu-rep
