#!/bin/bash
# ncu evidence for profiles/: launch list of exactly one bench step (second forward of tools/profile_step.py)
mkdir -p gpurun_out
PREC=${1:-fp16c8}
N=${2:-150}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s $N -c $N --csv --log-file gpurun_out/launches_${PREC}.csv \
    python tools/profile_step.py $PREC 32 1 > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log
