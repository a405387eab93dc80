#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bneck_tail_kernel -s 5 -c 5 -o gpurun_out/ncu_r02_bneck -f \
    python tools/profile_step.py fp16c8 32 1 > gpurun_out/ncu_r02_bneck.log 2>&1
tail -2 gpurun_out/ncu_r02_bneck.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r02u.csv \
    python tools/profile_step.py fp16c8 32 1 > gpurun_out/ncu_list_r02u.log 2>&1
tail -1 gpurun_out/ncu_list_r02u.log
