#!/bin/bash
# ncu evidence for profiles/: (1) launch list of one bench step, (2) --set full capture of the dominant GEMM launches
mkdir -p gpurun_out
PREC=${1:-fp16c8}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 190 -c 200 --csv --log-file gpurun_out/launches_${PREC}.csv \
    python tools/profile_step.py $PREC 32 1 > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log
# fpn convs + laterals are launches 53..60 of the umma kernel within a forward (second forward: skip 76 + 53)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:umma_gemm -s 129 -c 8 -o gpurun_out/umma_${PREC} -f \
    python tools/profile_step.py $PREC 32 1 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out/
