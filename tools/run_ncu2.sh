#!/bin/bash
# ncu --set full captures of chosen tcgen05 GEMM launches of one bench step (second forward of tools/profile_step.py)
mkdir -p gpurun_out
TAG=${1:-r02}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_gemm_kernel -s 81 -c 5 -o gpurun_out/ncu_${TAG}_layer1 -f \
    python tools/profile_step.py fp16c8 32 1 > gpurun_out/ncu_${TAG}_layer1.log 2>&1
tail -2 gpurun_out/ncu_${TAG}_layer1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_gemm_kernel -s 131 -c 2 -o gpurun_out/ncu_${TAG}_lat0_fpn0 -f \
    python tools/profile_step.py fp16c8 32 1 > gpurun_out/ncu_${TAG}_lat0.log 2>&1
tail -2 gpurun_out/ncu_${TAG}_lat0.log
ls -la gpurun_out/*.ncu-rep | tail -3
