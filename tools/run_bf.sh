#!/bin/bash
# first contact of the fused bottleneck kernel: guarded by timeouts (a hung kernel must not hang the box)
mkdir -p gpurun_out
TAG=${1:-bf}
{
timeout 240 python -m pytest tests/test_gpu_forward.py -x -q -m gpu -k "fused_bottleneck" 2>&1 | tail -25
echo "rc=$?"
echo "=== step"
timeout 120 python tools/step_time.py fp16c8 30 2>&1 | tail -2
MCG_TUNE_BF_PAIR=0 timeout 120 python tools/step_time.py fp16c8 30 2>&1 | tail -1
timeout 200 python tools/layer_times.py fp16c8 32 detail > gpurun_out/layer_times_${TAG}.json 2>&1
} > gpurun_out/${TAG}.log 2>&1
tail -40 gpurun_out/${TAG}.log
