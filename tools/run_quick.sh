#!/bin/bash
# quick GPU iteration: selected tests, step time, per-kernel event profile
mkdir -p gpurun_out
TAG=${1:-q}
TESTS=${2:-tests/test_gpu_forward.py}
{
timeout 900 python -m pytest $TESTS -x -q -m gpu 2>&1 | tail -15
echo "=== step"
timeout 300 python tools/step_time.py fp16c8 30 2>&1 | tail -2
timeout 300 python tools/layer_times.py fp16c8 32 detail > gpurun_out/layer_times_${TAG}.json 2>&1
} > gpurun_out/${TAG}.log 2>&1
tail -25 gpurun_out/${TAG}.log
