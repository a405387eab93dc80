#!/bin/bash
# One gpurun session: gpu tests, bench line, ncu launch list of one eager step, per-kernel event profile.
mkdir -p gpurun_out
TAG=${1:-r02a}
{
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
echo "=== bench"
timeout 900 python bench.py 2>&1 | tail -3
} > gpurun_out/${TAG}.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python tools/profile_step.py fp16c8 32 1 > gpurun_out/ncu_list_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_list_${TAG}.log
timeout 300 python tools/layer_times.py fp16c8 32 detail > gpurun_out/layer_times_${TAG}.json 2>&1
tail -30 gpurun_out/${TAG}.log
