#!/bin/bash
# One gpurun session; everything lands in gpurun_out/<tag>.log
mkdir -p gpurun_out
TAG=${1:-exp}
{
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
echo "=== layer times"
timeout 300 python tools/layer_times.py fp16x3 32
timeout 300 python tools/layer_times.py fp16 32
echo "=== step times"
timeout 300 python tools/gpu_diag.py bench_fp16x3_graph bench_fp16_graph
} > gpurun_out/${TAG}.log 2>&1
tail -40 gpurun_out/${TAG}.log
