#!/bin/bash
# One gpurun session; everything lands in gpurun_out/<tag>.*
mkdir -p gpurun_out
TAG=${1:-exp}
{
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
echo "=== layer times"
timeout 300 python tools/layer_times.py fp16x3 32 detail
timeout 300 python tools/layer_times.py fp16 32 detail
echo "=== bench"
timeout 600 python bench.py 2>&1 | tail -3
} > gpurun_out/${TAG}.log 2>&1
tail -40 gpurun_out/${TAG}.log
