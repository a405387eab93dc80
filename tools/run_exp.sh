#!/bin/bash
# One gpurun session; everything lands in gpurun_out/<tag>.*
mkdir -p gpurun_out
TAG=${1:-exp}
{
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
echo "=== bench"
timeout 900 python bench.py 2>&1 | tail -3
} > gpurun_out/${TAG}.log 2>&1
tail -12 gpurun_out/${TAG}.log
