#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-pp}
{
timeout 600 python -m pytest tests/test_preprocess.py -x -q -m gpu 2>&1 | tail -15
echo "=== bench_preprocess"
timeout 300 python tools/bench_preprocess.py 400 400 2>&1 | tail -2
timeout 300 python tools/bench_preprocess.py 720 1280 2>&1 | tail -2
timeout 300 python tools/bench_preprocess.py 160 160 2>&1 | tail -2
} > gpurun_out/${TAG}.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:preprocess_kernel -s 3 -c 1 -o gpurun_out/preprocess_${TAG} \
   python tools/bench_preprocess.py 400 400 2 > gpurun_out/ncu_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_${TAG}.log
cat gpurun_out/${TAG}.log
