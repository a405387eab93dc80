#!/bin/bash
# GPU iteration for the evaluation driver: evaluate / metric / preprocess tests + test-split stand-in
mkdir -p gpurun_out
TAG=${1:-q2}
{
timeout 900 python -m pytest tests/test_evaluate.py tests/test_metric.py tests/test_preprocess.py -x -q -m gpu 2>&1 | tail -15
echo "=== testsplit video"
timeout 300 python tools/bench_testsplit.py video 32 8 2>&1 | tail -1
echo "=== testsplit clip"
timeout 300 python tools/bench_testsplit.py clip 32 8 2>&1 | tail -1
} > gpurun_out/${TAG}.log 2>&1
tail -25 gpurun_out/${TAG}.log
