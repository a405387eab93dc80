#!/bin/bash
# One gpurun session for the round's evidence: gpu tests, bench line, ncu launch list of one eager step.
mkdir -p gpurun_out
TAG=${1:-r01e}
{
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
echo "=== bench"
timeout 600 python bench.py 2>&1 | tail -3
} > gpurun_out/${TAG}.log 2>&1
# first forward (warm-up, plans) + one profiled step: keep every launch, the post-processing takes the last step
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python tools/profile_step.py fp16c8 32 1 > gpurun_out/ncu_list_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_list_${TAG}.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:preprocess_kernel -s 3 -c 1 -o gpurun_out/preprocess_${TAG} \
   python tools/bench_preprocess.py 320 320 2 > gpurun_out/ncu_pp_${TAG}.log 2>&1
tail -12 gpurun_out/${TAG}.log
timeout 300 python tools/bench_testsplit.py 32 8 2>&1 | tail -1 > gpurun_out/testsplit_${TAG}.json
cat gpurun_out/testsplit_${TAG}.json | cut -c1-400
