#!/bin/bash
# A/B experiments on one box: step time under env knobs
mkdir -p gpurun_out
TAG=${1:-exp}
{
for v in 0 40 20; do
  MCG_TUNE_PAIR_MIN_MTILES=$v timeout 300 python tools/step_time.py fp16c8 30 2>&1 | tail -1
done
MCG_TUNE_PAIR_MIN_MTILES=40 timeout 300 python tools/layer_times.py fp16c8 32 detail 2>&1 | tail -1 > gpurun_out/layer_times_${TAG}_pm40.json
} > gpurun_out/${TAG}.log 2>&1
cat gpurun_out/${TAG}.log
