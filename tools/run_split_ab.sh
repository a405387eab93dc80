#!/bin/bash
# two-chain trunk region (option split_layers): bit-identity test, then same-box A/B of the replayed step
mkdir -p gpurun_out
{
timeout 400 python -m pytest tests/test_gpu_forward.py -x -q -m gpu -k "two_chain" 2>&1 | tail -8
for rep in 1 2; do
  for f in "MCG_TUNE_SPLIT_LAYERS=0" "MCG_TUNE_SPLIT_LAYERS=4" "MCG_TUNE_SPLIT_LAYERS=12" "MCG_TUNE_SPLIT_LAYERS=8"; do
    echo -n "[$f] "
    env $f timeout 200 python tools/step_time.py fp16c8 30 2>&1 | tail -1
  done
done
} > gpurun_out/split_ab.log 2>&1
cat gpurun_out/split_ab.log | cut -c1-400
