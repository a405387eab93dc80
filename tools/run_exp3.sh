#!/bin/bash
# Round 2, last session: (1) the new generic-tester GPU test, (2) same-box A/B of the epilogue staging knobs of the
# residual layers (ring stages vs output staging sets vs identity prefetch depth): per-layer eager times + replayed step.
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_generic_tester.py tests/test_reference_tools.py -x -q -m gpu 2>&1 | tail -5
for f in "" "MCG_TUNE_RES_BUFS=1" "MCG_TUNE_OUT_SETS=1" "MCG_TUNE_RES_BUFS=1 MCG_TUNE_OUT_SETS=1"; do
  echo "=== layer_times [$f]"
  env $f timeout 300 python tools/layer_times.py fp16c8 32 detail 2>&1 | tail -1
done
for rep in 1 2; do
  for f in "" "MCG_TUNE_RES_BUFS=1" "MCG_TUNE_OUT_SETS=1" "MCG_TUNE_RES_BUFS=1 MCG_TUNE_OUT_SETS=1"; do
    echo -n "[$f] "
    env $f timeout 300 python tools/step_time.py fp16c8 30 2>&1 | tail -1
  done
done
} > gpurun_out/exp3.log 2>&1
tail -12 gpurun_out/exp3.log
