#!/bin/bash
for e in "A=1" "MCG_TUNE_RES_BN=128" "MCG_TUNE_NO_RES_UP=1" "MCG_TUNE_RES_BN=128 MCG_TUNE_NO_RES_UP=1"; do
  env $e timeout 300 python tools/step_time.py fp16c8 40 2>&1 | tail -1
done
