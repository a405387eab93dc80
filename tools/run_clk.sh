#!/bin/bash
mkdir -p gpurun_out
for prec in fp16c8 fp16x3 fp16; do
  echo "### $prec"
  MCG_DEBUG_FLAGS=128 timeout 300 python tools/profile_step.py $prec 32 2 2>&1 | grep -E "K=2304|M=702464 N=256 K=64|M=702464 N=64 K=256" | tail -8
done 2>&1 | tee gpurun_out/clk.log
