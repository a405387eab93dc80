#!/bin/bash
# in-kernel wait counters of selected GEMM launches (needs a -DMCG_KERNEL_DEBUG=1 build: MCG_LIB_PATH)
mkdir -p gpurun_out
for spec in "$@"; do
  echo "### $spec"
  env $spec timeout 300 python tools/profile_step.py ${PREC:-fp16c8} 32 1 2>&1 | grep -E "^umma|^prod" | grep -E "${PAT:-M=702464 N=64 K=576|M=702464 N=256 K=2304|M=702464 N=256 K=64|M=702464 N=64 K=256|M=175616 N=128 K=1152}" | tail -12
done 2>&1 | tee gpurun_out/clk.log
