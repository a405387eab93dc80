#!/bin/bash
# ncu --set full captures of the launches round 2 had not profiled yet: layer3.1 / layer4.1 GEMMs (second forward of
# tools/profile_step.py: 70 umma launches per forward, layer3.1 conv1 = launch 84, layer4.1 conv1 = launch 102 of the process),
# the fused stem and the head's kernels.  The reports stay on the box (a full set is > 64 MiB); the raw pages come back as CSV.
mkdir -p gpurun_out
R=/tmp/ncu4; mkdir -p $R
cap() {  # name, kernel regex, skip, count, extra
  timeout 400 ncu --set full --clock-control none $5 -k "regex:$2" -s $3 -c $4 -o $R/$1 -f python tools/profile_step.py fp16c8 32 1 > $R/$1.log 2>&1
  tail -1 $R/$1.log
  ncu -i $R/$1.ncu-rep --page raw --csv > gpurun_out/ncu_r02_$1_raw.csv 2>/dev/null
}
cap layer3 umma_gemm_kernel 84 3
cap layer4 umma_gemm_kernel 102 3
cap stem stem_fused_kernel 1 1 "--import-source on"
ncu -i $R/stem.ncu-rep --page source --csv > gpurun_out/ncu_r02_stem_source.csv 2>/dev/null
cap head 'dynconv_mma_kernel|linear256_ln_kernel|mlp_chain_kernel|roi_align_kernel|layernorm_kernel' 24 8
ls -la gpurun_out/
