#!/bin/bash
# attribution experiments: layer times of one precision under MCG_DEBUG_FLAGS / tuning variants
mkdir -p gpurun_out
PREC=${1:-fp16c8}
shift
: > gpurun_out/flags.log
for spec in "$@"; do
  echo "### $spec" >> gpurun_out/flags.log
  env $spec timeout 300 python tools/layer_times.py $PREC 32 detail >> gpurun_out/flags.log 2>&1
done
python - <<'PY'
import json
keys = ['umma:fpn0','umma:fpn1','umma:l0b1c1','umma:l0b1c2','umma:l0b1c3','umma:l1b1c1','umma:l1b1c2','umma:l1b1c3','umma:l2b1c1','umma:l2b1c2','umma:l2b1c3','umma:l3b1c1','umma:l3b1c2','umma:l3b1c3','umma:lat0','umma:l0b0ds']
spec = None
for line in open('gpurun_out/flags.log'):
    if line.startswith('###'):
        spec = line.strip(); continue
    if line.startswith('{'):
        d = json.loads(line)
        k = d['kernels_us']
        print(spec, 'total', d['total_us'], 'gemm', d['gemm_us'], ' '.join(f"{x.split(':')[1]}={k[x][0]:.0f}" for x in keys if x in k))
    elif 'Error' in line or 'error' in line:
        print(spec, line.strip()[:200])
PY
