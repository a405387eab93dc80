#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_forward.py -x -q 2>&1 | tail -5
echo "=== layer times"
for prec in ${PRECS:-fp16c8}; do timeout 300 python tools/layer_times.py $prec 32 detail; done
} > gpurun_out/c8.log 2>&1
python - <<'PY'
import json
for line in open('gpurun_out/c8.log'):
    if line.startswith('{'):
        d = json.loads(line); k = d['kernels_us']
        print(d['precision'], 'total', d['total_us'], 'gemm', d['gemm_us'], 'other', d['other_us'])
        print('  ' + ' '.join(f"{x.split(':')[-1]}={v[0]:.0f}" for x, v in k.items() if x.startswith('umma:l') or x.startswith('umma:f') or x.startswith('stem')))
    elif not line.startswith('='):
        print(line.rstrip()[:300])
PY
