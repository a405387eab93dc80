#!/bin/bash
timeout 200 python -m pytest tests/test_gpu_forward.py -x -q -m gpu -k "single_clip or headline or intermediates" 2>&1 | tail -3
for e in "A=1" "MCG_TUNE_HEAD_PAIR=0"; do env $e timeout 200 python tools/step_time.py fp16c8 40 | tail -1; done
for e in "A=1" "MCG_TUNE_HEAD_PAIR=0"; do env $e timeout 300 python tools/layer_times.py fp16c8 32 detail 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels_us']
print(json.dumps({'env':d['env'],'total':d['total_us'],'dyn':[k['umma:s%ddyn'%i][0] for i in range(4)], 'dynconv':k['dynconv_mma_kernel']}))"; done
