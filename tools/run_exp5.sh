#!/bin/bash
run() { env "$@" timeout 300 python tools/layer_times.py fp16c8 32 detail 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
k=d['kernels_us']
sel={n:v[0] for n,v in k.items() if n.startswith('bneck') }
print(json.dumps({'env':d['env'],'total':d['total_us'],'sel':sel}))"; }
run A=1
run MCG_TUNE_BF_OUT_SETS=1
run MCG_TUNE_BF_OUT_SETS=2
for e in "A=1" "MCG_TUNE_BF_OUT_SETS=1" "MCG_TUNE_BF_OUT_SETS=2"; do env $e timeout 200 python tools/step_time.py fp16c8 30 | tail -1; done
