#!/bin/bash
mkdir -p gpurun_out
run() { env "$@" timeout 300 python tools/layer_times.py fp16c8 32 detail 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
k=d['kernels_us']
sel={n:v[0] for n,v in k.items() if n in ('umma:lat0','umma:lat1','umma:lat2','umma:l0b1c3','umma:l1b1c3','umma:l2b1c3','umma:l0b1c1','umma:l1b1c1','umma:l0b0c3ds')}
print(json.dumps({'env':d['env'],'total':d['total_us'],'sel':sel}))"; }
run A=1
run MCG_TUNE_NO_RES_UP=1
run MCG_TUNE_RES_BN=128
run MCG_TUNE_OUT_SETS=1
run MCG_TUNE_RES_BN=128 MCG_TUNE_NO_RES_UP=1
run MCG_TUNE_KBS=1
