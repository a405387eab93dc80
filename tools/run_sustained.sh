#!/bin/bash
# sustained rate: the replayed step over ~12 s with the SM clock / power sampled beside it
mkdir -p gpurun_out
( for i in $(seq 1 40); do nvidia-smi --query-gpu=clocks.sm,power.draw,temperature.gpu,clocks_throttle_reasons.active --format=csv,noheader; sleep 0.5; done ) > gpurun_out/sustained_clocks.csv 2>&1 &
SM=$!
for n in 30 300 1200; do timeout 200 python tools/step_time.py fp16c8 $n 2>&1 | tail -1; done > gpurun_out/sustained.log 2>&1
kill $SM 2>/dev/null
cat gpurun_out/sustained.log; sort gpurun_out/sustained_clocks.csv | uniq -c | sort -rn | head -12
