#!/bin/bash
# Same-box A/B of library builds: tools/run_ab.sh <libA.so> <libB.so> ... (each timed twice, interleaved)
for rep in 1 2; do
  for l in "$@"; do
    echo -n "$l: "
    MCG_LIB_PATH=$PWD/$l timeout 300 python tools/step_time.py fp16c8 30 2>&1 | tail -1
  done
done
