#!/bin/bash
# same-box A/B of env knobs on the current build: tools/run_flags2.sh "VAR=1" "" ... (each timed twice, interleaved)
for rep in 1 2; do
  for f in "$@"; do
    echo -n "[$f] "
    env $f timeout 300 python tools/step_time.py fp16c8 30 2>&1 | tail -1
  done
done
