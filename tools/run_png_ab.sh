#!/bin/bash
# same-box A/B of PNG decoder builds: tools/run_png_ab.sh <lib.so> ...
for l in "$@"; do
  echo -n "$l: "
  MCG_LIB_PATH=$PWD/$l timeout 200 python tools/bench_png.py --batches ${PNG_BATCHES:-224} 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print(' '.join('%d img: %.2f ms (%.0f img/s)' % (r['images'], r['device_ms'], r['images_per_s']) for r in d['runs']))"
done
