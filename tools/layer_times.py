"""Device time of every kernel of the forward (CUDA events after each launch, warm, eager mode).
usage: python tools/layer_times.py <precision> [B] [detail]      (env MCG_* tuning / attribution flags apply)"""
import collections
import json
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mcgaze_b200 import lib  # noqa: E402
from oracle import mcgaze_oracle as O  # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else 'fp16x3'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
detail = len(sys.argv) > 3
eng = lib.Engine(O.make_state_dict(0), 0, precision)
img = torch.randn(B * 7, 3, 224, 224, device='cuda')
out = eng.forward(img, clip_length=7)
eng.set_option('time_kernels', 1)
acc = collections.OrderedDict()
cnt = collections.Counter()
reps = 3
for _ in range(reps):
    eng.forward_into(img, 7, out)
    torch.cuda.synchronize()
    for nm, ms in eng.kernel_profile():
        key = nm
        if not detail:
            key = re.sub(r'^umma:l(\d)b\d+', r'umma:l\1', key)
            key = re.sub(r'^umma:s\d', 'umma:head_', key)
        acc[key] = acc.get(key, 0.0) + ms * 1e3 / reps
        cnt[key] += 1
env = {k: v for k, v in os.environ.items() if k.startswith('MCG_')}
gemm = sum(v for k, v in acc.items() if k.startswith('umma:'))
print(json.dumps({'precision': precision, 'env': env, 'total_us': round(sum(acc.values())), 'gemm_us': round(gemm),
                  'other_us': round(sum(acc.values()) - gemm),
                  'kernels_us': {k: [round(v, 1), cnt[k] // reps] for k, v in acc.items()}}))
