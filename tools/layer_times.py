"""Per-layer device times of the tcgen05 GEMM launches (CUDA events, warm, eager).
usage: python tools/layer_times.py <precision> [B]      (env MCG_DEBUG_FLAGS for attribution experiments)"""
import collections
import json
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mcgaze_b200 import lib  # noqa: E402
from oracle import mcgaze_oracle as O  # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else 'fp16lo8'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
names = ['stem']
for l, n in enumerate((3, 4, 6, 3)):
    for b in range(n):
        names += [f'l{l+1}b{b}c1', f'l{l+1}b{b}c2'] + ([f'l{l+1}b{b}ds'] if b == 0 else []) + [f'l{l+1}b{b}c3']
names += ['lat3', 'lat2', 'lat1', 'lat0', 'fpn0', 'fpn1', 'fpn2', 'fpn3']
for s in range(4):
    names += [f's{s}dyn', f's{s}fc', f's{s}ffn1', f's{s}ffn2']
eng = lib.Engine(O.make_state_dict(0), 0, precision)
img = torch.randn(B * 7, 3, 224, 224, device='cuda')
out = eng.forward(img, clip_length=7)
eng.set_option('time_kernels', 1)
acc = collections.OrderedDict()
reps = 3
for _ in range(reps):
    eng.forward_into(img, 7, out)
    torch.cuda.synchronize()
    for nm, t in zip(names, eng.umma_times()):
        key = re.sub(r'b\d', '', nm)
        key = re.sub(r'^s\d', 'head_', key)
        acc[key] = acc.get(key, 0.0) + t * 1e3 / reps
print(json.dumps({'precision': precision, 'flags': os.environ.get('MCG_DEBUG_FLAGS', '0'),
                  'total_us': round(sum(acc.values())), 'layers_us': {k: round(v, 1) for k, v in acc.items()}}))
