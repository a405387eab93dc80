"""Device time of the bench step (32 clips x 7 x 224^2, CUDA graph replay) for A/B experiments under MCG_* env knobs.
usage: python tools/step_time.py <precision> [steps]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mcgaze_b200 import lib  # noqa: E402
from oracle import mcgaze_oracle as O  # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else 'fp16c8'
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
eng = lib.Engine(O.make_state_dict(0), 0, precision)
img = torch.randn(32 * 7, 3, 224, 224, device='cuda')
out = eng.forward(img, clip_length=7)
ref = out['gaze'].clone()
eng.set_graph_mode(True)
for _ in range(5):
    eng.forward_into(img, 7, out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    eng.forward_into(img, 7, out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
env = {k: v for k, v in os.environ.items() if k.startswith('MCG_')}
print(json.dumps({'precision': precision, 'env': env, 'ms_per_step': round(ms, 4), 'clips_per_s': round(32e3 / ms, 1),
                  'graph_equals_eager': bool(torch.equal(ref, out['gaze']))}))
