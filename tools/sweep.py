"""BASELINE configs[4] / configs[2] as throughput side measurements (not bench lines): clip length T in {1,3,7,15} x
resolution {224,320} and the l2cs setting (448 x 448, T = 7), device-resident inputs, CUDA-graph replay.
usage: python tools/sweep.py [precision]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mcgaze_b200 import lib  # noqa: E402
from oracle import mcgaze_oracle as O  # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else 'fp16c8'
eng = lib.Engine(O.make_state_dict(0), 0, precision)
GFLOP = {(224, 1): 14.22, (224, 3): 42.66, (224, 7): 99.55, (224, 15): 213.32, (320, 1): 28.64, (320, 3): 85.93,
         (320, 7): 200.51, (320, 15): 429.67, (448, 7): 390.57}       # SURVEY 8d, per clip
rows = []
for res, T, frames in [(224, 1, 224), (224, 3, 222), (224, 7, 224), (224, 15, 225), (320, 1, 112), (320, 3, 111),
                       (320, 7, 112), (320, 15, 105), (448, 7, 56)]:
    B = frames // T
    img = torch.randn(B * T, 3, res, res, device='cuda')
    out = eng.forward(img, clip_length=T)
    eng.set_graph_mode(True)
    for _ in range(3):
        eng.forward_into(img, T, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    steps = 10
    for _ in range(steps):
        eng.forward_into(img, T, out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    eng.set_graph_mode(False)
    cps = B / (ms * 1e-3)
    rows.append({'resolution': res, 'T': T, 'clips_per_step': B, 'ms_per_step': round(ms, 3), 'clips_per_s': round(cps, 1),
                 'frames_per_s': round(cps * T, 1), 'algorithmic_tflops': round(cps * GFLOP[(res, T)] / 1e3, 1)})
    del img, out
    torch.cuda.empty_cache()
print(json.dumps({'precision': precision, 'sweep': rows}))
