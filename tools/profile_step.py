"""Run N forwards of the bench workload (for ncu launch lists / captures).
usage: python tools/profile_step.py <precision> [B] [steps] [graph]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mcgaze_b200 import lib  # noqa: E402
from oracle import mcgaze_oracle as O  # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else 'fp16x3'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
graph = int(sys.argv[4]) if len(sys.argv) > 4 else 0
sd = O.make_state_dict(0)
eng = lib.Engine(sd, 0, precision)
img = torch.randn(B * 7, 3, 224, 224, device='cuda')
out = eng.forward(img, clip_length=7)
eng.set_graph_mode(bool(graph))
torch.cuda.synchronize()
for _ in range(steps):
    eng.forward_into(img, 7, out)
torch.cuda.synchronize()
print('launches per forward', eng.last_launch_count)
