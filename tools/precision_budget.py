"""Where does the (yaw, pitch) error budget go?  CPU-only: rounds the operands of chosen convolutions of the oracle
to fp16 and prints max |d(yaw,pitch)| against the fp64 oracle (dead ends: per-layer mixed precision)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mcgaze_oracle as O
torch.set_num_threads(8)
sd = O.make_state_dict(0)
clips = [O.make_clip(s, 7) for s in (1, 2, 3)]

def r16(t): return t.half().float()

class LayerHooks(O.Hooks):
    """quantise operands of the convs whose index is in `sel` (order of _conv calls)"""
    def __init__(self, sel):
        super().__init__()
        self.sel = sel; self.i = -1; self.on = False
        self.quant = self._q; self.quant_w = self._qw; self.quant_head = None
    def _q(self, t):
        self.i += 1
        self.on = self.i in self.sel
        return r16(t) if self.on else t
    def _qw(self, t):
        return r16(t) if self.on else t
    def qh(self, t): return t

# conv order: 0 stem; per block conv1,conv2,conv3,(ds); then lateral0-3, fpn0-3
names = ['stem']
for li, nb in enumerate(O.STAGE_BLOCKS):
    for b in range(nb):
        for c in ('c1', 'c2', 'c3'):
            names.append(f'l{li+1}.{b}.{c}')
        if b == 0: names.append(f'l{li+1}.{b}.ds')
names += [f'lat{i}' for i in range(4)] + [f'fpn{i}' for i in range(4)]
print(len(names))
sd64 = {k: v.double() for k, v in sd.items()}
t0 = time.time()
ref = [O.forward(sd64, c.double()) for c in clips]
print('ref', time.time() - t0)
keys = [k for k in ref[0] if 'gaze' in k]
print(keys)
def err(sel):
    e = 0.0
    for c, r in zip(clips, ref):
        hk = LayerHooks(sel)
        o = O.forward(sd, c, hk=hk)
        for k in keys:
            e = max(e, (O.vector_to_yaw_pitch(o[k].double()) - O.vector_to_yaw_pitch(r[k])).abs().max().item())
    return e
groups = {
 'none': set(),
 'all': set(range(len(names))),
 'stem': {0},
 'layer1': {i for i, n in enumerate(names) if n.startswith('l1')},
 'layer2': {i for i, n in enumerate(names) if n.startswith('l2')},
 'layer3': {i for i, n in enumerate(names) if n.startswith('l3')},
 'layer4': {i for i, n in enumerate(names) if n.startswith('l4')},
 'lat': {i for i, n in enumerate(names) if n.startswith('lat')},
 'fpn': {i for i, n in enumerate(names) if n.startswith('fpn')},
 'c2 (3x3)': {i for i, n in enumerate(names) if n.endswith('c2')},
 'c1': {i for i, n in enumerate(names) if n.endswith('c1')},
 'c3': {i for i, n in enumerate(names) if n.endswith('c3')},
 'fpn0': {names.index('fpn0')}, 'fpn1': {names.index('fpn1')}, 'fpn2': {names.index('fpn2')}, 'fpn3': {names.index('fpn3')},
}
for g, sel in groups.items():
    t0 = time.time()
    print(f'{g:10s} n={len(sel):3d} err={err(sel):.3e}  ({time.time()-t0:.1f}s)', flush=True)
