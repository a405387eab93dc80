import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), 'tests'))
import cv2, numpy as np, torch
import png_cases
from mcgaze_b200 import evaluate as ev
from mcgaze_b200.apis import init_detector
from mcgaze_b200.pipeline import GpuTestPipeline
from oracle import mcgaze_oracle as O
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
model = init_detector(os.path.join(ROOT, 'configs/multiclue_gaze/multiclue_gaze_r50_gaze360.py'), None, 'cuda:0')
model.load_state_dict(O.make_state_dict(0))
tmp = tempfile.mkdtemp()
videos = []
for vi, L in enumerate([7, 10, 4]):
    names = []
    size = (120, 100) if vi != 1 else (90, 128)
    for t in range(L):
        name = f'v{vi:03d}/{t:05d}.png'
        os.makedirs(os.path.join(tmp, os.path.dirname(name)), exist_ok=True)
        cv2.imwrite(os.path.join(tmp, name), png_cases._natural(size[0], size[1], 100 * vi + t))
        names.append(name)
    videos.append(dict(id=vi + 1, file_names=names))
anno = dict(videos=videos, annotations=[])
log = {}
class Spy(GpuTestPipeline):
    def batch(self, frames, **kw):
        fr = frames if hasattr(frames, 'shape') else list(frames)
        host = fr.cpu().numpy().copy() if hasattr(fr, 'cpu') else [np.asarray(f.cpu().numpy() if hasattr(f, 'cpu') else f).copy() for f in fr]
        out = super().batch(frames, **kw)
        log.setdefault(self.tag, []).append((host, kw.get('rands'), out['img'][0].cpu().numpy().copy()))
        return out
class Lazy(GpuTestPipeline):
    def batch(self, frames, **kw):
        tc = lambda f: f.clone() if hasattr(f, 'clone') else torch.from_numpy(np.ascontiguousarray(f)).cuda()
        fr = tc(frames) if hasattr(frames, 'shape') else [tc(f) for f in frames]
        out = super().batch(frames, **kw)
        log.setdefault(self.tag, []).append((fr, kw.get('rands'), out['img'][0].clone(), out['img'][0].data_ptr(),
                                             frames.data_ptr() if hasattr(frames, 'data_ptr') else -1))
        return out
def run(mode, workers, kind):
    ds = ev.Gaze360ClipDataset(anno, img_prefix=tmp, decode=mode)
    pipe = kind(model.cfg.data.test.pipeline, seed=3)
    pipe.tag = mode
    return ev.single_gpu_test(model, ds, pipe, clips_per_batch=2, workers=workers)
ref = run('host', 0, Spy)
log.clear()
h = run('host', 0, Lazy)
g = run('gpu', 0, Lazy)
torch.cuda.synchronize()
print('host == ref', [bool(np.array_equal(a, b)) for a, b in zip(h, ref)])
print('gpu == ref', [bool(np.array_equal(a, b)) for a, b in zip(g, ref)])
for k, (a, b) in enumerate(zip(log['host'], log['gpu'])):
    fa = a[0] if isinstance(a[0], list) else list(a[0])
    fb = b[0] if isinstance(b[0], list) else list(b[0])
    print(' batch', k, 'frames equal', all(bool(torch.equal(x.cuda(), y.cuda())) for x, y in zip(fa, fb)), 'img equal', bool(torch.equal(a[2], b[2])),
          'img ptr', hex(a[3]), hex(b[3]), 'frames ptr', hex(a[4]), hex(b[4]), tuple(a[2].shape))
d = np.abs(g[2] - ref[2])
print('clip 2 per-frame max diff', d.max(axis=1))
