"""What the reference's own speed loop measures on this backend: tools/analysis_tools/benchmark.py:100-133 - one clip per
model call, `torch.cuda.synchronize()` around every call, 5 warm-up iterations - over the loader handle of
mcgaze_b200.datasets (host-decoded 300x300 frames from RAM, GPU test pipeline, forward).  Beside it: the same clips through
the batched driver.  usage: python tools/bench_loader_loop.py [clips]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from mcgaze_b200 import evaluate as ev  # noqa: E402
from mcgaze_b200.apis import init_detector  # noqa: E402
from mcgaze_b200.datasets import Gaze360Dataset, build_dataloader, build_dataset  # noqa: E402
from oracle import mcgaze_oracle as O  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfg_path = os.path.join(ROOT, 'configs/multiclue_gaze/multiclue_gaze_r50_gaze360.py')
model = init_detector(cfg_path, None, device='cuda:0')
model.load_state_dict(O.make_state_dict(0))
bank = np.random.default_rng(0).integers(0, 256, (16, 300, 300, 3), dtype=np.uint8)
Gaze360Dataset.frame_loader = staticmethod(lambda path: bank[hash(path) % 16])
anno = dict(videos=[dict(id=v + 1, file_names=[f'v{v:04d}/{t:05d}.png' for t in range(7)]) for v in range(n_clips)])
ds = build_dataset(dict(type='Gaze360Dataset', ann_file=anno, img_prefix='', pipeline=model.cfg.data.test.pipeline, clip_length=7,
                        test_mode=True, seed=0))
loader = build_dataloader(ds, samples_per_gpu=1, workers_per_gpu=0, dist=False, shuffle=False)
pure, n = 0.0, 0
for i, data in enumerate(loader):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.no_grad():
        model(return_loss=False, rescale=True, **data)
    torch.cuda.synchronize()
    if i >= 5:
        pure += time.perf_counter() - t0
        n += 1
t0 = time.perf_counter()
ev.single_gpu_test(model, ds, ds.make_pipeline(0), clips_per_batch=32, workers=4)
torch.cuda.synchronize()
batched = time.perf_counter() - t0
print(json.dumps({'clips': n_clips, 'reference_loop_clips_per_s': round(n / pure, 1), 'reference_loop_ms_per_clip': round(1e3 * pure / n, 3),
                  'batched_driver_clips_per_s_wall': round(n_clips / batched, 1),
                  'note': 'reference loop = one 7-frame clip per model call with a sync on both sides (model time only, like '
                          'benchmark.py); batched driver = 32 clips per forward, wall clock incl. host frame hand-over AND the cold start of the batch-32 plans / graph (use >= 2000 clips, or bench.py testsplit, for its steady rate)'}))
