"""TEST INFRASTRUCTURE ONLY — generates tests/golden/golden_preprocess.npz by running the REFERENCE's own
transform classes (mmdet/datasets/pipelines/transforms.py: CenterCrop, Resize, RandomFlip, Normalize, Pad),
built from the reference's own config dicts (configs/_base_/datasets/gaze360.py:27-36,
configs/multiclue_gaze/multiclue_gaze_r50_l2cs.py:31-39), on synthetic decoded frames.

    python oracle/gen_golden_preprocess.py        # needs /root/reference and cv2

mmcv's image functions come from oracle/refshim.py (published 1.4.8 wrappers over cv2), the pixel arithmetic
is OpenCV's (cv2 4.13 in this image).  LoadImageFromFile (PNG decode) and DefaultFormatBundle / Collect (DataContainer
wrapping) are not run: the frame is handed over decoded, and the HWC -> CHW transpose of formatting.py:96 is applied
here.  The random draw of CenterCrop is pinned by seeding numpy and recording np.random.rand(1) per frame.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')

# name, pipeline ('gaze360' | 'l2cs'), source (h, w), seed
CASES = [
    ('gaze360_120x100', 'gaze360', (120, 100), 11),       # upscale, portrait -> padded width
    ('gaze360_231x187', 'gaze360', (231, 187), 12),
    ('gaze360_480x640', 'gaze360', (480, 640), 13),       # landscape, downscale > 2x
    ('gaze360_448x448', 'gaze360', (448, 448), 14),
    ('gaze360_33x57', 'gaze360', (33, 57), 15),           # tiny source, big upscale
    ('l2cs_150x130', 'l2cs', (150, 130), 16),             # no crop, 448 target
]


def synthetic_frame(h: int, w: int, seed: int) -> np.ndarray:
    """uint8 BGR frame: smooth gradients + noise (every rounding case gets exercised)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([(xx * 255 // max(w - 1, 1)), (yy * 255 // max(h - 1, 1)), ((xx + yy) * 255 // max(h + w - 2, 1))], -1)
    noise = rng.integers(-96, 97, (h, w, 3))
    return np.clip(base + noise, 0, 255).astype(np.uint8)


def reference_pipeline(name: str):
    refshim.install_image_ops()
    from mmcv import Config
    from mmdet.datasets.pipelines import transforms as T
    cfg_path = {'gaze360': 'configs/multiclue_gaze/multiclue_gaze_r50_gaze360.py',
                'l2cs': 'configs/multiclue_gaze/multiclue_gaze_r50_l2cs.py'}[name]
    cfg = Config.fromfile(os.path.join(refshim.REFERENCE_ROOT, cfg_path))
    steps = []
    for d in cfg.data.test.pipeline:
        d = dict(d)
        t = d.pop('type')
        if t in ('LoadImageFromFile', 'DefaultFormatBundle', 'Collect'):
            continue
        steps.append(getattr(T, t)(**d))
    return steps


def main() -> None:
    out = {}
    for name, pipe, (h, w), seed in CASES:
        img = synthetic_frame(h, w, seed)
        steps = reference_pipeline(pipe)
        np.random.seed(seed)
        state = np.random.get_state()
        rand = float(np.random.rand(1)[0])          # the draw CenterCrop._get_crop_size will make (transforms.py:1129)
        np.random.set_state(state)
        results = dict(img=img.copy(), img_shape=img.shape, ori_shape=img.shape, img_fields=['img'],
                       bbox_fields=[], mask_fields=[], seg_fields=[])
        for s in steps:
            results = s(results)
        assert results['flip'] is False or not results['flip']
        chw = np.ascontiguousarray(results['img'].transpose(2, 0, 1))
        out[name + '.src'] = img
        out[name + '.rand'] = np.float64(rand)
        out[name + '.img'] = chw.astype(np.float32)
        out[name + '.img_shape'] = np.asarray(results['img_shape'], dtype=np.int64)
        out[name + '.pad_shape'] = np.asarray(results['pad_shape'], dtype=np.int64)
        out[name + '.scale_factor'] = np.asarray(results['scale_factor'], dtype=np.float32)
        print(name, img.shape, '->', results['img_shape'], results['pad_shape'], results['scale_factor'], 'rand', rand)
    np.savez_compressed(os.path.join(GOLD, 'golden_preprocess.npz'), **out)
    print('wrote', os.path.join(GOLD, 'golden_preprocess.npz'),
          os.path.getsize(os.path.join(GOLD, 'golden_preprocess.npz')) // 1024, 'KiB')


if __name__ == '__main__':
    main()
