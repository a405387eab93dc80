"""The reference's test-time image pipeline on the GPU.

`GpuTestPipeline(cfg.data.test.pipeline)` takes the SAME list of `dict(type=...)` entries the reference feeds to
`mmdet.datasets.pipelines.Compose` (tools/test_gaze360_gaze.py:58; configs/_base_/datasets/gaze360.py:27-36,
configs/multiclue_gaze/multiclue_gaze_r50_l2cs.py:31-39) and produces the same `img` / `img_metas` the reference's
pipeline + collate produce, but the pixel work (resize, colour swap, normalisation, padding, HWC -> CHW, batching)
runs in one CUDA kernel behind `mcg_preprocess` on uint8 frames already in HBM: 1 byte per sample crosses PCIe
instead of the 4 of the reference's fp32 tensors.

Host side (this file) = only the per-frame geometry, in the reference's own python arithmetic:
  CenterCrop._get_crop_size / _crop_data   mmdet/datasets/pipelines/transforms.py:1101-1130, 1036-1047
  Resize._resize_img (mmcv.imrescale / imresize, rescale_size)   transforms.py:213-241
  Pad._pad_img (mmcv.impad_to_multiple / impad)                  transforms.py:665-681
The crop of the reference's test run is random (one np.random.rand(1) per pipeline call, i.e. per frame in
tools/test_gaze360_gaze.py:96-100); here the draws come from `np.random` as well (so np.random.seed pins them), from
an own seeded generator (`seed=`), or from the caller (`rands=`).

There is no CPU fallback: without the CUDA library / a GPU the calls raise.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import lib

_SKIPPED = ('LoadImageFromFile', 'DefaultFormatBundle', 'Collect', 'ImageToTensor')


class GpuTestPipeline:
    def __init__(self, pipeline_cfg: Sequence[Dict[str, Any]], device: int = 0, seed: Optional[int] = None):
        self.device = int(device)
        self.crop: Optional[Tuple[str, Tuple[float, float]]] = None
        self.scale: Optional[Tuple[int, int]] = None
        self.keep_ratio = True
        self.mean = self.std = None
        self.to_rgb = True
        self.size_divisor: Optional[int] = None
        self.pad_size: Optional[Tuple[int, int]] = None
        self._rng = np.random.RandomState(seed) if seed is not None else None
        for step in pipeline_cfg:
            step = dict(step)
            t = step.pop('type')
            if t in _SKIPPED:
                continue                      # frames arrive decoded; bundling / collecting is what batch() returns
            if t == 'CenterCrop':
                ctype = step.get('crop_type', 'absolute')
                if ctype not in ('relative_range', 'relative', 'absolute'):
                    raise NotImplementedError(f'CenterCrop crop_type={ctype!r} is not supported by the GPU pipeline')
                self.crop = (ctype, tuple(step['crop_size']))
            elif t == 'Resize':
                scale = step.get('img_scale')
                if isinstance(scale, list):
                    if len(scale) != 1:
                        raise NotImplementedError('multi-scale Resize is a training-time feature')
                    scale = scale[0]
                if scale is None or step.get('ratio_range') is not None:
                    raise NotImplementedError('Resize needs one fixed img_scale')
                if step.get('backend', 'cv2') != 'cv2' or step.get('interpolation', 'bilinear') != 'bilinear':
                    raise NotImplementedError('only cv2 bilinear Resize is implemented on the GPU')
                self.scale = (int(scale[0]), int(scale[1]))
                self.keep_ratio = bool(step.get('keep_ratio', True))
            elif t == 'RandomFlip':
                if step.get('flip_ratio') not in (None, 0, 0.0):
                    raise NotImplementedError('test pipelines use RandomFlip(flip_ratio=0.0); flipping is not implemented')
            elif t == 'Normalize':
                # Normalize.__init__ stores float32 arrays (transforms.py:735-736)
                self.mean = np.array(step['mean'], dtype=np.float32)
                self.std = np.array(step['std'], dtype=np.float32)
                self.to_rgb = bool(step.get('to_rgb', True))
            elif t == 'Pad':
                if step.get('pad_to_square') or step.get('pad_val', 0) not in (0, dict(img=0, masks=0, seg=255)):
                    raise NotImplementedError('only zero padding to a size / size_divisor is implemented')
                self.size_divisor = step.get('size_divisor')
                self.pad_size = tuple(step['size']) if step.get('size') is not None else None
            else:
                raise NotImplementedError(f'pipeline step {t!r} is not part of the reference test pipelines')
        if self.scale is None or self.mean is None:
            raise ValueError('the GPU pipeline needs a Resize and a Normalize step')

    # ---------------------------------------------------------------------------- geometry (host, python floats)
    def _draw(self) -> float:
        return float((self._rng.rand(1) if self._rng is not None else np.random.rand(1))[0])

    def draw(self, n: int) -> np.ndarray:
        """The CenterCrop draws of n frames, in frame order (zeros when the pipeline has no random crop): n draws of
        rand(1) and one draw of rand(n) consume the generator identically."""
        if self.crop is not None and self.crop[0] == 'relative_range':
            return (self._rng if self._rng is not None else np.random).rand(n)
        return np.zeros(n)

    def crop_window(self, h: int, w: int, rand: float) -> Tuple[int, int, int, int]:
        """-> (y, x, crop_h, crop_w) of the CenterCrop slice, clipped to the image like the numpy slice is."""
        if self.crop is None:
            return 0, 0, h, w
        ctype, size = self.crop
        if ctype == 'absolute':
            ch, cw = min(size[0], h), min(size[1], w)
        elif ctype == 'relative':
            ch, cw = int(h * size[0] + 0.5), int(w * size[1] + 0.5)
        else:   # relative_range: ONE draw for both sides (transforms.py:1127-1130), float32 ratio + float64 draw
            cs = np.asarray(size, dtype=np.float32)
            rh, rw = cs + np.asarray([rand], dtype=np.float64) * (1 - cs)
            ch, cw = int(h * rh + 0.5), int(w * rw + 0.5)
        if ch <= 0 or cw <= 0:
            raise ValueError('CenterCrop produced an empty window')
        y = int(max(h - ch, 0) / 2 + 0.5)
        x = int(max(w - cw, 0) / 2 + 0.5)
        return y, x, min(y + ch, h) - y, min(x + cw, w) - x

    def resized_size(self, h: int, w: int) -> Tuple[int, int]:
        """-> (new_h, new_w) of Resize (mmcv.rescale_size when keep_ratio, else img_scale = (w, h))."""
        if not self.keep_ratio:
            return self.scale[1], self.scale[0]
        f = min(max(self.scale) / max(h, w), min(self.scale) / min(h, w))
        return int(h * float(f) + 0.5), int(w * float(f) + 0.5)

    def padded_size(self, h: int, w: int) -> Tuple[int, int]:
        if self.pad_size is not None:
            return max(self.pad_size[0], h), max(self.pad_size[1], w)
        if self.size_divisor:
            d = self.size_divisor
            return int(np.ceil(h / d)) * d, int(np.ceil(w / d)) * d
        return h, w

    def clip_canvases(self, shapes: Sequence[Tuple[int, int]], rands: Sequence[float], T: int) -> List[Tuple[int, int]]:
        """Padded canvas (Hp, Wp) of every T-frame clip of a frame list: what the reference's per-clip collate pads the
        clip's frames to (tools/test_gaze360_gaze.py:102-105)."""
        ph, pw = self._arrays(shapes, rands)[-2:]
        Hp = ph.reshape(-1, T).max(1)
        Wp = pw.reshape(-1, T).max(1)
        Wp = Wp + (-Wp) % 4
        return list(zip(Hp.tolist(), Wp.tolist()))

    def plan(self, shapes: Sequence[Tuple[int, int]], rands: Optional[Sequence[float]] = None):
        """Per-frame geometry + metas for frames of the given (h, w); -> (geometry, metas, (Hp, Wp)).
        Vectorised over the frames in float64 / int64 numpy: the same IEEE operations, in the same order, as the
        scalar methods above (`plan_scalar` keeps the literal per-frame form; the tests compare the two)."""
        n = len(shapes)
        h, w, y, x, ch, cw, nh, nw, ph, pw = self._arrays(shapes, rands)
        Hp, Wp = int(ph.max()), int(pw.max())
        geometry = np.stack([y, x, ch, cw, nh, nw], 1)
        ws, hs = nw / cw, nh / ch
        scale = np.stack([ws, hs, ws, hs], 1).astype(np.float32)
        norm_cfg = dict(mean=self.mean, std=self.std, to_rgb=self.to_rgb)
        hl, wl, nhl, nwl, phl, pwl = (a.tolist() for a in (h, w, nh, nw, ph, pw))
        metas = [dict(filename=None, ori_filename=None, ori_shape=(hl[i], wl[i], 3), img_shape=(nhl[i], nwl[i], 3),
                      pad_shape=(phl[i], pwl[i], 3), scale_factor=scale[i], flip=False, flip_direction=None,
                      img_norm_cfg=norm_cfg) for i in range(n)]
        if Wp % 4:
            Wp += 4 - Wp % 4                   # the kernel stores float4; configs pad to 32 anyway
        return list(map(tuple, geometry.tolist())), metas, (Hp, Wp)

    def _arrays(self, shapes: Sequence[Tuple[int, int]], rands: Optional[Sequence[float]] = None):
        """-> int64 arrays (h, w, crop_y, crop_x, crop_h, crop_w, new_h, new_w, pad_h, pad_w), one entry per frame."""
        n = len(shapes)
        if rands is None:
            rands = self.draw(n)
        hw = np.asarray(shapes, dtype=np.int64).reshape(n, 2)
        h, w = hw[:, 0], hw[:, 1]
        r = np.asarray(rands, dtype=np.float64).reshape(n)
        if self.crop is None:
            y = x = np.zeros(n, np.int64)
            ch, cw = h, w
        else:
            ctype, size = self.crop
            if ctype == 'absolute':
                ch, cw = np.minimum(int(size[0]), h), np.minimum(int(size[1]), w)
            elif ctype == 'relative':
                ch, cw = (h * float(size[0]) + 0.5).astype(np.int64), (w * float(size[1]) + 0.5).astype(np.int64)
            else:
                cs = np.asarray(size, dtype=np.float32)
                rest = (1 - cs).astype(np.float64)                       # the subtraction happens in float32
                ch = (h * (np.float64(cs[0]) + r * rest[0]) + 0.5).astype(np.int64)
                cw = (w * (np.float64(cs[1]) + r * rest[1]) + 0.5).astype(np.int64)
            if np.any(ch <= 0) or np.any(cw <= 0):
                raise ValueError('CenterCrop produced an empty window')
            y = (np.maximum(h - ch, 0) / 2 + 0.5).astype(np.int64)
            x = (np.maximum(w - cw, 0) / 2 + 0.5).astype(np.int64)
            ch, cw = np.minimum(y + ch, h) - y, np.minimum(x + cw, w) - x
        if self.keep_ratio:
            f = np.minimum(max(self.scale) / np.maximum(ch, cw), min(self.scale) / np.minimum(ch, cw))
            nh, nw = (ch * f + 0.5).astype(np.int64), (cw * f + 0.5).astype(np.int64)
        else:
            nh, nw = np.full(n, self.scale[1], np.int64), np.full(n, self.scale[0], np.int64)
        if self.pad_size is not None:
            ph, pw = np.maximum(self.pad_size[0], nh), np.maximum(self.pad_size[1], nw)
        elif self.size_divisor:
            d = self.size_divisor
            ph, pw = np.ceil(nh / d).astype(np.int64) * d, np.ceil(nw / d).astype(np.int64) * d
        else:
            ph, pw = nh, nw
        return h, w, y, x, ch, cw, nh, nw, ph, pw

    def plan_scalar(self, shapes: Sequence[Tuple[int, int]], rands: Sequence[float]):
        """The literal per-frame form of plan() (python floats, as the reference's transforms compute)."""
        geometry, metas = [], []
        Hp = Wp = 0
        for (h, w), r in zip(shapes, rands):
            y, x, ch, cw = self.crop_window(h, w, r)
            nh, nw = self.resized_size(ch, cw)
            ph, pw = self.padded_size(nh, nw)
            Hp, Wp = max(Hp, ph), max(Wp, pw)
            geometry.append((y, x, ch, cw, nh, nw))
            w_scale, h_scale = nw / cw, nh / ch
            metas.append(dict(
                filename=None, ori_filename=None, ori_shape=(h, w, 3), img_shape=(nh, nw, 3), pad_shape=(ph, pw, 3),
                scale_factor=np.array([w_scale, h_scale, w_scale, h_scale], dtype=np.float32),
                flip=False, flip_direction=None,
                img_norm_cfg=dict(mean=self.mean, std=self.std, to_rgb=self.to_rgb)))
        if Wp % 4:
            Wp += 4 - Wp % 4
        return geometry, metas, (Hp, Wp)

    # ---------------------------------------------------------------------------- device work
    def batch(self, frames: Sequence[Any], rands: Optional[Sequence[float]] = None, out=None,
              filenames: Optional[Sequence[str]] = None) -> Dict[str, Any]:
        """frames: decoded BGR uint8 [h, w, 3] arrays (numpy -> copied to the device) or CUDA uint8 tensors, or one
        [n, h, w, 3] array / CUDA tensor when all frames have one size.
        -> dict(img=[Tensor[n, 3, Hp, Wp]], img_metas=[[meta] * n]) ready for `model(return_loss=False, **data)`:
        the collate of the reference (tools/test_gaze360_gaze.py:102-105) pads every frame to the largest one."""
        import torch
        if not torch.cuda.is_available():
            raise lib.McgError('mcgaze_b200 needs a CUDA device (B200, sm_100a); there is no CPU path')
        dev = torch.device('cuda', self.device)
        if hasattr(frames, 'data_ptr') or (isinstance(frames, np.ndarray) and frames.ndim == 4):
            # one [n, h, w, 3] block (frames of one size): a single H2D copy, descriptors without a python loop
            dframes = torch.from_numpy(np.ascontiguousarray(frames)) if isinstance(frames, np.ndarray) else frames
            if not dframes.is_cuda:
                dframes = dframes.to(dev, non_blocking=True)       # asynchronous when the block is pinned
            shapes = [(int(dframes.shape[1]), int(dframes.shape[2]))] * int(dframes.shape[0])
        else:
            dframes = []
            for f in frames:
                if isinstance(f, np.ndarray):
                    if f.dtype != np.uint8 or f.ndim != 3 or f.shape[2] != 3:
                        raise ValueError('frames must be uint8 [h, w, 3] (BGR as decoded by LoadImageFromFile)')
                    f = torch.from_numpy(np.ascontiguousarray(f)).to(dev, non_blocking=True)
                dframes.append(f)
            shapes = [(int(f.shape[0]), int(f.shape[1])) for f in dframes]
        geometry, metas, (Hp, Wp) = self.plan(shapes, rands)
        if filenames is not None:
            for m, name in zip(metas, filenames):
                m['filename'] = m['ori_filename'] = name
        n = len(dframes)
        if out is None:
            out = torch.empty((n, 3, Hp, Wp), dtype=torch.float32, device=dev)
        elif tuple(out.shape) != (n, 3, Hp, Wp):
            raise ValueError(f'out must have shape {(n, 3, Hp, Wp)}')
        with torch.cuda.device(dev):
            lib.preprocess(dframes, geometry, self.mean, self.std, self.to_rgb, out)
        return dict(img=[out], img_metas=[metas])

    def __call__(self, data: Dict[str, Any]) -> Dict[str, Any]:
        """Per-frame call of the reference's Compose (tools/test_gaze360_gaze.py:48-49): `data['img']` holds the
        decoded frame.  -> dict(img=Tensor[3, Hp, Wp] on the device, img_metas=meta)."""
        if 'img' not in data:
            raise KeyError("the GPU pipeline takes decoded frames: pass dict(img=uint8 BGR array)")
        name = (data.get('img_info') or {}).get('filename')
        res = self.batch([data['img']], filenames=[name] if name is not None else None)
        return dict(img=res['img'][0][0], img_metas=res['img_metas'][0][0])
