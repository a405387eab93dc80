"""CPU oracle for the test-time image pipeline that feeds the MCGaze forward.

TEST INFRASTRUCTURE ONLY.  Nothing under ``mcgaze_b200/`` may import this module.

numpy restatement (integer / float32 arithmetic spelled out, no cv2) of what the reference's
``test_pipeline`` does to one decoded frame (configs/_base_/datasets/gaze360.py:27-36,
configs/multiclue_gaze/multiclue_gaze_r50_l2cs.py:31-39):

    LoadImageFromFile -> [CenterCrop(0.68, relative_range)] -> Resize(keep_ratio) -> RandomFlip(0.0)
    -> Normalize(to_rgb) -> Pad(size_divisor=32) -> DefaultFormatBundle (HWC -> CHW)

The reference's transforms (mmdet/datasets/pipelines/transforms.py) delegate the pixel arithmetic to
mmcv-full 1.4.8 (``imrescale``, ``imnormalize``, ``impad_to_multiple``), which is not under
/root/reference and is itself a thin layer over OpenCV (``cv2.resize(INTER_LINEAR)``,
``cv2.cvtColor`` / ``cv2.subtract`` / ``cv2.multiply``, ``cv2.copyMakeBorder``).  Pinning:
  * ``resize_linear_u8`` / ``imnormalize`` are checked bit-exact against cv2 4.13 itself
    (tests/test_preprocess.py, live when cv2 is importable);
  * the whole chain is checked bit-exact against golden fixtures produced by the REFERENCE's own
    transform classes (CenterCrop / Resize / RandomFlip / Normalize / Pad, built from the reference's
    config dicts) running on mmcv's published image functions over cv2
    (oracle/gen_golden_preprocess.py -> tests/golden/golden_preprocess.npz).
  * "parity unpinned" only at the mmcv wrapper boundary (its few lines are restated in
    oracle/refshim.py from the published 1.4.8 source).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import numpy as np

COEF_BITS = 11                      # OpenCV INTER_RESIZE_COEF_BITS
COEF_SCALE = 1 << COEF_BITS


# ----------------------------------------------------------------------------------------
# geometry (python float / int arithmetic exactly as the reference's python code does it)
# ----------------------------------------------------------------------------------------
def center_crop_size(h: int, w: int, crop_size: Sequence[float], rand: float) -> Tuple[int, int]:
    """CenterCrop._get_crop_size, crop_type='relative_range' (transforms.py:1125-1130): ONE uniform
    draw `rand` (np.random.rand(1)[0]) scales both sides: ratio = crop_size + rand * (1 - crop_size),
    computed on a float32 array plus a float64 array (-> float64)."""
    cs = np.asarray(crop_size, dtype=np.float32)
    ch, cw = cs + np.asarray([rand], dtype=np.float64) * (1 - cs)
    return int(h * ch + 0.5), int(w * cw + 0.5)


def center_crop_window(h: int, w: int, crop_h: int, crop_w: int) -> Tuple[int, int, int, int]:
    """CenterCrop._crop_data (transforms.py:1036-1047): (y1, x1, y2, x2) of the numpy slice; the slice
    clips at the image border like numpy does."""
    margin_h = max(h - crop_h, 0)
    margin_w = max(w - crop_w, 0)
    y1 = int(margin_h / 2 + 0.5)
    x1 = int(margin_w / 2 + 0.5)
    return y1, x1, min(y1 + crop_h, h), min(x1 + crop_w, w)


def rescale_size(w: int, h: int, scale: Tuple[int, int]) -> Tuple[int, int]:
    """mmcv.image.geometric.rescale_size with a tuple scale (the call of Resize._resize_img,
    transforms.py:217-221): factor = min(long/max(h,w), short/min(h,w)); new = int(x*factor+0.5)."""
    max_long, max_short = max(scale), min(scale)
    f = min(max_long / max(h, w), max_short / min(h, w))
    return int(w * float(f) + 0.5), int(h * float(f) + 0.5)


def pad_size(h: int, w: int, divisor: int = 32) -> Tuple[int, int]:
    """mmcv.impad_to_multiple (Pad._pad_img, transforms.py:677-679): ceil to a multiple, pad bottom/right."""
    return int(np.ceil(h / divisor)) * divisor, int(np.ceil(w / divisor)) * divisor


# ----------------------------------------------------------------------------------------
# cv2.resize(INTER_LINEAR) on uint8 — OpenCV's fixed-point algorithm (imgproc/src/resize.cpp:
# resize() coefficient tables, HResizeLinear<uchar,int,short>, VResizeLinear<uchar,int,short>)
# ----------------------------------------------------------------------------------------
def _linear_taps(src: int, dst: int) -> Tuple[np.ndarray, np.ndarray]:
    """Per destination index: integer source index (may be -1 / src-1) and the float32 fraction.
    scale = 1 / (dst / src) in double (NOT src / dst), f = float((d + 0.5) * scale - 0.5)."""
    inv = np.float64(dst) / np.float64(src)
    scale = np.float64(1.0) / inv
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    return s, (f - s.astype(np.float32)).astype(np.float32)


def _coef(f: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """saturate_cast<short>(c * 2048) for c = 1 - f and c = f (float32 product, round half to even)."""
    c1 = np.rint(f * np.float32(COEF_SCALE)).astype(np.int32)
    c0 = np.rint((np.float32(1) - f) * np.float32(COEF_SCALE)).astype(np.int32)
    return c0, c1


def resize_linear_u8(img: np.ndarray, dst_w: int, dst_h: int) -> np.ndarray:
    """Bit-exact cv2.resize(img, (dst_w, dst_h), interpolation=cv2.INTER_LINEAR) for uint8 HWC input.
    Horizontal pass in int32 (a0*S[sx] + a1*S[sx+1], 11-bit coefficients; taps outside the row are
    folded onto the border pixel with fx = 0), vertical pass
    ((b0*(r0>>4))>>16) + ((b1*(r1>>4))>>16) + 2) >> 2 with the row index clipped to the image."""
    sh, sw = img.shape[:2]
    if (sh, sw) == (dst_h, dst_w):
        return img.copy()
    sx, fx = _linear_taps(sw, dst_w)
    neg = sx < 0
    fx[neg] = 0
    sx[neg] = 0
    top = sx >= sw - 1
    fx[top] = 0
    sx[top] = sw - 1
    a0, a1 = _coef(fx)
    sx1 = np.minimum(sx + 1, sw - 1)
    sy, fy = _linear_taps(sh, dst_h)
    b0, b1 = _coef(fy)
    y0 = np.clip(sy, 0, sh - 1)
    y1 = np.clip(sy + 1, 0, sh - 1)
    src = img.astype(np.int32)
    rows = src[:, sx] * a0[None, :, None] + src[:, sx1] * a1[None, :, None]          # [sh, dst_w, c]
    r0, r1 = rows[y0], rows[y1]
    out = (((b0[:, None, None] * (r0 >> 4)) >> 16) + ((b1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


# ----------------------------------------------------------------------------------------
# mmcv.imnormalize (photometric.py, 1.4.8): float32 image, float64 scalars, two cv2 array-scalar ops
# ----------------------------------------------------------------------------------------
def imnormalize(img_u8_bgr: np.ndarray, mean: Sequence[float], std: Sequence[float], to_rgb: bool = True) -> np.ndarray:
    """cv2.subtract / cv2.multiply of a float32 array with a float64 scalar evaluate in double and round
    to float32 after EACH op: t = f32(f64(x) - mean), y = f32(f64(t) * (1/std)); mean / std are the
    float32 config values widened to double (Normalize.__init__, transforms.py:735-736)."""
    mean64 = np.asarray(mean, dtype=np.float32).astype(np.float64)
    stdinv64 = 1 / np.asarray(std, dtype=np.float32).astype(np.float64)
    x = img_u8_bgr[..., ::-1] if to_rgb else img_u8_bgr
    t = (x.astype(np.float64) - mean64).astype(np.float32)
    return (t.astype(np.float64) * stdinv64).astype(np.float32)


# ----------------------------------------------------------------------------------------
# the chain for one frame
# ----------------------------------------------------------------------------------------
def preprocess_frame(img_bgr: np.ndarray, scale: Tuple[int, int], mean: Sequence[float], std: Sequence[float],
                     to_rgb: bool = True, crop_size: Optional[Sequence[float]] = None, rand: float = 0.0,
                     divisor: int = 32) -> Dict[str, object]:
    """uint8 BGR [h, w, 3] -> dict(img float32 [3, Hp, Wp], img_shape, pad_shape, ori_shape, scale_factor)
    with the meta keys the forward consumes (Resize._resize_img transforms.py:231-241, Pad._pad_img :680)."""
    ori_shape = img_bgr.shape
    if crop_size is not None:
        ch, cw = center_crop_size(img_bgr.shape[0], img_bgr.shape[1], crop_size, rand)
        y1, x1, y2, x2 = center_crop_window(img_bgr.shape[0], img_bgr.shape[1], ch, cw)
        img_bgr = img_bgr[y1:y2, x1:x2]
    h, w = img_bgr.shape[:2]
    new_w, new_h = rescale_size(w, h, scale)
    resized = resize_linear_u8(img_bgr, new_w, new_h)
    scale_factor = np.array([new_w / w, new_h / h, new_w / w, new_h / h], dtype=np.float32)
    norm = imnormalize(resized, mean, std, to_rgb)
    Hp, Wp = pad_size(new_h, new_w, divisor)
    out = np.zeros((Hp, Wp, 3), dtype=np.float32)
    out[:new_h, :new_w] = norm
    return dict(img=np.ascontiguousarray(out.transpose(2, 0, 1)),      # DefaultFormatBundle, formating.py
                img_shape=(new_h, new_w, 3), pad_shape=(Hp, Wp, 3), ori_shape=tuple(ori_shape),
                scale_factor=scale_factor)
