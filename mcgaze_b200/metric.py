"""Gaze360 / l2cs scorer: restatement of the reference's tools/calculate_mae_gaze360.py
(smooth_filter :16-29, vector_to_yaw_pitch :60-66, compute_yaw_angular :69-74,
compute_angular_error :77-94, gaze_error :110-188) as vectorised numpy on the host, and of its l2cs
sibling tools/calculate_mae_l2cs.py (`variant='l2cs'`: front-20 also needs |pitch(gt)| <= 20 deg, :139;
the ground truth of video k is annotations[3 * k], :110 -> `l2cs_ground_truth`; the smoothing IS applied,
:124, whatever the comment at :17 says).
Kept bit-compatible in behaviour: smoothing alpha 0.6 with re-normalisation (not for 1-frame
videos), only the TARGET is normalised in the angular error, per-video mean weighted by frames,
front = |yaw(gt)| <= 90 deg, front-20 = |yaw(gt)| <= 20 deg.  The reference computes in fp32
torch; float64 here changes the printed 2-decimal numbers by < 0.005."""
from __future__ import annotations

from typing import Dict, Sequence

import numpy as np


def smooth_filter(x: np.ndarray, alpha: float = 0.6) -> np.ndarray:
    if x.shape[0] < 2:
        return x
    out = alpha * x
    out[0] += (1 - alpha) * x[1]
    out[-1] += (1 - alpha) * x[-2]
    out[1:-1] += (1 - alpha) * (x[:-2] + x[2:]) / 2
    return out / np.linalg.norm(out, axis=1, keepdims=True)


def vector_to_yaw_pitch(v: np.ndarray) -> np.ndarray:
    v = v.reshape(-1, 3)
    v = v / np.linalg.norm(v, axis=1, keepdims=True)
    return np.stack([np.arctan2(v[:, 0], -v[:, 2]), np.arcsin(v[:, 1])], 1)


def angular_error_deg(pred: np.ndarray, gt: np.ndarray) -> np.ndarray:
    """per-frame angle in degrees; pred is used as-is, gt is normalised (:84)."""
    gt = gt / np.linalg.norm(gt, axis=1, keepdims=True)
    return np.degrees(np.arccos((pred * gt).sum(1)))


def l2cs_ground_truth(anno: Dict) -> list:
    """calculate_mae_l2cs.py:110: the ground truth of result video k is annotations[3 * k]['gaze']."""
    return [np.asarray(a['gaze'], dtype=np.float64).reshape(-1, 3) for a in anno['annotations'][::3]]


def gaze_error(pred_videos: Sequence[np.ndarray], gt_videos: Sequence[np.ndarray], variant: str = 'gaze360') -> Dict[str, float]:
    if variant not in ('gaze360', 'l2cs'):
        raise ValueError(f'unknown scorer variant {variant!r}')
    tot = {'360': [0.0, 0], 'front90': [0.0, 0], 'front20': [0.0, 0]}
    for p, g in zip(pred_videos, gt_videos):
        p = smooth_filter(np.asarray(p, dtype=np.float64).reshape(-1, 3).copy())
        g = np.asarray(g, dtype=np.float64).reshape(-1, 3)
        assert len(p) == len(g)
        err = angular_error_deg(p, g)
        yp = np.degrees(np.abs(vector_to_yaw_pitch(g)))
        yaw = yp[:, 0]
        front20 = (yaw <= 20) & (yp[:, 1] <= 20) if variant == 'l2cs' else yaw <= 20
        for key, mask in (('360', np.ones(len(g), bool)), ('front90', yaw <= 90), ('front20', front20)):
            n = int(mask.sum())
            if n:
                tot[key][0] += float(err[mask].mean()) * n     # per-video mean x frames (:157-158)
                tot[key][1] += n
    return {f'mae_{k}': v[0] / max(v[1], 1) for k, v in tot.items()} | {f'frames_{k}': v[1] for k, v in tot.items()}
