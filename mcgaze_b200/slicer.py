"""Clip slicing and overlap merging of the reference's evaluation driver
(tools/test_gaze360_gaze.py:60-269), restated so that MANY clips go through one forward.

`plan_clips` reproduces :73-86 (7-frame windows, stride 4, last window right-aligned with its own
overlap); `merge_video` reproduces :129-201 (non-overlapping frames copied, overlapping frames
averaged -- gaze vectors are NOT re-normalised, boxes are zeroed where either score < 0.5);
`video_record` reproduces the JSON schema of :210-260 (boxes xywh, None when zeroed).
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np

CLIP_LEN = 7
STRIDE = 4
PERSON_THRESHOLD = 0.5


def plan_clips(video_length: int, clip_len: int = CLIP_LEN, stride: int = STRIDE) -> List[Tuple[int, int, int]]:
    """-> [(start_frame, n_frames, overlap_with_previous)] in clip order."""
    if video_length <= clip_len:
        return [(0, video_length, 0)]
    n = math.ceil((video_length - clip_len) / stride) + 1
    plan = []
    for i in range(n):
        if i != n - 1:
            plan.append((i * stride, clip_len, clip_len - stride if i else 0))
        else:
            rem = (video_length - clip_len) % stride
            plan.append((video_length - clip_len, clip_len, clip_len - rem if rem else clip_len - stride))
    return plan


def merge_video(plan: Sequence[Tuple[int, int, int]], boxes: Sequence[np.ndarray], scores: Sequence[np.ndarray],
                gaze: Sequence[np.ndarray]) -> Dict[str, np.ndarray]:
    """boxes[i] [T,3,4], scores[i] [T,3], gaze[i] [T,4,3] per clip -> per-video arrays
    det [L,3,5] (x1,y1,x2,y2,score), gaze [L,4,3] (fusion, face, eyes, head)."""
    det_v = gz_v = None
    for (start, n, overlap), b, s, g in zip(plan, boxes, scores, gaze):
        b = np.asarray(b, dtype=np.float32).copy()
        s = np.asarray(s, dtype=np.float32)
        g = np.asarray(g, dtype=np.float32)
        b[s < PERSON_THRESHOLD] = 0.0                               # :135-141 / :188-194
        det = np.concatenate([b, s[..., None]], -1)                  # [T,3,5]
        if det_v is None:
            det_v, gz_v = det, g
            continue
        new = n - overlap
        det_v = np.concatenate([det_v, det[-new:]], 0)               # non-overlapping tail  :157-160
        gz_v = np.concatenate([gz_v, g[-new:]], 0)
        o_prev = slice(det_v.shape[0] - n, det_v.shape[0] - new)     # overlapping frames
        prev, cur = det_v[o_prev], det[:overlap]
        bad = (prev[..., 4] < PERSON_THRESHOLD) | (cur[..., 4] < PERSON_THRESHOLD)     # :170-174
        avg = (prev + cur) / 2
        avg[..., :4][bad] = 0.0
        det_v[o_prev] = avg
        gz_v[o_prev] = (gz_v[o_prev] + g[:overlap]) / 2              # :182-183, not re-normalised
    return {'det': det_v, 'gaze': gz_v}


def video_record(video_id: int, merged: Dict[str, np.ndarray]) -> Dict[str, object]:
    det, gz = merged['det'], merged['gaze']
    rec: Dict[str, object] = dict(video_id=video_id, category_id=1, fusion_gazes=[])
    for ci, name in enumerate(('face', 'eyes', 'head')):
        rec[f'{name}_bboxes'], rec[f'{name}_gazes'], rec[f'{name}_score'] = [], [], []
    for t in range(det.shape[0]):
        rec['fusion_gazes'].append(gz[t, 0].tolist())
        for ci, name in enumerate(('face', 'eyes', 'head')):
            x1, y1, x2, y2 = (float(v) for v in det[t, ci, :4])
            rec[f'{name}_bboxes'].append(None if (x1 + y1 + x2 + y2) == 0 else [x1, y1, x2 - x1, y2 - y1])
            rec[f'{name}_gazes'].append(gz[t, 1 + ci].tolist())
            rec[f'{name}_score'].append(float(det[t, ci, 4]))
    return rec
