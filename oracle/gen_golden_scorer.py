"""Golden numbers of the reference's two scorers on a seeded synthetic result / annotation pair.

    python oracle/gen_golden_scorer.py        # needs /root/reference (this container only)

Executes the reference's OWN scripts (tools/calculate_mae_gaze360.py, tools/calculate_mae_l2cs.py: plain torch, no mmcv)
on synthetic per-video predictions and ground truth and stores what they print, next to the inputs' seed, in
tests/golden/golden_scorer.json.  TEST INFRASTRUCTURE: pins mcgaze_b200/metric.py and mcg_gaze_error (both variants).
The l2cs annotation file is not derivable offline, hence synthetic data: video lengths 1..40, ground truth
spread over the whole sphere so that all three categories are populated, predictions = ground truth + noise."""
from __future__ import annotations

import contextlib
import importlib.util
import io
import json
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get('MCGAZE_REFERENCE', '/root/reference')


def synthetic(seed: int = 0, n_videos: int = 60):
    """-> (results as the eval json, annotations in the Gaze360 layout, annotations in the l2cs layout)."""
    rng = np.random.RandomState(seed)
    results, ann360, annl2 = [], [], []
    for v in range(n_videos):
        L = int(rng.randint(1, 41))
        yaw = rng.uniform(-np.pi, np.pi, L) * (0.25 if v % 3 == 0 else 1.0)
        pitch = rng.uniform(-1.0, 1.0, L) * (0.3 if v % 2 == 0 else 1.0)
        gt = np.stack([np.cos(pitch) * np.sin(yaw), np.sin(pitch), -np.cos(pitch) * np.cos(yaw)], 1)
        gt = gt * rng.uniform(0.5, 2.0, (L, 1))                     # the scorers normalise the target themselves
        pred = gt / np.linalg.norm(gt, axis=1, keepdims=True) + rng.normal(0, 0.15, (L, 3))
        pred = pred / np.linalg.norm(pred, axis=1, keepdims=True)
        results.append({'video_id': v + 1, 'fusion_gazes': pred.astype(np.float32).tolist()})
        ann360.append({'gaze': gt.astype(np.float32).tolist()})
        for k in range(3):                                          # l2cs: annotation 3k belongs to video k (:110)
            annl2.append({'gaze': (gt if k == 0 else gt[::-1]).astype(np.float32).tolist()})
    return results, {'annotations': ann360}, {'annotations': annl2}


def run_reference(script: str, results, anno):
    spec = importlib.util.spec_from_file_location('ref_scorer', os.path.join(REF, 'tools', script))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        mod.gaze_error(results, anno, 'fusion_gazes')
    nums = [float(x) for x in re.findall(r':\s*(-?[0-9.]+)', buf.getvalue())]
    assert len(nums) == 3, buf.getvalue()
    return dict(mae_360=nums[0], mae_front90=nums[1], mae_front20=nums[2])


def main():
    results, a360, al2 = synthetic(0)
    out = {'seed': 0, 'n_videos': len(results),
           'gaze360': run_reference('calculate_mae_gaze360.py', results, a360),
           'l2cs': run_reference('calculate_mae_l2cs.py', results, al2)}
    path = os.path.join(ROOT, 'tests', 'golden', 'golden_scorer.json')
    json.dump(out, open(path, 'w'), indent=1)
    print(path, out)


if __name__ == '__main__':
    sys.exit(main())
