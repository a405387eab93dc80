"""CPU: the scorer restatement reproduces what the reference's own calculate_mae_gaze360.py prints
for the shipped results JSON (fixtures made by oracle/gen_golden.py)."""
import json
import os

import numpy as np

from mcgaze_b200 import metric


def _videos(d, key):
    off = np.concatenate([[0], np.cumsum(d['lengths'])])
    return [d[key][off[i]:off[i + 1]] for i in range(len(d['lengths']))]


def test_mae_known_answer(golden_dir):
    d = np.load(os.path.join(golden_dir, 'golden_gaze360_results.npz'))
    gold = json.load(open(os.path.join(golden_dir, 'golden_mae_gaze360.json')))
    assert len(d['lengths']) == 517 and int(d['lengths'].sum()) == 25969
    gt = _videos(d, 'gt')
    for name in ('fusion_gazes', 'face_gazes', 'eyes_gazes', 'head_gazes'):
        r = metric.gaze_error(_videos(d, name), gt)
        for k in ('mae_360', 'mae_front90', 'mae_front20'):
            assert abs(r[k] - gold[name][k]) < 0.0075, (name, k, r[k], gold[name][k])   # printed with %.2f
    assert gold['fusion_gazes'] == {'mae_360': 12.99, 'mae_front90': 10.72, 'mae_front20': 9.99}


def test_smooth_filter_edges():
    one = np.array([[0.0, 0.0, -2.0]])
    assert (metric.smooth_filter(one.copy()) == one).all()          # 1-frame videos are not normalised
    x = np.array([[0, 0, -1.0], [0, 1.0, 0], [1.0, 0, 0]])
    y = metric.smooth_filter(x.copy())
    assert np.allclose(np.linalg.norm(y, axis=1), 1.0)
    assert np.allclose(y[0] * np.linalg.norm(0.6 * x[0] + 0.4 * x[1]), 0.6 * x[0] + 0.4 * x[1])
