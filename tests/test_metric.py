"""CPU: the scorer restatement reproduces what the reference's own calculate_mae_gaze360.py prints
for the shipped results JSON (fixtures made by oracle/gen_golden.py)."""
import json
import os

import numpy as np
import pytest

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


def test_device_scorer_has_no_cpu_fallback():
    import ctypes

    import torch
    from mcgaze_b200 import lib
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    so = lib.load_library()
    assert so.mcg_gaze_error(None, None, None, 1, None, None) == -2
    assert b'no CPU fallback' in so.mcg_last_error()
    with pytest.raises(lib.McgError):
        lib.gaze_error(torch.zeros(2, 3), torch.zeros(2, 3), [2])
    del ctypes


@pytest.mark.gpu
def test_gpu_scorer_reproduces_the_published_mae(golden_dir):
    """mcg_gaze_error on the predictions of the reference's shipped results JSON: the three numbers the reference's
    calculate_mae_gaze360.py prints (12.99 / 10.72 / 9.99 for the fused gaze), the frame counts of the categories,
    and <= 1e-3 deg to the float64 restatement (the reference itself computes in fp32)."""
    import torch
    from mcgaze_b200 import lib
    d = np.load(os.path.join(golden_dir, 'golden_gaze360_results.npz'))
    gold = json.load(open(os.path.join(golden_dir, 'golden_mae_gaze360.json')))
    gt_dev = torch.from_numpy(d['gt'].astype(np.float32)).cuda()
    gt = _videos(d, 'gt')
    for name in ('fusion_gazes', 'face_gazes', 'eyes_gazes', 'head_gazes'):
        got = lib.gaze_error(torch.from_numpy(d[name].astype(np.float32)).cuda(), gt_dev, d['lengths'])
        ref = metric.gaze_error(_videos(d, name), gt)
        for k in ('mae_360', 'mae_front90', 'mae_front20'):
            assert abs(got[k] - ref[k]) < 1e-3, (name, k, got[k], ref[k])
            assert abs(got[k] - gold[name][k]) < 0.0075, (name, k, got[k], gold[name][k])
        for k in ('frames_360', 'frames_front90', 'frames_front20'):
            assert got[k] == ref[k], (name, k)


@pytest.mark.gpu
def test_gpu_scorer_edge_cases():
    """1-frame videos are not smoothed (nor normalised), 2-frame videos use the two end rules, long videos span
    several loop iterations of a CTA."""
    import torch
    from mcgaze_b200 import lib
    rng = np.random.default_rng(3)
    lengths = [1, 2, 3, 1, 300, 2, 129]
    n = sum(lengths)
    gt = rng.normal(size=(n, 3))
    pred = gt + 0.2 * rng.normal(size=(n, 3))
    pred /= np.linalg.norm(pred, axis=1, keepdims=True)
    off = np.concatenate([[0], np.cumsum(lengths)])
    ref = metric.gaze_error([pred[off[i]:off[i + 1]] for i in range(len(lengths))],
                            [gt[off[i]:off[i + 1]] for i in range(len(lengths))])
    got = lib.gaze_error(torch.from_numpy(pred.astype(np.float32)).cuda(), torch.from_numpy(gt.astype(np.float32)).cuda(), lengths)
    for k in ref:
        assert abs(got[k] - ref[k]) < 1e-3, (k, got[k], ref[k])
    with pytest.raises(lib.McgError):
        lib.gaze_error(torch.zeros(3, 3).cuda(), torch.zeros(3, 3).cuda(), [2])          # lengths do not add up
