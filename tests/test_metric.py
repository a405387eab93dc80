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
    assert so.mcg_gaze_error(None, None, None, 1, 0, None, None) == -2
    assert so.mcg_merge_clips(None, None, None, 1, 7, 4, None, None, None) == -2
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


# ------------------------------------------------------------------------------------------------------------------
# both scorer variants against what the reference's OWN scripts print (oracle/gen_golden_scorer.py executes
# tools/calculate_mae_gaze360.py and tools/calculate_mae_l2cs.py on seeded synthetic videos)
# ------------------------------------------------------------------------------------------------------------------
def _synthetic_scorer_case():
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('gen_golden_scorer', os.path.join(root, 'oracle', 'gen_golden_scorer.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize('variant', ['gaze360', 'l2cs'])
def test_scorer_variants_match_the_reference_scripts(golden_dir, variant):
    gen = _synthetic_scorer_case()
    gold = json.load(open(os.path.join(golden_dir, 'golden_scorer.json')))
    results, a360, al2 = gen.synthetic(gold['seed'], gold['n_videos'])
    pred = [np.asarray(r['fusion_gazes']) for r in results]
    gt = metric.l2cs_ground_truth(al2) if variant == 'l2cs' else [np.asarray(a['gaze']) for a in a360['annotations']]
    got = metric.gaze_error(pred, gt, variant=variant)
    for k, want in gold[variant].items():
        assert abs(got[k] - want) < 0.0075, (variant, k, got[k], want)          # printed with %.2f
    assert gold['l2cs']['mae_front20'] != gold['gaze360']['mae_front20']         # the pitch condition matters here
    if os.path.isdir(gen.REF):                                                   # live against the reference script
        script = 'calculate_mae_l2cs.py' if variant == 'l2cs' else 'calculate_mae_gaze360.py'
        assert gen.run_reference(script, results, al2 if variant == 'l2cs' else a360) == gold[variant]


@pytest.mark.gpu
@pytest.mark.parametrize('variant', ['gaze360', 'l2cs'])
def test_gpu_scorer_variants(golden_dir, variant):
    import torch
    from mcgaze_b200 import lib
    gen = _synthetic_scorer_case()
    gold = json.load(open(os.path.join(golden_dir, 'golden_scorer.json')))
    results, a360, al2 = gen.synthetic(gold['seed'], gold['n_videos'])
    pred = [np.asarray(r['fusion_gazes'], dtype=np.float32) for r in results]
    gt = metric.l2cs_ground_truth(al2) if variant == 'l2cs' else [np.asarray(a['gaze']) for a in a360['annotations']]
    lengths = [len(p) for p in pred]
    got = lib.gaze_error(torch.from_numpy(np.concatenate(pred)).cuda(),
                         torch.from_numpy(np.concatenate(gt).astype(np.float32)).cuda(), lengths, variant=variant)
    ref = metric.gaze_error(pred, gt, variant=variant)
    for k, want in gold[variant].items():
        assert abs(got[k] - want) < 0.0075 and abs(got[k] - ref[k]) < 1e-3, (variant, k, got[k], want, ref[k])
    for k in ('frames_360', 'frames_front90', 'frames_front20'):
        assert got[k] == ref[k], k
    # sums of disjoint shards add up to the whole: the one all-reduce of a sharded run
    cut = len(lengths) // 3
    off = int(np.sum(lengths[:cut]))
    p_all = torch.from_numpy(np.concatenate(pred)).cuda()
    g_all = torch.from_numpy(np.concatenate(gt).astype(np.float32)).cuda()
    a = lib.gaze_error_sums(p_all[:off].contiguous(), g_all[:off].contiguous(), lengths[:cut], variant)
    b = lib.gaze_error_sums(p_all[off:].contiguous(), g_all[off:].contiguous(), lengths[cut:], variant)
    whole = lib.sums_to_mae((a + b).cpu().numpy())
    for k in ref:
        assert abs(whole[k] - got[k]) < 1e-9 * max(1.0, abs(got[k])), k


@pytest.mark.gpu
def test_gpu_overlap_merge_matches_the_host_merger():
    """mcg_merge_clips against mcgaze_b200.slicer.merge_video (itself pinned against the reference's loop) on videos of
    every length class: single short clip, exact window, right-aligned last window with overlap 4 / 5 / 6 / 3 -
    including frames covered by THREE clips - and scores on both sides of the 0.5 threshold.  Bit-exact."""
    import torch
    from mcgaze_b200 import lib, slicer
    rng = np.random.default_rng(5)
    lengths = [1, 3, 7, 8, 9, 10, 11, 12, 15, 23, 50, 101]
    rows, cpv = [], []
    want_det, want_gaze = [], []
    for L in lengths:
        plan = slicer.plan_clips(L)
        boxes = [rng.uniform(0, 200, (n, 3, 4)).astype(np.float32) for _, n, _ in plan]
        scores = [rng.uniform(0.2, 0.9, (n, 3)).astype(np.float32) for _, n, _ in plan]
        gaze = [rng.normal(size=(n, 4, 3)).astype(np.float32) for _, n, _ in plan]
        m = slicer.merge_video(plan, boxes, scores, gaze)
        want_det.append(m['det'])
        want_gaze.append(m['gaze'])
        cpv.append(len(plan))
        for (_, n, _), b, s, g in zip(plan, boxes, scores, gaze):
            r = np.zeros((7, 27), np.float32)
            r[:n] = np.concatenate([b.reshape(n, 12), s, g.reshape(n, 12)], 1)
            rows.append(r)
    det, gz = lib.merge_clips(torch.from_numpy(np.stack(rows)).cuda(), cpv, lengths)
    assert np.array_equal(det.cpu().numpy(), np.concatenate(want_det))
    assert np.array_equal(gz.cpu().numpy(), np.concatenate(want_gaze))
