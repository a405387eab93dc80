"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.npz|json by running the REFERENCE's own
python code in this container (it cannot travel to the GPU box, the fixtures can).

    python oracle/gen_golden.py            # needs /root/reference

What is pinned:
  golden_forward_*.npz   outputs of the reference's `MultiClueGaze` (built by the reference's
                         `build_detector` from the reference's own config, weights =
                         oracle.make_state_dict(seed), strict key match) called exactly as
                         tools/test_gaze360_gaze.py:107-111 calls it.  mmcv primitives come from
                         oracle/refshim.py (restated; see its header).
  golden_coder.json      the reference's DeltaXYWHBBoxCoder on the inputs of its own known-answer
                         test (tests/test_utils/test_coder.py:27-75) + the expected values there.
  golden_mae_gaze360.json / golden_gaze360_results.npz
                         the reference's tools/calculate_mae_gaze360.py:gaze_error run on the shipped
                         results/results_multiclue_gaze_r50_gaze360_test.json against GT rebuilt from
                         tools/dataset_converters/gaze360/test.txt, plus the
                         predictions/GT themselves in compact form so the scorer restatement can be
                         re-checked without /root/reference.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import mcgaze_oracle as O  # noqa: E402
from oracle import refshim  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')
REF = refshim.REFERENCE_ROOT

FORWARD_CASES = [
    # name, seed, T, H, W, img_shape(h,w), scale_factor
    dict(name='t7_224', seed=0, T=7, H=224, W=224, img_hw=(224, 224), scale=(1.0, 1.0, 1.0, 1.0)),
    dict(name='t3_192x224_rescale', seed=5, T=3, H=192, W=224, img_hw=(180, 224), scale=(0.7, 0.75, 0.7, 0.75)),
    dict(name='t1_224', seed=7, T=1, H=224, W=224, img_hw=(224, 224), scale=(1.0, 1.0, 1.0, 1.0)),
]


def build_reference_model(sd):
    refshim.install()
    from mmcv import Config
    from mmdet.models import build_detector
    cfg = Config.fromfile(os.path.join(REF, 'configs/multiclue_gaze/multiclue_gaze_r50_gaze360.py'))
    cfg.model.backbone.init_cfg = None          # no torchvision:// download offline
    cfg.model.train_cfg = None
    model = build_detector(cfg.model, test_cfg=cfg.get('test_cfg'))
    own = model.state_dict()
    assert set(own.keys()) == set(sd.keys()), (sorted(set(own) ^ set(sd))[:10])
    for k in own:
        assert tuple(own[k].shape) == tuple(sd[k].shape), k
    model.load_state_dict(sd, strict=True)
    model.eval()
    return model


def gen_forward():
    sd = O.make_state_dict(0)
    model = build_reference_model(sd)
    for c in FORWARD_CASES:
        img = O.make_clip(c['seed'], c['T'], c['H'], c['W'])
        metas = [dict(img_shape=(c['img_hw'][0], c['img_hw'][1], 3), ori_shape=(c['H'], c['W'], 3),
                      pad_shape=(c['H'], c['W'], 3), scale_factor=np.array(c['scale'], dtype=np.float32),
                      flip=False, filename=f'{i:05d}.png') for i in range(c['T'])]
        with torch.no_grad():
            (det_bboxes, det_labels), gaze = model(return_loss=False, rescale=True, format=False, img=[img],
                                                   img_metas=[metas])
        out = {k: v.numpy() for k, v in gaze.items()}
        out['det_bboxes'] = torch.stack(det_bboxes).numpy()      # [T,3,5]
        assert all(l == [0, 1, 2] for l in det_labels)
        np.savez_compressed(os.path.join(GOLD, f'golden_forward_{c["name"]}.npz'),
                            seed=c['seed'], T=c['T'], H=c['H'], W=c['W'], img_hw=np.array(c['img_hw']),
                            scale=np.array(c['scale'], dtype=np.float32), **out)
        print('forward', c['name'], out['gaze_score'][0], out['det_bboxes'][0, 0])


def gen_coder():
    refshim.install()
    from mmdet.core.bbox.coder import DeltaXYWHBBoxCoder
    rois = [[0., 0., 1., 1.], [0., 0., 1., 1.], [0., 0., 1., 1.], [5., 5., 5., 5.]]
    deltas = [[0., 0., 0., 0.], [1., 1., 1., 1.], [0., 0., 2., -1.], [0.7, -1.9, -0.5, 0.3]]
    expected = [[0.0000, 0.0000, 1.0000, 1.0000], [0.1409, 0.1409, 2.8591, 2.8591],
                [0.0000, 0.3161, 4.1945, 0.6839], [5.0000, 5.0000, 5.0000, 5.0000]]   # test_coder.py:34-37
    out = DeltaXYWHBBoxCoder().decode(torch.tensor(rois), torch.tensor(deltas), max_shape=(32, 32))
    assert torch.tensor(expected).allclose(out, atol=1e-4)
    # the gaze config's coder: stds (.5,.5,1,1), clip_border=False (cfg :69-73), random boxes
    g = torch.Generator().manual_seed(3)
    r2 = torch.rand(64, 4, generator=g) * 200
    r2[:, 2:] += r2[:, :2]
    d2 = torch.randn(64, 4, generator=g) * 1.5
    d2[0, 2:] = torch.tensor([9.0, -9.0])                        # exercises the wh clamp
    out2 = DeltaXYWHBBoxCoder(target_stds=[0.5, 0.5, 1., 1.], clip_border=False).decode(r2, d2, max_shape=(224, 224))
    json.dump({'kat_rois': rois, 'kat_deltas': deltas, 'kat_expected': expected, 'kat_reference_out': out.tolist(),
               'cfg_rois': r2.tolist(), 'cfg_deltas': d2.tolist(), 'cfg_out': out2.tolist()},
              open(os.path.join(GOLD, 'golden_coder.json'), 'w'))
    print('coder ok')


def rebuild_gaze360_gt():
    """tools/dataset_converters/gaze360/generate_json_from_ori.py:40-60: sort lines, new video when the
    frame number is not consecutive or the recording / person directory changes."""
    lines = open(os.path.join(REF, 'tools/dataset_converters/gaze360/test.txt')).read().strip().split('\n')
    lines.sort()
    videos, cur, prev = [], [], None
    for ln in lines:
        parts = ln.split(' ')
        path = parts[0].split('/')
        key = (path[0], path[2], int(path[3].replace('.jpg', '')))
        if prev is not None and not (key[2] == prev[2] + 1 and key[1] == prev[1] and key[0] == prev[0]):
            videos.append(cur)
            cur = []
        cur.append([float(parts[1]), float(parts[2]), float(parts[3])])
        prev = key
    videos.append(cur)
    return videos


def gen_mae():
    sys.path.insert(0, os.path.join(REF, 'tools'))
    import calculate_mae_gaze360 as ref_mae
    ev = json.load(open(os.path.join(REF, 'results/results_multiclue_gaze_r50_gaze360_test.json')))
    gt = rebuild_gaze360_gt()
    assert len(gt) == len(ev) and all(len(g) == len(v['fusion_gazes']) for g, v in zip(gt, ev))
    anno = {'annotations': [{'gaze': g} for g in gt]}
    res = {}
    for name in ('fusion_gazes', 'face_gazes', 'eyes_gazes', 'head_gazes'):
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            ref_mae.gaze_error(ev, anno, name)
        vals = [float(l.split(':')[1]) for l in buf.getvalue().strip().split('\n') if ':' in l]
        res[name] = {'mae_360': vals[0], 'mae_front90': vals[1], 'mae_front20': vals[2]}
        print(name, res[name])
    json.dump(res, open(os.path.join(GOLD, 'golden_mae_gaze360.json'), 'w'), indent=1)
    lengths = np.array([len(g) for g in gt], dtype=np.int32)
    pack = {'lengths': lengths, 'gt': np.concatenate([np.array(g, dtype=np.float64) for g in gt])}
    for name in ('fusion_gazes', 'face_gazes', 'eyes_gazes', 'head_gazes'):
        pack[name] = np.concatenate([np.array(v[name], dtype=np.float64).reshape(-1, 3) for v in ev])
    np.savez_compressed(os.path.join(GOLD, 'golden_gaze360_results.npz'), **pack)
    # clip bookkeeping of the reference slicer (tools/test_gaze360_gaze.py:73-86) for the scaling bench
    hist = {}
    for L in lengths.tolist():
        n = 1 if L <= 7 else int(np.ceil((L - 7) / 4)) + 1
        for i in range(n):
            T = min(L, 7)
            hist[T] = hist.get(T, 0) + 1
    json.dump({'clip_length_histogram': hist, 'videos': int(len(lengths)), 'frames': int(lengths.sum())},
              open(os.path.join(GOLD, 'golden_gaze360_clip_hist.json'), 'w'))
    print('clip hist', hist)


if __name__ == '__main__':
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    which = sys.argv[1:] or ['forward', 'coder', 'mae']
    if 'coder' in which:
        gen_coder()
    if 'mae' in which:
        gen_mae()
    if 'forward' in which:
        gen_forward()
