"""CPU: the oracle restatement against (a) fixtures produced by the reference's own python code
(oracle/gen_golden.py), (b) the reference's known-answer test for the box coder, (c) torch /
torchvision primitives it restates."""
import glob
import json
import math
import os

import numpy as np
import pytest
import torch

from oracle import mcgaze_oracle as O

TOL = 2e-5   # fp32 CPU, different op order than the reference's nn.Modules


@pytest.mark.parametrize('name', ['t7_224', 't3_192x224_rescale', 't1_224'])
def test_forward_matches_reference_fixture(golden_dir, synthetic_sd, name):
    g = np.load(os.path.join(golden_dir, f'golden_forward_{name}.npz'))
    T, H, W = int(g['T']), int(g['H']), int(g['W'])
    img = O.make_clip(int(g['seed']), T, H, W)
    img_hw = torch.tensor([[float(g['img_hw'][0]), float(g['img_hw'][1])]] * T)
    scale = torch.tensor(g['scale'])[None].repeat(T, 1)
    out = O.forward(synthetic_sd, img, clip_length=T, img_hw=img_hw, scale_factor=scale)
    for k in ('gaze_score', 'face_gaze_score', 'eyes_gaze_score', 'head_gaze_score'):
        assert np.abs(out[k].numpy() - g[k]).max() < TOL, k
    det = g['det_bboxes']
    assert np.abs(out['boxes'].numpy() - det[..., :4]).max() < 2e-3      # pixels, values up to ~300
    assert np.abs(out['scores'].numpy() - det[..., 4]).max() < TOL


def test_delta2bbox_reference_known_answer(golden_dir):
    g = json.load(open(os.path.join(golden_dir, 'golden_coder.json')))
    out = O.delta2bbox(torch.tensor(g['kat_rois']), torch.tensor(g['kat_deltas']), max_shape=(32, 32))
    assert torch.tensor(g['kat_expected']).allclose(out, atol=1e-4)      # tests/test_utils/test_coder.py:27-42
    out2 = O.delta2bbox(torch.tensor(g['cfg_rois']), torch.tensor(g['cfg_deltas']), stds=O.BBOX_STDS,
                        clip_border=False)
    assert torch.tensor(g['cfg_out']).allclose(out2, atol=1e-4, rtol=1e-6)
    assert O.delta2bbox(torch.zeros(0, 4), torch.zeros(0, 4)).shape == (0, 4)


def test_roi_align_restatement_vs_torchvision():
    from torchvision.ops import roi_align
    g = torch.Generator().manual_seed(0)
    feat = torch.randn(3, 16, 20, 24, generator=g)
    rois = torch.tensor([[0, 1.0, 2.0, 60.0, 50.0], [1, -20.0, -8.0, 30.0, 200.0], [2, 10.0, 10.0, 10.5, 10.2],
                         [2, 70.0, 60.0, 120.0, 90.0], [0, 0.0, 0.0, 96.0, 80.0]])
    for scale in (0.25, 0.125):
        a = O.roi_align_ref(feat, rois, scale)
        b = roi_align(feat, rois, (7, 7), scale, 2, aligned=True)
        assert (a - b).abs().max() < 1e-5


def test_mha_restatement_vs_torch(synthetic_sd):
    p = 'roi_head.bbox_head.0.attention'
    m = torch.nn.MultiheadAttention(256, 8)
    m.load_state_dict({'in_proj_weight': synthetic_sd[p + '.attn.in_proj_weight'],
                       'in_proj_bias': synthetic_sd[p + '.attn.in_proj_bias'],
                       'out_proj.weight': synthetic_sd[p + '.attn.out_proj.weight'],
                       'out_proj.bias': synthetic_sd[p + '.attn.out_proj.bias']})
    x = torch.randn(5, 4, 256, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        ref = x + m(x, x, x, need_weights=False)[0]
    assert (O.mha_residual(x, synthetic_sd, p) - ref).abs().max() < 1e-5


def test_batched_clips_equal_per_clip(synthetic_sd):
    """B > 1 with clip_length = T (forward_train's convention, multiclue_gaze_roi_head.py:229) must
    equal running the reference's one-clip test path per clip."""
    T = 3
    a, b = O.make_clip(11, T, 64, 96), O.make_clip(12, T, 64, 96)
    both = O.forward(synthetic_sd, torch.cat([a, b]), clip_length=T)
    for i, clip in enumerate((a, b)):
        one = O.forward(synthetic_sd, clip)
        for k in ('gaze_score', 'face_gaze_score', 'boxes', 'scores'):
            assert (both[k][i * T:(i + 1) * T] - one[k]).abs().max() < 1e-4, k


def test_level_mapping_edges():
    b = torch.tensor([[0, 0, 111.9, 111.9], [0, 0, 112.0, 112.0], [0, 0, 224, 224], [0, 0, 448, 448],
                      [0, 0, 2000, 2000], [5, 5, 5, 5]], dtype=torch.float32)
    assert O.map_roi_levels(b).tolist() == [0, 1, 2, 3, 3, 0]


def test_vector_to_yaw_pitch_matches_reference_formula():
    v = torch.tensor([[0.0, 0.0, -1.0], [1.0, 0.0, 0.0], [0.3, 0.4, -0.5]])
    yp = O.vector_to_yaw_pitch(v)
    assert abs(yp[0, 0]) < 1e-7 and abs(yp[0, 1]) < 1e-7
    assert abs(yp[1, 0] - math.pi / 2) < 1e-6
    n = v[2] / v[2].norm()
    assert abs(yp[2, 1] - math.asin(n[1])) < 1e-6


def test_all_forward_fixtures_present(golden_dir):
    assert len(glob.glob(os.path.join(golden_dir, 'golden_forward_*.npz'))) == 3
