"""Test-time image pipeline: oracle vs the reference's transforms (goldens) and vs cv2,
host geometry of GpuTestPipeline, and -- on the GPU -- mcg_preprocess bit-exact against the oracle / goldens."""
import os

import numpy as np
import pytest

from mcgaze_b200.compat import Config
from mcgaze_b200.pipeline import GpuTestPipeline
from oracle import preprocess_oracle as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MEAN = [123.675, 116.28, 103.53]
STD = [58.395, 57.12, 57.375]
CFG = {'gaze360': 'configs/multiclue_gaze/multiclue_gaze_r50_gaze360.py',
       'l2cs': 'configs/multiclue_gaze/multiclue_gaze_r50_l2cs.py'}


@pytest.fixture(scope='module')
def golden(golden_dir):
    g = np.load(os.path.join(golden_dir, 'golden_preprocess.npz'))
    return {n: {k.split('.', 1)[1]: g[k] for k in g.files if k.startswith(n + '.')}
            for n in sorted({k.split('.')[0] for k in g.files})}


def pipeline_cfg(name):
    return Config.fromfile(os.path.join(ROOT, CFG[name])).data.test.pipeline


def oracle_frame(src, name, rand):
    l2 = name.startswith('l2cs')
    return P.preprocess_frame(src, (448, 448) if l2 else (224, 224), MEAN, STD, True,
                              None if l2 else (0.68, 0.68), rand)


# ------------------------------------------------------------------------------------------ CPU: oracle pinning
def test_oracle_matches_reference_transforms(golden):
    """goldens = the reference's own CenterCrop/Resize/RandomFlip/Normalize/Pad classes over cv2
    (oracle/gen_golden_preprocess.py); the restatement must agree bit for bit."""
    assert len(golden) >= 6
    for name, g in golden.items():
        o = oracle_frame(g['src'], name, float(g['rand']))
        assert tuple(g['img_shape']) == o['img_shape'], name
        assert tuple(g['pad_shape']) == o['pad_shape'], name
        assert np.array_equal(g['scale_factor'], o['scale_factor']), name
        assert o['img'].dtype == np.float32 and np.array_equal(g['img'], o['img']), name


def test_oracle_resize_and_normalize_match_cv2_live():
    cv2 = pytest.importorskip('cv2')
    rng = np.random.default_rng(0)
    for it in range(60):
        sh, sw = (int(v) for v in rng.integers(1, 400, 2))
        dh, dw = (int(v) for v in rng.integers(1, 300, 2))
        if it % 5 == 0:
            dh, dw = max(sh // 2, 1), max(sw // 2, 1)          # OpenCV's exact-2x shortcut (INTER_AREA path)
        if it % 7 == 0:
            dh, dw = sh, sw                                      # identity
        img = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        ref = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(ref, P.resize_linear_u8(img, dw, dh)), ((sh, sw), (dh, dw))
    allv = np.repeat(np.arange(256, dtype=np.uint8)[:, None, None], 3, 2).reshape(16, 16, 3)
    for to_rgb in (True, False):
        f = allv.copy().astype(np.float32)
        if to_rgb:
            cv2.cvtColor(f, cv2.COLOR_BGR2RGB, f)
        cv2.subtract(f, np.float64(np.float32(MEAN).reshape(1, -1)), f)
        cv2.multiply(f, 1 / np.float64(np.float32(STD).reshape(1, -1)), f)
        assert np.array_equal(f, P.imnormalize(allv, MEAN, STD, to_rgb))


def test_oracle_edge_cases():
    one = np.full((1, 1, 3), 200, np.uint8)
    assert np.all(P.resize_linear_u8(one, 5, 4) == 200)                      # 1x1 source
    img = np.arange(2 * 3 * 3, dtype=np.uint8).reshape(2, 3, 3)
    assert np.array_equal(P.resize_linear_u8(img, 3, 2), img)                # identity
    assert P.center_crop_window(10, 10, 20, 20) == (0, 0, 10, 10)            # crop larger than the image
    assert P.pad_size(224, 186) == (224, 192) and P.pad_size(224, 224) == (224, 224)
    assert P.rescale_size(100, 120, (224, 224)) == (187, 224)


# ------------------------------------------------------------------------------------------ CPU: host logic
def test_pipeline_geometry_matches_goldens(golden):
    for name, g in golden.items():
        pipe = GpuTestPipeline(pipeline_cfg('l2cs' if name.startswith('l2cs') else 'gaze360'))
        h, w = g['src'].shape[:2]
        geometry, metas, (Hp, Wp) = pipe.plan([(h, w)], [float(g['rand'])])
        assert metas[0]['img_shape'] == tuple(g['img_shape']) and metas[0]['pad_shape'] == tuple(g['pad_shape'])
        assert np.array_equal(metas[0]['scale_factor'], g['scale_factor'])
        assert (Hp, Wp) == tuple(g['pad_shape'][:2]) and metas[0]['ori_shape'] == (h, w, 3)
        y, x, ch, cw, nh, nw = geometry[0]
        if name.startswith('l2cs'):
            assert (y, x, ch, cw) == (0, 0, h, w)
        else:
            ech, ecw = P.center_crop_size(h, w, (0.68, 0.68), float(g['rand']))
            y1, x1, y2, x2 = P.center_crop_window(h, w, ech, ecw)
            assert (y, x, ch, cw) == (y1, x1, y2 - y1, x2 - x1)


@pytest.mark.skipif(not os.path.exists('/root/reference/configs/multiclue_gaze'), reason='reference tree not mounted')
def test_reference_pipeline_configs_build_unchanged():
    """the reference's own config files (Gaze360- and l2cs-setting) give the same pipeline as this repo's copies"""
    for name in ('gaze360', 'l2cs'):
        ref = Config.fromfile(os.path.join('/root/reference', CFG[name])).data.test.pipeline
        ours = pipeline_cfg(name)
        assert [dict(d) for d in ref] == [dict(d) for d in ours]
        p = GpuTestPipeline(ref)
        assert p.scale == ((448, 448) if name == 'l2cs' else (224, 224)) and p.size_divisor == 32 and p.to_rgb
        assert (p.crop is None) == (name == 'l2cs')


def test_vectorised_plan_equals_the_literal_per_frame_arithmetic():
    rng = np.random.default_rng(7)
    shapes = [(int(a), int(b)) for a, b in rng.integers(8, 2000, (500, 2))]
    rands = [float(v) for v in rng.random(500)]
    cfgs = [pipeline_cfg('gaze360'), pipeline_cfg('l2cs'),
            [dict(type='CenterCrop', crop_size=(0.5, 0.8), crop_type='relative'), dict(type='Resize', img_scale=(320, 200), keep_ratio=False),
             dict(type='Normalize', mean=MEAN, std=STD), dict(type='Pad', size=(256, 352))],
            [dict(type='CenterCrop', crop_size=(100, 300), crop_type='absolute'), dict(type='Resize', img_scale=(333, 200), keep_ratio=True),
             dict(type='Normalize', mean=MEAN, std=STD)]]
    for cfg in cfgs:
        p = GpuTestPipeline(cfg)
        g1, m1, c1 = p.plan(shapes, rands)
        g2, m2, c2 = p.plan_scalar(shapes, rands)
        assert g1 == g2 and c1 == c2
        for a, b in zip(m1, m2):
            assert a['img_shape'] == b['img_shape'] and a['pad_shape'] == b['pad_shape'] and a['ori_shape'] == b['ori_shape']
            assert np.array_equal(a['scale_factor'], b['scale_factor'])


def test_pipeline_random_draws_follow_numpy_like_the_reference():
    """CenterCrop draws np.random.rand(1) per call (transforms.py:1129): seeding numpy pins the GPU pipeline's crops
    exactly as it pins the reference's."""
    pipe = GpuTestPipeline(pipeline_cfg('gaze360'))
    np.random.seed(3)
    r = [float(np.random.rand(1)[0]) for _ in range(3)]
    np.random.seed(3)
    g1, _, _ = pipe.plan([(300, 280)] * 3)
    g2, _, _ = pipe.plan([(300, 280)] * 3, r)
    assert g1 == g2 and len({g[:4] for g in g1}) > 1
    a = GpuTestPipeline(pipeline_cfg('gaze360'), seed=5).plan([(300, 280)] * 3)[0]
    b = GpuTestPipeline(pipeline_cfg('gaze360'), seed=5).plan([(300, 280)] * 3)[0]
    assert a == b


def test_pipeline_rejects_what_it_does_not_implement():
    base = [dict(type='Resize', img_scale=(224, 224), keep_ratio=True), dict(type='Normalize', mean=MEAN, std=STD)]
    with pytest.raises(NotImplementedError):
        GpuTestPipeline(base + [dict(type='RandomFlip', flip_ratio=0.5)])
    with pytest.raises(NotImplementedError):
        GpuTestPipeline(base + [dict(type='PhotoMetricDistortion')])
    with pytest.raises(ValueError):
        GpuTestPipeline([dict(type='Pad', size_divisor=32)])
    p = GpuTestPipeline([dict(type='Resize', img_scale=(320, 200), keep_ratio=False),
                         dict(type='Normalize', mean=MEAN, std=STD, to_rgb=False)])
    assert p.resized_size(50, 60) == (200, 320) and p.padded_size(200, 320) == (200, 320)


def test_pipeline_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    pipe = GpuTestPipeline(pipeline_cfg('gaze360'))
    with pytest.raises(Exception, match='CUDA'):
        pipe.batch([np.zeros((32, 32, 3), np.uint8)])


# ------------------------------------------------------------------------------------------ GPU: parity (bit-exact)
@pytest.mark.gpu
def test_gpu_matches_goldens_bit_exact(golden):
    for name, g in golden.items():
        pipe = GpuTestPipeline(pipeline_cfg('l2cs' if name.startswith('l2cs') else 'gaze360'))
        res = pipe.batch([g['src']], rands=[float(g['rand'])])
        got = res['img'][0][0].cpu().numpy()
        assert got.shape == g['img'].shape and np.array_equal(got, g['img']), name
        assert res['img_metas'][0][0]['img_shape'] == tuple(g['img_shape'])


@pytest.mark.gpu
def test_gpu_ragged_batch_matches_oracle_and_pads_with_zeros(golden):
    """All Gaze360 goldens in ONE call: frames of different sizes share a canvas (the reference's collate pads to
    the largest frame); outside each frame's own padded area there are zeros as well."""
    names = [n for n in golden if n.startswith('gaze360')]
    pipe = GpuTestPipeline(pipeline_cfg('gaze360'))
    res = pipe.batch([golden[n]['src'] for n in names], rands=[float(golden[n]['rand']) for n in names])
    out = res['img'][0].cpu().numpy()
    assert out.shape[2:] == (224, 224)
    for i, n in enumerate(names):
        ref = golden[n]['img']
        canvas = np.zeros(out.shape[1:], np.float32)
        canvas[:, :ref.shape[1], :ref.shape[2]] = ref
        assert np.array_equal(out[i], canvas), n


@pytest.mark.gpu
def test_gpu_random_geometry_vs_oracle():
    import torch
    from mcgaze_b200 import lib
    rng = np.random.default_rng(4)
    frames, geo, refs = [], [], []
    for it in range(40):
        sh, sw = (int(v) for v in rng.integers(1, 300, 2))
        src = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        ch, cw = int(rng.integers(1, sh + 1)), int(rng.integers(1, sw + 1))
        y, x = int(rng.integers(0, sh - ch + 1)), int(rng.integers(0, sw - cw + 1))
        dh, dw = int(rng.integers(1, 257)), int(rng.integers(1, 321))
        if it % 6 == 0:
            dh, dw = ch, cw
        if it % 9 == 0:
            dh, dw = max(ch // 2, 1), max(cw // 2, 1)
        frames.append(src)
        geo.append((y, x, ch, cw, dh, dw))
        refs.append(P.resize_linear_u8(src[y:y + ch, x:x + cw], dw, dh))
    for to_rgb in (True, False):
        # sources live inside wider device buffers: the row stride differs from 3 * w
        dev = [torch.from_numpy(np.pad(f, ((0, 0), (0, 5), (0, 0)))).cuda()[:, :f.shape[1]] for f in frames]
        out = torch.full((len(frames), 3, 256, 320), 7.0, device='cuda')
        lib.preprocess(dev, geo, MEAN, STD, to_rgb, out)
        got = out.cpu().numpy()
        for i, (r, g) in enumerate(zip(refs, geo)):
            exp = np.zeros((3, 256, 320), np.float32)
            exp[:, :g[4], :g[5]] = P.imnormalize(r, MEAN, STD, to_rgb).transpose(2, 0, 1)
            assert np.array_equal(got[i], exp), (i, g, to_rgb)


@pytest.mark.gpu
def test_gpu_invalid_arguments_are_rejected_with_error_codes():
    """C-ABI error behaviour: negative code + message, nothing launched, nothing thrown across the ABI."""
    import torch
    from mcgaze_b200 import lib
    src = torch.zeros((40, 50, 3), dtype=torch.uint8, device='cuda')
    out = torch.zeros((1, 3, 32, 64), device='cuda')
    good = (0, 0, 40, 50, 26, 32)
    lib.preprocess([src], [good], MEAN, STD, True, out)
    for bad in [(0, 0, 41, 50, 26, 32),        # crop window leaves the source
                (-1, 0, 40, 50, 26, 32),
                (0, 0, 40, 50, 33, 32),        # resized frame does not fit the canvas
                (0, 0, 40, 50, 26, 0)]:
        with pytest.raises(lib.McgError, match='code -1'):
            lib.preprocess([src], [bad], MEAN, STD, True, out)
    with pytest.raises(lib.McgError):
        lib.preprocess([src], [good], MEAN, STD, True, torch.zeros((1, 3, 32, 62), device='cuda'))   # Wp % 4
    with pytest.raises(lib.McgError):
        lib.preprocess([src.float()], [good], MEAN, STD, True, out)                                  # not uint8
    torch.cuda.synchronize()


@pytest.mark.gpu
def test_gpu_full_batch_properties():
    """BASELINE configs[1] size (224 frames -> 224^2): batch independence (a frame's result does not depend on its
    neighbours or on the launch chunking, bit-exact), constant images map to the normalisation table, identity
    geometry equals plain normalisation."""
    import torch
    rng = np.random.default_rng(9)
    pipe = GpuTestPipeline(pipeline_cfg('gaze360'), seed=1)
    srcs = [rng.integers(0, 256, (int(rng.integers(200, 420)), int(rng.integers(200, 420)), 3), dtype=np.uint8)
            for _ in range(8)]
    frames = [srcs[i % 8] for i in range(224)]
    rands = [float(v) for v in rng.random(224)]
    full = pipe.batch(frames, rands=rands)['img'][0]
    assert full.shape[0] == 224
    for i in (0, 95, 96, 200, 223):
        alone = pipe.batch([frames[i]], rands=[rands[i]])['img'][0][0]
        assert torch.equal(full[i, :, :alone.shape[1], :alone.shape[2]], alone)
    block = np.stack([srcs[0]] * 5)                                    # one [n, h, w, 3] block == a list of frames
    a = pipe.batch(block, rands=rands[:5])['img'][0]
    b = pipe.batch([srcs[0]] * 5, rands=rands[:5])['img'][0]
    assert torch.equal(a, b) and torch.equal(pipe.batch(torch.from_numpy(block).cuda(), rands=rands[:5])['img'][0], a)
    const = np.full((300, 300, 3), 0, np.uint8)
    const[..., 0], const[..., 1], const[..., 2] = 10, 100, 250          # B, G, R
    c = pipe.batch([const], rands=[0.5])['img'][0][0].cpu().numpy()
    lut = P.imnormalize(const[:1, :1], MEAN, STD, True)[0, 0]
    assert all(np.all(c[p] == lut[p]) for p in range(3))
    ident = GpuTestPipeline([dict(type='Resize', img_scale=(96, 64), keep_ratio=False),
                             dict(type='Normalize', mean=MEAN, std=STD, to_rgb=True)])
    img = rng.integers(0, 256, (64, 96, 3), dtype=np.uint8)
    got = ident.batch([img])['img'][0][0].cpu().numpy()
    assert np.array_equal(got, P.imnormalize(img, MEAN, STD, True).transpose(2, 0, 1))


@pytest.mark.gpu
def test_gpu_pipeline_feeds_the_forward(synthetic_sd):
    """uint8 frames -> GpuTestPipeline -> MultiClueGaze forward, against oracle pipeline -> oracle forward:
    (yaw, pitch) within 1e-3 rad (BASELINE north_star tolerance)."""
    import torch
    from mcgaze_b200 import lib
    from oracle import mcgaze_oracle as O
    rng = np.random.default_rng(2)
    T = 3
    yy, xx = np.mgrid[0:260, 0:240]
    frames = [np.clip(np.stack([xx, yy, (xx + yy) // 2], -1) + rng.integers(-60, 60, (260, 240, 3)), 0, 255).astype(np.uint8)
              for _ in range(T)]
    rands = [0.3] * T
    pipe = GpuTestPipeline(pipeline_cfg('gaze360'))
    data = pipe.batch(frames, rands=rands)
    metas = data['img_metas'][0]
    ref_frames = [oracle_frame(f, 'gaze360', r) for f, r in zip(frames, rands)]
    img_ref = torch.from_numpy(np.stack([r['img'] for r in ref_frames]))
    assert torch.equal(data['img'][0].cpu(), img_ref)
    img_hw = torch.tensor([[m['img_shape'][0], m['img_shape'][1]] for m in metas], dtype=torch.float32)
    scale = torch.from_numpy(np.stack([m['scale_factor'] for m in metas]))
    ref = O.forward(synthetic_sd, img_ref, img_hw=img_hw, scale_factor=scale)
    eng = lib.Engine(synthetic_sd, 0, 'fp16c8')
    out = eng.forward(data['img'][0], img_hw=img_hw.numpy(), scale_factor=scale.numpy())
    torch.cuda.synchronize()
    for i, k in enumerate(('gaze_score', 'face_gaze_score', 'eyes_gaze_score', 'head_gaze_score')):
        d = (O.vector_to_yaw_pitch(out['gaze'][:, i].cpu()) - O.vector_to_yaw_pitch(ref[k])).abs()
        assert float(torch.minimum(d, 2 * torch.pi - d).max()) < 1e-3, k
