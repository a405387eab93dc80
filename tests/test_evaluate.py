"""Batched / sharded evaluation driver (SURVEY.md section 8 rows f1 + f2).

CPU: a stub model with the reference's call contract checks the driver logic (batching many clips per forward,
DistributedSampler-style sharding over a world_size-2 gloo group, one gather, overlap merge, JSON records, scorer)
against a literal one-clip-per-forward loop in the reference's order.  GPU: the real backend, batched vs per-clip."""
import os
import socket

import numpy as np
import pytest
import torch

from mcgaze_b200 import evaluate as ev
from mcgaze_b200 import slicer

LENGTHS = [1, 5, 7, 8, 12, 23, 9, 7, 15]


def make_anno(lengths=LENGTHS, seed=0):
    rng = np.random.default_rng(seed)
    videos, anns = [], []
    for vi, L in enumerate(lengths):
        videos.append(dict(id=vi + 1, file_names=[f'v{vi:03d}/{t:05d}.png' for t in range(L)]))
        g = rng.normal(size=(L, 3))
        anns.append(dict(video_id=vi + 1, gaze=(g / np.linalg.norm(g, axis=1, keepdims=True)).tolist()))
    return dict(videos=videos, annotations=anns)


def fake_loader(path):
    """A tiny decoded 'frame' whose content encodes the file name."""
    v, t = path.split('/')[-2:]
    code = int(v[1:]) * 1000 + int(t.split('.')[0])
    img = np.zeros((4, 4, 3), np.uint8)
    img[..., 0], img[..., 1], img[..., 2] = code % 251, (code // 251) % 251, 7
    return img


class StubPipeline:
    """frames -> [n, 3, 4, 4] float tensor (no resize), metas with identity scale."""

    def batch(self, frames, filenames=None):
        img = torch.from_numpy(np.asarray(frames).astype(np.float32)).permute(0, 3, 1, 2).contiguous()
        metas = [dict(img_shape=(4, 4, 3), scale_factor=np.ones(4, np.float32), filename=f) for f in (filenames or [None] * len(frames))]
        return dict(img=[img], img_metas=[metas])


class StubModel:
    """Reference call contract; outputs depend on the frame AND on its position in its clip (like the temporal
    attention makes them), so clip batching / ordering mistakes change the result."""

    def __init__(self):
        self.calls = []

    def __call__(self, return_loss, rescale, format, img, img_metas, clip_length=None):
        x = img[0]
        n = x.shape[0]
        T = clip_length or n
        self.calls.append((n, T))
        code = x[:, 0, 0, 0] + 251 * x[:, 1, 0, 0]
        pos = torch.arange(n) % T
        clip_sum = code.view(-1, T).sum(1).repeat_interleave(T)
        base = (code * 0.37 + pos * 1.3 + clip_sum * 0.011)
        boxes = torch.stack([base + c * 10 + k for c in range(3) for k in range(4)], 1).view(n, 3, 4)
        scores = torch.sigmoid(torch.stack([torch.sin(base + c) * 3 for c in range(3)], 1))
        gaze = torch.stack([torch.cos(base * (i + 1) * 0.1) for i in range(12)], 1).view(n, 4, 3)
        gaze = gaze / gaze.norm(dim=-1, keepdim=True)
        det = [torch.cat([boxes[i], scores[i][:, None]], 1) for i in range(n)]
        return (det, [[0, 1, 2]] * n), {'gaze_score': gaze[:, 0], 'face_gaze_score': gaze[:, 1],
                                        'eyes_gaze_score': gaze[:, 2], 'head_gaze_score': gaze[:, 3]}


def reference_loop(model, dataset, pipeline):
    """tools/test_gaze360_gaze.py:60-260 in its own order: one clip per forward, merge, records."""
    rows = []
    for i in range(len(dataset)):
        rows.append(ev.run_clips(model, dataset, pipeline, [i], clips_per_batch=1)[i])
    return ev.videos_from_clips(dataset, rows)


def test_prefetching_loader_handles_ragged_frames():
    """frames of one size travel as one staged block, mixed sizes as a list; both give the per-frame results"""
    seen = []

    class ShapePipeline:
        def batch(self, frames, filenames=None):
            seen.append('block' if hasattr(frames, 'shape') else 'list')
            imgs = [torch.from_numpy(np.asarray(f)[:4, :4].astype(np.float32)).permute(2, 0, 1) for f in frames]
            metas = [dict(img_shape=(4, 4, 3), scale_factor=np.ones(4, np.float32), filename=n) for n in filenames]
            return dict(img=[torch.stack(imgs)], img_metas=[metas])

    def ragged_loader(path):
        img = fake_loader(path)
        return np.pad(img, ((0, 2), (0, 0), (0, 0))) if path.startswith('v003') else img

    ds_a = ev.Gaze360ClipDataset(make_anno(), loader=ragged_loader)
    ds_b = ev.Gaze360ClipDataset(make_anno(), loader=fake_loader)
    a = ev.single_gpu_test(StubModel(), ds_a, ShapePipeline(), clips_per_batch=4, workers=2)
    assert 'block' in seen and 'list' in seen
    b = ev.single_gpu_test(StubModel(), ds_b, StubPipeline(), clips_per_batch=4)
    assert all(np.allclose(x, y, rtol=1e-6, atol=1e-6) for x, y in zip(a, b))


def test_dataset_plans_clips_like_the_reference():
    ds = ev.Gaze360ClipDataset(make_anno(), loader=fake_loader)
    assert len(ds) == sum(len(slicer.plan_clips(L)) for L in LENGTHS)
    info = ds.clip_info(len(ds) - 1)
    assert info['video'] == len(LENGTHS) - 1 and info['start'] + info['n'] == LENGTHS[-1]
    item = ds[0]
    assert len(item['frames']) == item['n'] == 1 and item['frames'][0].dtype == np.uint8
    with pytest.raises(NotImplementedError):
        ev.Gaze360ClipDataset(make_anno(), test_mode=False)


def test_batched_driver_equals_one_clip_per_forward():
    ds = ev.Gaze360ClipDataset(make_anno(), loader=fake_loader)
    ref_records, ref_merged = reference_loop(StubModel(), ds, StubPipeline())
    model = StubModel()
    rows = ev.single_gpu_test(model, ds, StubPipeline(), clips_per_batch=4)
    rows_w = ev.single_gpu_test(StubModel(), ds, StubPipeline(), clips_per_batch=4, workers=3)     # prefetching loader
    assert all(np.array_equal(a, b) for a, b in zip(rows, rows_w))
    records, merged = ev.videos_from_clips(ds, rows)
    assert len(model.calls) < len(ds)                                  # many clips per forward
    assert all(n % T == 0 for n, T in model.calls)
    for a, b in zip(merged, ref_merged):       # (torch's vectorised sin / cos differ in the last bit with the batch shape)
        assert np.allclose(a['det'], b['det'], rtol=1e-6, atol=1e-6) and np.allclose(a['gaze'], b['gaze'], rtol=1e-6, atol=1e-6)
    for r, q in zip(records, ref_records):
        assert r.keys() == q.keys() and r['video_id'] == q['video_id']
        assert [b is None for b in r['face_bboxes']] == [b is None for b in q['face_bboxes']]
        assert np.allclose(r['fusion_gazes'], q['fusion_gazes'], atol=1e-6)
    assert [r['video_id'] for r in records] == list(range(1, len(LENGTHS) + 1))
    assert all(len(r['fusion_gazes']) == L for r, L in zip(records, LENGTHS))
    m = ev.evaluate(ds, records)
    assert m['frames_360'] == sum(LENGTHS) and 0 < m['mae_360'] < 180


def _worker(rank, world, port, out_q):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    ds = ev.Gaze360ClipDataset(make_anno(), loader=fake_loader)
    model = StubModel()
    rows = ev.multi_gpu_test(model, ds, StubPipeline(), clips_per_batch=3)
    out_q.put((rank, [r.tolist() for r in rows], sum(n // T for n, T in model.calls)))
    dist.destroy_process_group()


def test_two_rank_gloo_multi_gpu_test():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(2):
        r, rows, nclips = q.get(timeout=180)
        got[r] = (rows, nclips)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ds = ev.Gaze360ClipDataset(make_anno(), loader=fake_loader)
    want = ev.single_gpu_test(StubModel(), ds, StubPipeline(), clips_per_batch=5)
    per = -(-len(ds) // 2)
    for r in range(2):
        rows, nclips = got[r]
        assert nclips <= per                                            # each rank ran only its shard
        assert len(rows) == len(ds)
        for a, b in zip(rows, want):
            assert np.allclose(np.asarray(a, dtype=np.float32), b, rtol=1e-6, atol=1e-6)


@pytest.mark.gpu
def test_gpu_batched_evaluation_matches_per_clip(synthetic_sd):
    """Real backend + GPU pipeline: many clips per forward == one clip per forward (clips are independent units)."""
    from mcgaze_b200.apis import init_detector
    from mcgaze_b200.pipeline import GpuTestPipeline
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    model = init_detector(os.path.join(root, 'configs/multiclue_gaze/multiclue_gaze_r50_gaze360.py'), None, 'cuda:0')
    model.load_state_dict(synthetic_sd)
    cfg = model.cfg
    rng = np.random.default_rng(0)
    base = rng.integers(0, 256, (96, 88, 3), dtype=np.uint8)

    def loader(path):
        v, t = path.split('/')[-2:]
        k = int(v[1:]) * 37 + int(t.split('.')[0])
        return np.roll(base, k, axis=1) ^ np.uint8(k & 31)

    anno = make_anno([7, 11, 3, 8])
    ds = ev.Gaze360ClipDataset(anno, loader=loader)
    rands = {}

    class FixedCrops(GpuTestPipeline):              # same crop for a frame whichever batch it is in
        def batch(self, frames, filenames=None, **kw):
            r = [rands.setdefault(f, 0.1 + 0.8 * ((hash(f) % 97) / 97)) for f in filenames]
            return super().batch(frames, rands=r, filenames=filenames)

    pipe = FixedCrops(cfg.data.test.pipeline)
    batched = ev.single_gpu_test(model, ds, pipe, clips_per_batch=3)
    single = [ev.run_clips(model, ds, pipe, [i], clips_per_batch=1)[i] for i in range(len(ds))]
    for a, b in zip(batched, single):
        assert a.shape == b.shape and np.allclose(a, b, atol=1e-4, rtol=1e-4)
        assert np.allclose(np.linalg.norm(a[:, 15:].reshape(-1, 3), axis=1), 1, atol=1e-4)
    records, _ = ev.videos_from_clips(ds, batched)
    assert [len(r['fusion_gazes']) for r in records] == [7, 11, 3, 8]
    assert ev.evaluate(ds, records)['frames_360'] == 29


@pytest.mark.gpu
def test_gpu_cli_writes_reference_schema_json(synthetic_sd, tmp_path):
    """tools/test_gaze360.py end to end: PNG frames on disk + a checkpoint file -> results JSON in the schema of
    tools/test_gaze360_gaze.py:210-260 (what calculate_mae_gaze360.py reads) + the MAE lines."""
    import json
    import subprocess
    import sys
    cv2 = pytest.importorskip('cv2')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    rng = np.random.default_rng(1)
    anno = make_anno([3, 9])
    for v in anno['videos']:
        for f in v['file_names']:
            os.makedirs(tmp_path / 'frames' / os.path.dirname(f), exist_ok=True)
            assert cv2.imwrite(str(tmp_path / 'frames' / f), rng.integers(0, 256, (120, 100, 3), dtype=np.uint8))
    json.dump(anno, open(tmp_path / 'test.json', 'w'))
    torch.save({'state_dict': synthetic_sd, 'meta': {}}, tmp_path / 'ckpt.pth')
    r = subprocess.run([sys.executable, os.path.join(root, 'tools', 'test_gaze360.py'),
                        os.path.join(root, 'configs/multiclue_gaze/multiclue_gaze_r50_gaze360.py'), str(tmp_path / 'ckpt.pth'),
                        '--json', str(tmp_path / 'test.json'), '--root', str(tmp_path / 'frames'), '--seed', '0',
                        '--clips-per-batch', '2'], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert 'fusion_gazes: MAE 360' in r.stdout
    res = json.load(open(tmp_path / 'results' / 'results_multiclue_gaze_r50_gaze360_test.json'))
    assert [v['video_id'] for v in res] == [1, 2] and [len(v['fusion_gazes']) for v in res] == [3, 9]
    for v in res:
        assert set(v) == {'video_id', 'category_id', 'fusion_gazes', 'face_bboxes', 'face_gazes', 'face_score',
                          'eyes_bboxes', 'eyes_gazes', 'eyes_score', 'head_bboxes', 'head_gazes', 'head_score'}
        assert all(b is None or len(b) == 4 for b in v['head_bboxes'])


@pytest.mark.gpu
def test_gpu_device_scorer_matches_host_scorer_on_driver_output():
    """evaluate_on_device (mcg_gaze_error) == evaluate (host restatement of calculate_mae_gaze360.py) on the merged
    output of the driver (stub forward: the scorer only sees the merged gaze arrays)."""
    ds = ev.Gaze360ClipDataset(make_anno(), loader=fake_loader)
    rows = ev.single_gpu_test(StubModel(), ds, StubPipeline(), clips_per_batch=4)
    records, merged = ev.videos_from_clips(ds, rows)
    host = ev.evaluate(ds, records)
    dev = ev.evaluate_on_device(ds, merged)
    for k in host:
        assert abs(host[k] - dev[k]) < 2e-3, (k, host[k], dev[k])


# ------------------------------------------------------------------------------------------------------------------
# video-sharded run: merge + score where the results are, one all-reduce of 6 doubles
# ------------------------------------------------------------------------------------------------------------------
def _video_worker(rank, world, port, out_q, lengths=None):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    ds = ev.Gaze360ClipDataset(make_anno(lengths or LENGTHS), loader=fake_loader)
    model = StubModel()
    out = ev.multi_gpu_test_videos(model, ds, StubPipeline(), clips_per_batch=3)
    out_q.put((rank, out['mae'], out['sums'], out['videos_local'], [(m['det'].tolist(), m['gaze'].tolist()) for m in out['merged']],
               sum(n // T for n, T in model.calls)))
    dist.destroy_process_group()


def test_two_rank_gloo_video_sharded_mae():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_video_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(2):
        item = q.get(timeout=180)
        got[item[0]] = item[1:]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ds = ev.Gaze360ClipDataset(make_anno(), loader=fake_loader)
    rows = ev.single_gpu_test(StubModel(), ds, StubPipeline(), clips_per_batch=5)
    records, merged = ev.videos_from_clips(ds, rows)
    want = ev.evaluate(ds, records)
    assert got[0][2] == [0, 2, 4, 6, 8] and got[1][2] == [1, 3, 5, 7]           # videos, not clips, are sharded
    assert got[0][4] + got[1][4] == len(ds) and max(got[0][4], got[1][4]) < len(ds)
    for r in range(2):
        mae, sums, _, merged_r, _ = got[r]
        assert len(sums) == 6 and sums[1] == sum(LENGTHS)
        for k in want:
            assert abs(mae[k] - want[k]) < 1e-6, (r, k, mae[k], want[k])
        for (det, gz), m in zip(merged_r, merged):
            assert np.allclose(np.asarray(det, np.float32), m['det'], rtol=1e-6, atol=1e-6)
            assert np.allclose(np.asarray(gz, np.float32), m['gaze'], rtol=1e-6, atol=1e-6)


def test_three_rank_gloo_video_sharding_with_skewed_shards():
    """Round-robin over videos of very different lengths: one rank holds far more frames than total / world + longest
    video (the bound a first version of the gather buffer used: a rank failed before the all-gather and the others
    waited for it until the NCCL timeout).  Every rank must come back with all videos."""
    import torch.multiprocessing as mp
    lengths = [30, 1, 1, 30, 2, 1, 30, 1, 3, 23]
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_video_worker, args=(r, 3, port, q, lengths)) for r in range(3)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(3):
        item = q.get(timeout=180)
        got[item[0]] = item[1:]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ds = ev.Gaze360ClipDataset(make_anno(lengths), loader=fake_loader)
    rows = ev.single_gpu_test(StubModel(), ds, StubPipeline(), clips_per_batch=5)
    records, merged = ev.videos_from_clips(ds, rows)
    want = ev.evaluate(ds, records)
    for r in range(3):
        mae, sums, mine, merged_r, _ = got[r]
        assert mine == list(range(r, len(lengths), 3)) and sums[1] == sum(lengths)
        assert all(abs(mae[k] - want[k]) < 1e-6 for k in want)
        assert all(np.allclose(np.asarray(det, np.float32), m['det'], rtol=1e-6, atol=1e-6) for (det, _), m in zip(merged_r, merged))


def test_video_sharded_run_without_a_process_group_and_l2cs_variant():
    """world = 1 (no torch.distributed): same numbers as the clip-level driver; the l2cs variant reads annotation 3k."""
    anno = make_anno()
    ds = ev.Gaze360ClipDataset(anno, loader=fake_loader)
    out = ev.multi_gpu_test_videos(StubModel(), ds, StubPipeline(), clips_per_batch=4)
    rows = ev.single_gpu_test(StubModel(), ds, StubPipeline(), clips_per_batch=4)
    records, merged = ev.videos_from_clips(ds, rows)
    want = ev.evaluate(ds, records)
    assert all(abs(out['mae'][k] - want[k]) < 1e-6 for k in want)
    recs = ev.records_from_merged(ds, out['merged'])
    assert [r['video_id'] for r in recs] == [r['video_id'] for r in records]
    assert all(np.allclose(a['fusion_gazes'], b['fusion_gazes'], atol=1e-6) for a, b in zip(recs, records))
    l2 = dict(videos=anno['videos'], annotations=[a if k == 0 else dict(gaze=a['gaze'][::-1]) for a in anno['annotations'] for k in range(3)])
    ds2 = ev.Gaze360ClipDataset(l2, loader=fake_loader)
    o2 = ev.multi_gpu_test_videos(StubModel(), ds2, StubPipeline(), clips_per_batch=4, variant='l2cs')
    w2 = ev.evaluate(ds2, records, variant='l2cs')
    assert all(abs(o2['mae'][k] - w2[k]) < 1e-6 for k in w2)
    assert o2['mae']['mae_360'] == pytest.approx(want['mae_360'], abs=1e-6)          # same GT for the 360 class
    assert o2['mae']['frames_front20'] <= want['frames_front20']                      # the extra pitch condition


def test_batches_keep_every_clip_on_its_own_canvas():
    """ADVICE (round 1, high): a clip must be padded to ITS OWN largest frame, never to its batch neighbours'.  A
    pipeline stand-in whose canvas depends on the frame size shows the driver splitting a mixed batch per canvas."""
    seen = []

    class CanvasPipeline(StubPipeline):
        device = 0

        def draw(self, n):
            return np.zeros(n)

        def clip_canvases(self, shapes, rands, T):
            hw = np.asarray(shapes).reshape(-1, T, 2).max(1)
            return [(int(h), int(w)) for h, w in hw]

        def batch(self, frames, rands=None, filenames=None):
            shapes = {tuple(np.asarray(f).shape[:2]) for f in frames}
            seen.append(shapes)
            imgs = [torch.from_numpy(np.asarray(f)[:4, :4].astype(np.float32)).permute(2, 0, 1) for f in frames]
            metas = [dict(img_shape=(4, 4, 3), scale_factor=np.ones(4, np.float32), filename=n) for n in filenames]
            return dict(img=[torch.stack(imgs)], img_metas=[metas])

    def loader(path):                    # videos 1 and 3 have taller frames
        img = fake_loader(path)
        return np.pad(img, ((0, 2), (0, 0), (0, 0))) if path[:4] in ('v001', 'v003') else img

    lengths = [7, 7, 7, 7]
    ds = ev.Gaze360ClipDataset(make_anno(lengths), loader=loader)
    rows = ev.single_gpu_test(StubModel(), ds, CanvasPipeline(), clips_per_batch=4)
    assert all(len(s) == 1 for s in seen) and len(seen) == 2            # one call per canvas, never mixed
    ref = ev.single_gpu_test(StubModel(), ev.Gaze360ClipDataset(make_anno(lengths), loader=fake_loader), StubPipeline(), clips_per_batch=1)
    assert all(np.allclose(a, b, rtol=1e-6, atol=1e-6) for a, b in zip(rows, ref))


@pytest.mark.gpu
def test_gpu_mixed_aspect_ratios_batched_equals_per_clip(synthetic_sd):
    """Videos of different aspect ratios in one batch: every clip still runs on its own padded canvas, so batched ==
    one clip per forward (the reference's collate pads a clip to its own largest frame)."""
    from mcgaze_b200.apis import init_detector
    from mcgaze_b200.pipeline import GpuTestPipeline
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    model = init_detector(os.path.join(root, 'configs/multiclue_gaze/multiclue_gaze_r50_gaze360.py'), None, 'cuda:0')
    model.load_state_dict(synthetic_sd)
    rng = np.random.default_rng(0)
    bases = {0: rng.integers(0, 256, (96, 96, 3), dtype=np.uint8), 1: rng.integers(0, 256, (120, 96, 3), dtype=np.uint8),
             2: rng.integers(0, 256, (90, 130, 3), dtype=np.uint8)}

    def loader(path):
        v, t = path.split('/')[-2:]
        k = int(v[1:]) * 37 + int(t.split('.')[0])
        return np.roll(bases[int(v[1:]) % 3], k, axis=1) ^ np.uint8(k & 31)

    ds = ev.Gaze360ClipDataset(make_anno([7, 7, 7, 11]), loader=loader)
    rands = {}

    class FixedCrops(GpuTestPipeline):
        def draw(self, n):
            return np.full(n, 0.5)

    pipe = FixedCrops(model.cfg.data.test.pipeline)
    canv = {pipe.clip_canvases([bases[v % 3].shape[:2]] * 7, np.full(7, 0.5), 7)[0] for v in range(3)}
    assert len(canv) == 3, canv                                                      # three different canvases
    batched = ev.single_gpu_test(model, ds, pipe, clips_per_batch=8)
    single = [ev.run_clips(model, ds, pipe, [i], clips_per_batch=1)[i] for i in range(len(ds))]
    for a, b in zip(batched, single):
        assert a.shape == b.shape and np.allclose(a, b, atol=1e-4, rtol=1e-4)
    del rands


@pytest.mark.gpu
def test_gpu_video_sharded_run_merges_and_scores_on_the_device(synthetic_sd):
    """multi_gpu_test_videos on one GPU (world = 1): device merge + device scorer == host merge + host scorer."""
    from mcgaze_b200.apis import init_detector
    from mcgaze_b200.pipeline import GpuTestPipeline
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    model = init_detector(os.path.join(root, 'configs/multiclue_gaze/multiclue_gaze_r50_gaze360.py'), None, 'cuda:0')
    model.load_state_dict(synthetic_sd)
    rng = np.random.default_rng(0)
    base = rng.integers(0, 256, (96, 88, 3), dtype=np.uint8)

    def loader(path):
        v, t = path.split('/')[-2:]
        k = int(v[1:]) * 37 + int(t.split('.')[0])
        return np.roll(base, k, axis=1) ^ np.uint8(k & 31)

    class FixedCrops(GpuTestPipeline):
        def draw(self, n):
            return np.full(n, 0.3)

    ds = ev.Gaze360ClipDataset(make_anno([7, 12, 3, 9]), loader=loader)
    pipe = FixedCrops(model.cfg.data.test.pipeline)
    out = ev.multi_gpu_test_videos(model, ds, pipe, clips_per_batch=3, workers=2)
    rows = ev.single_gpu_test(model, ds, pipe, clips_per_batch=3)
    records, merged = ev.videos_from_clips(ds, rows)
    want = ev.evaluate(ds, records)
    for k in want:
        assert abs(out['mae'][k] - want[k]) < 2e-3, (k, out['mae'][k], want[k])
    for a, b in zip(out['merged'], merged):
        assert np.allclose(a['det'], b['det'], atol=1e-4, rtol=1e-4) and np.allclose(a['gaze'], b['gaze'], atol=1e-5)
