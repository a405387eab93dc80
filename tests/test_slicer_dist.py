"""CPU: clip slicer / merger against a literal torch restatement of the reference's loop, and the
multi-rank sharding + gather over a world_size-2 gloo group."""
import math
import os
import socket

import numpy as np
import pytest
import torch

from mcgaze_b200 import dist as mdist
from mcgaze_b200 import slicer


def reference_merge(L, clip_out, clip_len=7, stride=4, thr=0.5):
    """Literal restatement of tools/test_gaze360_gaze.py:73-201 on precomputed per-clip outputs.
    clip_out(frames) -> (det_bboxes [T,3,5], fusion [T,1,3], other [T,3,3]) torch tensors."""
    clip_num = 1 if L <= clip_len else math.ceil((L - clip_len) / stride) + 1
    imgs = list(range(L))
    for ci in range(clip_num):
        if ci != clip_num - 1:
            cur = imgs[ci * stride:ci * stride + clip_len]
            ov = clip_len - stride
        else:
            cur = imgs[-clip_len:]
            ov = clip_len - (L - clip_len) % stride if (L - clip_len) % stride else clip_len - stride
        det, fus, oth = clip_out(cur)
        det, fus, oth = det.permute(1, 0, 2), fus.permute(1, 0, 2), oth.permute(1, 0, 2)
        c, s = torch.split(det, [4, 1], dim=-1)
        det = torch.cat([torch.where(s < thr, torch.zeros_like(c), c), s], -1)
        if ci == 0:
            vd, vo, vf = det, oth, fus
            continue
        new = clip_len - ov
        vd = torch.cat((vd, torch.zeros(vd.size(0), new, 5)), 1)
        vo = torch.cat((vo, torch.zeros(vo.size(0), new, 3)), 1)
        vf = torch.cat((vf, torch.zeros(vf.size(0), new, 3)), 1)
        vd[:, -new:], vo[:, -new:], vf[:, -new:] = det[:, -new:], oth[:, -new:], fus[:, -new:]
        o1, o2 = vd[:, -clip_len:-new], det[:, -clip_len:-new]
        c1, s1 = torch.split(o1, [4, 1], dim=-1)
        c2, s2 = torch.split(o2, [4, 1], dim=-1)
        m = torch.logical_or(s1 < thr, s2 < thr)
        vd[:, -clip_len:-new] = torch.cat([torch.where(m, torch.zeros_like(c1), (c1 + c2) / 2), (s1 + s2) / 2], -1)
        vo[:, -clip_len:-new] = (vo[:, -clip_len:-new] + oth[:, -clip_len:-new]) / 2
        vf[:, -clip_len:-new] = (vf[:, -clip_len:-new] + fus[:, -clip_len:-new]) / 2
    return vd.permute(1, 0, 2), vf.permute(1, 0, 2), vo.permute(1, 0, 2)


@pytest.mark.parametrize('L', [1, 3, 7, 8, 11, 12, 15, 23, 50])
def test_slicer_matches_reference_loop(L):
    g = torch.Generator().manual_seed(L)
    frame_boxes = torch.rand(64, L, 3, 4, generator=g) * 100
    frame_scores = torch.rand(64, L, 3, generator=g)
    frame_gaze = torch.randn(64, L, 4, 3, generator=g)
    calls = {'n': 0}

    def clip_out(frames):
        k = calls['n']
        calls['n'] += 1
        f = torch.tensor(frames)
        det = torch.cat([frame_boxes[k, f], frame_scores[k, f][..., None]], -1)
        return det, frame_gaze[k, f][:, :1], frame_gaze[k, f][:, 1:]

    vd, vf, vo = reference_merge(L, clip_out)
    plan = slicer.plan_clips(L)
    assert len(plan) == calls['n']
    boxes = [frame_boxes[k, s:s + n].numpy() for k, (s, n, _) in enumerate(plan)]
    scores = [frame_scores[k, s:s + n].numpy() for k, (s, n, _) in enumerate(plan)]
    gaze = [frame_gaze[k, s:s + n].numpy() for k, (s, n, _) in enumerate(plan)]
    merged = slicer.merge_video(plan, boxes, scores, gaze)
    assert merged['det'].shape == (L, 3, 5)
    assert np.allclose(merged['det'], vd.numpy(), atol=1e-6)
    assert np.allclose(merged['gaze'][:, 0], vf.numpy()[:, 0], atol=1e-6)
    assert np.allclose(merged['gaze'][:, 1:], vo.numpy(), atol=1e-6)
    rec = slicer.video_record(5, merged)
    assert len(rec['fusion_gazes']) == L and rec['video_id'] == 5
    zeroed = merged['det'][:, 0, :4].sum(-1) == 0
    assert all((b is None) == bool(z) for b, z in zip(rec['face_bboxes'], zeroed))


def test_gaze360_clip_count_matches_survey(golden_dir):
    import json
    lengths = np.load(os.path.join(golden_dir, 'golden_gaze360_results.npz'))['lengths']
    hist = {}
    for L in lengths.tolist():
        for (_, n, _) in slicer.plan_clips(L):
            hist[n] = hist.get(n, 0) + 1
    gold = json.load(open(os.path.join(golden_dir, 'golden_gaze360_clip_hist.json')))['clip_length_histogram']
    assert hist == {int(k): v for k, v in gold.items()} and sum(hist.values()) == 6365 and hist[7] == 6237


def test_shard_and_interleave_roundtrip():
    for n, w in [(10, 2), (7, 4), (6365, 8), (3, 8)]:
        items = np.arange(n * 2, dtype=np.float32).reshape(n, 2)
        per = -(-n // w)
        parts = []
        for r in range(w):
            idx = mdist.padded_shard(n, r, w)
            assert len(idx) == per and idx[:len(mdist.shard_indices(n, r, w))] == mdist.shard_indices(n, r, w)
            parts.append(items[idx])
        assert np.array_equal(mdist.interleave(parts, n), items)


def _worker(rank, world, port, n_items, out_q):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    idx = mdist.padded_shard(n_items, rank, world)
    local = torch.tensor([[i, i * i] for i in idx], dtype=torch.float32)      # "forward" of my clips
    full = mdist.gather_results(local, n_items)
    out_q.put((rank, full.numpy()))
    dist.destroy_process_group()


def test_two_rank_gloo_gather():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    n = 11
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.array([[i, i * i] for i in range(n)], dtype=np.float32)
    for r in range(2):
        assert np.array_equal(got[r], want)


# ---- properties over arbitrary lengths (hypothesis): the slicing of tools/test_gaze360_gaze.py:73-86 and the merge of :129-201
from hypothesis import given, settings  # noqa: E402
from hypothesis import strategies as st  # noqa: E402


@settings(max_examples=200, deadline=None)
@given(L=st.integers(1, 600), clip_len=st.integers(2, 16), stride_frac=st.integers(1, 15))
def test_plan_covers_every_frame_and_counts_its_own_overlap(L, clip_len, stride_frac):
    stride = 1 + stride_frac % (clip_len - 1) if clip_len > 2 else 1                     # 1 <= stride < clip_len
    plan = slicer.plan_clips(L, clip_len, stride)
    covered = np.zeros(L, dtype=int)
    end_prev = 0
    for i, (start, n, overlap) in enumerate(plan):
        assert 0 <= start and start + n <= L and 1 <= n <= clip_len
        assert (n == clip_len) or len(plan) == 1                                         # only a short video has a short clip
        covered[start:start + n] += 1
        # `overlap` is what the reference's merge treats as already present: the frames before the running end
        assert overlap == (end_prev - start if i else 0) and 0 <= overlap < n
        end_prev = max(end_prev, start + n)
    assert covered.min() >= 1 and end_prev == L
    if L > clip_len:
        assert len(plan) == math.ceil((L - clip_len) / stride) + 1 and plan[-1][0] == L - clip_len


@settings(max_examples=60, deadline=None)
@given(L=st.integers(1, 90), seed=st.integers(0, 2 ** 16))
def test_merge_of_frame_functions_is_the_function(L, seed):
    """If a clip's output for a frame depends on the frame only (scores all >= 0.5), averaging the overlaps changes nothing:
    the merged video equals the per-frame values, whatever the length class."""
    rng = np.random.default_rng(seed)
    boxes = rng.uniform(1, 100, (L, 3, 4)).astype(np.float32)
    scores = rng.uniform(0.5, 1.0, (L, 3)).astype(np.float32)
    gaze = rng.normal(size=(L, 4, 3)).astype(np.float32)
    plan = slicer.plan_clips(L)
    m = slicer.merge_video(plan, [boxes[s:s + n] for s, n, _ in plan], [scores[s:s + n] for s, n, _ in plan],
                           [gaze[s:s + n] for s, n, _ in plan])
    assert m['det'].shape == (L, 3, 5) and m['gaze'].shape == (L, 4, 3)
    assert np.allclose(m['det'][..., :4], boxes, rtol=0, atol=1e-5) and np.allclose(m['det'][..., 4], scores, rtol=0, atol=1e-6)
    assert np.allclose(m['gaze'], gaze, rtol=0, atol=1e-6)
    rec = slicer.video_record(7, m)
    assert rec['video_id'] == 7 and len(rec['fusion_gazes']) == L and all(b is not None and len(b) == 4 for b in rec['head_bboxes'])


@settings(max_examples=200, deadline=None)
@given(n=st.integers(1, 500), world=st.integers(1, 9))
def test_shards_partition_the_items_and_interleave_restores_them(n, world):
    parts = [mdist.shard_indices(n, r, world) for r in range(world)]
    assert sorted(i for p in parts for i in p) == list(range(n))
    padded = [mdist.padded_shard(n, r, world) for r in range(world)]
    per = -(-n // world)
    assert all(len(p) == per for p in padded) and all(p[:len(q)] == q for p, q in zip(padded, parts))
    data = [np.asarray(p, dtype=np.float32)[:, None] * 2.0 for p in padded]            # item i carries 2 i
    assert np.array_equal(mdist.interleave(data, n)[:, 0], 2.0 * np.arange(n, dtype=np.float32))
