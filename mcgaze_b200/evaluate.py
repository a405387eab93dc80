"""Batched, multi-GPU evaluation driver (SURVEY.md section 8, rows f1 + f2).

The reference evaluates with tools/test_gaze360_gaze.py: one clip per forward, T python threads running the CPU
pipeline per clip, `.cpu().tolist()` per frame (:60-269); its generic `tools/test.py` / `dist_test.sh` path cannot run
the gaze configs because `Gaze360Dataset.__getitem__` raises NotImplementedError in test mode
(mmdet/datasets/gaze360.py:310-312).  This module provides what that path needs, on the new backend:

  Gaze360ClipDataset   test-mode dataset: item i = one clip (7-frame window, stride 4, right-aligned last window:
                       tools/test_gaze360_gaze.py:73-86) of one video of `test.json`
  single_gpu_test /    mmdet/apis/test.py:17-78 / :81-209 for clips: every rank takes the clips
  multi_gpu_test       DistributedSampler(shuffle=False) would give it (samplers/distributed_sampler.py:53), MANY
                       clips go through one forward (`clips_per_batch`), results are all-gathered ONCE and put back
                       in dataset order (the `zip(*part_list)` re-ordering of :204-206)
  videos_from_clips    overlap merge + JSON records exactly as tools/test_gaze360_gaze.py:129-260
  evaluate             tools/calculate_mae_gaze360.py / calculate_mae_l2cs.py on those records
  multi_gpu_test_videos  the sharded run SURVEY section 8e describes: VIDEOS are sharded over the ranks, each rank
                       merges the overlaps of its videos and scores them ON ITS DEVICE (mcg_merge_clips,
                       mcg_gaze_error), the MAE needs one all-reduce of 6 doubles; per-video arrays are gathered only
                       when the caller wants the JSON

Batches hold clips of ONE length and ONE padded canvas: the reference collates one clip at a time, so a clip's frames are
padded to that clip's own largest frame (tools/test_gaze360_gaze.py:102-105) - a clip must never inherit a larger canvas
from its batch neighbours (zero padding changes the convolutions' border behaviour and the proposal boxes).

`model` is anything with the reference's call contract `model(return_loss=False, rescale=True, format=False,
img=[Tensor], img_metas=[[meta...]], clip_length=T)` (mcgaze_b200.detector.MultiClueGaze); `pipeline` is a
mcgaze_b200.pipeline.GpuTestPipeline (or any object with `.batch(frames, filenames=...) -> dict(img, img_metas)`).
"""
from __future__ import annotations

import json
import os
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import dist as mdist
from . import metric, slicer

ROW = 3 * 4 + 3 + 4 * 3          # per frame: boxes [3,4], scores [3], gaze [4,3]


def _default_loader(path: str) -> np.ndarray:
    """LoadImageFromFile (mmdet/datasets/pipelines/loading.py:58-69): cv2 decode, BGR uint8.  Decoding stays on the host."""
    import cv2
    img = cv2.imread(path, cv2.IMREAD_COLOR)
    if img is None:
        raise FileNotFoundError(path)
    return img


class Gaze360ClipDataset:
    def __init__(self, ann_file, img_prefix: str = '', clip_len: int = slicer.CLIP_LEN, stride: int = slicer.STRIDE,
                 loader: Optional[Callable[[str], np.ndarray]] = None, test_mode: bool = True, decode: str = 'host',
                 check_crc: bool = False):
        """decode='host': LoadImageFromFile on the host (cv2, in the loader threads); decode='gpu': the batched drivers
        below only read the files and decode them on the device (mcgaze_b200/png.py: mcg_png_parse + mcg_png_decode);
        batches holding a file that decoder does not take (JPEG, 16-bit or interlaced PNG) go through the host loader.
        check_crc: also verify every chunk CRC on the host (libpng does; ~0.2 ms per frame and core) - the device always
        verifies the Adler-32 of the pixel data, which is what a damaged file most likely breaks."""
        if not test_mode:
            raise NotImplementedError('training is out of scope for the B200 inference backend')
        if decode not in ('host', 'gpu'):
            raise ValueError("decode must be 'host' or 'gpu'")
        if decode == 'gpu' and loader is not None:
            raise ValueError("decode='gpu' reads the files itself; it does not combine with a custom loader")
        self.decode = decode
        self.check_crc = bool(check_crc)
        self.host_decoded_batches = 0             # decode='gpu': batches that fell back to the host loader
        self.anno = json.load(open(ann_file)) if isinstance(ann_file, (str, os.PathLike)) else ann_file
        self.img_prefix = img_prefix
        self.clip_len, self.stride = clip_len, stride
        self.loader = loader or _default_loader
        self.videos: List[List[str]] = [list(v['file_names']) for v in self.anno['videos']]
        self.plans = [slicer.plan_clips(len(v), clip_len, stride) for v in self.videos]
        # flat clip index -> (video index, clip index within the video)
        self.index: List[Tuple[int, int]] = [(vi, ci) for vi, p in enumerate(self.plans) for ci in range(len(p))]

    def __len__(self) -> int:
        return len(self.index)

    def clip_info(self, i: int) -> Dict[str, Any]:
        vi, ci = self.index[i]
        start, n, overlap = self.plans[vi][ci]
        return dict(video=vi, clip=ci, start=start, n=n, overlap=overlap, filenames=self.videos[vi][start:start + n])

    def __getitem__(self, i: int) -> Dict[str, Any]:
        info = self.clip_info(i)
        info['frames'] = [self.loader(os.path.join(self.img_prefix, f) if self.img_prefix else f) for f in info['filenames']]
        return info


def _batches(dataset: Gaze360ClipDataset, indices: Sequence[int], clips_per_batch: int) -> List[List[int]]:
    """Clips of one length share a forward (the temporal attention needs one clip_length per call)."""
    by_len: Dict[int, List[int]] = {}
    for i in indices:
        by_len.setdefault(dataset.clip_info(i)['n'], []).append(i)
    out = []
    for n in sorted(by_len, reverse=True):
        idx = by_len[n]
        out += [idx[k:k + clips_per_batch] for k in range(0, len(idx), clips_per_batch)]
    return out


def _load_batch(dataset: Gaze360ClipDataset, batch: Sequence[int], pool, staging: Optional[Dict[Any, Any]] = None,
                slot: int = 0) -> Tuple[int, Any, List[str]]:
    """Decode the frames of a batch of clips.  With a thread pool and a `staging` cache every worker decodes its frame
    and writes it straight into ONE [n, h, w, 3] block in pinned host memory (three slots per shape), so the batch
    crosses PCIe in a single asynchronous copy; frames of different sizes fall back to a list."""
    infos = [dataset.clip_info(i) for i in batch]
    names = [f for it in infos for f in it['filenames']]
    paths = [os.path.join(dataset.img_prefix, f) if dataset.img_prefix else f for f in names]
    n = len(paths)
    if pool is None or staging is None:
        frames = [dataset.loader(p) for p in paths]
        return infos[0]['n'], (np.stack(frames) if len({f.shape for f in frames}) == 1 else frames), names
    import torch
    first = dataset.loader(paths[0])
    key = (slot, n) + first.shape
    if key not in staging:
        staging[key] = torch.empty((n,) + first.shape, dtype=torch.uint8, pin_memory=torch.cuda.is_available())
    block = staging[key].numpy()
    block[0] = first

    def work(lo_hi: Tuple[int, int]):
        odd = []
        for k in range(*lo_hi):                 # a contiguous run of frames per task: 224 one-frame tasks cost more
            f = dataset.loader(paths[k])        # in executor overhead than the copies themselves
            if f.shape != first.shape or f.dtype != np.uint8:
                odd.append(f)
            else:
                block[k] = f
                odd.append(None)
        return odd

    tasks = max(1, min(n - 1, 2 * getattr(pool, '_max_workers', 4)))
    cuts = np.linspace(1, n, tasks + 1).astype(int)
    odd = [None] + [o for part in pool.map(work, zip(cuts[:-1], cuts[1:])) for o in part]
    if all(o is None for o in odd):
        return infos[0]['n'], staging[key], names
    return infos[0]['n'], [block[k].copy() if o is None else o for k, o in enumerate(odd)], names


def _stage_unit(dataset: Gaze360ClipDataset, unit: Sequence[Sequence[int]], pool):
    """decode='gpu': the host part for SEVERAL batches at once - read the files and walk their chunks into one pinned
    block (mcgaze_b200/png.py).  One mcg_png_decode launch then covers the whole unit: a warp needs ~10 ms for a frame
    whatever else runs, so the decode is paid once per unit instead of once per batch.
    -> (StagedPngs of the unit's DISTINCT files, [(clip length, frame names, first frame, end frame) per batch], index of
    every frame into the distinct files) or None when a file is not a PNG the device decoder takes (the unit then goes
    through the host loader)."""
    from . import png
    paths, parts, index, where = [], [], [], {}
    for batch in unit:
        infos = [dataset.clip_info(i) for i in batch]
        names = [f for it in infos for f in it['filenames']]
        first = len(index)
        for f in names:                       # overlapping clips share frames (stride 4 of 7): a file is staged and decoded once
            k = where.get(f)
            if k is None:
                k = where[f] = len(paths)
                paths.append(os.path.join(dataset.img_prefix, f) if dataset.img_prefix else f)
            index.append(k)
        parts.append((infos[0]['n'], names, first, len(index)))
    try:
        return png.GpuPngDecoder(check_crc=dataset.check_crc).stage(paths, pool), parts, np.asarray(index, dtype=np.int64)
    except png.UnsupportedPng:
        dataset.host_decoded_batches += len(unit)
        return None


def _canvas_groups(pipeline, frames, names: Sequence[str], T: int):
    """Split the clips of a loaded batch by their OWN padded canvas.  -> [(frames, names, rands, clip positions)], one
    entry per canvas; `rands` are the CenterCrop draws of those frames (drawn once per batch, in frame order, so the
    grouping does not change which draw a frame gets)."""
    n = len(names)
    nclips = n // T
    if not hasattr(pipeline, 'clip_canvases') or not hasattr(pipeline, 'draw'):
        return [(frames, names, None, list(range(nclips)))]          # foreign pipeline object: nothing to plan with
    if hasattr(frames, 'shape'):
        shapes = [(int(frames.shape[1]), int(frames.shape[2]))] * n
    else:
        shapes = [(int(f.shape[0]), int(f.shape[1])) for f in frames]
    rands = pipeline.draw(n)
    canvases = pipeline.clip_canvases(shapes, rands, T)
    if len(set(canvases)) == 1:
        return [(frames, names, rands, list(range(nclips)))]
    groups = []
    for cv in sorted(set(canvases)):
        clips = [k for k in range(nclips) if canvases[k] == cv]
        idx = [k * T + t for k in clips for t in range(T)]
        sub = frames[idx] if hasattr(frames, 'shape') else [frames[i] for i in idx]
        groups.append((sub, [names[i] for i in idx], rands[idx], clips))
    return groups


def run_clips(model, dataset: Gaze360ClipDataset, pipeline, indices: Sequence[int], clips_per_batch: int = 32,
              workers: int = 0, to_host: bool = True, batch_sink: Optional[List[Any]] = None,
              decode_group: int = 8) -> Dict[int, Any]:
    """-> {clip index: float32 [n_frames, ROW]} for the given clips (numpy; torch tensors on the model's device with
    to_host=False); frames of a batch go through ONE pipeline call and ONE forward per padded canvas.  With `workers` > 0
    the frames of batch k+1 are decoded on a thread pool while the GPU works on batch k (the role of the DataLoader
    workers in mmdet/apis/test.py:107-109); on a GPU the uint8 block of batch k+1 crosses PCIe on a copy stream while
    batch k computes, and results come back through pinned buffers gated by per-batch events, so the host never waits
    for the batch it has just queued.  With a decode='gpu' dataset the workers only read files; the PNG streams of
    `decode_group` batches cross PCIe compressed and are decoded by ONE mcg_png_decode launch."""
    import torch
    from concurrent.futures import ThreadPoolExecutor
    results: Dict[int, Any] = {}
    batches = _batches(dataset, indices, clips_per_batch)
    pool = ThreadPoolExecutor(workers) if workers > 0 else None
    feeder = ThreadPoolExecutor(1) if workers > 0 else None
    # pinned staging, three slots: batch k+2 is decoded into slot (k+2) % 3 while batch k+1 is being launched and batch
    # k still runs; the upload of the last user of that slot, batch k-1, was issued (copy stream) before batch k's
    # forward was queued and is complete once batch k-1's results have been collected
    staging: Optional[Dict[Any, Any]] = {} if workers > 0 else None
    pending = None                                # results of the previous batch: (clip ids, T, rows, event)
    cuda = torch.cuda.is_available()
    copy_stream = None

    upload_done: Dict[int, Any] = {}              # staging slot -> event of the last H2D copy that read it
    decoder = None
    decode_check = None                           # decode='gpu': (staged files, pinned status copy, event) of the batch in upload()

    def upload(frames, slot):
        """pinned uint8 block -> device on a side stream; the compute stream waits for the copy only.  Staged PNG files
        (decode='gpu'): compressed bytes -> device + mcg_png_decode on the compute stream, status checked at collect()."""
        nonlocal copy_stream, decoder, decode_check
        decode_check = None
        if hasattr(frames, 'zoff'):               # png.StagedPngs
            from . import png
            if decoder is None:
                decoder = png.GpuPngDecoder(getattr(pipeline, 'device', 0))
            with torch.cuda.device(decoder.device):
                if copy_stream is None:
                    copy_stream = torch.cuda.Stream(device=decoder.device)
                dframes, status = decoder.launch(frames, copy_stream=copy_stream)
                decode_check = (frames,) + decoder.status_async(status)
            return dframes
        if not (cuda and hasattr(frames, 'is_pinned') and frames.is_pinned() and hasattr(pipeline, 'device')):
            return frames
        dev = torch.device('cuda', pipeline.device)
        if copy_stream is None:
            copy_stream = torch.cuda.Stream(device=dev)
        cur = torch.cuda.current_stream(dev)
        with torch.cuda.stream(copy_stream):
            d = frames.to(dev, non_blocking=True)
            done = torch.cuda.Event()
            done.record(copy_stream)
        upload_done[slot] = done
        cur.wait_event(done)
        d.record_stream(cur)
        return d

    def collect(p):
        ids, T, rows, ev, chk = p
        if chk is not None:                       # per-image status of this batch's PNG decode (queued before its forward)
            chk[2].synchronize()
            decoder.check(chk[0], chk[1])
        if ev is not None:
            ev.synchronize()                      # this batch's read-back only; later batches keep running
        if batch_sink is not None:                # whole batches for the caller (no per-clip slicing of device tensors)
            batch_sink.append((ids, T, rows))
            return
        if to_host:
            rows = rows.numpy() if hasattr(rows, 'numpy') else np.asarray(rows)
            rows = rows.astype(np.float32, copy=False)
        for k, i in enumerate(ids):
            results[i] = rows[k * T:(k + 1) * T]

    gpu_decode = getattr(dataset, 'decode', 'host') == 'gpu'
    per_unit = max(1, int(decode_group)) if gpu_decode else 1
    units = [batches[k:k + per_unit] for k in range(0, len(batches), per_unit)]

    def load_unit(ui: int):
        """-> ('gpu', (staged files, parts)) or ('host', [(T, frames, names) per batch])"""
        if gpu_decode:
            r = _stage_unit(dataset, units[ui], pool)
            if r is not None:
                return 'gpu', r
            return 'host', [_load_batch(dataset, b, pool) for b in units[ui]]       # rare: plain arrays, no pinned slots
        return 'host', [_load_batch(dataset, units[ui][0], pool, staging, ui % 3)]

    try:
        nxt = feeder.submit(load_unit, 0) if feeder and units else None
        bi = -1
        for ui, unit in enumerate(units):
            if feeder:
                kind, payload = nxt.result()
                if ui + 1 < len(units):
                    # (host decode: units are single batches) the loader is about to overwrite pinned slot (ui + 1) % 3:
                    # the copy of its last user (batch ui - 2) must have left it - nothing else orders the two when the
                    # results stay on the device
                    prev = upload_done.pop((ui + 1) % 3, None)
                    if prev is not None:
                        prev.synchronize()
                    nxt = feeder.submit(load_unit, ui + 1)
                else:
                    nxt = None
            else:
                kind, payload = load_unit(ui)
            if kind == 'gpu':
                staged, descr, index = payload
                dframes = upload(staged, 0)                                    # one H2D copy + one decode launch per unit
                if hasattr(dframes, 'shape'):                                  # frames of one size: gather on the device
                    didx = torch.from_numpy(index).to(dframes.device, non_blocking=True)
                    items = [(T, dframes.index_select(0, didx[a:b]), names) for T, names, a, b in descr]
                else:
                    items = [(T, [dframes[i] for i in index[a:b]], names) for T, names, a, b in descr]
            else:
                items = payload
            for k, (T, frames, names) in enumerate(items):
                bi += 1
                batch = unit[k]
                if kind != 'gpu':
                    frames = upload(frames, bi % 3)
                elif k > 0:
                    decode_check = None                                            # the unit's status rides on its first batch
                parts = []
                for sub, sub_names, rands, clips in _canvas_groups(pipeline, frames, names, T):
                    data = pipeline.batch(sub, filenames=sub_names) if rands is None else \
                        pipeline.batch(sub, rands=rands, filenames=sub_names)
                    (det_bboxes, _), gz = model(return_loss=False, rescale=True, format=False, clip_length=T, **data)
                    n = len(sub_names)
                    det = torch.stack(list(det_bboxes)).float()                              # [B*T, 3, 5]
                    gaze = torch.stack([gz['gaze_score'], gz['face_gaze_score'], gz['eyes_gaze_score'], gz['head_gaze_score']], 1)
                    rows = torch.cat([det[..., :4].reshape(n, 12), det[..., 4], gaze.reshape(n, 12).float()], 1)
                    parts.append(([batch[k] for k in clips], rows))
                ids = [i for p_ids, _ in parts for i in p_ids]
                rows = parts[0][1] if len(parts) == 1 else torch.cat([r for _, r in parts])
                ev = None
                if rows.is_cuda and to_host:
                    host = torch.empty(rows.shape, dtype=rows.dtype, pin_memory=True)
                    host.copy_(rows, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(torch.cuda.current_stream(rows.device))
                    rows = host
                # hand the PREVIOUS batch over only now, with this batch already queued: the GPU never waits for the host
                if pending is not None:
                    collect(pending)
                pending = (ids, T, rows, ev, decode_check)
        if pending is not None:
            collect(pending)
    finally:
        for ex in (feeder, pool):
            if ex is not None:
                ex.shutdown(wait=True)
    return results


def single_gpu_test(model, dataset: Gaze360ClipDataset, pipeline, clips_per_batch: int = 32, workers: int = 0) -> List[np.ndarray]:
    res = run_clips(model, dataset, pipeline, range(len(dataset)), clips_per_batch, workers)
    return [res[i] for i in range(len(dataset))]


def multi_gpu_test(model, dataset: Gaze360ClipDataset, pipeline, clips_per_batch: int = 32, group=None,
                   device=None, workers: int = 0) -> List[np.ndarray]:
    """Every rank returns the results of ALL clips in dataset order (one all-gather; NCCL on GPUs, gloo on CPU)."""
    import torch
    import torch.distributed as tdist
    rank, world = tdist.get_rank(group), tdist.get_world_size(group)
    n = len(dataset)
    mine = mdist.padded_shard(n, rank, world)
    res = run_clips(model, dataset, pipeline, sorted(set(mine)), clips_per_batch, workers)
    packed = np.zeros((len(mine), dataset.clip_len, ROW), dtype=np.float32)
    for k, i in enumerate(mine):
        packed[k, :res[i].shape[0]] = res[i]
    local = torch.from_numpy(packed)
    if device is not None:
        local = local.to(device)
    full = mdist.gather_results(local, n, group=group).numpy()
    return [full[i, :dataset.clip_info(i)['n']] for i in range(n)]


def videos_from_clips(dataset: Gaze360ClipDataset, clip_rows: Sequence[np.ndarray]) -> Tuple[List[Dict[str, Any]], List[Dict[str, np.ndarray]]]:
    """Per-clip rows in dataset order -> (JSON records of tools/test_gaze360_gaze.py:210-260, merged per-video arrays)."""
    records, merged_all = [], []
    k = 0
    for vi, plan in enumerate(dataset.plans):
        rows = clip_rows[k:k + len(plan)]
        k += len(plan)
        boxes = [r[:, :12].reshape(-1, 3, 4) for r in rows]
        scores = [r[:, 12:15] for r in rows]
        gaze = [r[:, 15:].reshape(-1, 4, 3) for r in rows]
        merged = slicer.merge_video(plan, boxes, scores, gaze)
        merged_all.append(merged)
        vid = dataset.anno['videos'][vi].get('id', vi + 1)
        records.append(slicer.video_record(vid, merged))
    return records, merged_all


def ground_truth_videos(dataset: Gaze360ClipDataset) -> Optional[List[np.ndarray]]:
    """Per-video GT gaze vectors: annotation k belongs to video k and carries one vector per frame
    (tools/calculate_mae_gaze360.py:118-121)."""
    ann = dataset.anno.get('annotations')
    if not ann:
        return None
    gt = [np.asarray(a['gaze'], dtype=np.float64).reshape(-1, 3) for a in ann]
    if len(gt) != len(dataset.videos) or any(len(g) != len(v) for g, v in zip(gt, dataset.videos)):
        raise ValueError('annotations do not match the videos (one annotation per video, one gaze per frame)')
    return gt


def ground_truth(dataset: Gaze360ClipDataset, variant: str = 'gaze360') -> Optional[List[np.ndarray]]:
    """Per-video ground truth for a scorer variant: Gaze360 = annotation k, l2cs = annotation 3k
    (tools/calculate_mae_l2cs.py:110)."""
    if variant == 'l2cs':
        if not dataset.anno.get('annotations'):
            return None
        gt = metric.l2cs_ground_truth(dataset.anno)
        if len(gt) < len(dataset.videos) or any(len(g) != len(v) for g, v in zip(gt, dataset.videos)):
            raise ValueError('l2cs annotations do not match the videos (annotation 3k belongs to video k)')
        return gt[:len(dataset.videos)]
    return ground_truth_videos(dataset)


def evaluate(dataset: Gaze360ClipDataset, records: Sequence[Dict[str, Any]], key: str = 'fusion_gazes',
             variant: str = 'gaze360') -> Dict[str, float]:
    gt = ground_truth(dataset, variant)
    if gt is None:
        raise ValueError('the annotation file carries no ground-truth gazes')
    return metric.gaze_error([np.asarray(r[key], dtype=np.float64) for r in records], gt, variant=variant)


def evaluate_on_device(dataset: Gaze360ClipDataset, merged: Sequence[Dict[str, np.ndarray]], clue: int = 0,
                       device: str = 'cuda:0', variant: str = 'gaze360') -> Dict[str, float]:
    """The same numbers from mcg_gaze_error: the merged per-video gaze arrays of `videos_from_clips`
    (clue 0 = fused, 1 / 2 / 3 = face / eyes / head) are scored by one kernel launch on the device."""
    import torch
    from . import lib
    gt = ground_truth(dataset, variant)
    if gt is None:
        raise ValueError('the annotation file carries no ground-truth gazes')
    pred = np.concatenate([m['gaze'][:, clue] for m in merged]).astype(np.float32)
    gt_all = np.concatenate(gt).astype(np.float32)
    return lib.gaze_error(torch.from_numpy(pred).to(device), torch.from_numpy(gt_all).to(device), [len(g) for g in gt],
                          variant=variant)


def multi_gpu_test_videos(model, dataset: Gaze360ClipDataset, pipeline, clips_per_batch: int = 32, group=None,
                          workers: int = 0, variant: str = 'gaze360', clue: int = 0, gather_videos: bool = True,
                          merge: str = 'auto') -> Dict[str, Any]:
    """The sharded evaluation of SURVEY section 8e: rank r takes VIDEOS r, r + W, ... (smoothing and the overlap merge
    need whole videos on one rank), runs their clips in batches, merges the overlaps and scores its videos where the
    results are - on its device through mcg_merge_clips / mcg_gaze_error when the forward returns CUDA tensors
    (`merge='device'`), with the host restatements otherwise (`'host'`: CPU stand-in models in the gloo tests) - and the
    MAE of the whole split costs ONE all-reduce of 6 doubles ({sum, frames} x {360, front-180, front-20}).
    -> dict(mae=..., sums=[6], merged=[per-video dict(det [L,3,5], gaze [L,4,3])] in dataset order when gather_videos
    (one all-gather of the merged rows), else only this rank's videos under 'merged_local' / 'videos_local')."""
    import torch
    import torch.distributed as tdist
    dist_on = tdist.is_available() and tdist.is_initialized()
    rank = tdist.get_rank(group) if dist_on else 0
    world = tdist.get_world_size(group) if dist_on else 1
    nv = len(dataset.videos)
    mine = mdist.shard_indices(nv, rank, world)
    first = np.concatenate([[0], np.cumsum([len(p) for p in dataset.plans])]).astype(np.int64)
    clip_ids = [int(c) for v in mine for c in range(first[v], first[v + 1])]
    done: List[Any] = []
    if clip_ids:
        run_clips(model, dataset, pipeline, clip_ids, clips_per_batch, workers, to_host=False, batch_sink=done)
    T, L = dataset.clip_len, [len(dataset.videos[v]) for v in mine]
    on_device = bool(done) and hasattr(done[0][2], 'is_cuda') and done[0][2].is_cuda
    if merge == 'auto':
        merge = 'device' if on_device else 'host'
    gt = ground_truth(dataset, variant)
    sums = torch.zeros(6, dtype=torch.float64)
    det = gz = None
    # per-clip rows [n_clips, clip_len, ROW] in dataset order from the batches: one cat + one index_select
    order, blocks = [], []
    for ids, t, r in done:
        r = torch.as_tensor(r).reshape(len(ids), t, ROW)
        blocks.append(r if t == T else torch.nn.functional.pad(r, (0, 0, 0, T - t)))
        order += ids
    rows = None
    if blocks:
        pos = {c: k for k, c in enumerate(order)}
        rows = torch.cat(blocks).index_select(0, torch.tensor([pos[c] for c in clip_ids], device=blocks[0].device))
    if mine and merge == 'device':
        from . import lib
        dev = rows.device
        det, gz = lib.merge_clips(rows.contiguous(), [len(dataset.plans[v]) for v in mine], L, T, dataset.stride)
        if gt is not None:
            gt_dev = torch.from_numpy(np.concatenate([gt[v] for v in mine]).astype(np.float32)).to(dev)
            sums = lib.gaze_error_sums(gz[:, clue].contiguous(), gt_dev, L, variant)
    elif mine:
        merged_local = []
        host_rows = rows.cpu().numpy()
        at = 0
        for v in mine:
            r = [host_rows[at + k, :n] for k, (_, n, _) in enumerate(dataset.plans[v])]
            at += len(dataset.plans[v])
            merged_local.append(slicer.merge_video(dataset.plans[v], [x[:, :12].reshape(-1, 3, 4) for x in r],
                                                   [x[:, 12:15] for x in r], [x[:, 15:].reshape(-1, 4, 3) for x in r]))
        det = torch.from_numpy(np.concatenate([m['det'] for m in merged_local]))
        gz = torch.from_numpy(np.concatenate([m['gaze'] for m in merged_local]))
        if gt is not None:
            m = metric.gaze_error([x['gaze'][:, clue] for x in merged_local], [gt[v] for v in mine], variant=variant)
            sums = torch.tensor([m[f'mae_{k}'] * m[f'frames_{k}'] if j == 0 else m[f'frames_{k}']
                                 for k in ('360', 'front90', 'front20') for j in (0, 1)], dtype=torch.float64)
    if dist_on and world > 1:
        backend = tdist.get_backend(group)
        sums = sums.cuda() if backend == 'nccl' and not sums.is_cuda else (sums.cpu() if backend != 'nccl' else sums)
        tdist.all_reduce(sums, group=group)                         # THE collective of a sharded MAE run
    out: Dict[str, Any] = dict(sums=[float(v) for v in sums.cpu()], videos_local=mine)
    if gt is not None:
        o = out['sums']
        out['mae'] = {f'mae_{k}': o[2 * j] / max(o[2 * j + 1], 1.0) for j, k in enumerate(('360', 'front90', 'front20'))} | \
                     {f'frames_{k}': int(o[2 * j + 1]) for j, k in enumerate(('360', 'front90', 'front20'))}
    off = np.concatenate([[0], np.cumsum(L)]).astype(np.int64)
    local = [dict(det=det[off[k]:off[k + 1]], gaze=gz[off[k]:off[k + 1]]) for k in range(len(mine))] if mine else []
    if not gather_videos:
        out['merged_local'] = local
        return out
    # JSON wanted: ONE all-gather of the merged per-frame rows (27 floats per frame), back to dataset order
    # same buffer size on every rank, computed from what every rank knows (lengths + sharding rule): the largest shard
    fmax = max(sum(len(dataset.videos[v]) for v in mdist.shard_indices(nv, r, world)) for r in range(world))
    packed = torch.zeros(fmax, 27, dtype=torch.float32, device=det.device if det is not None else 'cpu')
    if mine:
        packed[:off[-1], :15] = det.reshape(-1, 15)
        packed[:off[-1], 15:] = gz.reshape(-1, 12)
    if dist_on and world > 1:
        if tdist.get_backend(group) != 'nccl':
            packed = packed.cpu()
        elif not packed.is_cuda:           # a rank without videos still takes part in the NCCL collective
            packed = packed.cuda()
        bufs = [torch.empty_like(packed) for _ in range(world)]
        tdist.all_gather(bufs, packed, group=group)
    else:
        bufs = [packed]
    merged: List[Optional[Dict[str, np.ndarray]]] = [None] * nv
    for r, b in enumerate(bufs):
        b = b.cpu().numpy()
        o = 0
        for v in mdist.shard_indices(nv, r, world):
            n = len(dataset.videos[v])
            merged[v] = dict(det=b[o:o + n, :15].reshape(n, 3, 5).copy(), gaze=b[o:o + n, 15:].reshape(n, 4, 3).copy())
            o += n
    out['merged'] = merged
    return out


def records_from_merged(dataset: Gaze360ClipDataset, merged: Sequence[Dict[str, np.ndarray]]) -> List[Dict[str, Any]]:
    """Merged per-video arrays -> the JSON records of tools/test_gaze360_gaze.py:210-260."""
    return [slicer.video_record(dataset.anno['videos'][vi].get('id', vi + 1), m) for vi, m in enumerate(merged)]
