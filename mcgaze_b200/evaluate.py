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
  evaluate             tools/calculate_mae_gaze360.py on those records

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
                 loader: Optional[Callable[[str], np.ndarray]] = None, test_mode: bool = True):
        if not test_mode:
            raise NotImplementedError('training is out of scope for the B200 inference backend')
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


def run_clips(model, dataset: Gaze360ClipDataset, pipeline, indices: Sequence[int], clips_per_batch: int = 32,
              workers: int = 0) -> Dict[int, np.ndarray]:
    """-> {clip index: float32 [n_frames, ROW]} for the given clips; frames of a batch go through ONE pipeline call
    and ONE forward.  With `workers` > 0 the frames of batch k+1 are decoded on a thread pool while the GPU works on
    batch k (the role of the DataLoader workers in mmdet/apis/test.py:107-109)."""
    import torch
    from concurrent.futures import ThreadPoolExecutor
    results: Dict[int, np.ndarray] = {}
    batches = _batches(dataset, indices, clips_per_batch)
    pool = ThreadPoolExecutor(workers) if workers > 0 else None
    feeder = ThreadPoolExecutor(1) if workers > 0 else None
    # pinned staging, three slots: batch k+2 is decoded into slot (k+2) % 3 while batch k+1 is being launched and batch
    # k still runs; the last user of that slot, batch k-1, has been read back (a stream sync) by then
    staging: Optional[Dict[Any, Any]] = {} if workers > 0 else None
    pending = None                                # results of the previous batch, still on the device

    def collect(p):
        batch, T, rows = p
        rows = rows.cpu().numpy().astype(np.float32)                                 # ONE device->host read per batch
        for k, i in enumerate(batch):
            results[i] = rows[k * T:(k + 1) * T]

    try:
        nxt = feeder.submit(_load_batch, dataset, batches[0], pool, staging, 0) if feeder and batches else None
        for bi, batch in enumerate(batches):
            if feeder:
                T, frames, names = nxt.result()
                nxt = feeder.submit(_load_batch, dataset, batches[bi + 1], pool, staging, (bi + 1) % 3) \
                    if bi + 1 < len(batches) else None
            else:
                T, frames, names = _load_batch(dataset, batch, None)
            data = pipeline.batch(frames, filenames=names)
            (det_bboxes, _), gz = model(return_loss=False, rescale=True, format=False, clip_length=T, **data)
            n = len(names)
            det = torch.stack(list(det_bboxes)).float()                              # [B*T, 3, 5]
            gaze = torch.stack([gz['gaze_score'], gz['face_gaze_score'], gz['eyes_gaze_score'], gz['head_gaze_score']], 1)
            rows = torch.cat([det[..., :4].reshape(n, 12), det[..., 4], gaze.reshape(n, 12).float()], 1)
            # read the PREVIOUS batch back only now, with this batch already queued: the GPU never waits for the host
            if pending is not None:
                collect(pending)
            pending = (batch, T, rows)
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


def evaluate(dataset: Gaze360ClipDataset, records: Sequence[Dict[str, Any]], key: str = 'fusion_gazes') -> Dict[str, float]:
    gt = ground_truth_videos(dataset)
    if gt is None:
        raise ValueError('the annotation file carries no ground-truth gazes')
    return metric.gaze_error([np.asarray(r[key], dtype=np.float64) for r in records], gt)


def evaluate_on_device(dataset: Gaze360ClipDataset, merged: Sequence[Dict[str, np.ndarray]], clue: int = 0,
                       device: str = 'cuda:0') -> Dict[str, float]:
    """The same numbers from mcg_gaze_error: the merged per-video gaze arrays of `videos_from_clips`
    (clue 0 = fused, 1 / 2 / 3 = face / eyes / head) are scored by one kernel launch on the device."""
    import torch
    from . import lib
    gt = ground_truth_videos(dataset)
    if gt is None:
        raise ValueError('the annotation file carries no ground-truth gazes')
    pred = np.concatenate([m['gaze'][:, clue] for m in merged]).astype(np.float32)
    gt_all = np.concatenate(gt).astype(np.float32)
    return lib.gaze_error(torch.from_numpy(pred).to(device), torch.from_numpy(gt_all).to(device), [len(g) for g in gt])
