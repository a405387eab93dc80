"""Test-mode `Gaze360Dataset` + `build_dataset` / `build_dataloader` + the reference-signature `single_gpu_test` /
`multi_gpu_test`, so that the reference's GENERIC tester - `tools/test.py:120-240` and
`tools/dist_test.sh` - runs the gaze configs unmodified on the B200 backend.

In the reference that path is dead for gaze: `Gaze360Dataset.__getitem__` raises NotImplementedError in test mode
(mmdet/datasets/gaze360.py:310-312, `prepare_test_clip` :382-383) and `evaluate` writes a video-segmentation JSON
(:397-404), so only the bespoke `tools/test_gaze360_gaze.py` can evaluate.  Here:

  Gaze360Dataset       same registry name and constructor arguments (gaze360.py:23-34); item i = one clip of the slicing
                       tools/test_gaze360_gaze.py:73-86 uses; `evaluate` / `format_results` merge the overlaps, write the
                       results JSON in the schema of tools/test_gaze360_gaze.py:210-260 (what calculate_mae_*.py read) and
                       return the three MAE numbers of every clue when the annotation file carries ground truth
  build_dataset        mmdet/datasets/builder.py:57-80 (plain datasets only; the wrappers are training-side)
  build_dataloader     mmdet/datasets/builder.py:83-180 -> ClipLoader: carries dataset / samples_per_gpu / workers_per_gpu;
                       the batched driver of mcgaze_b200.evaluate does the loading (decode threads, pinned staging, copy
                       stream) instead of torch DataLoader worker processes
  single_gpu_test      mmdet/apis/test.py:17-78  (model, data_loader, show, out_dir, show_score_thr)
  multi_gpu_test       mmdet/apis/test.py:81-126 (model, data_loader, tmpdir, gpu_collect): clips sharded like
                       DistributedSampler(shuffle=False), ONE all-gather, dataset order restored (:204-206)

Results are per-clip float32 rows [n_frames, 27] (boxes 3x4, scores 3, gazes 4x3) in dataset order - the `outputs` that
tools/test.py pickles with --out and hands to `dataset.evaluate`.  A clip's result does not depend on its batch
neighbours (tests/test_gpu_forward.py: bit-exact batch independence), so `samples_per_gpu` only sets how many clips share
a forward: at least MCG_CLIPS_PER_BATCH (default 32) are used, because tools/test.py reads `data.test.samples_per_gpu`
(absent in the gaze configs -> 1)."""
from __future__ import annotations

import json
import os
import os.path as osp
from typing import Dict, Optional

import numpy as np

from . import evaluate as ev
from .compat import Registry, build_from_cfg

DATASETS = Registry('dataset')
PIPELINES = Registry('pipeline')

_METRICS = ('mae', 'gaze')
_KEYS = ('fusion_gazes', 'face_gazes', 'eyes_gazes', 'head_gazes')


@DATASETS.register_module()
class Gaze360Dataset(ev.Gaze360ClipDataset):
    """reference: mmdet/datasets/gaze360.py:18-84 (constructor contract), :310-312 / :382-383 (the missing test mode)."""

    CLASSES = ('person_face')                     # gaze360.py:21 (a plain string there too)

    # hooks for stand-ins (CPU tests): frame loader and pipeline factory of the batched driver
    frame_loader = None
    pipeline_factory = None

    def __init__(self, ann_file, pipeline, clip_length=7, gaze_dim=3, classes=None, data_root=None, img_prefix='',
                 seg_prefix=None, proposal_file=None, test_mode=False, filter_empty_gt=True, decode: Optional[str] = None,
                 stride: Optional[int] = None, scorer: Optional[str] = None, seed: Optional[int] = None):
        if not test_mode:
            raise NotImplementedError('Gaze360Dataset: only test_mode=True exists on the B200 inference backend '
                                      '(training is out of scope)')
        if gaze_dim != 3:
            raise NotImplementedError('gaze_dim must be 3 (unit gaze vectors)')
        if proposal_file is not None:
            raise NotImplementedError('proposal files are not used by MultiClueGaze (fixed embedding proposals)')
        if data_root is not None:                                    # gaze360.py:46-56
            if isinstance(ann_file, str) and not osp.isabs(ann_file):
                ann_file = osp.join(data_root, ann_file)
            if not (img_prefix is None or osp.isabs(img_prefix)):
                img_prefix = osp.join(data_root, img_prefix)
        self.ann_file = ann_file
        self.data_root = data_root
        self.test_mode = True
        self.pipeline_cfg = [dict(t) for t in pipeline]
        self.seed = seed
        loader = type(self).frame_loader
        if decode is None:
            decode = os.environ.get('MCG_DECODE') or ('gpu' if loader is None and _cuda() else 'host')
        super().__init__(ann_file, img_prefix=img_prefix or '', clip_len=int(clip_length),
                         stride=int(stride) if stride is not None else ev.slicer.STRIDE,
                         loader=loader, decode=decode)
        if classes is not None:
            self.CLASSES = tuple(classes) if not isinstance(classes, str) else classes
        name = ann_file if isinstance(ann_file, str) else ''
        self.scorer = scorer or ('l2cs' if 'l2cs' in name.lower() else 'gaze360')
        self.pipeline = None                                         # built per device by make_pipeline()
        print(f'origin__num = {sum(len(v) for v in self.videos)}')   # gaze360.py:82 prints the frame count

    def make_pipeline(self, device: int = 0):
        """The test pipeline of the config on `device` (mcgaze_b200.pipeline.GpuTestPipeline: geometry on the host in the
        reference's arithmetic, pixels in one mcg_preprocess launch per batch)."""
        factory = type(self).pipeline_factory
        if factory is not None:
            return factory(self.pipeline_cfg)
        from .pipeline import GpuTestPipeline
        return GpuTestPipeline(self.pipeline_cfg, device=device, seed=self.seed)

    # --- mmdet dataset surface used by tools/test.py ------------------------------------------------------------
    def format_results(self, results, results_file: str = 'results.json', **kwargs):
        """-> (records, results_file): overlaps merged, JSON written in the schema of tools/test_gaze360_gaze.py:210-260."""
        if len(results) != len(self):
            raise ValueError(f'The length of results is not equal to the dataset len: {len(results)} != {len(self)}')
        records, merged = ev.videos_from_clips(self, [np.asarray(r, dtype=np.float32) for r in results])
        d = osp.dirname(osp.abspath(results_file))
        os.makedirs(d, exist_ok=True)
        with open(results_file, 'w') as f:
            json.dump(records, f)
        self._merged = merged
        return records, results_file

    def evaluate(self, results, metric='mae', results_file: str = 'results.json', logger=None, jsonfile_prefix=None,
                 classwise: bool = False, scorer: Optional[str] = None, **kwargs) -> Dict[str, float]:
        """Same signature as gaze360.py:397-404.  Writes the results JSON (like the reference does) and returns
        {'<clue>_mae_360' / '_mae_front180' / '_mae_front20': degrees} from the scorer of tools/calculate_mae_gaze360.py
        (or calculate_mae_l2cs.py: `scorer='l2cs'`, default by the annotation file's name); an empty dict when the
        annotation file has no ground-truth gazes."""
        metrics = [metric] if isinstance(metric, str) else list(metric)
        for m in metrics:
            if m not in _METRICS:
                raise KeyError(f'metric {m} is not supported (use one of {_METRICS})')
        if jsonfile_prefix is not None:
            results_file = jsonfile_prefix + '.json'
        records, path = self.format_results(results, results_file)
        out: Dict[str, float] = {}
        if ev.ground_truth(self, scorer or self.scorer) is None:
            print(f'wrote {path}; the annotation file carries no ground-truth gazes, nothing to score')
            return out
        for key in _KEYS:
            m = ev.evaluate(self, records, key, variant=scorer or self.scorer)
            clue = key.split('_')[0]
            out[f'{clue}_mae_360'] = round(m['mae_360'], 4)
            out[f'{clue}_mae_front180'] = round(m['mae_front90'], 4)
            out[f'{clue}_mae_front20'] = round(m['mae_front20'], 4)
        return out


def build_dataset(cfg, default_args=None):
    """mmdet/datasets/builder.py:57-80.  ConcatDataset / RepeatDataset / ClassBalancedDataset wrap TRAINING sets."""
    if isinstance(cfg, (list, tuple)):
        raise NotImplementedError('concatenated test datasets are not supported')
    if cfg.get('type') in ('ConcatDataset', 'RepeatDataset', 'ClassBalancedDataset', 'MultiImageMixDataset'):
        raise NotImplementedError(f"{cfg['type']} is a training-side dataset wrapper")
    args = cfg.to_dict() if hasattr(cfg, 'to_dict') else dict(cfg)
    return build_from_cfg(args, DATASETS, default_args)


class ClipLoader:
    """What `build_dataloader` returns: the handle tools/test.py passes to single_gpu_test / multi_gpu_test, which hand its
    `dataset` to the batched driver (loading overlapped with the forward: decode threads, pinned blocks, copy stream).

    Iterating it gives what a loop like tools/analysis_tools/benchmark.py:105-112 expects from an mmdet DataLoader - the
    keyword arguments of one model call, `model(return_loss=False, rescale=True, **data)`: `img=[Tensor[n, 3, H, W]]` on the
    device, `img_metas=[[meta] * n]`, already unwrapped (there are no DataContainers to scatter), one clip per item at
    `samples_per_gpu=1`; with more samples per GPU an item holds several clips of one length and one canvas and carries
    `clip_length`.  With `dist=True` inside a process group a rank iterates the clips DistributedSampler(shuffle=False)
    gives it.  Frames are decoded on the host here (the synchronous path; the batched driver has the device decoder)."""

    def __init__(self, dataset, samples_per_gpu: int = 1, workers_per_gpu: int = 0, dist: bool = False, shuffle: bool = False):
        if shuffle:
            raise NotImplementedError('shuffle=True is a training-side option')
        self.dataset = dataset
        self.samples_per_gpu = max(int(samples_per_gpu), 1)
        self.workers_per_gpu = int(workers_per_gpu)
        self.dist = bool(dist)
        self.batch_size = self.samples_per_gpu

    @property
    def clips_per_batch(self) -> int:
        return max(self.samples_per_gpu, int(os.environ.get('MCG_CLIPS_PER_BATCH', '32')))

    def _indices(self):
        n = len(self.dataset)
        if self.dist:
            import torch.distributed as tdist
            if tdist.is_available() and tdist.is_initialized() and tdist.get_world_size() > 1:
                return ev.mdist.padded_shard(n, tdist.get_rank(), tdist.get_world_size())
        return list(range(n))

    def __len__(self):
        return -(-len(self._indices()) // self.samples_per_gpu)

    def __iter__(self):
        import torch
        ds = self.dataset
        pipe = ds.make_pipeline(torch.cuda.current_device() if torch.cuda.is_available() else 0)
        idx = self._indices()
        batches = [[i] for i in idx] if self.samples_per_gpu == 1 else ev._batches(ds, idx, self.samples_per_gpu)
        for batch in batches:
            T, frames, names = ev._load_batch(ds, batch, None)
            for sub, sub_names, rands, clips in ev._canvas_groups(pipe, frames, names, T):
                data = pipe.batch(sub, filenames=sub_names) if rands is None else pipe.batch(sub, rands=rands, filenames=sub_names)
                if len(clips) > 1:
                    data = dict(data, clip_length=T)
                yield data


def build_dataloader(dataset, samples_per_gpu, workers_per_gpu, num_gpus=1, dist=True, shuffle=True, seed=None,
                     runner_type='EpochBasedRunner', persistent_workers=False, class_aware_sampler=None, **kwargs):
    """mmdet/datasets/builder.py:83-180 for test-time use (`shuffle=False`)."""
    return ClipLoader(dataset, samples_per_gpu, workers_per_gpu, dist=dist, shuffle=shuffle)


def _cuda() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def _unwrap(model):
    return model.module if hasattr(model, 'module') else model


def _driver_args(model, data_loader):
    if not hasattr(data_loader, 'dataset'):
        raise TypeError('data_loader must come from build_dataloader (it carries the dataset)')
    ds = data_loader.dataset
    if not isinstance(ds, ev.Gaze360ClipDataset):
        raise TypeError(f'the batched test driver needs a clip dataset, got {type(ds).__name__}')
    m = _unwrap(model)
    make = getattr(ds, 'make_pipeline', None)
    if make is None:
        raise TypeError('dataset has no make_pipeline(); build it with build_dataset(cfg.data.test)')
    pipe = make(getattr(m, 'device_index', 0))
    cpb = getattr(data_loader, 'clips_per_batch', None) or max(int(getattr(data_loader, 'samples_per_gpu', 1)), 1)
    return m, ds, pipe, cpb, int(getattr(data_loader, 'workers_per_gpu', 0))


def single_gpu_test(model, data_loader, show=False, out_dir=None, show_score_thr=0.3):
    """mmdet/apis/test.py:17-78 with the reference's signature -> per-clip rows in dataset order."""
    if show or out_dir:
        raise NotImplementedError('--show / --show-dir (painting detections) is not part of the inference backend')
    m, ds, pipe, cpb, workers = _driver_args(model, data_loader)
    return ev.single_gpu_test(m, ds, pipe, clips_per_batch=cpb, workers=workers)


def multi_gpu_test(model, data_loader, tmpdir=None, gpu_collect=False):
    """mmdet/apis/test.py:81-126 with the reference's signature.  `tmpdir` / `gpu_collect` choose HOW the reference
    collects the parts (pickles in a shared directory :129-173, or pickled byte tensors :176-209); here it is always ONE
    all-gather of the packed float rows over the process group's backend, and every rank returns the full list."""
    import torch.distributed as tdist
    m, ds, pipe, cpb, workers = _driver_args(model, data_loader)
    if not (tdist.is_available() and tdist.is_initialized()) or tdist.get_world_size() == 1:
        return ev.single_gpu_test(m, ds, pipe, clips_per_batch=cpb, workers=workers)
    device = f"cuda:{getattr(m, 'device_index', 0)}" if tdist.get_backend() == 'nccl' else None
    return ev.multi_gpu_test(m, ds, pipe, clips_per_batch=cpb, device=device, workers=workers)
