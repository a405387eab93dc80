"""mmcv.runner / mmcv.parallel / mmcv.fileio pieces that the reference's generic tester touches
(tools/test.py:9-13, :186-221): process-group set-up, rank info, the
(Distributed)DataParallel wrappers - which have nothing to scatter or synchronise for an inference engine with one replica
per process - and dump / load / mkdir_or_exist."""
from __future__ import annotations

import json
import os
import pickle
from typing import Any, Optional, Tuple


def get_dist_info() -> Tuple[int, int]:
    """mmcv.runner.get_dist_info: (rank, world_size), (0, 1) outside a process group."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def init_dist(launcher: str, backend: str = 'nccl', **kwargs) -> None:
    """mmcv.runner.init_dist for the `pytorch` launcher (what tools/dist_test.sh uses): RANK / WORLD_SIZE / MASTER_* come
    from torch.distributed.launch / torchrun; the process takes GPU LOCAL_RANK (mmcv: rank % device_count)."""
    import torch
    import torch.distributed as dist
    if launcher != 'pytorch':
        raise NotImplementedError(f"launcher {launcher!r}: only 'pytorch' (torch.distributed.launch / torchrun) is supported")
    if 'RANK' not in os.environ:
        raise RuntimeError("init_dist('pytorch'): RANK is not set - start the script with torch.distributed.launch / torchrun")
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    rank = int(os.environ['RANK'])
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    if n > 0:
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', rank % n)) % n)
    elif backend == 'nccl':
        raise RuntimeError('init_dist: backend nccl needs CUDA devices (pass dist_params.backend=gloo for a CPU dry run)')
    if not dist.is_initialized():
        dist.init_process_group(backend=backend, **kwargs)


def wrap_fp16_model(model) -> None:
    """tools/test.py:198-200 calls this when the config has an `fp16` key.  The engine's arithmetic is fixed by its
    precision mode (fp16 + e4m3 corrected tensor-core products, fp32 accumulation); there is nothing to wrap."""
    raise NotImplementedError("config key 'fp16' is not supported: choose the engine precision instead "
                              "(MultiClueGaze(precision='fp16c8' | 'fp16x3' | 'fp16'))")


def fuse_conv_bn(model):
    """mmcv.cnn.fuse_conv_bn (--fuse-conv-bn): BatchNorm is always folded into the convolution weights when the engine
    packs a checkpoint (mcg_create), so the model is returned as it is."""
    return model


class _Replica:
    """Shared behaviour of the two wrappers: `.module`, call-through, attribute fall-through."""

    def __init__(self, module, device_ids=None, **kwargs):
        self.module = module
        self.device_ids = list(device_ids) if device_ids is not None else None
        if self.device_ids:
            dev = self.device_ids[0]
            if hasattr(module, 'to'):
                module.to(f'cuda:{dev}' if isinstance(dev, int) else dev)

    def __call__(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def eval(self):
        if hasattr(self.module, 'eval'):
            self.module.eval()
        return self

    def __getattr__(self, name):
        if name == 'module':
            raise AttributeError(name)
        return getattr(self.module, name)


class MMDataParallel(_Replica):
    """mmcv.parallel.MMDataParallel(model, device_ids=[gpu]) (tools/test.py:210): one device; torch's DataParallel moves a
    single-device module to device_ids[0], so does this."""

    def __init__(self, module, device_ids=None, dim: int = 0, **kwargs):
        if device_ids is not None and len(device_ids) > 1:
            raise NotImplementedError('MMDataParallel over several GPUs in one process: start one process per GPU instead')
        super().__init__(module, device_ids if device_ids is not None else [0])


class MMDistributedDataParallel(_Replica):
    """mmcv.parallel.MMDistributedDataParallel(model.cuda(), device_ids=[current], broadcast_buffers=False)
    (tools/test.py:214-217): inference replicas share nothing, so there are no gradients or buffers to synchronise."""

    def __init__(self, module, device_ids=None, broadcast_buffers: bool = False, find_unused_parameters: bool = False, **kwargs):
        super().__init__(module, device_ids)


def mkdir_or_exist(dir_name: str, mode: int = 0o777) -> None:
    if dir_name:
        os.makedirs(os.path.expanduser(dir_name), mode=mode, exist_ok=True)


def _np_default(o: Any):
    import numpy as np
    if isinstance(o, np.ndarray):
        return o.tolist()
    if isinstance(o, np.generic):
        return o.item()
    raise TypeError(f'{type(o)} is not JSON serialisable')


def dump(obj: Any, file: Optional[str] = None, file_format: Optional[str] = None, **kwargs):
    """mmcv.dump: format by extension (.pkl / .pickle, .json); without `file` the serialised string / bytes."""
    fmt = file_format or (os.path.splitext(str(file))[1].lstrip('.').lower() if file is not None else None)
    if fmt in ('pkl', 'pickle'):
        kwargs.setdefault('protocol', 2)
        if file is None:
            return pickle.dumps(obj, **kwargs)
        with open(file, 'wb') as f:
            pickle.dump(obj, f, **kwargs)
    elif fmt == 'json':
        kwargs.setdefault('default', _np_default)
        if file is None:
            return json.dumps(obj, **kwargs)
        with open(file, 'w') as f:
            json.dump(obj, f, **kwargs)
    else:
        raise TypeError(f'Unsupported format: {fmt}')


def load(file: str, file_format: Optional[str] = None, **kwargs):
    fmt = file_format or os.path.splitext(str(file))[1].lstrip('.').lower()
    if fmt in ('pkl', 'pickle'):
        with open(file, 'rb') as f:
            return pickle.load(f, **kwargs)
    if fmt == 'json':
        with open(file) as f:
            return json.load(f, **kwargs)
    raise TypeError(f'Unsupported format: {fmt}')
