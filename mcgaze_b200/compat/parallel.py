"""mmcv.parallel DataContainer / collate / scatter work-alikes, restricted to what
tools/test_gaze360_gaze.py:98-101 does with them."""
from __future__ import annotations

from collections.abc import Mapping, Sequence
from typing import Any, List


class DataContainer:
    def __init__(self, data, stack: bool = False, padding_value: int = 0, cpu_only: bool = False, pad_dims: int = 2):
        self._data = data
        self._stack = stack
        self._padding_value = padding_value
        self._cpu_only = cpu_only
        self._pad_dims = pad_dims

    data = property(lambda self: self._data)
    stack = property(lambda self: self._stack)
    cpu_only = property(lambda self: self._cpu_only)
    padding_value = property(lambda self: self._padding_value)
    pad_dims = property(lambda self: self._pad_dims)

    def __repr__(self):
        return f'DataContainer({self._data!r})'


def collate(batch: Sequence[Any], samples_per_gpu: int = 1):
    """Groups of `samples_per_gpu`; stacked tensors are right/bottom zero-padded to the group max."""
    import torch
    import torch.nn.functional as F
    if not isinstance(batch, Sequence):
        raise TypeError(f'{type(batch)} is not supported.')
    first = batch[0]
    if isinstance(first, DataContainer):
        out: List[Any] = []
        if first.cpu_only:
            for i in range(0, len(batch), samples_per_gpu):
                out.append([s.data for s in batch[i:i + samples_per_gpu]])
            return DataContainer(out, first.stack, first.padding_value, cpu_only=True)
        if first.stack:
            for i in range(0, len(batch), samples_per_gpu):
                grp = [s.data for s in batch[i:i + samples_per_gpu]]
                nd = first.pad_dims or 0
                if nd:
                    mx = [max(t.shape[-d] for t in grp) for d in range(1, nd + 1)]
                    padded = []
                    for t in grp:
                        pad = []
                        for d in range(1, nd + 1):
                            pad += [0, mx[d - 1] - t.shape[-d]]
                        padded.append(F.pad(t, pad, value=first.padding_value))
                    grp = padded
                out.append(torch.stack(grp, 0))
            return DataContainer(out, True, first.padding_value)
        for i in range(0, len(batch), samples_per_gpu):
            out.append([s.data for s in batch[i:i + samples_per_gpu]])
        return DataContainer(out, False, first.padding_value)
    if isinstance(first, Mapping):
        return {k: collate([d[k] for d in batch], samples_per_gpu) for k in first}
    if isinstance(first, Sequence) and not isinstance(first, (str, bytes)):
        return [collate(list(s), samples_per_gpu) for s in zip(*batch)]
    return torch.utils.data.dataloader.default_collate(list(batch))


def scatter(inputs, target_gpus, dim: int = 0):
    """Single-target scatter: unwrap DataContainers and move tensors to target_gpus[0]
    (-1 or 'cpu' keeps them on the host).  Returns a 1-element list like mmcv."""
    import torch
    dev = target_gpus[0]
    device = None if dev in (-1, 'cpu') else (torch.device('cuda', dev) if isinstance(dev, int) else torch.device(dev))

    def go(x):
        if isinstance(x, DataContainer):
            if x.cpu_only:
                return x.data[0]
            return go(x.data[0])
        if torch.is_tensor(x):
            return x.to(device, non_blocking=True) if device is not None else x
        if isinstance(x, Mapping):
            return {k: go(v) for k, v in x.items()}
        if isinstance(x, (list, tuple)):
            return type(x)(go(v) for v in x)
        return x

    return [go(inputs)]
