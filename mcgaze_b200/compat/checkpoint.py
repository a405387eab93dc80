"""mmcv.runner.load_checkpoint work-alike (mmdet/apis/inference.py:45):
torch pickle, `state_dict` sub-key if present, regex `revise_keys`, NON-strict load that reports
missing / unexpected keys, returns the checkpoint dict (so callers can read meta.CLASSES)."""
from __future__ import annotations

import re
from collections import OrderedDict
from typing import Any, Dict, List, Optional, Sequence, Tuple


def revise_state_dict(state_dict: Dict[str, Any], revise_keys: Sequence[Tuple[str, str]]) -> 'OrderedDict[str, Any]':
    out = OrderedDict(state_dict)
    for p, r in revise_keys:
        out = OrderedDict((re.sub(p, r, k), v) for k, v in out.items())
    return out


def load_state_dict(module, state_dict: Dict[str, Any], strict: bool = False, logger=None) -> Dict[str, List[str]]:
    """Non-strict load into anything exposing `load_state_dict` / `state_dict` (torch modules
    and the mcgaze_b200 backend detector).  Returns {'missing': [...], 'unexpected': [...]}."""
    own = set(module.state_dict().keys())
    given = set(state_dict.keys())
    report = {'missing': sorted(k for k in own - given if 'num_batches_tracked' not in k),
              'unexpected': sorted(given - own)}
    module.load_state_dict(state_dict, strict=False)
    own_report = getattr(module, 'load_report', None)
    if isinstance(own_report, dict):      # the backend detector knows which reference keys its engine ignores
        report = {'missing': sorted(own_report['missing']), 'unexpected': sorted(own_report['unexpected'])}
    if strict and (report['missing'] or report['unexpected']):
        raise RuntimeError(f"state_dict mismatch: missing {report['missing'][:8]}, unexpected {report['unexpected'][:8]}")
    if logger is not None and (report['missing'] or report['unexpected']):
        logger.warning('checkpoint/model key mismatch: %s', report)
    return report


def load_checkpoint(model, filename: str, map_location: Optional[str] = 'cpu', strict: bool = False, logger=None,
                    revise_keys: Sequence[Tuple[str, str]] = ((r'^module\.', ''),)) -> Dict[str, Any]:
    import torch
    ckpt = torch.load(filename, map_location=map_location, weights_only=False)
    if not isinstance(ckpt, dict):
        raise RuntimeError(f'No state_dict found in checkpoint file {filename}')
    sd = ckpt['state_dict'] if 'state_dict' in ckpt else ckpt
    sd = revise_state_dict(sd, revise_keys)
    load_state_dict(model, sd, strict, logger)
    return ckpt
