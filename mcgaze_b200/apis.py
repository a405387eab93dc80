"""`init_detector` of the reference (mmdet/apis/inference.py:17-56) for the B200 backend."""
from __future__ import annotations

import warnings
from typing import Optional, Union

from .compat import Config, load_checkpoint
from .registry import build_detector
from . import detector as _detector  # noqa: F401  (registers the model classes)


def init_detector(config: Union[str, Config], checkpoint: Optional[str] = None, device: str = 'cuda:0',
                  cfg_options: Optional[dict] = None, precision: Optional[str] = None):
    if isinstance(config, str):
        config = Config.fromfile(config)
    elif not isinstance(config, Config):
        raise TypeError(f'config must be a filename or Config object, but got {type(config)}')
    if cfg_options is not None:
        config.merge_from_dict(cfg_options)
    if 'pretrained' in config.model:
        config.model.pretrained = None
    if 'init_cfg' in config.model.get('backbone', {}):
        config.model.backbone.init_cfg = None
    config.model.train_cfg = None
    model_cfg = dict(config.model.to_dict())
    # engine precision: the argument, else the config's `engine = dict(precision=...)` (configs/_base_/default_runtime.py),
    # else the parity mode fp16c8
    model_cfg['precision'] = precision or (config.get('engine') or {}).get('precision', 'fp16c8')
    model = build_detector(model_cfg, test_cfg=config.get('test_cfg'))
    if checkpoint is not None:
        ckpt = load_checkpoint(model, checkpoint, map_location='cpu',
                               revise_keys=[(r'^module\.', ''), ('mask_head', 'blink_head')])
        meta = ckpt.get('meta', {}) if isinstance(ckpt, dict) else {}
        if 'CLASSES' in meta:
            model.CLASSES = meta['CLASSES']
        else:
            warnings.simplefilter('once')
            warnings.warn("Class names are not saved in the checkpoint's meta data, keeping face/eyes/head.")
    model.cfg = config
    model.to(device)
    model.eval()
    return model
