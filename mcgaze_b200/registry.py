"""mmdet-style registries for the B200 backend (mirrors mmdet/models/builder.py:1-59 and
mmdet/models/utils/builder.py:5-10 of the reference): the `type='...'` strings of the
reference's configs resolve to the classes in mcgaze_b200.detector."""
from __future__ import annotations

from .compat import Registry, build_from_cfg

MODELS = Registry('models')
BACKBONES = NECKS = ROI_EXTRACTORS = SHARED_HEADS = HEADS = LOSSES = DETECTORS = MODELS
TRANSFORMER = Registry('Transformer')


def build_backbone(cfg):
    return BACKBONES.build(cfg)


def build_neck(cfg):
    return NECKS.build(cfg)


def build_roi_extractor(cfg):
    return ROI_EXTRACTORS.build(cfg)


def build_head(cfg):
    return HEADS.build(cfg)


def build_loss(cfg):
    return LOSSES.build(cfg)


def build_transformer(cfg, default_args=None):
    return build_from_cfg(cfg, TRANSFORMER, default_args)


def build_detector(cfg, train_cfg=None, test_cfg=None):
    """mmdet.models.build_detector (builder.py:49-59)."""
    assert cfg.get('train_cfg') is None or train_cfg is None
    assert cfg.get('test_cfg') is None or test_cfg is None
    return DETECTORS.build(cfg, default_args=dict(train_cfg=train_cfg, test_cfg=test_cfg))
