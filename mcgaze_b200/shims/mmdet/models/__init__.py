"""mmdet.models: registries and builders (mmdet/models/builder.py:1-59) resolving to the backend's classes."""
from mcgaze_b200 import detector as _detector  # noqa: F401  (registers MultiClueGaze and its sub-module specs)
from mcgaze_b200.registry import (BACKBONES, DETECTORS, HEADS, LOSSES, MODELS, NECKS, ROI_EXTRACTORS,  # noqa: F401
                                  SHARED_HEADS, TRANSFORMER, build_backbone, build_detector, build_head, build_loss,
                                  build_neck, build_roi_extractor, build_transformer)
