"""mmdet.datasets: what the inference tools import (tools/test_gaze360_gaze.py:14-15, tools/test.py:16-17)."""
from mcgaze_b200.datasets import (DATASETS, PIPELINES, Gaze360Dataset, build_dataloader,  # noqa: F401
                                  build_dataset)

from . import pipelines  # noqa: F401
from .pipelines import Compose  # noqa: F401


def replace_ImageToTensor(pipelines_cfg):
    """mmdet/datasets/utils.py: ImageToTensor -> DefaultFormatBundle in a (possibly nested) pipeline config, needed
    when a test pipeline is used with batch size > 1.  The GPU pipeline skips both steps (its output IS the bundle)."""
    import copy
    pipelines_cfg = copy.deepcopy(pipelines_cfg)
    for i, step in enumerate(pipelines_cfg):
        if step['type'] == 'MultiScaleFlipAug':
            step['transforms'] = replace_ImageToTensor(step['transforms'])
        elif step['type'] == 'ImageToTensor':
            pipelines_cfg[i] = {'type': 'DefaultFormatBundle'}
    return pipelines_cfg
