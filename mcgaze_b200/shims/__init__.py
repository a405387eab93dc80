"""Import-name shims: `mmdet` and `mmcv` packages that re-export the B200 backend under the module paths the
reference's inference tools import (tools/test_gaze360_gaze.py:8-16, mmdet/apis/inference.py:4-14), so those tools run
UNMODIFIED with this directory on PYTHONPATH:

    PYTHONPATH=$(python -m mcgaze_b200.shims) python /path/to/MCGaze/tools/test_gaze360_gaze.py <config> <checkpoint>

Only the inference surface exists (init_detector, Compose for test pipelines, collate / scatter / DataContainer, Config /
DictAction, registries, load_checkpoint); everything training-side raises.  No compute lives here: the model is
mcgaze_b200.detector.MultiClueGaze (libmcgaze_b200.so), the image pipeline is mcgaze_b200.pipeline.GpuTestPipeline."""
import os

PATH = os.path.dirname(os.path.abspath(__file__))


def path() -> str:
    return PATH
