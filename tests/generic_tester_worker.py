"""TEST INFRASTRUCTURE: runs the reference's generic tester (tools/test.py, UNMODIFIED, via runpy) on this repo's `mmdet` /
`mmcv` import shims with the CPU stand-in detector / frame loader / pipeline of oracle/stub_clip_model.py.

    python tests/generic_tester_worker.py <reference root> [tool path relative to it] <args of the tool ...>

(the tool defaults to tools/test.py; tools/analysis_tools/benchmark.py is the other one run this way)

Also the per-rank entry of the 2-process gloo run (started by tests/test_reference_tools.py with RANK / WORLD_SIZE /
MASTER_* set): on a machine without GPUs `torch.cuda.current_device()` (tools/test.py:216) is answered with 0."""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(ref: str, argv):
    sys.path.insert(0, ROOT)
    from mcgaze_b200 import shims
    sys.path.insert(0, shims.PATH)
    import torch
    from mcgaze_b200.datasets import Gaze360Dataset
    from mcgaze_b200.registry import DETECTORS
    from oracle import stub_clip_model as S
    DETECTORS.register_module(name='StubClipDetector', module=S.StubDetector, force=True)
    saved = (Gaze360Dataset.frame_loader, Gaze360Dataset.pipeline_factory, sys.argv, torch.cuda.current_device)
    Gaze360Dataset.frame_loader = staticmethod(S.encode_frame)
    Gaze360Dataset.pipeline_factory = staticmethod(lambda cfg: S.StubBatchPipeline())
    if not torch.cuda.is_available():
        torch.cuda.current_device = lambda: 0
    tool = 'tools/test.py'
    if argv and argv[0].endswith('.py') and os.path.isfile(os.path.join(ref, argv[0])) and argv[0].startswith('tools/'):
        tool, argv = argv[0], argv[1:]
    sync = torch.cuda.synchronize
    if not torch.cuda.is_available():
        torch.cuda.synchronize = lambda *a, **k: None
    sys.argv = [os.path.join(ref, tool)] + list(argv)
    try:
        runpy.run_path(sys.argv[0], run_name='__main__')
    finally:
        Gaze360Dataset.frame_loader, Gaze360Dataset.pipeline_factory, sys.argv, torch.cuda.current_device = saved
        torch.cuda.synchronize = sync
        sys.path.remove(shims.PATH)


if __name__ == '__main__':
    run(sys.argv[1], sys.argv[2:])
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()
