"""Golden output of the reference's OWN evaluation driver (tools/test_gaze360_gaze.py, executed UNMODIFIED).

    python oracle/gen_golden_slicer.py        # needs /root/reference (this container only)

The script is run in-process (runpy) with this repo's `mmdet` / `mmcv` import shims (mcgaze_b200/shims) on sys.path -
which is exactly how a user runs it on the B200 backend - but with `init_detector` and `Compose` swapped for the
deterministic CPU stand-ins of oracle/stub_clip_model.py, so it needs neither a GPU nor image files.  Everything else is
the reference's code: clip slicing (:60-86), the per-clip thread / sort / collate / scatter sequence (:88-101), the
model call (:107-111), the overlap merge (:129-201) and the JSON records (:210-260).  The JSON it writes is stored as
tests/golden/golden_slicer_reference.json and pins mcgaze_b200.slicer / evaluate / mcg_merge_clips.
TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import json
import os
import runpy
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get('MCGAZE_REFERENCE', '/root/reference')
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def run_reference_driver(anno: dict, config: str = None):
    """-> the list of per-video records tools/test_gaze360_gaze.py writes for `anno` with the stand-in model."""
    from mcgaze_b200 import shims
    from mcgaze_b200.compat import Config
    from oracle import stub_clip_model as S
    if shims.PATH not in sys.path:
        sys.path.insert(0, shims.PATH)
    import mmdet.apis
    import mmdet.datasets.pipelines
    config = config or os.path.join(ROOT, 'configs/multiclue_gaze/multiclue_gaze_r50_gaze360.py')
    models = []

    def init_detector(cfg_path, checkpoint, device='cuda:0', cfg_options=None):
        m = S.StubModel(Config.fromfile(cfg_path))
        models.append(m)
        return m

    saved = (mmdet.apis.init_detector, mmdet.datasets.pipelines.Compose, sys.argv, os.getcwd())
    with tempfile.TemporaryDirectory() as tmp:
        json.dump(anno, open(os.path.join(tmp, 'test.json'), 'w'))
        try:
            mmdet.apis.init_detector = init_detector
            mmdet.datasets.pipelines.Compose = S.StubCompose
            sys.argv = ['test_gaze360_gaze.py', config, 'none.pth', '--json', os.path.join(tmp, 'test.json'), '--root', 'frames',
                        '--device', 'cpu']
            os.chdir(tmp)
            runpy.run_path(os.path.join(REF, 'tools', 'test_gaze360_gaze.py'), run_name='__main__')
            out = [f for f in os.listdir(os.path.join(tmp, 'results'))]
            assert len(out) == 1, out
            records = json.load(open(os.path.join(tmp, 'results', out[0])))
        finally:
            mmdet.apis.init_detector, mmdet.datasets.pipelines.Compose, sys.argv = saved[:3]
            os.chdir(saved[3])
    return records, models[0].calls


def main():
    from oracle import stub_clip_model as S
    records, calls = run_reference_driver(S.make_anno())
    assert all(n == T for n, T in calls)                       # the reference runs ONE clip per forward
    path = os.path.join(ROOT, 'tests', 'golden', 'golden_slicer_reference.json')
    json.dump(dict(lengths=S.LENGTHS, forwards=len(calls), records=records), open(path, 'w'))
    print(path, os.path.getsize(path), 'bytes;', len(calls), 'forwards')


if __name__ == '__main__':
    main()
