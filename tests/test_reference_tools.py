"""The reference's own tools on this backend (north_star: "tools ... run unchanged").

CPU: tests/golden/golden_slicer_reference.json is what the reference's tools/test_gaze360_gaze.py - executed UNMODIFIED
on the `mmdet` / `mmcv` import shims of mcgaze_b200/shims, with a deterministic stand-in model (oracle/gen_golden_slicer.py)
- writes for 11 videos of every length class.  The batched driver of this repo (slicer + evaluate) must reproduce it
record for record; when /root/reference is present the script is run again, live.  GPU: the device merge kernel against
the same golden, and the shims' real `init_detector` / `Compose` in the script's own per-clip call sequence against the
batched driver."""
import importlib.util
import json
import os
import sys

import numpy as np
import pytest
import torch

from mcgaze_b200 import evaluate as ev
from oracle import stub_clip_model as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get('MCGAZE_REFERENCE', '/root/reference')


@pytest.fixture(scope='module')
def golden(golden_dir):
    return json.load(open(os.path.join(golden_dir, 'golden_slicer_reference.json')))


def _dataset():
    return ev.Gaze360ClipDataset(S.make_anno(), img_prefix='frames', loader=S.encode_frame)


def _same_records(ours, theirs, atol=0.0):
    assert len(ours) == len(theirs)
    for a, b in zip(ours, theirs):
        assert list(a.keys()) == list(b.keys()) and a['video_id'] == b['video_id'] and a['category_id'] == b['category_id']
        for k in a:
            if k in ('video_id', 'category_id'):
                continue
            assert len(a[k]) == len(b[k]), k
            for x, y in zip(a[k], b[k]):
                assert (x is None) == (y is None), (a['video_id'], k)
                if x is not None:
                    assert np.allclose(np.asarray(x, np.float64), np.asarray(y, np.float64), rtol=0, atol=atol), (a['video_id'], k)


def test_batched_driver_reproduces_the_reference_scripts_json(golden):
    assert golden['lengths'] == S.LENGTHS
    ds = _dataset()
    model = S.StubModel()
    rows = ev.single_gpu_test(model, ds, S.StubBatchPipeline(), clips_per_batch=4)
    records, _ = ev.videos_from_clips(ds, rows)
    assert len(model.calls) < golden['forwards'] == len(ds)              # many clips per forward vs one
    _same_records(records, golden['records'], atol=1e-6)                 # (the stand-in's sin / cos are batch-shape sensitive)
    # one clip per forward, like the script: bit-identical records
    rows1 = ev.single_gpu_test(S.StubModel(), ds, S.StubBatchPipeline(), clips_per_batch=1)
    _same_records(ev.videos_from_clips(ds, rows1)[0], golden['records'], atol=0.0)


@pytest.mark.skipif(not os.path.isdir(REF), reason='the reference checkout only exists in the build container')
def test_reference_script_runs_unmodified_on_the_shims(golden):
    spec = importlib.util.spec_from_file_location('gen_golden_slicer', os.path.join(ROOT, 'oracle', 'gen_golden_slicer.py'))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    records, calls = gen.run_reference_driver(S.make_anno())
    assert records == golden['records'] and len(calls) == golden['forwards']


def test_shim_packages_expose_what_the_reference_tools_import():
    """tools/test_gaze360_gaze.py:8-16 and mmdet/apis/inference.py:4-14, by name."""
    from mcgaze_b200 import shims
    sys.path.insert(0, shims.PATH)
    try:
        import mmcv
        import mmdet
        from mmcv import Config, DictAction  # noqa: F401
        from mmcv.cnn.utils.flops_counter import add_flops_counting_methods, flops_to_string, params_to_string  # noqa: F401
        from mmcv.parallel import collate, scatter  # noqa: F401
        from mmcv.runner import load_checkpoint  # noqa: F401
        from mmdet.apis import init_detector, multi_gpu_test, single_gpu_test  # noqa: F401
        from mmdet.core.bbox import bbox_overlaps
        from mmdet.datasets import replace_ImageToTensor
        from mmdet.datasets.pipelines import Compose  # noqa: F401
        from mmdet.models import DETECTORS, build_detector  # noqa: F401
        assert 'MultiClueGaze' in DETECTORS.module_dict and mmcv.__version__ and mmdet.__version__
        iou = bbox_overlaps(torch.tensor([[0., 0., 10., 10.]]), torch.tensor([[5., 5., 15., 15.], [0., 0., 10., 10.]]))
        assert torch.allclose(iou, torch.tensor([[25 / 175, 1.0]]))
        assert replace_ImageToTensor([dict(type='ImageToTensor', keys=['img'])]) == [dict(type='DefaultFormatBundle')]
        m = add_flops_counting_methods(S.StubModel())
        m.start_flops_count(), m.stop_flops_count()
    finally:
        sys.path.remove(shims.PATH)


@pytest.mark.gpu
def test_gpu_merge_kernel_reproduces_the_reference_scripts_json(golden):
    """mcg_merge_clips on the stand-in's per-clip outputs == the merge the reference script did (bit-exact)."""
    from mcgaze_b200 import lib, slicer
    ds = _dataset()
    rows = ev.single_gpu_test(S.StubModel(), ds, S.StubBatchPipeline(), clips_per_batch=1)
    packed = np.zeros((len(ds), 7, 27), np.float32)
    for i, r in enumerate(rows):
        packed[i, :r.shape[0]] = r
    det, gz = lib.merge_clips(torch.from_numpy(packed).cuda(), [len(p) for p in ds.plans], S.LENGTHS)
    det, gz = det.cpu().numpy(), gz.cpu().numpy()
    off = np.concatenate([[0], np.cumsum(S.LENGTHS)])
    records = [slicer.video_record(vi + 1, dict(det=det[off[vi]:off[vi + 1]], gaze=gz[off[vi]:off[vi + 1]]))
               for vi in range(len(S.LENGTHS))]
    _same_records(records, golden['records'], atol=0.0)


@pytest.mark.gpu
def test_gpu_shims_in_the_reference_scripts_call_sequence(synthetic_sd, tmp_path):
    """init_detector + Compose + collate + scatter from the SHIM packages, driven exactly like
    tools/test_gaze360_gaze.py:88-111 drives them (one thread per frame, sort by filename, collate, scatter, model
    call), on PNG files, against the batched driver on the same crops."""
    cv2 = pytest.importorskip('cv2')
    from threading import Thread
    from mcgaze_b200 import shims
    sys.path.insert(0, shims.PATH)
    try:
        from mmcv.parallel import collate, scatter
        from mmdet.apis import init_detector
        from mmdet.datasets.pipelines import Compose
    finally:
        sys.path.remove(shims.PATH)
    rng = np.random.default_rng(2)
    names = [f'v000/{t:05d}.png' for t in range(7)]
    os.makedirs(tmp_path / 'v000')
    for n in names:
        assert cv2.imwrite(str(tmp_path / n), rng.integers(0, 256, (110, 96, 3), dtype=np.uint8))
    cfg = os.path.join(ROOT, 'configs/multiclue_gaze/multiclue_gaze_r50_gaze360.py')
    model = init_detector(cfg, None, device='cuda:0')
    model.load_state_dict(synthetic_sd)
    test_pipeline = Compose(model.cfg.data.test.pipeline)
    np.random.seed(5)
    draws = np.random.rand(7)
    np.random.seed(5)
    datas, threads = [], []
    for img in names:                       # sequential start/join keeps the crop draws in frame order for the comparison
        data = dict(img_info=dict(filename=img), img_prefix=str(tmp_path))
        t = Thread(target=lambda d=data: datas.append(test_pipeline(d)))
        t.start()
        t.join()
    datas = sorted(datas, key=lambda x: x['img_metas'].data['filename'])
    datas = collate(datas, samples_per_gpu=len(names))
    datas['img_metas'] = datas['img_metas'].data
    datas['img'] = datas['img'].data
    datas = scatter(datas, ['cuda:0'])[0]
    (det_bboxes, det_labels), det_gazes = model(return_loss=False, rescale=True, format=False, **datas)
    det = torch.stack(det_bboxes)
    assert det.shape == (7, 3, 5) and det_gazes['gaze_score'].shape == (7, 3)

    ds = ev.Gaze360ClipDataset(dict(videos=[dict(id=1, file_names=names)]), img_prefix=str(tmp_path))

    class FixedDraws(type(test_pipeline.pipeline)):
        def draw(self, n):
            return draws[:n]

    pipe = FixedDraws(model.cfg.data.test.pipeline)
    rows = ev.single_gpu_test(model, ds, pipe, clips_per_batch=4)[0]
    assert np.allclose(rows[:, :12].reshape(7, 3, 4), det[..., :4].cpu().numpy(), atol=1e-4)
    assert np.allclose(rows[:, 15:18], det_gazes['gaze_score'].cpu().numpy(), atol=1e-6)


@pytest.mark.gpu
def test_gpu_demo_notebook_call_sequence(synthetic_sd):
    """MCGaze_demo/demo.ipynb cells 2-4 on the shims: init_detector with the l2cs config, Compose(pipeline[1:]) on head crops
    the caller decoded itself (`dict(filename=j, ori_filename=111, img=crop, img_shape=..., ori_shape=..., img_fields=['img'])`,
    crops of DIFFERENT sizes), sort by meta['filename'], collate(samples_per_gpu=<clip length>), scatter, ONE model call over
    the whole track - against GpuTestPipeline.batch + the same model on the same crops."""
    from mcgaze_b200 import shims
    sys.path.insert(0, shims.PATH)
    try:
        from mmcv.parallel import collate, scatter
        from mmdet.apis import init_detector
        from mmdet.datasets.pipelines import Compose
    finally:
        sys.path.remove(shims.PATH)
    cfg_path = os.path.join(ROOT, 'configs/multiclue_gaze/multiclue_gaze_r50_l2cs.py')
    model = init_detector(cfg_path, None, device='cuda:0', cfg_options=None)
    model.load_state_dict(synthetic_sd)
    cfg = model.cfg
    test_pipeline = Compose(cfg.data.test.pipeline[1:])
    rng = np.random.default_rng(8)
    sizes = [(90, 90), (96, 88), (90, 90), (101, 97), (84, 90)]
    crops = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in sizes]
    datas = []
    for j in reversed(range(len(crops))):                                      # out of order: the sort must restore it
        c = crops[j]
        datas.append(test_pipeline(dict(filename=j, ori_filename=111, img=c, img_shape=c.shape, ori_shape=(2 * 45, 2 * 45, 3),
                                        img_fields=['img'])))
    datas = sorted(datas, key=lambda x: x['img_metas'].data['filename'])
    assert [d['img_metas'].data['filename'] for d in datas] == list(range(len(crops)))
    assert datas[0]['img_metas'].data['ori_filename'] == 111 and datas[0]['img_metas'].data['ori_shape'] == (90, 90, 3)
    datas = collate(datas, samples_per_gpu=len(crops))
    datas['img_metas'] = datas['img_metas'].data
    datas['img'] = datas['img'].data
    datas = scatter(datas, ['cuda:0'])[0]
    with torch.no_grad():
        (det_bboxes, det_labels), det_gazes = model(return_loss=False, rescale=True, format=False, **datas)
    g = det_gazes['gaze_score']
    assert g.shape == (len(crops), 3) and len(det_bboxes) == len(crops) and det_bboxes[0].shape == (3, 5)
    assert torch.allclose(g.norm(dim=1), torch.ones(len(crops), device=g.device), atol=1e-4)
    # the same crops through the batched pipeline call (no crop step in the l2cs pipeline: nothing random)
    from mcgaze_b200.pipeline import GpuTestPipeline
    data = GpuTestPipeline(cfg.data.test.pipeline, device=0).batch(crops, filenames=list(range(len(crops))))
    assert torch.equal(data['img'][0], datas['img'][0])
    (det2, _), gz2 = model(return_loss=False, rescale=True, format=False, **data)
    assert torch.equal(gz2['gaze_score'], g) and all(torch.equal(a, b) for a, b in zip(det2, det_bboxes))
