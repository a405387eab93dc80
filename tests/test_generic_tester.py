"""the reference's GENERIC tester (tools/test.py, tools/dist_test.sh) on this backend.

In the reference that path cannot evaluate the gaze configs (Gaze360Dataset raises in test mode,
mmdet/datasets/gaze360.py:310-312).  mcgaze_b200/datasets.py + the shim packages give it a test-mode clip dataset, a loader
handle, the (Distributed)DataParallel wrappers and `single_gpu_test` / `multi_gpu_test` with the reference's signatures.

CPU: tools/test.py is executed UNMODIFIED (runpy, in a subprocess; single process and 2 ranks over gloo) with the stand-in
detector of oracle/stub_clip_model.py; the JSON it makes `dataset.evaluate` write must equal the golden JSON that the
reference's OWN bespoke driver tools/test_gaze360_gaze.py wrote for the same videos (tests/golden/golden_slicer_reference.json)
- bit-identical at one clip per forward.  GPU: the same call sequence (tools/test.py:186-240) on the real engine with PNG files."""
import json
import os
import pickle
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

from mcgaze_b200 import evaluate as ev
from mcgaze_b200 import shims
from oracle import stub_clip_model as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get('MCGAZE_REFERENCE', '/root/reference')
CFG = os.path.join(ROOT, 'configs/multiclue_gaze/multiclue_gaze_r50_gaze360.py')
WORKER = os.path.join(ROOT, 'tests', 'generic_tester_worker.py')
needs_ref = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, 'tools', 'test.py')),
                               reason='the reference checkout only exists in the build container')


@pytest.fixture(scope='module')
def golden(golden_dir):
    return json.load(open(os.path.join(golden_dir, 'golden_slicer_reference.json')))


@pytest.fixture()
def shim_path():
    sys.path.insert(0, shims.PATH)
    yield shims.PATH
    sys.path.remove(shims.PATH)


def _flat(records):
    out = []
    for rec in records:
        for k, v in rec.items():
            if isinstance(v, list):
                for x in v:
                    out += [0.0] * 4 if x is None else (list(x) if isinstance(x, list) else [x])
            else:
                out.append(v)
    return np.asarray(out, dtype=np.float64)


def _files(tmp_path):
    json.dump(S.make_anno_with_gt(), open(tmp_path / 'test.json', 'w'))
    torch.save({'state_dict': {}, 'meta': {}}, tmp_path / 'none.pth')
    return [CFG, str(tmp_path / 'none.pth'), '--out', str(tmp_path / 'out.pkl'), '--eval', 'mae', '--eval-options',
            f"results_file={tmp_path / 'res.json'}", '--cfg-options', 'model.type=StubClipDetector',
            f"data.test.ann_file={tmp_path / 'test.json'}", 'data.test.img_prefix=frames']


def _expected_metrics(golden):
    ds = ev.Gaze360ClipDataset(S.make_anno_with_gt(), img_prefix='frames', loader=S.encode_frame)
    want = {}
    for key in ('fusion_gazes', 'face_gazes', 'eyes_gazes', 'head_gazes'):
        m = ev.evaluate(ds, golden['records'], key)
        want[key.split('_')[0]] = (m['mae_360'], m['mae_front90'], m['mae_front20'])
    return want


def _check_outputs(tmp_path, golden, stdout, atol):
    outputs = pickle.load(open(tmp_path / 'out.pkl', 'rb'))
    assert len(outputs) == golden['forwards'] and all(np.asarray(o).shape[1] == 27 for o in outputs)
    records = json.load(open(tmp_path / 'res.json'))
    assert [r['video_id'] for r in records] == [r['video_id'] for r in golden['records']]
    assert [list(r) for r in records] == [list(r) for r in golden['records']]          # same schema, same key order
    a, b = _flat(records), _flat(golden['records'])
    assert a.shape == b.shape and np.abs(a - b).max() <= atol
    printed = eval(stdout.strip().splitlines()[-1])                                    # tools/test.py:240 print(metric)
    for clue, (m360, m180, m20) in _expected_metrics(golden).items():
        assert abs(printed[f'{clue}_mae_360'] - m360) < 1e-3 and abs(printed[f'{clue}_mae_front180'] - m180) < 1e-3
        assert abs(printed[f'{clue}_mae_front20'] - m20) < 1e-3


@needs_ref
def test_reference_generic_tester_runs_unmodified_single_process(golden, tmp_path):
    env = dict(os.environ, MCG_CLIPS_PER_BATCH='1')
    for k in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK'):
        env.pop(k, None)
    r = subprocess.run([sys.executable, WORKER, REF] + _files(tmp_path), cwd=tmp_path, env=env, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    _check_outputs(tmp_path, golden, r.stdout, atol=0.0)               # one clip per forward: bit-identical


@needs_ref
def test_reference_generic_tester_runs_unmodified_two_ranks_gloo(golden, tmp_path):
    """`tools/dist_test.sh <cfg> <ckpt> 2 --eval mae` = torch.distributed.launch of tools/test.py --launcher pytorch."""
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    args = _files(tmp_path) + ['dist_params.backend=gloo', '--launcher', 'pytorch']
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE='2', MASTER_ADDR='127.0.0.1',
                   MASTER_PORT=str(port), MCG_CLIPS_PER_BATCH='3')
        procs.append(subprocess.Popen([sys.executable, WORKER, REF] + args, cwd=tmp_path, env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=600) for p in procs]
    for p, (_, err) in zip(procs, outs):
        assert p.returncode == 0, err[-3000:]
    _check_outputs(tmp_path, golden, outs[0][0], atol=1e-6)
    assert 'mae_360' not in outs[1][0]                                 # only rank 0 evaluates (tools/test.py:223)


def test_test_py_call_sequence_with_stand_ins(golden, shim_path, tmp_path):
    """The calls of tools/test.py:186-240 against the shim packages (no reference checkout needed)."""
    import mmcv
    from mmcv.cnn import fuse_conv_bn
    from mmcv.parallel import MMDataParallel, MMDistributedDataParallel
    from mmcv.runner import get_dist_info, load_checkpoint
    from mmdet.apis import multi_gpu_test, single_gpu_test
    from mmdet.datasets import Gaze360Dataset, build_dataloader, build_dataset, replace_ImageToTensor
    from mmdet.utils import setup_multi_processes, update_data_root
    cfg = mmcv.Config.fromfile(CFG)
    json.dump(S.make_anno_with_gt(), open(tmp_path / 'test.json', 'w'))
    cfg.merge_from_dict({'data.test.ann_file': str(tmp_path / 'test.json'), 'data.test.img_prefix': 'frames'})
    update_data_root(cfg)
    omp = {k: os.environ.get(k) for k in ('OMP_NUM_THREADS', 'MKL_NUM_THREADS')}
    try:
        with pytest.warns(UserWarning) if any(v is None for v in omp.values()) else _nullcontext():
            setup_multi_processes(cfg)
    finally:
        for k, v in omp.items():
            os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)
    cfg.data.test.test_mode = True
    samples_per_gpu = cfg.data.test.pop('samples_per_gpu', 1)
    assert samples_per_gpu == 1 and replace_ImageToTensor(cfg.data.test.pipeline) == cfg.data.test.pipeline.copy()
    saved = (Gaze360Dataset.frame_loader, Gaze360Dataset.pipeline_factory)
    Gaze360Dataset.frame_loader = staticmethod(S.encode_frame)
    Gaze360Dataset.pipeline_factory = staticmethod(lambda c: S.StubBatchPipeline())
    try:
        dataset = build_dataset(cfg.data.test)
        assert isinstance(dataset, Gaze360Dataset) and dataset.test_mode and dataset.decode == 'host'
        assert len(dataset) == golden['forwards'] and dataset.CLASSES == 'person_face'
        loader = build_dataloader(dataset, samples_per_gpu=samples_per_gpu, workers_per_gpu=cfg.data.workers_per_gpu,
                                  dist=False, shuffle=False)
        assert loader.dataset is dataset and len(loader) == len(dataset) and loader.clips_per_batch == 32
        first = next(iter(loader))
        assert first['img'][0].shape == (1, 3, 4, 4) and len(first['img_metas'][0]) == 1         # video 0 has one frame
        assert dataset[0]['n'] == 1 and dataset[0]['frames'][0].shape == (4, 4, 3)
        model = S.StubDetector()
        torch.save({'state_dict': {}, 'meta': {'CLASSES': ('face', 'eyes', 'head')}}, tmp_path / 'c.pth')
        ckpt = load_checkpoint(model, str(tmp_path / 'c.pth'), map_location='cpu')
        assert ckpt['meta']['CLASSES'] == ('face', 'eyes', 'head') and fuse_conv_bn(model) is model
        wrapped = MMDataParallel(model, device_ids=[0])
        assert wrapped.module is model and wrapped.CLASSES == model.CLASSES and get_dist_info() == (0, 1)
        outputs = single_gpu_test(wrapped, loader, False, None, 0.3)
        assert len(model.calls) < len(dataset)                                           # batched, not one clip per forward
        assert MMDistributedDataParallel(model, device_ids=[0], broadcast_buffers=False).module is model
        out2 = multi_gpu_test(wrapped, loader, None, False)                              # no process group: same thing
        assert all(np.array_equal(a, b) for a, b in zip(outputs, out2))
        mmcv.dump(outputs, str(tmp_path / 'o.pkl'))
        assert all(np.array_equal(a, b) for a, b in zip(mmcv.load(str(tmp_path / 'o.pkl')), outputs))
        metric = dataset.evaluate(outputs, metric=['mae'], results_file=str(tmp_path / 'r' / 'res.json'))
        assert np.abs(_flat(json.load(open(tmp_path / 'r' / 'res.json'))) - _flat(golden['records'])).max() <= 1e-6
        for clue, (m360, m180, m20) in _expected_metrics(golden).items():
            assert abs(metric[f'{clue}_mae_360'] - m360) < 1e-3 and abs(metric[f'{clue}_mae_front20'] - m20) < 1e-3
        with pytest.raises(KeyError):
            dataset.evaluate(outputs, metric='bbox')
        with pytest.raises(ValueError):
            dataset.format_results(outputs[:-1])
        with pytest.raises(NotImplementedError):
            single_gpu_test(wrapped, loader, show=True)
        with pytest.raises(NotImplementedError):
            build_dataset(dict(cfg.data.test.to_dict(), test_mode=False))
        with pytest.raises(NotImplementedError):
            build_dataloader(dataset, 1, 0, shuffle=True)
        # no ground truth in the annotation file: the JSON is written, nothing is scored
        nogt = build_dataset(dict(cfg.data.test.to_dict(), ann_file=S.make_anno()))
        assert nogt.evaluate(outputs, metric='mae', results_file=str(tmp_path / 'nogt.json')) == {}
        assert os.path.getsize(tmp_path / 'nogt.json') > 0
    finally:
        Gaze360Dataset.frame_loader, Gaze360Dataset.pipeline_factory = saved


class _nullcontext:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def test_update_data_root_and_dataset_paths(shim_path, tmp_path, monkeypatch):
    import mmcv
    from mmdet.datasets import build_dataset
    from mmdet.utils import update_data_root
    cfg = mmcv.Config.fromfile(CFG)
    monkeypatch.setenv('MMDET_DATASETS', str(tmp_path) + '/')
    update_data_root(cfg)
    assert cfg.data_root == str(tmp_path) + '/' and cfg.data.test.ann_file == str(tmp_path / 'test.json')
    assert cfg.data.test.img_prefix == str(tmp_path / 'test_rawframes') + '/'
    # data_root joins relative paths (gaze360.py:46-56); l2cs annotation files select the l2cs scorer
    os.makedirs(tmp_path / 'l2cs')
    json.dump(S.make_anno([3]), open(tmp_path / 'l2cs' / 'test.json', 'w'))
    ds = build_dataset(dict(type='Gaze360Dataset', ann_file='l2cs/test.json', img_prefix='frames/', data_root=str(tmp_path),
                            pipeline=cfg.data.test.pipeline, clip_length=7, test_mode=True, decode='host'))
    assert ds.img_prefix == str(tmp_path / 'frames') + '/' and ds.scorer == 'l2cs' and len(ds) == 1


@pytest.mark.gpu
def test_gpu_generic_tester_call_sequence_on_png_files(synthetic_sd, shim_path, tmp_path):
    """tools/test.py:186-240 on the real engine: build_dataset / build_dataloader / build_detector / load_checkpoint /
    MMDataParallel / single_gpu_test / dataset.evaluate on PNG rawframes (decoded on the device), against the batched
    driver called directly with the same crop draws."""
    cv2 = pytest.importorskip('cv2')
    import mmcv
    from mmcv.parallel import MMDataParallel
    from mmcv.runner import load_checkpoint
    from mmdet.apis import single_gpu_test
    from mmdet.datasets import build_dataloader, build_dataset
    from mmdet.models import build_detector
    from mcgaze_b200.pipeline import GpuTestPipeline
    rng = np.random.default_rng(3)
    anno = S.make_anno_with_gt([3, 9, 12])
    for v in anno['videos']:
        for f in v['file_names']:
            os.makedirs(tmp_path / 'frames' / os.path.dirname(f), exist_ok=True)
            assert cv2.imwrite(str(tmp_path / 'frames' / f), rng.integers(0, 256, (120, 100, 3), dtype=np.uint8))
    json.dump(anno, open(tmp_path / 'test.json', 'w'))
    torch.save({'state_dict': synthetic_sd, 'meta': {}}, tmp_path / 'ckpt.pth')
    cfg = mmcv.Config.fromfile(CFG)
    cfg.merge_from_dict({'data.test.ann_file': str(tmp_path / 'test.json'), 'data.test.img_prefix': str(tmp_path / 'frames'),
                         'data.test.seed': 0})
    cfg.data.test.test_mode = True
    samples_per_gpu = cfg.data.test.pop('samples_per_gpu', 1)
    dataset = build_dataset(cfg.data.test)
    assert dataset.decode == 'gpu'
    loader = build_dataloader(dataset, samples_per_gpu=samples_per_gpu, workers_per_gpu=2, dist=False, shuffle=False)
    cfg.model.train_cfg = None
    model = build_detector(cfg.model, test_cfg=cfg.get('test_cfg'))
    load_checkpoint(model, str(tmp_path / 'ckpt.pth'), map_location='cpu')
    model = MMDataParallel(model, device_ids=[0])
    outputs = single_gpu_test(model, loader, False, None, 0.3)
    assert len(outputs) == len(dataset) == 1 + 2 + 3 and dataset.host_decoded_batches == 0
    want = ev.single_gpu_test(model.module, ev.Gaze360ClipDataset(anno, img_prefix=str(tmp_path / 'frames')),
                              GpuTestPipeline(cfg.data.test.pipeline, device=0, seed=0), clips_per_batch=32, workers=2)
    for a, b in zip(outputs, want):
        assert np.array_equal(a, b)                                   # device PNG decode == cv2 decode, same draws
    metric = dataset.evaluate(outputs, metric=['mae'], results_file=str(tmp_path / 'res.json'))
    assert set(metric) == {f'{c}_mae_{k}' for c in ('fusion', 'face', 'eyes', 'head') for k in ('360', 'front180', 'front20')}
    res = json.load(open(tmp_path / 'res.json'))
    assert [len(v['fusion_gazes']) for v in res] == [3, 9, 12]


@needs_ref
@pytest.mark.parametrize('tool', ['tools/test.py', 'tools/test_gaze360_gaze.py'])
def test_shims_export_every_mmcv_mmdet_name_the_reference_tools_import(tool, shim_path):
    """Parsed from the reference scripts themselves: every `from mmcv... import x` / `from mmdet... import y` resolves."""
    import ast
    import importlib
    tree = ast.parse(open(os.path.join(REF, tool)).read())
    seen = 0
    for node in ast.walk(tree):
        if isinstance(node, ast.ImportFrom) and node.module and node.module.split('.')[0] in ('mmcv', 'mmdet'):
            mod = importlib.import_module(node.module)
            for a in node.names:
                assert hasattr(mod, a.name), f'{tool}: from {node.module} import {a.name}'
                seen += 1
        elif isinstance(node, ast.Import):
            for a in node.names:
                if a.name.split('.')[0] in ('mmcv', 'mmdet'):
                    importlib.import_module(a.name)
                    seen += 1
    assert seen >= 8


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs (gpurun --gpus 2)')
def test_gpu_two_ranks_nccl_generic_tester(synthetic_sd, tmp_path):
    """The distributed branch of tools/test.py (dist_test.sh): init_dist over NCCL, MMDistributedDataParallel(model.cuda()),
    multi_gpu_test - one process per GPU under torch.distributed.run; the gathered per-clip rows equal a one-GPU run."""
    cv2 = pytest.importorskip('cv2')
    from mcgaze_b200.apis import init_detector
    from mcgaze_b200.pipeline import GpuTestPipeline
    rng = np.random.default_rng(4)
    anno = S.make_anno_with_gt([3, 9, 12, 23, 7])
    for v in anno['videos']:
        for f in v['file_names']:
            os.makedirs(tmp_path / 'frames' / os.path.dirname(f), exist_ok=True)
            assert cv2.imwrite(str(tmp_path / 'frames' / f), rng.integers(0, 256, (120, 100, 3), dtype=np.uint8))
    json.dump(anno, open(tmp_path / 'test.json', 'w'))
    torch.save({'state_dict': synthetic_sd, 'meta': {}}, tmp_path / 'ckpt.pth')
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
                        '127.0.0.1', '--master-port', str(port), os.path.join(ROOT, 'tests', 'dist_generic_tester.py'),
                        str(tmp_path)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    outputs = pickle.load(open(tmp_path / 'dist_out.pkl', 'rb'))
    model = init_detector(CFG, str(tmp_path / 'ckpt.pth'), device='cuda:0')
    ds = ev.Gaze360ClipDataset(anno, img_prefix=str(tmp_path / 'frames'))
    pipe_cfg = [dict(t) for t in model.cfg.data.test.pipeline]
    pipe_cfg[1]['crop_type'] = 'relative'                             # the deterministic crop the ranks used
    want = ev.single_gpu_test(model, ds, GpuTestPipeline(pipe_cfg, device=0), clips_per_batch=32)
    assert len(outputs) == len(want) == len(ds)
    for a, b in zip(outputs, want):                                   # other batches, other GPU: a clip's rows do not change
        assert np.allclose(np.asarray(a), b, rtol=0, atol=1e-4)
    res = json.load(open(tmp_path / 'dist_res.json'))
    assert [len(v['fusion_gazes']) for v in res] == [3, 9, 12, 23, 7]
    assert 'fusion_mae_360' in r.stdout


@needs_ref
def test_reference_benchmark_tool_runs_unmodified(tmp_path):
    """tools/analysis_tools/benchmark.py (tracing / profiling): deep-copies the config, builds dataset /
    loader / detector, wraps it in MMDistributedDataParallel and times `model(return_loss=False, rescale=True, **data)`
    over the loader - one clip per item."""
    args = _files(tmp_path)
    cfg_opts = args[args.index('--cfg-options') + 1:]
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    env = dict(os.environ, RANK='0', LOCAL_RANK='0', WORLD_SIZE='1', MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    r = subprocess.run([sys.executable, WORKER, REF, 'tools/analysis_tools/benchmark.py', CFG, str(tmp_path / 'none.pth'),
                        '--launcher', 'pytorch', '--max-iter', '12', '--log-interval', '4', '--repeat-num', '2',
                        '--cfg-options'] + cfg_opts + ['dist_params.backend=gloo'],
                       cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    assert r.stdout.count('Overall fps') == 3 and 'Done image [12 ' in r.stdout


def test_loader_iterates_model_ready_clips(golden):
    """`for data in data_loader: model(return_loss=False, rescale=True, **data)` - the DataLoader protocol of
    mmdet/apis/test.py:26-34 and tools/analysis_tools/benchmark.py:105-112."""
    import copy
    from mcgaze_b200.compat import Config
    from mcgaze_b200.datasets import Gaze360Dataset, build_dataloader, build_dataset
    cfg = copy.deepcopy(Config.fromfile(CFG))
    assert isinstance(cfg, Config) and cfg.filename == CFG
    saved = (Gaze360Dataset.frame_loader, Gaze360Dataset.pipeline_factory)
    Gaze360Dataset.frame_loader = staticmethod(S.encode_frame)
    Gaze360Dataset.pipeline_factory = staticmethod(lambda c: S.StubBatchPipeline())
    try:
        ds = build_dataset(dict(cfg.data.test.to_dict(), ann_file=S.make_anno(), img_prefix='frames', test_mode=True))
        want = ev.single_gpu_test(S.StubModel(), ds, S.StubBatchPipeline(), clips_per_batch=1)
        # one clip per item, dataset order, no clip_length (the reference's model call takes T from the tensor)
        model = S.StubModel()
        loader = build_dataloader(ds, samples_per_gpu=1, workers_per_gpu=0, dist=True, shuffle=False)
        assert len(loader) == len(ds) == golden['forwards']
        rows = []
        for data in loader:
            assert set(data) == {'img', 'img_metas'} and len(data['img']) == 1 and len(data['img_metas'][0]) == data['img'][0].shape[0]
            (det, _), gz = model(return_loss=False, rescale=True, **data)
            det = torch.stack(det)
            g = torch.stack([gz['gaze_score'], gz['face_gaze_score'], gz['eyes_gaze_score'], gz['head_gaze_score']], 1)
            rows.append(torch.cat([det[..., :4].reshape(-1, 12), det[..., 4], g.reshape(-1, 12)], 1).numpy())
        assert all(np.array_equal(a, b) for a, b in zip(rows, want))
        # several clips per item: one length per item, clip_length set, every clip exactly once
        loader = build_dataloader(ds, samples_per_gpu=4, workers_per_gpu=0, dist=False, shuffle=False)
        seen = 0
        for data in loader:
            n = data['img'][0].shape[0]
            T = data.get('clip_length', n)
            assert n % T == 0 and n // T <= 4 and (n // T == 1 or 'clip_length' in data)
            model(return_loss=False, rescale=True, **data)
            seen += n // T
        assert seen == len(ds)
    finally:
        Gaze360Dataset.frame_loader, Gaze360Dataset.pipeline_factory = saved


@pytest.mark.gpu
def test_gpu_loader_iteration_matches_the_batched_driver(synthetic_sd, tmp_path):
    """The synchronous loop of tools/analysis_tools/benchmark.py on the real engine: iterating the loader and calling the
    model per clip gives the rows the batched driver gives (same crop: deterministic `relative` CenterCrop)."""
    cv2 = pytest.importorskip('cv2')
    from mcgaze_b200.apis import init_detector
    from mcgaze_b200.datasets import build_dataloader, build_dataset
    rng = np.random.default_rng(6)
    anno = S.make_anno([3, 9])
    for v in anno['videos']:
        for f in v['file_names']:
            os.makedirs(tmp_path / 'frames' / os.path.dirname(f), exist_ok=True)
            assert cv2.imwrite(str(tmp_path / 'frames' / f), rng.integers(0, 256, (110, 96, 3), dtype=np.uint8))
    model = init_detector(CFG, None, device='cuda:0')
    model.load_state_dict(synthetic_sd)
    pipe_cfg = [dict(t) for t in model.cfg.data.test.pipeline]
    pipe_cfg[1]['crop_type'] = 'relative'
    ds = build_dataset(dict(type='Gaze360Dataset', ann_file=anno, img_prefix=str(tmp_path / 'frames'), pipeline=pipe_cfg,
                            clip_length=7, test_mode=True))
    want = ev.single_gpu_test(model, ds, ds.make_pipeline(0), clips_per_batch=8)
    rows = []
    for data in build_dataloader(ds, samples_per_gpu=1, workers_per_gpu=0, dist=False, shuffle=False):
        assert data['img'][0].is_cuda
        (det, _), gz = model(return_loss=False, rescale=True, **data)
        det = torch.stack(list(det)).float()
        g = torch.stack([gz['gaze_score'], gz['face_gaze_score'], gz['eyes_gaze_score'], gz['head_gaze_score']], 1)
        rows.append(torch.cat([det[..., :4].reshape(-1, 12), det[..., 4], g.reshape(-1, 12).float()], 1).cpu().numpy())
    assert len(rows) == len(want) == 3
    for a, b in zip(rows, want):
        assert np.array_equal(a, b)
