"""CPU: config / registry / checkpoint / collate stand-ins and the drop-in detector surface."""
import argparse
import os

import numpy as np
import pytest
import torch

from mcgaze_b200.compat import (Config, DataContainer, DictAction, Registry, build_from_cfg, collate,
                                load_checkpoint, scatter)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = os.path.join(ROOT, 'configs', 'multiclue_gaze', 'multiclue_gaze_r50_gaze360.py')
REF_CFG = '/root/reference/configs/multiclue_gaze/multiclue_gaze_r50_gaze360.py'


def test_config_inheritance_and_delete(tmp_path):
    (tmp_path / 'base.py').write_text("a = dict(x=1, y=dict(p=1, q=2))\nlst = [1, 2, 3]\nkeep = 'k'\n")
    (tmp_path / 'child.py').write_text(
        "_base_ = './base.py'\na = dict(y=dict(_delete_=True, r=3), z=5)\nlst = [9]\n")
    c = Config.fromfile(str(tmp_path / 'child.py'))
    assert c.a.x == 1 and c.a.z == 5 and c.a.y == {'r': 3} and c.lst == [9] and c.keep == 'k'
    c.merge_from_dict({'a.y.r': 7, 'new.k': 'v'})
    assert c.a.y.r == 7 and c.new.k == 'v'
    with pytest.raises(AttributeError):
        c.a.nope


def test_our_config_builds_the_detector():
    from mcgaze_b200.apis import init_detector
    m = init_detector(CFG, None, 'cuda:0')
    assert type(m).__name__ == 'MultiClueGaze' and len(m.roi_head.bbox_head) == 4
    assert m.cfg.data.samples_per_gpu == 32 and m.cfg.data.test.pipeline[2]['img_scale'] == (224, 224)
    l2 = init_detector(CFG.replace('gaze360.py', 'l2cs.py'), None, 'cuda:0', cfg_options={'data.samples_per_gpu': 4})
    assert l2.cfg.data.samples_per_gpu == 4 and l2.cfg.data.test.pipeline[1]['img_scale'] == (448, 448)
    with pytest.raises(RuntimeError):
        m.to('cpu')                                   # no CPU path
    with pytest.raises(NotImplementedError):
        m(img=[None], img_metas=[[{}]], return_loss=True)


@pytest.mark.skipif(not os.path.exists(REF_CFG), reason='reference tree not mounted')
def test_reference_config_file_builds_unchanged():
    from mcgaze_b200.apis import init_detector
    ours, ref = Config.fromfile(CFG), Config.fromfile(REF_CFG)
    assert ref.model.type == 'MultiClueGaze' and ref.optimizer.type == 'AdamW' and 'momentum' not in ref.optimizer
    a, b = ours.model.to_dict(), ref.model.to_dict()
    a.pop('train_cfg'), b.pop('train_cfg')
    assert a == b                                      # same model keys / type strings as the reference
    assert type(init_detector(REF_CFG, None, 'cuda:0')).__name__ == 'MultiClueGaze'


def test_unsupported_architecture_fails_loudly():
    from mcgaze_b200.registry import build_detector
    from mcgaze_b200 import detector  # noqa: F401
    cfg = Config.fromfile(CFG).model.to_dict()
    cfg['backbone']['depth'] = 101
    with pytest.raises(NotImplementedError):
        build_detector(cfg)


def test_registry_and_build_from_cfg():
    R = Registry('things')
    child = Registry('child', parent=R)

    @R.register_module()
    class A:
        def __init__(self, v=1):
            self.v = v

    assert R.build(dict(type='A', v=3)).v == 3 and child.get('A') is A and 'A' in child
    with pytest.raises(KeyError):
        R.build(dict(type='B'))
    with pytest.raises(KeyError):
        R.register_module()(A)
    with pytest.raises(TypeError):
        build_from_cfg([1], R)


def test_dict_action():
    p = argparse.ArgumentParser()
    p.add_argument('--cfg-options', nargs='+', action=DictAction)
    ns = p.parse_args(['--cfg-options', 'a.b=1', 'c=[1,2]', 'd=x', 'e=True', 'f=(1.5,2)'])
    assert ns.cfg_options == {'a.b': 1, 'c': [1, 2], 'd': 'x', 'e': True, 'f': (1.5, 2)}


def test_load_checkpoint_revise_keys_nonstrict(tmp_path):
    m = torch.nn.Linear(2, 2)
    sd = {'module.weight': torch.ones(2, 2), 'module.extra': torch.zeros(1)}
    f = str(tmp_path / 'c.pth')
    torch.save({'state_dict': sd, 'meta': {'CLASSES': ('a',)}}, f)
    ck = load_checkpoint(m, f, revise_keys=[(r'^module\.', '')])
    assert ck['meta']['CLASSES'] == ('a',) and (m.weight == 1).all()
    with pytest.raises(RuntimeError):
        load_checkpoint(m, f, strict=True)


def test_collate_scatter_like_the_slicer():
    """tools/test_gaze360_gaze.py:98-101: T per-frame dicts -> one clip batch, metas list-of-list."""
    T = 3
    datas = [dict(img=DataContainer(torch.full((3, 4 + i, 6), float(i)), stack=True),
                  img_metas=DataContainer(dict(filename=f'{i}.png'), cpu_only=True)) for i in range(T)]
    b = collate(datas, samples_per_gpu=T)
    b['img_metas'], b['img'] = b['img_metas'].data, b['img'].data
    out = scatter(b, [-1])[0]
    assert out['img'][0].shape == (T, 3, 6, 6) and out['img'][0][0, :, 4:, :].abs().sum() == 0
    assert [m['filename'] for m in out['img_metas'][0]] == ['0.png', '1.png', '2.png']


def test_detector_checks_the_checkpoint_against_what_the_engine_consumes(synthetic_sd):
    """ADVICE (round 1): load_state_dict reports missing / unexpected keys and refuses wrong shapes; unknown
    architecture options are rejected instead of loading a checkpoint into the wrong network."""
    import pytest
    import torch
    from mcgaze_b200 import detector as D
    from mcgaze_b200.compat import Config
    from mcgaze_b200.compat.checkpoint import load_state_dict
    from mcgaze_b200.registry import build_detector
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = Config.fromfile(os.path.join(root, 'configs/multiclue_gaze/multiclue_gaze_r50_gaze360.py'))
    model = build_detector(cfg.model.to_dict())
    assert set(model.state_dict()) == set(D.consumed_keys())                     # before a load: the consumed keys
    rep = load_state_dict(model, synthetic_sd, strict=True)                      # the reference layout: nothing to report
    assert rep == {'missing': [], 'unexpected': []} and model.load_report == dict(missing=[], unexpected=[])
    sd = dict(synthetic_sd)
    del sd['backbone.layer2.0.conv1.weight']
    sd['backbone.layer1.0.conv2.conv_offset.weight'] = torch.zeros(18, 64, 3, 3)   # a DCN checkpoint
    model2 = build_detector(cfg.model.to_dict())
    model2.load_state_dict(sd)
    assert model2.load_report == dict(missing=['backbone.layer2.0.conv1.weight'],
                                      unexpected=['backbone.layer1.0.conv2.conv_offset.weight'])
    with pytest.raises(RuntimeError, match='state_dict mismatch'):
        build_detector(cfg.model.to_dict()).load_state_dict(sd, strict=True)
    bad = dict(synthetic_sd)
    bad['neck.fpn_convs.0.conv.weight'] = torch.zeros(256, 256, 1, 1)
    with pytest.raises(RuntimeError, match='size mismatch'):
        build_detector(cfg.model.to_dict()).load_state_dict(bad)
    for key, value in (('dcn', dict(type='DCNv2')), ('deep_stem', True), ('norm_cfg', dict(type='GN', num_groups=32))):
        m = cfg.model.to_dict()
        m['backbone'][key] = value
        with pytest.raises(NotImplementedError):
            build_detector(m)
    m = cfg.model.to_dict()
    m['neck']['norm_cfg'] = dict(type='BN')
    with pytest.raises(NotImplementedError):
        build_detector(m)


def test_runner_and_parallel_stand_ins(tmp_path, monkeypatch):
    """mcgaze_b200/compat/runner.py: what tools/test.py:9-13 imports from mmcv.runner / mmcv.parallel / mmcv (fileio)."""
    import numpy as np
    from mcgaze_b200.compat import runner as R
    assert R.get_dist_info() == (0, 1)
    with pytest.raises(NotImplementedError):
        R.init_dist('slurm')
    monkeypatch.delenv('RANK', raising=False)
    with pytest.raises(RuntimeError, match='RANK'):
        R.init_dist('pytorch', backend='gloo')
    if not torch.cuda.is_available():
        monkeypatch.setenv('RANK', '0')
        with pytest.raises(RuntimeError, match='nccl'):
            R.init_dist('pytorch', backend='nccl')
    with pytest.raises(NotImplementedError):
        R.wrap_fp16_model(object())

    class M:
        CLASSES = ('a',)

        def __init__(self):
            self.dev, self.evals = None, 0

        def to(self, d):
            self.dev = d
            return self

        def eval(self):
            self.evals += 1
            return self

        def __call__(self, x, k=1):
            return x * k

    m = M()
    assert R.fuse_conv_bn(m) is m
    w = R.MMDataParallel(m, device_ids=[1])
    assert w.module is m and m.dev == 'cuda:1' and w(3, k=2) == 6 and w.CLASSES == ('a',) and w.eval() is w and m.evals == 1
    with pytest.raises(AttributeError):
        w.nonexistent
    with pytest.raises(NotImplementedError):
        R.MMDataParallel(m, device_ids=[0, 1])
    d = R.MMDistributedDataParallel(m, device_ids=[0], broadcast_buffers=False)
    assert d.module is m and m.dev == 'cuda:0'
    R.mkdir_or_exist(str(tmp_path / 'a' / 'b'))
    R.mkdir_or_exist(str(tmp_path / 'a' / 'b'))
    obj = dict(x=[np.float32(1.5), np.arange(3)], y='z')
    R.dump(obj, str(tmp_path / 'a' / 'o.json'))
    assert R.load(str(tmp_path / 'a' / 'o.json')) == dict(x=[1.5, [0, 1, 2]], y='z')
    R.dump([np.arange(4, dtype=np.float32)], str(tmp_path / 'a' / 'o.pkl'))
    assert np.array_equal(R.load(str(tmp_path / 'a' / 'o.pkl'))[0], np.arange(4, dtype=np.float32))
    assert R.load(str(tmp_path / 'a' / 'o.json')) and isinstance(R.dump(obj, file_format='json'), str)
    with pytest.raises(TypeError):
        R.dump(obj, str(tmp_path / 'o.yaml'))


def test_config_copies_and_pickles_like_mmcv_config():
    import copy
    import pickle
    cfg = Config.fromfile(os.path.join(ROOT, 'configs/multiclue_gaze/multiclue_gaze_r50_gaze360.py'))
    c2 = copy.deepcopy(cfg)
    c2.data.test.test_mode = True
    assert isinstance(c2, Config) and c2.filename == cfg.filename and 'test_mode' not in cfg.data.test
    c3 = pickle.loads(pickle.dumps(cfg))
    assert isinstance(c3, Config) and c3.to_dict() == cfg.to_dict()
    assert isinstance(copy.copy(cfg), Config)
    with pytest.raises(AttributeError):
        cfg.no_such_key
