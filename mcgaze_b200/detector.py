"""Drop-in `MultiClueGaze` for the reference's config / registry surface, backed by the CUDA
engine (include/mcgaze_b200.h).

The classes registered here mirror the `type=` names of configs/multiclue_gaze/*.py
(reference: mmdet/models/detectors/multiclue_gaze.py:8, backbones/resnet.py:305, necks/fpn.py:10,
dense_heads/fixed_embedding_rpn_head.py:10, roi_heads/multiclue_gaze_roi_head.py:9,
roi_extractors/single_level_roi_extractor.py, bbox_heads/gaze_stqi_head.py:17,
mask_heads/gaze_head.py:16, utils/transformer.py:1054).  The sub-module classes only VALIDATE
their config against what the sm_100a kernels implement (the architecture is compiled in) and
describe the parameter layout; all arithmetic happens in libmcgaze_b200.so.  There is no CPU path:
constructing the engine without a CUDA device raises.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

from .registry import (BACKBONES, DETECTORS, HEADS, LOSSES, NECKS, ROI_EXTRACTORS, TRANSFORMER, build_backbone,
                       build_head, build_neck, build_roi_extractor, build_transformer)


class _Spec:
    """Config-carrying node; `expect` pins the hyper-parameters the kernels hard-wire, `only` lists for options that
    change the architecture the one value (or values) the compiled kernels implement - anything else raises instead of
    loading a checkpoint into the wrong network."""
    expect: Dict[str, Any] = {}
    only: Dict[str, Any] = {}

    def __init__(self, **cfg):
        cfg.pop('init_cfg', None)
        cfg.pop('train_cfg', None)
        cfg.pop('test_cfg', None)
        for k, allowed in self.only.items():
            if k in cfg and cfg[k] not in allowed:
                raise NotImplementedError(f'{type(self).__name__}: {k}={cfg[k]!r} is not supported by the sm_100a kernels '
                                          f'(supported: {list(allowed)!r})')
        for k, want in self.expect.items():
            if k in cfg:
                got = cfg[k]
                got = list(got) if isinstance(got, (tuple, list)) else got
                w = list(want) if isinstance(want, (tuple, list)) else want
                if got != w:
                    raise NotImplementedError(
                        f'{type(self).__name__}: {k}={cfg[k]!r} is not supported by the sm_100a kernels '
                        f'(compiled for {k}={want!r})')
        self.cfg = cfg


@BACKBONES.register_module()
class ResNet(_Spec):
    expect = dict(depth=50, num_stages=4, out_indices=(0, 1, 2, 3), style='pytorch')
    # resnet.py:370-410: options that would change the computation (deformable convs, deep stem, avg-pool downsample,
    # plugins, other norms) must not pass silently; frozen_stages / norm_eval / with_cp only matter for training
    only = dict(dcn=(None,), stage_with_dcn=((False, False, False, False), [False, False, False, False]), deep_stem=(False,),
                avg_down=(False,), plugins=(None,), conv_cfg=(None,), in_channels=(3,), base_channels=(64,),
                strides=((1, 2, 2, 2), [1, 2, 2, 2]), dilations=((1, 1, 1, 1), [1, 1, 1, 1]), stem_channels=(None, 64))

    def __init__(self, **cfg):
        super().__init__(**cfg)
        norm = cfg.get('norm_cfg') or dict(type='BN')
        if norm.get('type') not in ('BN', 'SyncBN'):
            raise NotImplementedError(f'ResNet: norm_cfg={norm!r} is not supported (eval-mode BatchNorm is folded into the convs)')


@NECKS.register_module()
class FPN(_Spec):
    expect = dict(in_channels=[256, 512, 1024, 2048], out_channels=256, start_level=0, num_outs=4)
    # fpn.py:62-129: no extra levels, no norm / activation inside the ConvModules, nearest 2x top-down
    only = dict(add_extra_convs=(False, 'on_input'), relu_before_extra_convs=(False,), no_norm_on_lateral=(False,),
                conv_cfg=(None,), norm_cfg=(None,), act_cfg=(None,), end_level=(-1,),
                upsample_cfg=(dict(mode='nearest'),))


@HEADS.register_module()
class FixedEmbeddingRPNHead(_Spec):
    expect = dict(proposal_feature_channel=256, num_proposals=3)


@ROI_EXTRACTORS.register_module()
class SingleRoIExtractor(_Spec):
    expect = dict(out_channels=256, featmap_strides=[4, 8, 16, 32], finest_scale=56)

    def __init__(self, **cfg):
        super().__init__(**cfg)
        rl = cfg.get('roi_layer', {})
        if rl.get('type', 'RoIAlign') != 'RoIAlign' or rl.get('output_size', 7) != 7 or rl.get('sampling_ratio', 2) != 2:
            raise NotImplementedError(f'roi_layer {rl!r} not supported (RoIAlign 7x7, sampling_ratio=2 only)')


@TRANSFORMER.register_module()
class DynamicConv(_Spec):
    expect = dict(in_channels=256, feat_channels=64, out_channels=256, input_feat_shape=7, with_proj=True)


@LOSSES.register_module(name=['FocalLoss', 'L1Loss', 'GIoULoss', 'GazeArccosLoss', 'GazeTempLoss', 'CrossEntropyLoss'])
class _InferenceOnlyLoss(_Spec):
    """Losses are training-only (out of scope); kept so that reference configs build."""


@HEADS.register_module()
class GazeSTQIHead(_Spec):
    expect = dict(num_classes=3, num_ffn_fcs=2, num_heads=8, num_cls_fcs=1, num_reg_fcs=3, feedforward_channels=2048,
                  in_channels=256, dropout=0.0)

    def __init__(self, **cfg):
        super().__init__(**cfg)
        self.dynamic_conv = build_transformer(cfg.get('dynamic_conv_cfg', dict(type='DynamicConv')))
        lc = cfg.get('loss_cls', dict(use_sigmoid=True))
        if not lc.get('use_sigmoid', False):
            raise NotImplementedError('GazeSTQIHead: only sigmoid classification is implemented')
        bc = cfg.get('bbox_coder', {})
        if bc and (list(bc.get('target_stds', [0.5, 0.5, 1., 1.])) != [0.5, 0.5, 1., 1.] or bc.get('clip_border', False)
                   or list(bc.get('target_means', [0., 0., 0., 0.])) != [0., 0., 0., 0.]):
            raise NotImplementedError(f'bbox_coder {bc!r} not supported')


@HEADS.register_module()
class GazeHead(_Spec):
    expect = dict(in_channels=256, gaze_dim=3)


@HEADS.register_module()
class MultiClueGazeROIHead(_Spec):
    expect = dict(num_stages=4, proposal_feature_channel=256)

    def __init__(self, bbox_roi_extractor=None, bbox_head=None, gaze_head=None, **cfg):
        super().__init__(**cfg)
        n = cfg.get('num_stages', 4)
        self.bbox_roi_extractor = build_roi_extractor(bbox_roi_extractor)
        bbox_head = bbox_head if isinstance(bbox_head, (list, tuple)) else [bbox_head] * n
        gaze_head = gaze_head if isinstance(gaze_head, (list, tuple)) else [gaze_head] * n
        assert len(bbox_head) == n and len(gaze_head) == n
        self.bbox_head = [build_head(h) for h in bbox_head]
        self.gaze_head = [build_head(h) for h in gaze_head]


def consumed_keys() -> 'OrderedDict[str, tuple]':
    """Checkpoint keys (reference layout) and shapes the engine reads: the `need()` calls of
    Engine::load_weights (mcgaze_b200/csrc/mcg_api.cu)."""
    k: 'OrderedDict[str, tuple]' = OrderedDict()

    def bn(prefix, c):
        for n in ('weight', 'bias', 'running_mean', 'running_var'):
            k[f'{prefix}.{n}'] = (c,)

    def ln(prefix, c):
        k[f'{prefix}.weight'] = (c,)
        k[f'{prefix}.bias'] = (c,)

    k['backbone.conv1.weight'] = (64, 3, 7, 7)
    bn('backbone.bn1', 64)
    cin = 64
    for li, (planes, blocks) in enumerate(zip((64, 128, 256, 512), (3, 4, 6, 3))):
        for b in range(blocks):
            p = f'backbone.layer{li + 1}.{b}'
            k[f'{p}.conv1.weight'] = (planes, cin, 1, 1)
            bn(f'{p}.bn1', planes)
            k[f'{p}.conv2.weight'] = (planes, planes, 3, 3)
            bn(f'{p}.bn2', planes)
            k[f'{p}.conv3.weight'] = (planes * 4, planes, 1, 1)
            bn(f'{p}.bn3', planes * 4)
            if b == 0:
                k[f'{p}.downsample.0.weight'] = (planes * 4, cin, 1, 1)
                bn(f'{p}.downsample.1', planes * 4)
            cin = planes * 4
    for i, c in enumerate((256, 512, 1024, 2048)):
        k[f'neck.lateral_convs.{i}.conv.weight'] = (256, c, 1, 1)
        k[f'neck.lateral_convs.{i}.conv.bias'] = (256,)
        k[f'neck.fpn_convs.{i}.conv.weight'] = (256, 256, 3, 3)
        k[f'neck.fpn_convs.{i}.conv.bias'] = (256,)
    k['rpn_head.init_proposal_bboxes.weight'] = (3, 4)
    k['rpn_head.init_proposal_features.weight'] = (3, 256)
    for s in range(4):
        p = f'roi_head.bbox_head.{s}'
        k[f'{p}.attention.attn.in_proj_weight'] = (768, 256)
        k[f'{p}.attention.attn.in_proj_bias'] = (768,)
        k[f'{p}.attention.attn.out_proj.weight'] = (256, 256)
        k[f'{p}.attention.attn.out_proj.bias'] = (256,)
        ln(f'{p}.attention_norm', 256)
        q = f'{p}.instance_interactive_conv'
        k[f'{q}.dynamic_layer.weight'] = (32768, 256)
        k[f'{q}.dynamic_layer.bias'] = (32768,)
        ln(f'{q}.norm_in', 64)
        ln(f'{q}.norm_out', 256)
        k[f'{q}.fc_layer.weight'] = (256, 12544)
        k[f'{q}.fc_layer.bias'] = (256,)
        ln(f'{q}.fc_norm', 256)
        ln(f'{p}.instance_interactive_conv_norm', 256)
        k[f'{p}.ffn.layers.0.0.weight'] = (2048, 256)
        k[f'{p}.ffn.layers.0.0.bias'] = (2048,)
        k[f'{p}.ffn.layers.1.weight'] = (256, 2048)
        k[f'{p}.ffn.layers.1.bias'] = (256,)
        ln(f'{p}.ffn_norm', 256)
        k[f'{p}.cls_fcs.0.weight'] = (256, 256)
        ln(f'{p}.cls_fcs.1', 256)
        for j in range(3):
            k[f'{p}.reg_fcs.{3 * j}.weight'] = (256, 256)
            ln(f'{p}.reg_fcs.{3 * j + 1}', 256)
        for clue in ('face', 'eyes', 'head'):
            k[f'{p}.{clue}_fc_cls.weight'] = (1, 256)
            k[f'{p}.{clue}_fc_cls.bias'] = (1,)
            k[f'{p}.{clue}_fc_reg.weight'] = (4, 256)
            k[f'{p}.{clue}_fc_reg.bias'] = (4,)
    h = 'roi_head.gaze_head.3'          # only the last stage's gaze head runs at test time
    for clue in ('face', 'eyes', 'head'):
        for tower in (f'{h}.gaze_{clue}_fcs', f'{h}.gaze_{clue}_confidence'):
            for j in range(2):
                k[f'{tower}.{3 * j}.weight'] = (256, 256)
                ln(f'{tower}.{3 * j + 1}', 256)
        for fc in (f'{h}.fc_{clue}', f'{h}.fc_{clue}_confidence'):
            k[f'{fc}.weight'] = (3, 256)
            k[f'{fc}.bias'] = (3,)
    k[f'{h}.fc_gaze.weight'] = (3, 9)
    k[f'{h}.fc_gaze.bias'] = (3,)
    return k


def tolerated_key(key: str) -> bool:
    """Checkpoint keys of the reference model that the inference engine does not read: BN batch counters, the unused
    fc_cls / fc_reg every BBoxHead creates (bbox_head.py:66-81), the gaze heads of stages 0-2
    (multiclue_gaze_roi_head.py:367-378 uses the last one only)."""
    import re
    return bool(key.endswith('num_batches_tracked')
                or re.match(r'roi_head\.bbox_head\.\d\.fc_(cls|reg)\.(weight|bias)$', key)
                or re.match(r'roi_head\.gaze_head\.[0-2]\.', key))


@DETECTORS.register_module()
class MultiClueGaze:
    """reference: mmdet/models/detectors/multiclue_gaze.py:8-131 (+ base.py:112-174 dispatch)."""

    CLASSES = ('face', 'eyes', 'head')

    def __init__(self, backbone, rpn_head, roi_head, train_cfg=None, test_cfg=None, neck=None, pretrained=None,
                 init_cfg=None, precision: str = 'fp16c8'):
        self.backbone = build_backbone(backbone)
        self.neck = build_neck(neck) if neck is not None else None
        if self.neck is None:
            raise NotImplementedError('MultiClueGaze without an FPN neck is not supported')
        self.rpn_head = build_head(rpn_head)
        self.roi_head = build_head(roi_head)
        self.test_cfg = test_cfg
        self.precision = precision
        self.device_index = 0
        self.cfg = None
        self._sd: 'OrderedDict[str, Any]' = OrderedDict()
        self._engine = None
        self._ranges_checked = False
        self.training = False

    # --- torch.nn.Module-like surface used by init_detector / load_checkpoint ---------------
    def state_dict(self):
        """The loaded checkpoint tensors; before a load, the keys the engine consumes (values None) so that
        `load_checkpoint` can report missing / unexpected keys like it does for a torch module."""
        return self._sd if self._sd else OrderedDict((k, None) for k in consumed_keys())

    def load_state_dict(self, state_dict, strict: bool = False):
        """Checks the checkpoint against what the engine consumes: a consumed key with the wrong shape always raises
        (torch does too), missing consumed keys raise when `strict` (the engine raises on its first forward otherwise),
        keys the engine ignores (the BBoxHead's unused fc_cls / fc_reg, gaze heads of stages 0-2, BN counters) are
        tolerated, anything else is reported as unexpected."""
        want = consumed_keys()
        bad = [f'{k}: {tuple(state_dict[k].shape)} != {want[k]}' for k in want
               if k in state_dict and hasattr(state_dict[k], 'shape') and tuple(state_dict[k].shape) != want[k]]
        if bad:
            raise RuntimeError('size mismatch for ' + '; '.join(bad[:8]))
        missing = [k for k in want if k not in state_dict]
        unexpected = [k for k in state_dict if k not in want and not tolerated_key(k)]
        if strict and (missing or unexpected):
            raise RuntimeError(f'state_dict mismatch: missing {missing[:8]}, unexpected {unexpected[:8]}')
        self.load_report = dict(missing=missing, unexpected=unexpected)
        self._sd = OrderedDict(state_dict)
        self._engine = None
        self._ranges_checked = False
        return self

    def to(self, device):
        s = str(device)
        if s.startswith('cpu'):
            raise RuntimeError('mcgaze_b200 has no CPU path; use a cuda device')
        index = int(s.split(':')[1]) if ':' in s else 0
        if index != self.device_index:
            self._engine = None                   # packed weights and workspace live on one device
        self.device_index = index
        return self

    def cuda(self, device=None):
        """torch.nn.Module.cuda(): the current device unless one is named (tools/test.py:214 calls `model.cuda()` after
        init_dist has selected the rank's GPU)."""
        if device is None:
            import torch
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        return self.to(device if isinstance(device, str) and device.startswith('cuda') else f'cuda:{int(device)}')

    def eval(self):
        self.training = False
        return self

    def train(self, mode: bool = True):
        if mode:
            raise NotImplementedError('training is out of scope for the B200 inference backend')
        return self

    @property
    def engine(self):
        if self._engine is None:
            if not self._sd:
                raise RuntimeError('MultiClueGaze: no weights loaded (call load_state_dict / init_detector)')
            from .lib import Engine
            self._engine = Engine(self._sd, self.device_index, self.precision)
        return self._engine

    # --- forward ------------------------------------------------------------------------------
    def __call__(self, *args, **kwargs):
        return self.forward(*args, **kwargs)

    def forward(self, img, img_metas, return_loss: bool = True, **kwargs):
        if return_loss:
            raise NotImplementedError('forward_train is out of scope for the B200 inference backend')
        return self.forward_test(img, img_metas, **kwargs)

    def forward_test(self, imgs, img_metas, **kwargs):
        """base.py:112-154: list-of-augmentations wrapper; exactly one augmentation is supported."""
        for var, name in [(imgs, 'imgs'), (img_metas, 'img_metas')]:
            if not isinstance(var, list):
                raise TypeError(f'{name} must be a list, but got {type(var)}')
        if len(imgs) != len(img_metas):
            raise ValueError(f'num of augmentations ({len(imgs)}) != num of image meta ({len(img_metas)})')
        if len(imgs) != 1:
            raise NotImplementedError('test-time augmentation is not supported (the reference raises too)')
        for img, metas in zip(imgs, img_metas):
            for m in metas:
                m['batch_input_shape'] = tuple(img.shape[-2:])
        return self.simple_test(imgs[0], img_metas[0], **kwargs)

    def simple_test(self, img, img_metas: Sequence[Dict[str, Any]], rescale: bool = False, format: bool = False,
                    clip_length: Optional[int] = None):
        """multiclue_gaze.py:105-131 -> multiclue_gaze_roi_head.py:287-384.  `img` [T,3,H,W] CUDA
        fp32 is ONE clip unless `clip_length` says the batch holds several clips of that length."""
        import torch
        T = img.shape[0]
        img_hw = np.array([[m['img_shape'][0], m['img_shape'][1]] for m in img_metas], dtype=np.float32)
        scale = None
        if rescale:
            scale = np.stack([np.asarray(m['scale_factor'], dtype=np.float32).reshape(4) for m in img_metas])
        out = self.engine.forward(img, clip_length=clip_length or T, img_hw=img_hw, scale_factor=scale)
        if not self._ranges_checked:
            # first forward of a checkpoint: its activations / BN-folded weights must sit inside the operand window of
            # the fp16c8 corrections (lib.Engine.check_ranges raises otherwise); one scan + sync, then never again
            self._ranges_checked = True
            self.engine.check_ranges()
        gaze, boxes, scores = out['gaze'], out['boxes'], out['scores']
        det_bboxes = list(torch.cat([boxes, scores[..., None]], dim=2).unbind(0))     # T x [3, 5], one kernel
        det_labels = [[0, 1, 2] for _ in range(T)]
        if format:
            bbox_results = [[det_bboxes[i][c:c + 1].cpu().numpy() for c in range(3)] for i in range(T)]
        else:
            bbox_results = (det_bboxes, det_labels)
        gaze_results = {'gaze_score': gaze[:, 0], 'face_gaze_score': gaze[:, 1], 'eyes_gaze_score': gaze[:, 2],
                        'head_gaze_score': gaze[:, 3]}
        return bbox_results, gaze_results
