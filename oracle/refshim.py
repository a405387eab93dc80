"""TEST INFRASTRUCTURE ONLY — lets the reference's OWN python code (/root/reference/mmdet/...)
be imported and executed in this container although mmcv-full 1.4.8 cannot be installed.

`install()` puts a stand-in `mmcv` package (plus empty stubs for pycocotools / terminaltables /
lvis / cityscapesscripts) into `sys.modules` and prepends /root/reference to `sys.path`.

The stand-in is two-tier:
  * the mmcv symbols the MCGaze hot path really executes are restated on top of torch /
    torchvision from mmcv's published semantics: Config, Registry,
    BaseModule/ModuleList/Sequential, ConvModule, build_{norm,activation,conv}_layer,
    MultiheadAttention, FFN, ops.RoIAlign, force_fp32/auto_fp16 (no-ops), bias_init_with_prob;
  * every other attribute resolves to an inert placeholder that can be subclassed, called or
    used as a decorator, which is all the ~200 unrelated mmdet modules need at import time.
Only oracle/gen_golden.py and tests use this.  It is NOT a product component and not on any
measured path.  Because MultiheadAttention / FFN / RoIAlign / ConvModule are restated (their
source is not under /root/reference), parity at that boundary is "unpinned"; everything above
it (the reference's detector, roi head, STQI head, DynamicConv, GazeHead, box coder, FPN,
ResNet python code) runs unmodified.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import math
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

REFERENCE_ROOT = '/root/reference'


# ------------------------------------------------------------------------------- placeholders
class _PlaceholderMeta(type):
    def __getattr__(cls, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return _make_placeholder(f'{cls.__name__}.{name}')

    def __call__(cls, *args, **kwargs):
        if cls.__dict__.get('_is_placeholder', False):
            if len(args) == 1 and not kwargs and (isinstance(args[0], type) or callable(args[0])) \
                    and not isinstance(args[0], (str, int, float, dict, list, tuple)):
                return args[0]                      # bare decorator use
            if not args and kwargs:
                return lambda f=None, *a, **k: f    # decorator factory use (e.g. mmcv.jit(coderize=True))
        return super().__call__(*args, **kwargs)

    def __iter__(cls):
        return iter(())


def _make_placeholder(name: str):
    return _PlaceholderMeta(name.split('.')[-1], (), {
        '_is_placeholder': True,
        '__init__': lambda self, *a, **k: None,
        '__call__': lambda self, *a, **k: (a[0] if len(a) == 1 and callable(a[0]) else self),
        '__getattr__': lambda self, n: _make_placeholder(n)(),
        'register_module': staticmethod(lambda *a, **k: (k['module'] if k.get('module') is not None
                                                          else (lambda c: c))),
        'build': staticmethod(lambda *a, **k: None),
        'get': staticmethod(lambda *a, **k: None),
    })


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        ph = _make_placeholder(name)
        setattr(self, name, ph)
        return ph


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    PREFIXES = ('mmcv', 'pycocotools', 'terminaltables', 'lvis', 'cityscapesscripts', 'panopticapi', 'matplotlib',
                'seaborn', 'imagecorruptions', 'albumentations', 'onnx', 'onnxruntime', 'sklearn_never')

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split('.')[0] in self.PREFIXES:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


# ------------------------------------------------------------------------------- real pieces
class BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self.init_cfg = init_cfg
        self._is_init = False

    def init_weights(self):
        for m in self.children():
            if hasattr(m, 'init_weights'):
                m.init_weights()


class ModuleList(BaseModule, nn.ModuleList):
    def __init__(self, modules=None, init_cfg=None):
        BaseModule.__init__(self, init_cfg)
        nn.ModuleList.__init__(self, modules)


class Sequential(BaseModule, nn.Sequential):
    def __init__(self, *args, init_cfg=None):
        BaseModule.__init__(self, init_cfg)
        nn.Sequential.__init__(self, *args)


def _noop_decorator(*dargs, **dkwargs):
    if len(dargs) == 1 and callable(dargs[0]) and not dkwargs:
        return dargs[0]
    return lambda f: f


def build_norm_layer(cfg, num_features, postfix=''):
    t = cfg['type']
    if t in ('BN', 'BN2d'):
        layer, abbr = nn.BatchNorm2d(num_features, eps=cfg.get('eps', 1e-5)), 'bn'
    elif t == 'LN':
        layer, abbr = nn.LayerNorm(num_features, eps=cfg.get('eps', 1e-5)), 'ln'
    elif t == 'GN':
        layer, abbr = nn.GroupNorm(cfg['num_groups'], num_features), 'gn'
    else:
        raise KeyError(t)
    for p in layer.parameters():
        p.requires_grad = cfg.get('requires_grad', True)
    return abbr + str(postfix), layer


def build_activation_layer(cfg):
    cfg = dict(cfg)
    t = cfg.pop('type')
    return {'ReLU': nn.ReLU, 'GELU': nn.GELU, 'Sigmoid': nn.Sigmoid}[t](**cfg)


def build_conv_layer(cfg, *args, **kwargs):
    assert cfg is None or cfg.get('type', 'Conv2d') in ('Conv2d', 'Conv'), cfg
    return nn.Conv2d(*args, **kwargs)


class ConvModule(nn.Module):
    """conv -> norm -> act; bias='auto' == (norm_cfg is None); params under .conv / .bn"""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias='auto', conv_cfg=None, norm_cfg=None, act_cfg=dict(type='ReLU'), inplace=True, **kw):
        super().__init__()
        if bias == 'auto':
            bias = norm_cfg is None
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias)
        self.with_norm = norm_cfg is not None
        if self.with_norm:
            self.norm_name, norm = build_norm_layer(norm_cfg, out_channels)
            self.add_module(self.norm_name, norm)
        self.activate = build_activation_layer(dict(act_cfg, inplace=inplace)) if act_cfg else None

    def forward(self, x):
        x = self.conv(x)
        if self.with_norm:
            x = getattr(self, self.norm_name)(x)
        if self.activate is not None:
            x = self.activate(x)
        return x


class MultiheadAttention(BaseModule):
    """mmcv.cnn.bricks.transformer.MultiheadAttention(embed_dims, num_heads, attn_drop) with
    batch_first=False: forward(query) -> identity + nn.MultiheadAttention(q, k=q, v=q)[0]."""

    def __init__(self, embed_dims, num_heads, attn_drop=0., proj_drop=0., dropout_layer=None, init_cfg=None,
                 batch_first=False, **kwargs):
        super().__init__(init_cfg)
        assert not batch_first
        self.embed_dims, self.num_heads = embed_dims, num_heads
        self.attn = nn.MultiheadAttention(embed_dims, num_heads, attn_drop)

    def forward(self, query, key=None, value=None, identity=None, query_pos=None, key_pos=None, attn_mask=None,
                key_padding_mask=None, **kwargs):
        key = query if key is None else key
        value = key if value is None else value
        identity = query if identity is None else identity
        out = self.attn(query=query, key=key, value=value, attn_mask=attn_mask,
                        key_padding_mask=key_padding_mask)[0]
        return identity + out


class FFN(BaseModule):
    """mmcv FFN(embed_dims, feedforward_channels, num_fcs, act_cfg, ffn_drop/dropout, add_identity=True):
    layers = Sequential(Sequential(Linear, act, Dropout), Linear, Dropout); x + layers(x)."""

    def __init__(self, embed_dims=256, feedforward_channels=1024, num_fcs=2, act_cfg=dict(type='ReLU', inplace=True),
                 ffn_drop=0., dropout_layer=None, add_identity=True, init_cfg=None, **kwargs):
        super().__init__(init_cfg)
        ffn_drop = kwargs.get('dropout', ffn_drop)
        layers, cin = [], embed_dims
        for _ in range(num_fcs - 1):
            layers.append(Sequential(nn.Linear(cin, feedforward_channels), build_activation_layer(act_cfg),
                                     nn.Dropout(ffn_drop)))
            cin = feedforward_channels
        layers += [nn.Linear(feedforward_channels, embed_dims), nn.Dropout(ffn_drop)]
        self.layers = Sequential(*layers)
        self.add_identity = add_identity

    def forward(self, x, identity=None):
        out = self.layers(x)
        if not self.add_identity:
            return out
        return (x if identity is None else identity) + out


class RoIAlign(nn.Module):
    """mmcv.ops.RoIAlign(output_size, spatial_scale, sampling_ratio, pool_mode='avg', aligned=True).
    Uses torchvision.ops.roi_align, which mmcv itself offers as the equivalent path
    (`use_torchvision=True`)."""

    def __init__(self, output_size, spatial_scale=1.0, sampling_ratio=0, pool_mode='avg', aligned=True,
                 use_torchvision=False):
        super().__init__()
        self.output_size = (output_size, output_size) if isinstance(output_size, int) else tuple(output_size)
        self.spatial_scale, self.sampling_ratio, self.aligned = float(spatial_scale), int(sampling_ratio), aligned
        assert pool_mode == 'avg'

    def forward(self, input, rois):
        from torchvision.ops import roi_align
        return roi_align(input, rois, self.output_size, self.spatial_scale, self.sampling_ratio, self.aligned)


def bias_init_with_prob(prior_prob):
    return float(-math.log((1 - prior_prob) / prior_prob))


_installed = False


def install() -> None:
    global _installed
    if _installed:
        return
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    from mcgaze_b200.compat import Config, ConfigDict, DictAction, Registry, build_from_cfg, load_checkpoint

    sys.meta_path.insert(0, _StubFinder())
    import mmcv  # noqa: E402  (resolved by the stub finder)
    import mmcv.cnn
    import mmcv.cnn.bricks
    import mmcv.cnn.bricks.transformer as tr
    import mmcv.ops
    import mmcv.parallel
    import mmcv.runner
    import mmcv.utils

    mmcv.__version__ = '1.4.8'
    mmcv.Config, mmcv.ConfigDict, mmcv.DictAction = Config, ConfigDict, DictAction
    mmcv.jit = _noop_decorator
    mmcv.is_tuple_of = lambda seq, t: isinstance(seq, tuple) and all(isinstance(x, t) for x in seq)
    mmcv.is_list_of = lambda seq, t: isinstance(seq, list) and all(isinstance(x, t) for x in seq)
    mmcv.is_str = lambda x: isinstance(x, str)
    mmcv.utils.Registry, mmcv.utils.build_from_cfg = Registry, build_from_cfg
    mmcv.utils.ConfigDict, mmcv.utils.Config = ConfigDict, Config
    mmcv.utils.TORCH_VERSION = torch.__version__
    mmcv.utils.digit_version = lambda v: tuple(int(x) for x in v.split('+')[0].split('.')[:3] if x.isdigit())
    mmcv.utils.print_log = lambda *a, **k: None
    mmcv.utils.to_2tuple = lambda x: x if isinstance(x, tuple) else (x, x)
    mmcv.cnn.MODELS = Registry('model')
    mmcv.cnn.ConvModule = ConvModule
    mmcv.cnn.build_norm_layer = build_norm_layer
    mmcv.cnn.build_activation_layer = build_activation_layer
    mmcv.cnn.build_conv_layer = build_conv_layer
    mmcv.cnn.bias_init_with_prob = bias_init_with_prob
    for fn in ('constant_init', 'kaiming_init', 'normal_init', 'xavier_init', 'caffe2_xavier_init',
               'uniform_init', 'trunc_normal_init'):
        setattr(mmcv.cnn, fn, lambda *a, **k: None)
    mmcv.cnn.Linear, mmcv.cnn.Conv2d = nn.Linear, nn.Conv2d
    tr.MultiheadAttention, tr.FFN = MultiheadAttention, FFN
    mmcv.runner.BaseModule, mmcv.runner.ModuleList, mmcv.runner.Sequential = BaseModule, ModuleList, Sequential
    import mmcv.runner.base_module as bm
    bm.BaseModule, bm.ModuleList, bm.Sequential = BaseModule, ModuleList, Sequential
    mmcv.runner.auto_fp16 = mmcv.runner.force_fp32 = _noop_decorator
    mmcv.runner.load_checkpoint = load_checkpoint
    mmcv.runner.get_dist_info = lambda: (0, 1)
    mmcv.ops.RoIAlign = RoIAlign
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


# ----------------------------------------------------------------------------------------------------
# mmcv.image functions the reference's test pipeline calls (mmdet/datasets/pipelines/transforms.py:217-228,
# 671-679, 751-752), restated from the published mmcv 1.4.8 source (mmcv/image/geometric.py, photometric.py):
# thin wrappers over OpenCV, which IS available here (cv2 4.13) — so the pixel arithmetic of the goldens is
# OpenCV's own.  Used by oracle/gen_golden_preprocess.py only.
# ----------------------------------------------------------------------------------------------------
def _scale_size(size, scale):
    if isinstance(scale, (float, int)):
        scale = (scale, scale)
    w, h = size
    return int(w * float(scale[0]) + 0.5), int(h * float(scale[1]) + 0.5)


def rescale_size(old_size, scale, return_scale=False):
    w, h = old_size
    if isinstance(scale, (float, int)):
        scale_factor = scale
    else:
        max_long_edge, max_short_edge = max(scale), min(scale)
        scale_factor = min(max_long_edge / max(h, w), max_short_edge / min(h, w))
    new_size = _scale_size((w, h), scale_factor)
    return (new_size, scale_factor) if return_scale else new_size


def imresize(img, size, return_scale=False, interpolation='bilinear', out=None, backend=None):
    import cv2
    assert interpolation == 'bilinear' and backend in (None, 'cv2')
    h, w = img.shape[:2]
    resized = cv2.resize(img, size, dst=out, interpolation=cv2.INTER_LINEAR)
    if not return_scale:
        return resized
    return resized, size[0] / w, size[1] / h


def imrescale(img, scale, return_scale=False, interpolation='bilinear', backend=None):
    h, w = img.shape[:2]
    new_size, scale_factor = rescale_size((w, h), scale, return_scale=True)
    rescaled = imresize(img, new_size, interpolation=interpolation, backend=backend)
    return (rescaled, scale_factor) if return_scale else rescaled


def imnormalize(img, mean, std, to_rgb=True):
    import cv2
    img = img.copy().astype(np.float32)
    assert img.dtype != np.uint8
    mean = np.float64(mean.reshape(1, -1))
    stdinv = 1 / np.float64(std.reshape(1, -1))
    if to_rgb:
        cv2.cvtColor(img, cv2.COLOR_BGR2RGB, img)
    cv2.subtract(img, mean, img)
    cv2.multiply(img, stdinv, img)
    return img


def impad(img, *, shape=None, padding=None, pad_val=0, padding_mode='constant'):
    import cv2
    assert (shape is not None) ^ (padding is not None) and padding_mode == 'constant'
    if shape is not None:
        padding = (0, 0, shape[1] - img.shape[1], shape[0] - img.shape[0])
    return cv2.copyMakeBorder(img, padding[1], padding[3], padding[0], padding[2], cv2.BORDER_CONSTANT, value=pad_val)


def impad_to_multiple(img, divisor, pad_val=0):
    pad_h = int(np.ceil(img.shape[0] / divisor)) * divisor
    pad_w = int(np.ceil(img.shape[1] / divisor)) * divisor
    return impad(img, shape=(pad_h, pad_w), pad_val=pad_val)


def imflip(img, direction='horizontal'):
    return np.flip(img, axis=1) if direction == 'horizontal' else np.flip(img, axis=0)


def install_image_ops() -> None:
    """install() + the image functions above on the stand-in `mmcv` module."""
    install()
    import mmcv
    for fn in (rescale_size, imresize, imrescale, imnormalize, impad, impad_to_multiple, imflip):
        setattr(mmcv, fn.__name__, fn)
