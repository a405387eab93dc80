"""CPU oracle for the MCGaze per-clip inference forward (multiclue_gaze_r50).

TEST INFRASTRUCTURE ONLY.  Nothing under ``mcgaze_b200/`` may import this module; it is
used by ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` as the checker / CPU baseline, never as the product path.

It is a plain torch (fp32 or fp64, CPU) restatement of the reference's algorithm, written
against a flat ``state_dict`` in the reference's own checkpoint key layout.
Every function cites the reference file:line it follows (paths relative to /root/reference).

Parity pinning (see tests/golden/ and oracle/gen_golden.py):
  * the reference's OWN python code for this path (``mmdet/models/...``) is executed in this
    container on top of a small mmcv stand-in (oracle/refshim) and its outputs are stored as
    golden fixtures; this restatement is checked against them (tests/test_oracle_golden.py).
  * ``delta2bbox`` is checked against the reference's known-answer test
    (tests/test_utils/test_coder.py:27-75).
  * the mmcv-full 1.4.8 primitives themselves (MultiheadAttention, FFN, ConvModule, RoIAlign)
    are NOT under /root/reference; they are restated from their published semantics
     on top of torch / torchvision ops -> "parity unpinned" at the
    mmcv boundary only.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

NUM_STAGES = 4          # configs/multiclue_gaze/multiclue_gaze_r50_gaze360.py:6
NUM_CLUES = 3           # fixed_embedding_rpn_head.py:28  (0=face, 1=eyes, 2=head)
D_MODEL = 256
NUM_HEADS = 8
FEAT_CH = 64            # DynamicConv feat_channels, cfg :56
ROI_OUT = 7             # cfg :38
FPN_STRIDES = (4, 8, 16, 32)
FINEST_SCALE = 56       # single_level_roi_extractor.py:30
STAGE_BLOCKS = (3, 4, 6, 3)
WH_RATIO_CLIP = 16 / 1000
BBOX_STDS = (0.5, 0.5, 1.0, 1.0)   # cfg :72


class Hooks:
    """Optional instrumentation: ``quant`` is applied to both operands of every conv / linear
    / bmm (used to budget reduced-precision operand error); ``tap`` receives named
    intermediates (used by per-op parity tests)."""

    def __init__(self, quant: Optional[Callable[[Tensor], Tensor]] = None,
                 tap: Optional[Callable[[str, Tensor], None]] = None,
                 quant_head: Optional[Callable[[Tensor], Tensor]] = None,
                 quant_w: Optional[Callable[[Tensor], Tensor]] = None):
        self.quant = quant
        self.quant_w = quant_w if quant_w is not None else quant
        self.quant_head = quant_head if quant_head is not None else quant
        self.tap = tap

    def q(self, t: Tensor) -> Tensor:
        return t if self.quant is None else self.quant(t)

    def qw(self, t: Tensor) -> Tensor:
        return t if self.quant_w is None else self.quant_w(t)

    def qh(self, t: Tensor) -> Tensor:
        return t if self.quant_head is None else self.quant_head(t)

    def t(self, name: str, t: Tensor) -> None:
        if self.tap is not None:
            self.tap(name, t)


_NOHOOK = Hooks()


# ----------------------------------------------------------------------------------------
# backbone: mmdet/models/backbones/resnet.py:263-302 (Bottleneck.forward), :631-646
# (ResNet.forward), mmdet/models/utils/res_layer.py:39-104 (downsample = conv1x1(stride)+BN)
# ----------------------------------------------------------------------------------------
def _conv(x: Tensor, w: Tensor, b: Optional[Tensor], stride: int, pad: int, hk: Hooks) -> Tensor:
    return F.conv2d(hk.q(x), hk.qw(w), b, stride=stride, padding=pad)


def _bn(x: Tensor, sd: SD, p: str) -> Tensor:
    # eval-mode BatchNorm2d, eps=1e-5 (build_norm_layer default; norm_eval=True cfg :18)
    return F.batch_norm(x, sd[p + '.running_mean'], sd[p + '.running_var'],
                        sd[p + '.weight'], sd[p + '.bias'], training=False, eps=1e-5)


def bottleneck(x: Tensor, sd: SD, p: str, stride: int, hk: Hooks = _NOHOOK) -> Tensor:
    """resnet.py:263-302, style='pytorch' -> stride on the 3x3 conv (resnet.py:154-156)."""
    identity = x
    out = F.relu(_bn(_conv(x, sd[p + '.conv1.weight'], None, 1, 0, hk), sd, p + '.bn1'))
    out = F.relu(_bn(_conv(out, sd[p + '.conv2.weight'], None, stride, 1, hk), sd, p + '.bn2'))
    out = _bn(_conv(out, sd[p + '.conv3.weight'], None, 1, 0, hk), sd, p + '.bn3')
    if (p + '.downsample.0.weight') in sd:
        identity = _bn(_conv(x, sd[p + '.downsample.0.weight'], None, stride, 0, hk),
                       sd, p + '.downsample.1')
    return F.relu(out + identity)


def resnet50(x: Tensor, sd: SD, hk: Hooks = _NOHOOK, prefix: str = 'backbone') -> List[Tensor]:
    """resnet.py:631-646: conv1-bn-relu-maxpool, layer1..4, out_indices (0,1,2,3)."""
    x = F.relu(_bn(_conv(x, sd[prefix + '.conv1.weight'], None, 2, 3, hk), sd, prefix + '.bn1'))
    hk.t('stem', x)
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    hk.t('pool', x)
    outs = []
    for li, nblk in enumerate(STAGE_BLOCKS):
        for bi in range(nblk):
            stride = 2 if (bi == 0 and li > 0) else 1
            x = bottleneck(x, sd, f'{prefix}.layer{li + 1}.{bi}', stride, hk)
            hk.t(f'layer{li + 1}.{bi}', x)
        outs.append(x)
    return outs


# ----------------------------------------------------------------------------------------
# neck: mmdet/models/necks/fpn.py:151-204 (num_outs == num_ins -> no extra levels)
# ConvModule(norm_cfg=None, act_cfg=None) == conv + bias
# ----------------------------------------------------------------------------------------
def fpn(feats: Sequence[Tensor], sd: SD, hk: Hooks = _NOHOOK, prefix: str = 'neck') -> List[Tensor]:
    lat = [_conv(f, sd[f'{prefix}.lateral_convs.{i}.conv.weight'],
                 sd[f'{prefix}.lateral_convs.{i}.conv.bias'], 1, 0, hk)
           for i, f in enumerate(feats)]
    for i in range(len(lat) - 1, 0, -1):                       # fpn.py:165-174
        lat[i - 1] = lat[i - 1] + F.interpolate(lat[i], size=lat[i - 1].shape[2:], mode='nearest')
    outs = [_conv(l, sd[f'{prefix}.fpn_convs.{i}.conv.weight'],
                  sd[f'{prefix}.fpn_convs.{i}.conv.bias'], 1, 1, hk)
            for i, l in enumerate(lat)]
    for i, o in enumerate(outs):
        hk.t(f'fpn{i}', o)
    return outs


# ----------------------------------------------------------------------------------------
# rpn: mmdet/models/dense_heads/fixed_embedding_rpn_head.py:55-94
# ----------------------------------------------------------------------------------------
def init_proposals(sd: SD, img_hw: Tensor, prefix: str = 'rpn_head') -> Tuple[Tensor, Tensor]:
    """img_hw: [N,2] (h,w) of the UNPADDED image (meta['img_shape'], :80-82).
    Returns boxes [N,3,4] xyxy in pixels and features [N,3,256]."""
    b = sd[prefix + '.init_proposal_bboxes.weight']
    cx, cy, w, h = b.unbind(-1)                                  # transforms.py:245-256
    xyxy = torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], -1)
    whwh = torch.stack([img_hw[:, 1], img_hw[:, 0], img_hw[:, 1], img_hw[:, 0]], -1).to(b.dtype)
    boxes = xyxy[None] * whwh[:, None, :]
    feats = sd[prefix + '.init_proposal_features.weight'][None].expand(img_hw.shape[0], -1, -1)
    return boxes, feats.contiguous()


# ----------------------------------------------------------------------------------------
# RoI extractor: single_level_roi_extractor.py:36-115 + mmcv.ops.RoIAlign (aligned=True,
# avg, sampling_ratio=2) restated from the published kernel
# ----------------------------------------------------------------------------------------
def map_roi_levels(boxes: Tensor, num_levels: int = 4) -> Tensor:
    """single_level_roi_extractor.py:36-55; boxes [R,4] xyxy."""
    scale = torch.sqrt((boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1]))
    lvl = torch.floor(torch.log2(scale / FINEST_SCALE + 1e-6))
    return lvl.clamp(min=0, max=num_levels - 1).long()


def roi_align_ref(feat: Tensor, rois: Tensor, spatial_scale: float, out: int = ROI_OUT,
                  sampling: int = 2) -> Tensor:
    """Loop-free torch restatement of RoIAlign(aligned=True, avg).  feat [N,C,H,W];
    rois [K,5] (frame_idx,x1,y1,x2,y2) -> [K,C,out,out].
    start = coord*scale-0.5; bin = roi/out; samples at start + p*bin + (i+.5)*bin/2;
    a sample is 0 if y<-1 or y>H or x<-1 or x>W, else clamp to >=0, low=int(y),
    if low>=H-1: low=high=H-1,y=low; bilinear; mean over the 2x2 samples."""
    K = rois.shape[0]
    N, C, H, W = feat.shape
    dt = feat.dtype
    idx = rois[:, 0].long()
    x1 = rois[:, 1] * spatial_scale - 0.5
    y1 = rois[:, 2] * spatial_scale - 0.5
    x2 = rois[:, 3] * spatial_scale - 0.5
    y2 = rois[:, 4] * spatial_scale - 0.5
    bw = (x2 - x1) / out
    bh = (y2 - y1) / out
    g = (torch.arange(out * sampling, dtype=dt) // sampling).to(dt)          # bin index
    s = (torch.arange(out * sampling, dtype=dt) % sampling).to(dt)           # sample index
    ys = y1[:, None] + g[None] * bh[:, None] + (s[None] + 0.5) * bh[:, None] / sampling  # [K,14]
    xs = x1[:, None] + g[None] * bw[:, None] + (s[None] + 0.5) * bw[:, None] / sampling

    def prep(v: Tensor, size: int):
        valid = (v >= -1.0) & (v <= size)
        v = v.clamp(min=0)
        low = v.floor().long()
        top = low >= size - 1
        low = torch.where(top, torch.full_like(low, size - 1), low)
        high = torch.where(top, torch.full_like(low, size - 1), low + 1)
        v = torch.where(top, low.to(dt), v)
        l = v - low.to(dt)
        return valid, low, high, l, 1.0 - l

    vy, ylo, yhi, ly, hy = prep(ys, H)
    vx, xlo, xhi, lx, hx = prep(xs, W)
    fm = feat[idx]                                               # [K,C,H,W]
    P = out * sampling

    def gather(yi: Tensor, xi: Tensor) -> Tensor:               # -> [K,C,P,P]
        lin = (yi[:, :, None] * W + xi[:, None, :]).reshape(K, 1, P * P).expand(-1, C, -1)
        return fm.reshape(K, C, H * W).gather(2, lin).reshape(K, C, P, P)

    val = (gather(ylo, xlo) * (hy[:, :, None] * hx[:, None, :])[:, None]
           + gather(ylo, xhi) * (hy[:, :, None] * lx[:, None, :])[:, None]
           + gather(yhi, xlo) * (ly[:, :, None] * hx[:, None, :])[:, None]
           + gather(yhi, xhi) * (ly[:, :, None] * lx[:, None, :])[:, None])
    val = val * (vy[:, :, None] & vx[:, None, :])[:, None].to(dt)
    val = val.reshape(K, C, out, sampling, out, sampling).mean(dim=(3, 5))
    return val


def roi_extract(feats: Sequence[Tensor], rois: Tensor, use_torchvision: bool = False) -> Tensor:
    """single_level_roi_extractor.py:57-115: per-RoI level pick, per-level RoIAlign 7x7."""
    lvls = map_roi_levels(rois[:, 1:5], len(feats))
    C = feats[0].shape[1]
    out = feats[0].new_zeros(rois.shape[0], C, ROI_OUT, ROI_OUT)
    for i, f in enumerate(feats):
        inds = (lvls == i).nonzero(as_tuple=False).squeeze(1)
        if inds.numel() == 0:
            continue
        if use_torchvision:
            from torchvision.ops import roi_align
            r = roi_align(f, rois[inds], (ROI_OUT, ROI_OUT), 1.0 / FPN_STRIDES[i], 2, aligned=True)
        else:
            r = roi_align_ref(f, rois[inds], 1.0 / FPN_STRIDES[i])
        out[inds] = r
    return out


# ----------------------------------------------------------------------------------------
# GazeSTQIHead: mmdet/models/roi_heads/bbox_heads/gaze_stqi_head.py:119-202
# ----------------------------------------------------------------------------------------
def _linear(x: Tensor, w: Tensor, b: Optional[Tensor], hk: Hooks) -> Tensor:
    return F.linear(hk.qh(x), hk.qh(w), b)


def _ln(x: Tensor, sd: SD, p: str) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[p + '.weight'], sd[p + '.bias'], eps=1e-5)


def mha_residual(x: Tensor, sd: SD, p: str, hk: Hooks = _NOHOOK) -> Tensor:
    """mmcv MultiheadAttention(256, 8)(x) with key=value=identity=query, seq-first
    x [L, Bn, E]: returns x + out_proj(softmax(q k^T / sqrt(32)) v)  (SURVEY Appendix C)."""
    L, Bn, E = x.shape
    hd = E // NUM_HEADS
    qkv = _linear(x, sd[p + '.attn.in_proj_weight'], sd[p + '.attn.in_proj_bias'], hk)
    q, k, v = qkv.split(E, dim=-1)

    def heads(t: Tensor) -> Tensor:                              # [L,Bn,E] -> [Bn*h, L, hd]
        return t.reshape(L, Bn * NUM_HEADS, hd).transpose(0, 1)

    q, k, v = heads(q), heads(k), heads(v)
    att = torch.softmax((q * (1.0 / math.sqrt(hd))) @ k.transpose(1, 2), dim=-1)
    o = (att @ v).transpose(0, 1).reshape(L, Bn, E)
    o = _linear(o, sd[p + '.attn.out_proj.weight'], sd[p + '.attn.out_proj.bias'], hk)
    return x + o


def dynamic_conv(q: Tensor, roi_feat: Tensor, sd: SD, p: str, hk: Hooks = _NOHOOK) -> Tensor:
    """mmdet/models/utils/transformer.py:1116-1164.  q [R,256]; roi_feat [R,256,7,7]."""
    R = q.shape[0]
    x = roi_feat.flatten(2).permute(0, 2, 1)                     # [R,49,256]  (:1131-1133)
    params = _linear(q, sd[p + '.dynamic_layer.weight'], sd[p + '.dynamic_layer.bias'], hk)
    n_in = D_MODEL * FEAT_CH
    p_in = params[:, :n_in].reshape(R, D_MODEL, FEAT_CH)         # :1136-1137
    p_out = params[:, -n_in:].reshape(R, FEAT_CH, D_MODEL)       # :1138-1139
    f = torch.bmm(hk.qh(x), hk.qh(p_in))                         # :1144
    f = F.relu(_ln(f, sd, p + '.norm_in'))
    f = torch.bmm(hk.qh(f), hk.qh(p_out))                        # :1149
    f = F.relu(_ln(f, sd, p + '.norm_out'))
    f = f.flatten(1)                                             # (position, channel) order
    f = _linear(f, sd[p + '.fc_layer.weight'], sd[p + '.fc_layer.bias'], hk)
    return F.relu(_ln(f, sd, p + '.fc_norm'))


def stqi_head(roi_feat: Tensor, prop: Tensor, clip_length: int, sd: SD, p: str,
              hk: Hooks = _NOHOOK) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """gaze_stqi_head.py:119-202.  roi_feat [R,256,7,7]; prop [N,3,256], N = B*clip_length."""
    N, P, d = prop.shape
    T = clip_length
    x = prop.permute(1, 0, 2)                                    # [3,N,d]          :146
    x = _ln(mha_residual(x, sd, p + '.attention', hk), sd, p + '.attention_norm')   # :151
    x = x.permute(1, 0, 2)                                       # [N,3,d]          :154
    x = x.reshape(N // T, T, P, d).permute(1, 0, 2, 3).reshape(T, N * P // T, d)    # :156-161
    x = _ln(mha_residual(x, sd, p + '.attention', hk), sd, p + '.attention_norm')   # :162 (same weights)
    x = x.reshape(T, N // T, P, d).permute(1, 0, 2, 3).reshape(N, P, d)             # :163-166
    attn_feats = x
    hk.t(p + '.attn', attn_feats)
    q = attn_feats.reshape(-1, d)
    iic = dynamic_conv(q, roi_feat, sd, p + '.instance_interactive_conv', hk)
    obj = _ln(q + iic, sd, p + '.instance_interactive_conv_norm')                   # :173-176
    hk.t(p + '.iic', obj)
    h = F.relu(_linear(obj, sd[p + '.ffn.layers.0.0.weight'], sd[p + '.ffn.layers.0.0.bias'], hk))
    h = _linear(h, sd[p + '.ffn.layers.1.weight'], sd[p + '.ffn.layers.1.bias'], hk)
    obj = _ln(obj + h, sd, p + '.ffn_norm').reshape(N, P, d)                        # :179-180
    hk.t(p + '.obj', obj)
    cls_feat = F.relu(_ln(_linear(obj, sd[p + '.cls_fcs.0.weight'], None, hk), sd, p + '.cls_fcs.1'))
    reg_feat = obj
    for j in range(3):                                                              # :187-188
        reg_feat = F.relu(_ln(_linear(reg_feat, sd[f'{p}.reg_fcs.{3 * j}.weight'], None, hk),
                              sd, f'{p}.reg_fcs.{3 * j + 1}'))
    cls, delta = [], []
    for ci, name in enumerate(('face', 'eyes', 'head')):                            # :191-201
        cls.append(_linear(cls_feat[:, ci], sd[f'{p}.{name}_fc_cls.weight'],
                           sd[f'{p}.{name}_fc_cls.bias'], hk).reshape(N, 1, 1))
        delta.append(_linear(reg_feat[:, ci], sd[f'{p}.{name}_fc_reg.weight'],
                             sd[f'{p}.{name}_fc_reg.bias'], hk).reshape(N, 1, 4))
    return torch.cat(cls, 1), torch.cat(delta, 1), obj, attn_feats


# ----------------------------------------------------------------------------------------
# box decode: mmdet/core/bbox/coder/delta_xywh_bbox_coder.py:163-260
# ----------------------------------------------------------------------------------------
def delta2bbox(rois: Tensor, deltas: Tensor, means=(0., 0., 0., 0.), stds=(1., 1., 1., 1.),
               max_shape=None, wh_ratio_clip: float = WH_RATIO_CLIP, clip_border: bool = True) -> Tensor:
    if deltas.shape[0] == 0:
        return deltas
    ncls = deltas.shape[1] // 4
    d = deltas.reshape(-1, 4) * deltas.new_tensor(stds)[None] + deltas.new_tensor(means)[None]
    r = rois.repeat(1, ncls).reshape(-1, 4)
    pxy = (r[:, :2] + r[:, 2:]) * 0.5
    pwh = r[:, 2:] - r[:, :2]
    max_ratio = float(np.abs(np.log(wh_ratio_clip)))             # :240
    dwh = d[:, 2:].clamp(min=-max_ratio, max=max_ratio)
    gxy = pxy + pwh * d[:, :2]
    gwh = pwh * dwh.exp()
    bb = torch.cat([gxy - gwh * 0.5, gxy + gwh * 0.5], -1)
    if clip_border and max_shape is not None:                    # :255-257
        bb[..., 0::2].clamp_(min=0, max=max_shape[1])
        bb[..., 1::2].clamp_(min=0, max=max_shape[0])
    return bb.reshape(deltas.shape[0], -1)


# ----------------------------------------------------------------------------------------
# GazeHead: mmdet/models/roi_heads/mask_heads/gaze_head.py:138-202
# ----------------------------------------------------------------------------------------
def gaze_head(obj: Tensor, sd: SD, p: str, hk: Hooks = _NOHOOK) -> Dict[str, Tensor]:
    def tower(x: Tensor, name: str) -> Tensor:
        for j in (0, 3):
            x = F.relu(_ln(_linear(x, sd[f'{p}.{name}.{j}.weight'], None, hk), sd, f'{p}.{name}.{j + 1}'))
        return x

    per, weighted = {}, []
    for ci, c in enumerate(('face', 'eyes', 'head')):
        g = _linear(tower(obj[:, ci], f'gaze_{c}_fcs'), sd[f'{p}.fc_{c}.weight'], sd[f'{p}.fc_{c}.bias'], hk)
        conf = _linear(tower(obj[:, ci], f'gaze_{c}_confidence'),
                       sd[f'{p}.fc_{c}_confidence.weight'], sd[f'{p}.fc_{c}_confidence.bias'], hk)
        per[c] = g
        weighted.append(conf * g)                                # :186-190 (expand is a no-op)
    fused = _linear(torch.cat(weighted, 1), sd[p + '.fc_gaze.weight'], sd[p + '.fc_gaze.bias'], hk)

    def unit(v: Tensor) -> Tensor:                               # :197-200, no eps
        return v / torch.norm(v, dim=-1, keepdim=True)

    return {'gaze_score': unit(fused), 'face_gaze_score': unit(per['face']),
            'eyes_gaze_score': unit(per['eyes']), 'head_gaze_score': unit(per['head'])}


# ----------------------------------------------------------------------------------------
# roi head loop: mmdet/models/roi_heads/multiclue_gaze_roi_head.py:287-384 (+ :73-137)
# ----------------------------------------------------------------------------------------
def roi_head_simple_test(fpn_feats: Sequence[Tensor], boxes: Tensor, obj: Tensor, clip_length: int,
                         sd: SD, scale_factor: Optional[Tensor] = None, hk: Hooks = _NOHOOK,
                         use_torchvision: bool = False, prefix: str = 'roi_head') -> Dict[str, Tensor]:
    """boxes [N,3,4] xyxy, obj [N,3,256].  Returns dict with gaze (4x[N,3]), boxes [N,3,4]
    (divided by scale_factor if given, :360-362) and sigmoid scores [N,3]."""
    N = boxes.shape[0]
    cls = None
    for s in range(NUM_STAGES):
        idx = torch.arange(N, dtype=boxes.dtype).repeat_interleave(NUM_CLUES)[:, None]
        rois = torch.cat([idx, boxes.reshape(-1, 4)], 1)         # bbox2roi, transforms.py:75-94
        roi_feat = roi_extract(fpn_feats, rois, use_torchvision)
        hk.t(f'stage{s}.roi_feat', roi_feat)
        cls, delta, obj, _ = stqi_head(roi_feat, obj, clip_length, sd, f'{prefix}.bbox_head.{s}', hk)
        boxes = delta2bbox(rois[:, 1:], delta.reshape(-1, 4), stds=BBOX_STDS,
                           clip_border=False).reshape(N, NUM_CLUES, 4)               # bbox_head.py:380-497
        hk.t(f'stage{s}.boxes', boxes)
    scores = cls.sigmoid().reshape(N, NUM_CLUES)                 # :351-352
    if scale_factor is not None:
        boxes = boxes / scale_factor.to(boxes.dtype)[:, None, :]
    out = gaze_head(obj, sd, f'{prefix}.gaze_head.{NUM_STAGES - 1}', hk)   # :367,377-378 (last stage's obj)
    out['boxes'] = boxes
    out['scores'] = scores
    out['obj_feat'] = obj
    return out


def forward(sd: SD, img: Tensor, clip_length: Optional[int] = None, img_hw: Optional[Tensor] = None,
            scale_factor: Optional[Tensor] = None, hk: Hooks = _NOHOOK,
            use_torchvision: bool = False) -> Dict[str, Tensor]:
    """MultiClueGaze.simple_test (detectors/multiclue_gaze.py:105-131).
    img [N,3,H,W] (N = B*clip_length; the reference's test path is B=1, clip_length=N,
    multiclue_gaze_roi_head.py:340; B>1 follows forward_train's clip_length=T, :229)."""
    N, _, H, W = img.shape
    if clip_length is None:
        clip_length = N
    assert N % clip_length == 0
    if img_hw is None:
        img_hw = torch.tensor([[H, W]] * N, dtype=img.dtype)
    with torch.no_grad():
        feats = fpn(resnet50(img, sd, hk), sd, hk)
        boxes, obj = init_proposals(sd, img_hw)
        return roi_head_simple_test(feats, boxes.to(img.dtype), obj, clip_length, sd, scale_factor, hk,
                                    use_torchvision)


def vector_to_yaw_pitch(v: Tensor) -> Tensor:
    """tools/calculate_mae_gaze360.py:60-66 (vector_to_yaw_pitch): unit-normalise,
    pitch = asin(y), yaw = atan2(x, -z) -> [...,2] (yaw,pitch) in radians."""
    v = v / torch.norm(v, dim=-1, keepdim=True)
    pitch = torch.asin(v[..., 1])
    yaw = torch.atan2(v[..., 0], -v[..., 2])
    return torch.stack([yaw, pitch], -1)


# ----------------------------------------------------------------------------------------
# seeded synthetic checkpoint in the reference key layout
# ----------------------------------------------------------------------------------------
def make_state_dict(seed: int = 0, dtype: torch.dtype = torch.float32, include_unused: bool = True) -> SD:
    """Random-but-realistic weights: He-normal convs, BN with non-trivial running stats and
    a small last-BN gamma per bottleneck (so the residual stream keeps O(1) scale, like a
    trained net), Xavier-uniform head weights (gaze_stqi_head.py:102-117), LayerNorm
    affine near identity.  Deterministic for a given seed and torch version."""
    g = torch.Generator().manual_seed(seed)
    sd: SD = {}

    def randn(*shape, std=1.0):
        return torch.randn(*shape, generator=g, dtype=torch.float32) * std

    def rand(*shape, lo=0.0, hi=1.0):
        return torch.rand(*shape, generator=g, dtype=torch.float32) * (hi - lo) + lo

    def conv(name, cout, cin, k, bias=False, gain=2.0):
        sd[name + '.weight'] = randn(cout, cin, k, k, std=math.sqrt(gain / (cin * k * k)))
        if bias:
            sd[name + '.bias'] = randn(cout, std=0.05)

    def bn(name, c, gamma=1.0):
        sd[name + '.weight'] = rand(c, lo=0.7, hi=1.3) * gamma
        sd[name + '.bias'] = randn(c, std=0.1)
        sd[name + '.running_mean'] = randn(c, std=0.2)
        sd[name + '.running_var'] = rand(c, lo=0.6, hi=1.6)
        sd[name + '.num_batches_tracked'] = torch.tensor(1000, dtype=torch.long)

    def xavier(name, cout, cin, bias=True, bias_std=0.02):
        a = math.sqrt(6.0 / (cin + cout))
        sd[name + '.weight'] = rand(cout, cin, lo=-a, hi=a)
        if bias:
            sd[name + '.bias'] = randn(cout, std=bias_std)

    def ln(name, c):
        sd[name + '.weight'] = rand(c, lo=0.8, hi=1.2)
        sd[name + '.bias'] = randn(c, std=0.05)

    conv('backbone.conv1', 64, 3, 7)
    bn('backbone.bn1', 64)
    cin = 64
    for li, (nblk, planes) in enumerate(zip(STAGE_BLOCKS, (64, 128, 256, 512))):
        for bi in range(nblk):
            p = f'backbone.layer{li + 1}.{bi}'
            conv(p + '.conv1', planes, cin, 1); bn(p + '.bn1', planes)
            conv(p + '.conv2', planes, planes, 3); bn(p + '.bn2', planes)
            conv(p + '.conv3', planes * 4, planes, 1); bn(p + '.bn3', planes * 4, gamma=0.35)
            if bi == 0:
                conv(p + '.downsample.0', planes * 4, cin, 1, gain=1.0); bn(p + '.downsample.1', planes * 4)
            cin = planes * 4
    for i, c in enumerate((256, 512, 1024, 2048)):
        conv(f'neck.lateral_convs.{i}.conv', 256, c, 1, bias=True, gain=1.0)
        conv(f'neck.fpn_convs.{i}.conv', 256, 256, 3, bias=True, gain=1.0)
    # learned boxes: face / eyes / head start near the whole image but not identical, so the
    # three clues exercise different FPN levels (fixed_embedding_rpn_head.py:46-53 inits to
    # (.5,.5,1,1); trained values differ)
    sd['rpn_head.init_proposal_bboxes.weight'] = torch.tensor(
        [[0.50, 0.45, 0.55, 0.60], [0.50, 0.38, 0.30, 0.12], [0.50, 0.50, 0.95, 0.98]])
    sd['rpn_head.init_proposal_features.weight'] = randn(3, 256, std=1.0)
    for s in range(NUM_STAGES):
        p = f'roi_head.bbox_head.{s}'
        xavier(p + '.attention.attn.in_proj', 768, 256)
        sd[p + '.attention.attn.in_proj_weight'] = sd.pop(p + '.attention.attn.in_proj.weight')
        sd[p + '.attention.attn.in_proj_bias'] = sd.pop(p + '.attention.attn.in_proj.bias')
        xavier(p + '.attention.attn.out_proj', 256, 256)
        ln(p + '.attention_norm', 256)
        q = p + '.instance_interactive_conv'
        xavier(q + '.dynamic_layer', 2 * 256 * 64, 256)
        # trained dynamic filters are O(1/sqrt(fan_in)); xavier over 32768 outputs is too small
        sd[q + '.dynamic_layer.weight'] *= 8.0
        ln(q + '.norm_in', 64); ln(q + '.norm_out', 256)
        xavier(q + '.fc_layer', 256, 256 * 49)
        ln(q + '.fc_norm', 256)
        ln(p + '.instance_interactive_conv_norm', 256)
        xavier(p + '.ffn.layers.0.0', 2048, 256)
        xavier(p + '.ffn.layers.1', 256, 2048)
        ln(p + '.ffn_norm', 256)
        xavier(p + '.cls_fcs.0', 256, 256, bias=False); ln(p + '.cls_fcs.1', 256)
        for j in range(3):
            xavier(f'{p}.reg_fcs.{3 * j}', 256, 256, bias=False); ln(f'{p}.reg_fcs.{3 * j + 1}', 256)
        for c in ('face', 'eyes', 'head'):
            xavier(f'{p}.{c}_fc_cls', 1, 256)
            xavier(f'{p}.{c}_fc_reg', 4, 256)
            # a trained regressor emits small refinements; raw xavier deltas make the
            # 4-stage box recursion diverge out of the image (ill-conditioned, unlike the
            # released checkpoints whose boxes stay on the head crop)
            sd[f'{p}.{c}_fc_reg.weight'] *= 0.08
        if include_unused:                                       # BBoxHead.__init__ leftovers (bbox_head.py:66-81)
            xavier(p + '.fc_cls', 4, 12544)
            xavier(p + '.fc_reg', 4, 12544)
        h = f'roi_head.gaze_head.{s}'
        if s == NUM_STAGES - 1 or include_unused:
            for c in ('face', 'eyes', 'head'):
                for tw in (f'gaze_{c}_fcs', f'gaze_{c}_confidence'):
                    for j in (0, 3):
                        xavier(f'{h}.{tw}.{j}', 256, 256, bias=False); ln(f'{h}.{tw}.{j + 1}', 256)
                xavier(f'{h}.fc_{c}', 3, 256, bias_std=0.3)
                xavier(f'{h}.fc_{c}_confidence', 3, 256, bias_std=0.3)
                # gaze mostly frontal (-z) with moderate yaw/pitch like Gaze360 front-180 data,
                # so (yaw,pitch) is well-conditioned (not at the +-y pole)
                sd[f'{h}.fc_{c}.bias'] += torch.tensor([0.0, 0.0, -2.0])
                sd[f'{h}.fc_{c}_confidence.bias'] += torch.tensor([1.0, 1.0, 1.0])
            xavier(h + '.fc_gaze', 3, 9, bias_std=0.1)
            sd[h + '.fc_gaze.weight'] = sd[h + '.fc_gaze.weight'] * 0.3 + torch.eye(3).repeat(1, 3) / 3.0
    if dtype != torch.float32:
        sd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
    return sd


def make_clip(seed: int, n_frames: int, h: int = 224, w: int = 224, dtype=torch.float32) -> Tensor:
    """Synthetic pipeline output: mean/std-normalised pixels ~ N(0,1)."""
    g = torch.Generator().manual_seed(1000 + seed)
    return torch.randn(n_frames, 3, h, w, generator=g, dtype=torch.float32).to(dtype)
