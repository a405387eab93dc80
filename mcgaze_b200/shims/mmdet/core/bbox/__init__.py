"""mmdet.core.bbox: `bbox_overlaps` (imported by tools/test_gaze360_gaze.py:13; the script never calls it because the
datasets hold one person per video, :147) as a plain torch IoU / GIoU with mmdet's argument meaning
(mmdet/core/bbox/iou_calculators/iou2d_calculator.py)."""
from __future__ import annotations


def bbox_overlaps(bboxes1, bboxes2, mode: str = 'iou', is_aligned: bool = False, eps: float = 1e-6):
    import torch
    assert mode in ('iou', 'iof', 'giou'), f'Unsupported mode {mode}'
    a1 = (bboxes1[..., 2] - bboxes1[..., 0]) * (bboxes1[..., 3] - bboxes1[..., 1])
    a2 = (bboxes2[..., 2] - bboxes2[..., 0]) * (bboxes2[..., 3] - bboxes2[..., 1])
    if is_aligned:
        lt, rb = torch.max(bboxes1[..., :2], bboxes2[..., :2]), torch.min(bboxes1[..., 2:], bboxes2[..., 2:])
        elt, erb = torch.min(bboxes1[..., :2], bboxes2[..., :2]), torch.max(bboxes1[..., 2:], bboxes2[..., 2:])
        union = a1 + a2 if mode != 'iof' else a1
    else:
        lt = torch.max(bboxes1[..., :, None, :2], bboxes2[..., None, :, :2])
        rb = torch.min(bboxes1[..., :, None, 2:], bboxes2[..., None, :, 2:])
        elt = torch.min(bboxes1[..., :, None, :2], bboxes2[..., None, :, :2])
        erb = torch.max(bboxes1[..., :, None, 2:], bboxes2[..., None, :, 2:])
        union = a1[..., None] + a2[..., None, :] if mode != 'iof' else a1[..., None]
    wh = (rb - lt).clamp(min=0)
    overlap = wh[..., 0] * wh[..., 1]
    if mode != 'iof':
        union = union - overlap
    union = torch.max(union, union.new_tensor([eps]))
    ious = overlap / union
    if mode != 'giou':
        return ious
    ewh = (erb - elt).clamp(min=0)
    earea = torch.max(ewh[..., 0] * ewh[..., 1], union.new_tensor([eps]))
    return ious - (earea - union) / earea
