"""TEST INFRASTRUCTURE: a deterministic CPU stand-in for `MultiClueGaze` + the test pipeline with the reference's call
contract (tools/test_gaze360_gaze.py:96-111), used to run the reference's OWN evaluation script on this repo's `mmdet` /
`mmcv` import shims without a GPU (oracle/gen_golden_slicer.py, tests/test_reference_tools.py).  Outputs depend on the
frame, on its position inside its clip and on the other frames of the clip - like the temporal attention makes the real
model's - and box scores fall on both sides of the 0.5 person threshold, so slicing / ordering / merging mistakes show.
"""
from __future__ import annotations

import numpy as np
import torch

LENGTHS = [1, 5, 7, 8, 9, 10, 11, 12, 15, 23, 30]


def make_anno(lengths=LENGTHS):
    return dict(videos=[dict(id=vi + 1, file_names=[f'v{vi:03d}/{t:05d}.png' for t in range(L)])
                        for vi, L in enumerate(lengths)])


def frame_code(filename: str) -> int:
    v, t = filename.replace('\\', '/').split('/')[-2:]
    return int(v[1:]) * 1000 + int(t.split('.')[0])


class StubModel:
    def __init__(self, cfg=None):
        self.cfg = cfg
        self.calls = []

    def __call__(self, return_loss=True, rescale=False, format=False, img=None, img_metas=None, clip_length=None):
        assert return_loss is False and rescale is True and format is False
        x = img[0]
        n = x.shape[0]
        T = clip_length or n
        self.calls.append((n, T))
        code = x[:, 0, 0, 0].double() + 251.0 * x[:, 1, 0, 0].double()
        pos = (torch.arange(n) % T).double()
        clip_sum = code.view(-1, T).sum(1).repeat_interleave(T)
        base = code * 0.37 + pos * 1.3 + clip_sum * 0.011
        boxes = torch.stack([base + c * 10 + k * (3 + c) for c in range(3) for k in range(4)], 1).view(n, 3, 4)
        scores = torch.sigmoid(torch.stack([torch.sin(base + c) * 3 for c in range(3)], 1))
        gaze = torch.stack([torch.cos(base * (i + 1) * 0.1) for i in range(12)], 1).view(n, 4, 3)
        gaze = gaze / gaze.norm(dim=-1, keepdim=True)
        boxes, scores, gaze = boxes.float(), scores.float(), gaze.float()
        det = [torch.cat([boxes[i], scores[i][:, None]], 1) for i in range(n)]
        return (det, [[0, 1, 2]] * n), {'gaze_score': gaze[:, 0], 'face_gaze_score': gaze[:, 1],
                                        'eyes_gaze_score': gaze[:, 2], 'head_gaze_score': gaze[:, 3]}


def encode_frame(filename: str) -> np.ndarray:
    """The 'decoded frame' of a file name: a 4 x 4 BGR image whose first two channels carry the frame code."""
    code = frame_code(filename)
    img = np.zeros((4, 4, 3), np.uint8)
    img[..., 0], img[..., 1], img[..., 2] = code % 251, (code // 251) % 251, 7
    return img


class StubCompose:
    """Stands in for mmdet.datasets.pipelines.Compose: per-frame call, DataContainer outputs, `filename` in the meta."""

    def __init__(self, transforms=None):
        self.transforms = transforms

    def __call__(self, data):
        import os.path as osp
        from mcgaze_b200.compat.parallel import DataContainer
        name = osp.join(data['img_prefix'], data['img_info']['filename'])
        img = torch.from_numpy(encode_frame(name).astype(np.float32)).permute(2, 0, 1).contiguous()
        meta = dict(filename=name, ori_filename=data['img_info']['filename'], img_shape=(4, 4, 3), ori_shape=(4, 4, 3),
                    pad_shape=(4, 4, 3), scale_factor=np.ones(4, np.float32), flip=False)
        return dict(img_metas=DataContainer(meta, cpu_only=True), img=DataContainer(img, stack=True))


class StubBatchPipeline:
    """The same frames for mcgaze_b200.evaluate (its `.batch(frames, filenames=)` protocol)."""

    def batch(self, frames, filenames=None):
        img = torch.from_numpy(np.asarray(frames).astype(np.float32)).permute(0, 3, 1, 2).contiguous()
        metas = [dict(img_shape=(4, 4, 3), scale_factor=np.ones(4, np.float32), filename=f) for f in filenames]
        return dict(img=[img], img_metas=[metas])


class StubDetector(StubModel):
    """StubModel behind the surface tools/test.py:196-217 touches when it builds a detector from `cfg.model`
    (build_detector kwargs, load_checkpoint, `.cuda()`, `.eval()`, `.CLASSES`) - registered under a stub name by the tests."""

    CLASSES = ('face', 'eyes', 'head')

    def __init__(self, train_cfg=None, test_cfg=None, **model_cfg):
        super().__init__(cfg=None)
        self.model_cfg = model_cfg
        self.device_index = 0

    def state_dict(self):
        return {}

    def load_state_dict(self, state_dict, strict=False):
        return self

    def to(self, device):
        return self

    def cuda(self, device=None):
        return self

    def eval(self):
        return self


def make_anno_with_gt(lengths=LENGTHS, seed=0):
    """make_anno + one ground-truth gaze per frame (annotation k belongs to video k, tools/calculate_mae_gaze360.py:118-121)."""
    rng = np.random.default_rng(seed)
    anno = make_anno(lengths)
    anno['annotations'] = []
    for L in lengths:
        g = rng.normal(size=(L, 3))
        g /= np.linalg.norm(g, axis=1, keepdims=True)
        anno['annotations'].append(dict(gaze=g.tolist()))
    return anno
