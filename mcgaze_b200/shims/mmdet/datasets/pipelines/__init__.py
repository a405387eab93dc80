"""mmdet.datasets.pipelines.Compose for the reference's TEST pipelines (configs/_base_/datasets/gaze360.py:27-36,
multiclue_gaze_r50_l2cs.py:31-39), as tools/test_gaze360_gaze.py:58,96-105 drives it: one call per frame with
`dict(img_info=dict(filename=...), img_prefix=...)`, the results of a clip are sorted by `img_metas.data['filename']`,
collated and scattered - and as MCGaze_demo/demo.ipynb cells 3-4 drive it: `Compose(cfg.data.test.pipeline[1:])` on
`dict(filename=j, img=<decoded crop>, ...)`, up to ~101 frames of different sizes collated into one clip.

LoadImageFromFile (mmdet/datasets/pipelines/loading.py:58-69) runs on the host (cv2 decode); every later step is
mcgaze_b200.pipeline.GpuTestPipeline (one mcg_preprocess launch on the decoded uint8 frame).  The output has the
reference's structure: `img` = DataContainer(Tensor[3, H, W], stack=True), `img_metas` = DataContainer(meta,
cpu_only=True) (formatting.py:96, :331-333) - the tensor already lives on the GPU, so `scatter` has nothing to move."""
from __future__ import annotations

import os.path as osp
from typing import Any, Dict, Sequence

from mcgaze_b200.compat.parallel import DataContainer
from mcgaze_b200.pipeline import GpuTestPipeline


class Compose:
    def __init__(self, transforms: Sequence[Dict[str, Any]], device: int = 0):
        self.cfg = [dict(t) for t in transforms]
        self.load = any(t.get('type') == 'LoadImageFromFile' for t in self.cfg)
        self.pipeline = GpuTestPipeline(self.cfg, device=device)

    def __call__(self, data: Dict[str, Any]) -> Dict[str, Any]:
        data = dict(data)
        name = ori = None
        if 'img' not in data:
            if not self.load:
                raise KeyError("pipeline has no LoadImageFromFile step and the input carries no decoded 'img'")
            import cv2
            ori = data['img_info']['filename']
            name = osp.join(data['img_prefix'], ori) if data.get('img_prefix') is not None else ori
            img = cv2.imread(name, cv2.IMREAD_COLOR)
            if img is None:
                raise FileNotFoundError(name)
            data['img'] = img
        elif data.get('img_info'):
            name = ori = data['img_info'].get('filename')
        else:
            # a caller that decodes (and crops) itself, like MCGaze_demo/demo.ipynb cell 4: dict(filename=j, ori_filename=...,
            # img=head_crop, img_shape=..., ori_shape=..., img_fields=['img']) through Compose(pipeline[1:]); Collect copies
            # these keys into the meta (formatting.py:331-333) and the caller sorts the frames by meta['filename']
            name, ori = data.get('filename'), data.get('ori_filename')
        res = self.pipeline.batch([data['img']], filenames=[name])
        meta = res['img_metas'][0][0]
        meta['ori_filename'] = ori
        if 'img_info' not in data and data.get('ori_shape') is not None:
            meta['ori_shape'] = tuple(data['ori_shape'])
        return dict(img_metas=DataContainer(meta, cpu_only=True), img=DataContainer(res['img'][0][0], stack=True))

    def __repr__(self):
        return f'Compose(GpuTestPipeline, {len(self.cfg)} steps)'
