"""Freeze PNG decode fixtures: file bytes + what cv2.imdecode(IMREAD_COLOR) - the decoder LoadImageFromFile calls
(mmdet/datasets/pipelines/loading.py:58-69 via mmcv.imfrombytes) - returns for them in THIS container.
usage: python oracle/gen_golden_png.py   -> tests/golden/png_decode.npz"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import png_cases  # noqa: E402

KEEP = ['cv2_default', 'cv2_level9', 'cv2_gray', 'cv2_rgba', 'mixed_rgb', 'mixed_gray', 'mixed_ga', 'mixed_rgba', 'palette',
        'fixed', 'far_matches', 'long_codes', 'one_col', 'rows_33']

if __name__ == '__main__':
    cases = png_cases.cases()
    out = {}
    for n in KEEP:
        data = np.frombuffer(cases[n], np.uint8)
        out['file_' + n] = data
        out['bgr_' + n] = cv2.imdecode(data, cv2.IMREAD_COLOR)
    path = os.path.join(ROOT, 'tests', 'golden', 'png_decode.npz')
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), 'bytes; cv2', cv2.__version__)
