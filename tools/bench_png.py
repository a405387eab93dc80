"""PNG decode: device (mcg_png_parse + mcg_png_decode) against cv2.imdecode on the host cores.
Frames are synthetic photographs written by cv2.imwrite with its defaults (what tools/gaze360_img_reorganize.py:108
produces: Sub filter, Z_RLE, level 1).  usage: python tools/bench_png.py [--size 300] [--batches 224,2240] [--json out]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import cv2  # noqa: E402
import numpy as np  # noqa: E402
import torch  # noqa: E402

import png_cases  # noqa: E402
from mcgaze_b200.png import GpuPngDecoder  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--size', type=int, default=300)
    ap.add_argument('--batches', default='224,2240')
    ap.add_argument('--distinct', type=int, default=32)
    ap.add_argument('--json', default=None)
    args = ap.parse_args()
    files = [cv2.imencode('.png', png_cases._natural(args.size, args.size, k))[1].tobytes() for k in range(args.distinct)]
    raw = args.size * args.size * 3
    out = dict(size=args.size, file_bytes=int(np.mean([len(f) for f in files])), decoded_bytes=raw, runs=[])
    # host: cv2.imdecode, one thread
    arrs = [np.frombuffer(f, np.uint8) for f in files]
    t0 = time.perf_counter()
    reps = 0
    while time.perf_counter() - t0 < 2.0:
        for a in arrs:
            cv2.imdecode(a, cv2.IMREAD_COLOR)
        reps += len(arrs)
    out['cv2_ms_per_frame_one_core'] = 1e3 * (time.perf_counter() - t0) / reps
    dec = GpuPngDecoder(0)
    want = torch.from_numpy(cv2.imdecode(arrs[0], cv2.IMREAD_COLOR)).cuda()
    for n in [int(b) for b in args.batches.split(',')]:
        batch = [files[k % len(files)] for k in range(n)]
        reps = 5
        t0 = time.perf_counter()
        staged = dec.stage(batch)
        stage_ms = 1e3 * (time.perf_counter() - t0)
        stageds = [staged] + [dec.stage(batch) for _ in range(reps + 1)]     # a staged batch is launched once
        for sg in stageds[:2]:
            frames, status = dec.launch(sg)
        torch.cuda.synchronize()
        assert int(status.abs().sum()) == 0 and bool((frames[0] == want).all())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for sg in stageds[2:]:
            frames, status = dec.launch(sg)
        e1.record()
        torch.cuda.synchronize()
        assert int(status.abs().sum()) == 0 and bool((frames[0] == want).all())
        ms = e0.elapsed_time(e1) / reps
        out['runs'].append(dict(images=n, device_ms=ms, images_per_s=n / ms * 1e3, decoded_GBps=n * raw / ms / 1e6,
                                host_stage_ms=stage_ms, host_stage_us_per_image=1e3 * stage_ms / n,
                                h2d_bytes=int(staged.block.numel())))
    print(json.dumps(out))
    if args.json:
        json.dump(out, open(args.json, 'w'), indent=1)


if __name__ == '__main__':
    main()
