#!/usr/bin/env python
"""Benchmark of the MCGaze per-clip forward (BASELINE.json metric: clips/sec, 7-frame 224^2, R-50).

    python bench.py --gpus N --steps K --warmup W             # this repo's CUDA engine
    python bench.py --impl reference --gpus N --steps K --warmup W   # reference's CPU path (oracle port)

A "step" is one forward of `configs[1]`: 32 clips x 7 frames x 3 x 224 x 224 synthetic input per GPU
(weak scaling: every rank processes its own 32 clips per step; clips are independent units, so there
is no collective inside the forward; the per-rank results are all-gathered once at the end, inside
the timed region, like the reference's multi_gpu_test gather, mmdet/apis/test.py:179-209).
Rank 0 prints ONE JSON line.  For N > 1 launch with torchrun (one rank per GPU, NCCL).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CLIPS_PER_STEP = 32      # BASELINE.json configs[1]: bs=32 clips
T, H, W = 7, 224, 224
GFLOP_PER_CLIP = 99.55   # trunk 57.22 + FPN 39.79 + head 2.54 (2*MAC, T=7, 224^2)
METRIC = 'clips/sec (7-frame 224^2, R-50)'


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return {'tflops_sustained': d.get('bf16_tflops_sustained'), 'tflops_burst': d.get('bf16_tflops'),
                'hbm_gbs': d.get('hbm_gbs'), 'source': 'measured (MEASURED_PEAKS.json)'}
    return {'tflops_sustained': 1400.0, 'tflops_burst': 1590.0, 'hbm_gbs': 6650.0,
            'source': 'fallback (B200_PROFILING.md)'}


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""
    REASONS = {0x1: 'gpu_idle', 0x2: 'applications_clocks_setting', 0x4: 'sw_power_cap', 0x8: 'hw_slowdown',
               0x10: 'sync_boost', 0x20: 'sw_thermal_slowdown', 0x40: 'hw_thermal_slowdown',
               0x80: 'hw_power_brake_slowdown', 0x100: 'display_clock_setting'}

    def __init__(self, index: int, period: float = 0.05):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reason_bits, self.max_mhz, self.power = [], 0, None, []
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        while self.ok and not self._halt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                self.reason_bits |= int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                self.power.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            self._halt.wait(self.period)

    def finish(self):
        self._halt.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': ['unavailable'], 'samples': 0}
        s = sorted(self.samples)
        reasons = [n for b, n in self.REASONS.items() if self.reason_bits & b and n != 'gpu_idle']
        return {'sm_mhz': s[len(s) // 2], 'sm_max_mhz': self.max_mhz, 'reasons': reasons, 'samples': len(s),
                'power_w_max': max(self.power) if self.power else None}


def cpu_reference_throughput(budget_s: float, warm: int = 1):
    """The reference's PyTorch-CPU path (oracle port: same torch ops the reference calls) on all host
    threads, one 7-frame clip per forward like tools/test_gaze360_gaze.py, bounded to ~budget_s."""
    import torch
    from oracle import mcgaze_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    sd = O.make_state_dict(0)
    clips = [O.make_clip(100 + i, T, H, W) for i in range(4)]
    for i in range(warm):
        O.forward(sd, clips[i % 4])
    n, t0 = 0, time.perf_counter()
    while True:
        O.forward(sd, clips[n % 4])
        n += 1
        el = time.perf_counter() - t0
        if el >= budget_s or n >= 200:
            break
    return n / el, n, el, torch.get_num_threads()


def preprocess_extras(eng, dev, steps: int, graph: bool):
    """SURVEY section 8 row f3 beside the headline: (1) device time / achieved HBM bandwidth of mcg_preprocess on one
    step's worth of frames, L2 flushed before every timed launch; (2) the same end-to-end metric as `e2e`, but fed
    with decoded uint8 frames: H2D of uint8 -> mcg_preprocess -> mcg_forward -> D2H, copy of step i+1 overlapped
    with the forward of step i."""
    import numpy as np
    import torch
    from mcgaze_b200 import lib
    from mcgaze_b200.compat import Config
    from mcgaze_b200.pipeline import GpuTestPipeline
    root = os.path.dirname(os.path.abspath(__file__))
    cfg = Config.fromfile(os.path.join(root, 'configs/multiclue_gaze/multiclue_gaze_r50_gaze360.py'))
    pipe = GpuTestPipeline(cfg.data.test.pipeline, device=dev.index, seed=0)
    NB, SH, SW = CLIPS_PER_STEP * T, 320, 320
    gen = torch.Generator().manual_seed(99)
    host = [torch.randint(0, 256, (NB, SH, SW, 3), dtype=torch.uint8, generator=gen).pin_memory() for _ in range(2)]
    # one crop draw per clip here (every frame of a clip shares its window), so all frames land on one 224 x 224 canvas
    rands = np.repeat(np.random.RandomState(0).rand(CLIPS_PER_STEP), T)
    geometry, metas, (Hp, Wp) = pipe.plan([(SH, SW)] * NB, rands)
    assert (Hp, Wp) == (H, W)
    frames = host[0].to(dev)
    canvas = torch.empty(NB, 3, Hp, Wp, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        lib.preprocess(frames, geometry, pipe.mean, pipe.std, pipe.to_rgb, canvas)
    torch.cuda.synchronize()
    ms = []
    for _ in range(20):
        flush.zero_()
        torch.cuda._sleep(1_000_000)          # the host enqueues the timed launch while the GPU is still busy
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        lib.preprocess(frames, geometry, pipe.mean, pipe.std, pipe.to_rgb, canvas)
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    k_ms = float(np.median(ms))
    rd = int(sum(3 * g[2] * g[3] for g in geometry))
    wr = NB * 3 * Hp * Wp * 4
    peaks = load_peaks()
    kernel = {'kernel': 'mcg::preprocess_kernel (crop + cv2-exact bilinear resize + BGR->RGB + normalise + pad, u8 HWC -> fp32 NCHW)',
              'bound': 'hbm', 'achieved': (rd + wr) / k_ms / 1e6, 'peak': peaks.get('hbm_gbs'), 'unit': 'GB/s',
              'frac': (rd + wr) / k_ms / 1e6 / peaks['hbm_gbs'] if peaks.get('hbm_gbs') else None,
              'ms_per_launch': k_ms, 'frames_per_launch': NB, 'launches_per_step': 1,
              'algorithmic_bytes': {'read_crop_windows_u8': rd, 'write_canvas_f32': wr},
              'source': f'{NB} frames of {SH}x{SW}x3 uint8, CenterCrop(0.68..1) -> {Hp}x{Wp}', 'l2': 'flushed (256 MB memset) before every timed launch'}

    img_hw = np.array([[m['img_shape'][0], m['img_shape'][1]] for m in metas], dtype=np.float32)
    scale = np.stack([m['scale_factor'] for m in metas])
    eng.set_graph_mode(graph)
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)
    dbuf = [torch.empty_like(frames), torch.empty_like(frames)]
    copied = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    outs = [None, None]
    pinned = [{k: torch.empty(shape, dtype=torch.float32).pin_memory() for k, shape in
               (('gaze', (NB, 4, 3)), ('boxes', (NB, 3, 4)), ('scores', (NB, 3)))} for _ in range(2)]

    def upload(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i % 2])
            dbuf[i % 2].copy_(host[i % 2], non_blocking=True)
            copied[i % 2].record(copy_stream)

    def step(i):
        main_stream.wait_event(copied[i % 2])
        lib.preprocess(dbuf[i % 2], geometry, pipe.mean, pipe.std, pipe.to_rgb, canvas)
        consumed[i % 2].record(main_stream)
        o = eng.forward(canvas, clip_length=T, img_hw=img_hw, scale_factor=scale) if outs[0] is None else outs[0]
        if outs[0] is None:
            outs[0] = o
        else:
            eng.forward_into(canvas, T, o, img_hw=img_hw, scale_factor=scale)
        for k in pinned[i % 2]:
            pinned[i % 2][k].copy_(o[k], non_blocking=True)

    for ev in consumed:
        ev.record(main_stream)
    n_steps = max(3, min(steps, 10))
    for rep in range(2):                       # rep 0 = warm-up (plans, graph capture), rep 1 = timed
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        upload(0)
        for i in range(n_steps):
            if i + 1 < n_steps:
                upload(i + 1)
            step(i)
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
    # the reference's CPU pipeline on the same frames (its transforms' arithmetic = mmcv wrappers over cv2, restated
    # in oracle/refshim.py), one host thread, a bounded sample
    cpu = None
    try:
        import cv2  # noqa: F401
        from oracle import refshim
        cv2.setNumThreads(1)
        sample = host[0][:64].numpy()
        mean64, std32 = pipe.mean, pipe.std
        t0 = time.perf_counter()
        reps = 0
        while time.perf_counter() - t0 < 2.0:
            for k in range(sample.shape[0]):
                y, x, ch, cw, nh, nw = geometry[k]
                im = refshim.imresize(sample[k][y:y + ch, x:x + cw], (nw, nh))
                im = refshim.imnormalize(im, mean64, std32, pipe.to_rgb)
                im = refshim.impad_to_multiple(im, 32)
                np.ascontiguousarray(im.transpose(2, 0, 1))
            reps += 1
        el_cpu = time.perf_counter() - t0
        cpu = {'value': reps * sample.shape[0] / el_cpu, 'unit': 'frames/s', 'cores': 1, 'kind': 'port',
               'sample': f'{reps * sample.shape[0]} frames in {el_cpu:.1f} s: cv2.resize + mmcv.imnormalize + impad + transpose '
                         '(published mmcv wrappers over cv2, one thread), same frames and crop windows'}
    except Exception as e:      # the side measurement must never take the headline line down
        cpu = {'unavailable': f'{type(e).__name__}: {e}'}
    kernel['frames_per_s'] = NB / k_ms * 1e3
    return {'kernel': kernel, 'cpu_baseline': cpu,
            'e2e_u8': {'value': CLIPS_PER_STEP * n_steps / el, 'unit': 'clips/s', 'h2d_bytes_per_step': NB * SH * SW * 3,
                       'd2h_bytes_per_step': sum(v.numel() * 4 for v in pinned[0].values()), 'steps': n_steps,
                       'api': 'pinned uint8 frames -> H2D (copy stream, overlapped) -> mcg_preprocess -> mcg_forward -> D2H'}}


GFLOP_PER_CLIP_BY_SHAPE = {(224, 1): 14.22, (224, 3): 42.66, (224, 7): 99.55, (224, 15): 213.32, (320, 1): 28.64,
                           (320, 3): 85.93, (320, 7): 200.51, (320, 15): 429.67, (448, 7): 390.57}


def _time_shape(eng, dev, res: int, clip_len: int, clips: int, steps: int, graph: bool):
    """Device time of one forward of `clips` clips of clip_len x res x res (inputs resident, like `value`)."""
    import torch
    img = torch.randn(clips * clip_len, 3, res, res, device=dev)
    out = eng.forward(img, clip_length=clip_len)
    eng.set_graph_mode(graph)
    for _ in range(3):
        eng.forward_into(img, clip_len, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        eng.forward_into(img, clip_len, out)
    e1.record()
    torch.cuda.synchronize()
    eng.set_graph_mode(False)
    ms = e0.elapsed_time(e1) / steps
    cps = clips / (ms * 1e-3)
    del img, out
    torch.cuda.empty_cache()
    return {'resolution': res, 'T': clip_len, 'clips_per_step': clips, 'ms_per_step': ms, 'clips_per_s': cps,
            'frames_per_s': cps * clip_len, 'algorithmic_tflops': cps * GFLOP_PER_CLIP_BY_SHAPE[(res, clip_len)] / 1e3}


def config_extras(eng, dev, graph: bool):
    """BASELINE configs[2] (l2cs setting: bs = 32 clips x 7 frames x 448 x 448, multiclue_gaze_r50_l2cs.py:31-43) and
    configs[4] (clip-length sweep T in {1, 3, 7, 15} x resolution {224, 320}, ~224 / ~112 frames per step) as side
    measurements next to the headline: device-resident synthetic inputs, CUDA-graph replay, 5 timed steps each.
    Parity of these shapes: tests/test_gpu_forward.py (test_l2cs_batch_8x7x448_vs_oracle, test_config_sweep_shapes_vs_oracle)."""
    l2cs = _time_shape(eng, dev, 448, 7, 32, 5, graph)
    l2cs['workload'] = 'multiclue_gaze_r50 l2cs-setting inference, bs=32 clips x 7 frames x 448x448 (BASELINE configs[2])'
    sweep = [_time_shape(eng, dev, res, t, frames // t, 5, graph)
             for res, t, frames in [(224, 1, 224), (224, 3, 222), (224, 7, 224), (224, 15, 225), (320, 1, 112), (320, 3, 111),
                                    (320, 7, 112), (320, 15, 105)]]
    return {'l2cs_bs32_448': l2cs, 'sweep': {'workload': 'clip-length sweep T x resolution (BASELINE configs[4])', 'rows': sweep}}


def testsplit_extra(eng, sd, local_rank: int, world: int):
    """BASELINE configs[3]: the Gaze360 test split (517 videos with the REAL length list of the shipped results JSON =
    25 969 frames = 6365 clips, synthetic 300 x 300 uint8 frames from a RAM bank, PNG decoding excluded) through the
    sharded evaluation driver on ALL ranks: videos sharded round-robin, uint8 H2D on a copy stream, mcg_preprocess,
    batched forward, overlap merge + MAE on each device, ONE all-reduce of 6 doubles (+ one all-gather of the merged
    rows, as a JSON writer needs).  Wall clock, max over ranks."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from mcgaze_b200 import evaluate as ev
    from mcgaze_b200.apis import init_detector
    from mcgaze_b200.pipeline import GpuTestPipeline
    lengths = np.load(os.path.join(ROOT, 'tests/golden/golden_gaze360_results.npz'))['lengths'].tolist()
    grng = np.random.default_rng(1)
    anno = dict(videos=[dict(id=i + 1, file_names=[f'{i:04d}/{t:05d}.png' for t in range(L)]) for i, L in enumerate(lengths)],
                annotations=[dict(gaze=grng.normal(size=(L, 3)).tolist()) for L in lengths])
    rng = np.random.default_rng(0)
    bank = [rng.integers(0, 256, (300, 300, 3), dtype=np.uint8) for _ in range(64)]
    ds = ev.Gaze360ClipDataset(anno, loader=lambda p: bank[hash(p) % 64])
    model = init_detector(os.path.join(ROOT, 'configs/multiclue_gaze/multiclue_gaze_r50_gaze360.py'), None, f'cuda:{local_rank}')
    model.load_state_dict(sd)
    model._engine = eng                    # the engine of the headline measurement (same weights)
    eng.set_graph_mode(True)
    pipe = GpuTestPipeline(model.cfg.data.test.pipeline, device=local_rank, seed=0)
    ok = 1
    try:
        ev.run_clips(model, ds, pipe, list(range(64)), 32, 8)          # warm-up: plans, graph capture, staging buffers
        torch.cuda.synchronize()
    except Exception:
        ok = 0
    if world > 1:      # every rank enters the sharded run (it holds collectives) or none does
        t = torch.tensor([ok], device=f'cuda:{local_rank}')
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok = int(t.item())
    if not ok:
        eng.set_graph_mode(False)
        raise RuntimeError('warm-up of the evaluation driver failed on a rank')
    t0 = time.perf_counter()
    out = ev.multi_gpu_test_videos(model, ds, pipe, 32, workers=8, gather_videos=True)
    torch.cuda.synchronize()
    el = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([el], device=f'cuda:{local_rank}')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        el = float(t.item())
    eng.set_graph_mode(False)
    return {'workload': 'Gaze360 test split stand-in, sharded by video over the ranks (BASELINE configs[3])', 'videos': len(lengths),
            'frames': int(sum(lengths)), 'clips': len(ds), 'n_gpus': world, 'seconds': el, 'clips_per_s': len(ds) / el,
            'collectives': 'one all-reduce of 6 doubles (MAE) + one all-gather of the merged per-frame rows (JSON)',
            'mae_360_of_random_gt': out['mae']['mae_360'], 'frames_scored': out['mae']['frames_360'],
            'note': 'wall clock, max over ranks; host frame hand-over, uint8 H2D, GPU pipeline, forward, device merge + scorer, '
                    'collectives; PNG decoding excluded (frames come from a RAM bank)'}


def png_decode_extra(eng, sd, dev):
    """SURVEY section 8 row f3, the "decode" step: PNG files -> BGR frames in HBM (mcg_png_parse on the host, mcg_png_decode
    on the device) beside cv2.imdecode on one host core, on synthetic 300 x 300 photographs written by cv2.imwrite with
    its defaults (what tools/gaze360_img_reorganize.py:108 produces); and the Gaze360 test-split stand-in of
    `testsplit_extra` once more, this time from PNG FILES (64 distinct files, page-cached) with decode='gpu': file read,
    chunk walk, compressed H2D, device decode, mcg_preprocess, forward, device merge + scorer."""
    import tempfile
    import cv2
    import numpy as np
    import torch
    from mcgaze_b200 import evaluate as ev
    from mcgaze_b200.apis import init_detector
    from mcgaze_b200.pipeline import GpuTestPipeline
    from mcgaze_b200.png import GpuPngDecoder

    def photo(seed, h=300, w=300):
        rng = np.random.default_rng(seed)
        y, x = np.mgrid[0:h, 0:w].astype(np.float32)
        img = np.stack([128 + 90 * np.sin(x / 17 + c) * np.cos(y / 23 - c) + 12 * rng.standard_normal((h, w)) for c in range(3)], -1)
        return np.clip(img, 0, 255).astype(np.uint8)

    files = [cv2.imencode('.png', photo(k))[1].tobytes() for k in range(64)]
    arrs = [np.frombuffer(f, np.uint8) for f in files]
    t0 = time.perf_counter()
    reps = 0
    while time.perf_counter() - t0 < 1.5:
        for a in arrs[:16]:
            cv2.imdecode(a, cv2.IMREAD_COLOR)
        reps += 16
    cv2_ms = 1e3 * (time.perf_counter() - t0) / reps
    dec = GpuPngDecoder(dev.index)
    n = 2240
    batch = [files[k % 64] for k in range(n)]
    t0 = time.perf_counter()
    stageds = [dec.stage(batch) for _ in range(4)]
    stage_us = 1e6 * (time.perf_counter() - t0) / (4 * n)
    frames, status = dec.launch(stageds[0])
    torch.cuda.synchronize()
    want = torch.from_numpy(cv2.imdecode(arrs[0], cv2.IMREAD_COLOR)).to(dev)
    exact = bool((frames[0] == want).all()) and int(status.abs().sum()) == 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for sg in stageds[1:]:
        frames, status = dec.launch(sg)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    raw = 300 * 300 * 3
    out = {'kernel': 'mcg::png_inflate_kernel + mcg::png_unfilter_kernel (one warp per image, one launch per batch)',
           'images': n, 'frame': '300x300x3, cv2.imwrite defaults', 'file_bytes': int(np.mean([len(f) for f in files])),
           'device_ms': ms, 'images_per_s': n / ms * 1e3, 'decoded_GBps': n * raw / ms / 1e6,
           'h2d_bytes': int(stageds[0].block.numel()), 'host_parse_us_per_image_from_memory_with_crc_one_thread': stage_us,
           'cv2_imdecode_ms_per_image_one_core': cv2_ms, 'host_cores_equivalent': (n / ms * 1e3) * cv2_ms / 1e3,
           'bit_exact_vs_cv2': exact,
           'note': 'device time includes the H2D copy of the compressed bytes; timed with CUDA events over 3 launches'}
    del frames, status, stageds
    # test-split stand-in from PNG files
    lengths = np.load(os.path.join(ROOT, 'tests/golden/golden_gaze360_results.npz'))['lengths'].tolist()
    with tempfile.TemporaryDirectory() as tmp:
        for k, f in enumerate(files):
            with open(os.path.join(tmp, f'{k:02d}.png'), 'wb') as fh:
                fh.write(f)
        # host side of a unit of 8 batches from FILES: mcg_png_file_sizes + mcg_png_stage_files (read, chunk walk, copy into
        # the pinned block on 8 worker threads inside the C call)
        from concurrent.futures import ThreadPoolExecutor
        paths = [os.path.join(tmp, f'{k % 64:02d}.png') for k in range(1792)]
        with ThreadPoolExecutor(8) as tp:
            dec_nocrc = GpuPngDecoder(dev.index, check_crc=False)
            dec_nocrc.stage(paths, tp)
            t0 = time.perf_counter()
            for _ in range(3):
                dec_nocrc.stage(paths, tp)
            out['host_stage_files_ms_per_1792_8_threads'] = 1e3 * (time.perf_counter() - t0) / 3
        grng = np.random.default_rng(1)
        # one path per (video, frame) like data/gaze360/test_rawframes/<vid>/<00000>.png, hard-linked to the 64 files (the
        # driver decodes a file once per unit however many clips share it, so the names must be distinct to count)
        for i, L in enumerate(lengths):
            os.makedirs(os.path.join(tmp, f'{i:04d}'))
            for t in range(L):
                src, dst = os.path.join(tmp, f'{(i * 131 + t) % 64:02d}.png'), os.path.join(tmp, f'{i:04d}', f'{t:05d}.png')
                try:
                    os.link(src, dst)
                except OSError:
                    os.symlink(src, dst)
        anno = dict(videos=[dict(id=i + 1, file_names=[f'{i:04d}/{t:05d}.png' for t in range(L)]) for i, L in enumerate(lengths)],
                    annotations=[dict(gaze=grng.normal(size=(L, 3)).tolist()) for L in lengths])
        ds = ev.Gaze360ClipDataset(anno, img_prefix=tmp, decode='gpu')
        model = init_detector(os.path.join(ROOT, 'configs/multiclue_gaze/multiclue_gaze_r50_gaze360.py'), None, f'cuda:{dev.index}')
        model.load_state_dict(sd)
        model._engine = eng
        eng.set_graph_mode(True)
        try:
            pipe = GpuTestPipeline(model.cfg.data.test.pipeline, device=dev.index, seed=0)
            ev.run_clips(model, ds, pipe, list(range(64)), 32, 8)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res = ev.multi_gpu_test_videos(model, ds, pipe, 32, workers=8, gather_videos=True)
            torch.cuda.synchronize()
            el = time.perf_counter() - t0
        finally:
            eng.set_graph_mode(False)
    out['testsplit_from_png_files'] = {'clips': len(ds), 'frames': int(sum(lengths)), 'seconds': el, 'clips_per_s': len(ds) / el,
                                       'clip_frames_per_s': sum(c['n'] for c in (ds.clip_info(i) for i in range(len(ds)))) / el,
                                       'host_decoded_batches': ds.host_decoded_batches, 'frames_scored': res['mae']['frames_360'],
                                       'note': 'wall clock on one GPU; 25 969 distinct paths (hard links to 64 files); a file is read and '
                                               'decoded once per unit of 8 batches however many of its clips share it'}
    return out


def run_reference(args, rank, world):
    if rank != 0:
        return
    per_step_budget = max(1.0, min(20.0, 120.0 / max(args.steps + args.warmup, 1)))
    import torch
    from oracle import mcgaze_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    sd = O.make_state_dict(0)
    clips = [O.make_clip(100 + i, T, H, W) for i in range(4)]
    n_per_step = None
    times = []
    for s in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        if n_per_step is None:            # size the bounded sample on the first (warm-up) step
            n = 0
            while time.perf_counter() - t0 < per_step_budget and n < 64:
                O.forward(sd, clips[n % 4])
                n += 1
            n_per_step = max(n, 1)
        else:
            for i in range(n_per_step):
                O.forward(sd, clips[i % 4])
        if s >= args.warmup:
            times.append(time.perf_counter() - t0)
    if not times:                          # warmup only
        times = [per_step_budget]
    total = sum(times)
    value = n_per_step * len(times) / total
    cores = torch.get_num_threads()
    sample = f'{n_per_step} clips of 7x3x224x224 per step, one clip per forward, fp32, {cores} threads'
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'clips/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / len(times),
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'multiclue_gaze_r50 Gaze360-setting inference, 7-frame 224x224 clips '
                                   '(bounded CPU sample of the bs=32 workload)', 'clips_per_step': n_per_step,
                       'clip_length': T, 'height': H, 'width': W,
                       'note': 'mmcv cannot be installed offline, so the reference arm is the oracle port of the '
                               "reference's forward (same torch CPU ops), see DESIGN.md"},
            'cpu_baseline': {'value': value, 'unit': 'clips/s', 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': value, 'unit': 'clips/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


# measured DRAM traffic (bytes) of the dominant GEMM launch, from the ncu capture committed under profiles/
NCU_TRAFFIC = {'fp16c8': 722660864 + 504889856}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='mcgaze_b200', choices=['mcgaze_b200', 'reference'])
    ap.add_argument('--precision', default='fp16c8', choices=['fp16c8', 'fp16x3', 'fp16', 'simt'])
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--cpu-budget', type=float, default=12.0, help='seconds of CPU work for cpu_baseline')
    ap.add_argument('--skip-extras', action='store_true', help='skip fast-mode / cpu_baseline side measurements')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))

    if args.impl == 'reference':
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from mcgaze_b200 import lib
    from oracle import mcgaze_oracle as O      # only for the seeded synthetic checkpoint + cpu_baseline leg

    assert torch.cuda.is_available(), 'bench.py needs a CUDA device: the product has no CPU path'
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        import datetime
        # a rank that dies must not cost the others the default 10-minute collective timeout
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank), timeout=datetime.timedelta(seconds=180))
    dev = torch.device('cuda', local_rank)

    sd = O.make_state_dict(0)
    eng = lib.Engine(sd, local_rank, args.precision)
    NB = CLIPS_PER_STEP * T
    g = torch.Generator().manual_seed(1234 + rank)
    host_img = torch.randn(NB, 3, H, W, generator=g).pin_memory()
    img = host_img.to(dev)
    out = eng.forward(img, clip_length=T)
    torch.cuda.synchronize()
    launches_per_step = eng.last_launch_count

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput (`value`) ----------------
    eng.set_graph_mode(not args.no_graph)
    for _ in range(args.warmup):
        eng.forward_into(img, T, out)
    if world > 1:      # the collective of the timed region: communicator and buffers exist before the clock starts
        warm = torch.zeros(args.steps, NB, 4, 3, device=dev)
        dist.all_gather([torch.empty_like(warm) for _ in range(world)], warm)
    # everything with a rank-dependent host cost (NVML init, thread start, event creation) happens BEFORE the
    # barrier: between the barrier and e0.record() there is nothing but the record itself, so the ranks start
    # within the skew of one barrier and the end-of-run all-gather does not absorb start skew (round 1: N = 4)
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_mid = torch.cuda.Event(enable_timing=True)
    keep = []
    barrier()
    e0.record()
    for _ in range(args.steps):
        eng.forward_into(img, T, out)
        keep.append(out['gaze'].clone())
    e_mid.record()     # end of this rank's own forwards (reported per rank; the timed region goes on to e1)
    if world > 1:      # the one collective of the sharded test path: gather every rank's results
        mine = torch.stack(keep)
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
    e1.record()
    barrier()
    clocks = sampler.finish()
    ms_total = e0.elapsed_time(e1)
    rank_ms = None
    if world > 1:
        t = torch.tensor([ms_total, e0.elapsed_time(e_mid)], device=dev)
        all_t = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(all_t, t)
        per_rank = sorted(float(x[0].item()) for x in all_t)
        fwd = [float(x[1].item()) / args.steps for x in all_t]       # by rank: forwards only, before the all-gather
        rank_ms = {'min': per_rank[0] / args.steps, 'median': per_rank[len(per_rank) // 2] / args.steps,
                   'max': per_rank[-1] / args.steps, 'forwards_only_by_rank': fwd}
        ms_total = per_rank[-1]            # max over ranks
    ms_per_step = ms_total / args.steps
    value = world * CLIPS_PER_STEP * args.steps / (ms_total * 1e-3)
    # two more samples of the same K-step region (same barriers, same all-gather, max over ranks), reported beside the
    # headline sample: the region is ~0.2 s long and boxes / power states differ by a few per cent
    repeats_ms = [ms_per_step]
    for _ in range(2):
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        keep = []
        barrier()
        r0.record()
        for _ in range(args.steps):
            eng.forward_into(img, T, out)
            keep.append(out['gaze'].clone())
        if world > 1:
            mine = torch.stack(keep)
            dist.all_gather([torch.empty_like(mine) for _ in range(world)], mine)
        r1.record()
        barrier()
        t = torch.tensor([r0.elapsed_time(r1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        repeats_ms.append(float(t.item()) / args.steps)

    # ---------------- end to end through the host-buffer C-ABI calls (`e2e`) ----------------
    # every step: H2D copy of that step's 135 MB of pinned input, forward, D2H read of the results.
    # mcg_submit_host / mcg_wait_host keep two submissions in flight so the copy of step i+1 overlaps
    # the forward of step i (a DataLoader with pinned memory gives the reference the same overlap).
    host_imgs = [host_img, torch.randn(NB, 3, H, W, generator=g).pin_memory()]
    eng.set_graph_mode(not args.no_graph)
    for i in range(3):
        eng.forward_host(host_imgs[i % 2], clip_length=T)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(3, min(args.steps, 10))
    ticket = eng.submit_host(host_imgs[0], clip_length=T)
    for i in range(1, e2e_steps):
        nxt = eng.submit_host(host_imgs[i % 2], clip_length=T)
        res = eng.wait_host(ticket)
        ticket = nxt
    res = eng.wait_host(ticket)
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * CLIPS_PER_STEP * e2e_steps / e2e_s
    h2d = host_img.numel() * 4
    d2h = sum(v.numel() * 4 for v in res.values())

    # ---------------- dominant kernel roofline (tcgen05 GEMM), measured live with CUDA events ----------------
    peaks = load_peaks()
    eng.set_graph_mode(False)
    eng.set_option('time_kernels', 1)
    um_ms, um_fl, um_n = 0.0, 0.0, 0
    for _ in range(3):
        eng.forward_into(img, T, out)
        n_, fl_, ms_ = eng.umma_stats()
        um_ms, um_fl, um_n = um_ms + ms_, um_fl + fl_, um_n + n_
    eng.set_option('time_kernels', 0)
    achieved = um_fl / (um_ms * 1e-3) / 1e12 if um_ms > 0 else None
    step_tflops = value / world * GFLOP_PER_CLIP / 1e3
    roofline = {'bound': 'tensor', 'kernel': 'mcg::umma_gemm_kernel (tcgen05 implicit-GEMM conv / linear)',
                'achieved': achieved, 'peak': peaks['tflops_sustained'], 'unit': 'TFLOP/s',
                'frac': (achieved / peaks['tflops_sustained']) if achieved else None,
                # dram__bytes_read.sum + dram__bytes_write.sum of the largest launch of the step (FPN 3x3 on P2, 26 % of
                # the step's FLOPs) from the committed `ncu --set full` capture; its algorithmic bytes are 1259 MB
                'traffic': NCU_TRAFFIC.get(args.precision),
                'traffic_source': 'profiles/r02_ncu_summary.md (fpn0 launch, the largest of the step: 723 MB read + 505 MB written; algorithmic bytes 1259 MB)'
                if args.precision in NCU_TRAFFIC else None,
                'peak_source': peaks['source'] + ', bf16 dense sustained (kernel timed inside a long step)',
                'launches_per_step': um_n // 3, 'algorithmic_gflop_per_step': um_fl / 3 / 1e9,
                'kernel_ms_per_step': um_ms / 3, 'kernel_share_of_step': (um_ms / 3) / ms_per_step,
                'executed_flop_multiplier': {'fp16x3': 3, 'fp16c8': 2}.get(args.precision, 1),
                'executed_flop_multiplier_note': 'tensor-pipe time per algorithmic FLOP in fp16-MMA units: fp16x3 = 3 fp16 '
                                                 'MMAs, fp16c8 = 1 fp16 + 2 e4m3 MMAs at twice the rate',
                'step_achieved': step_tflops, 'step_frac': step_tflops / peaks['tflops_sustained'],
                'step_formula': 'clips_per_sec_per_gpu x 99.55 GFLOP/clip (SURVEY 8d)'}

    extras = {}
    cpu_baseline = None
    if not args.skip_extras:
        # BASELINE configs[3] on every rank (it contains the sharded run's collectives); never takes the headline down
        try:
            ts = testsplit_extra(eng, sd, local_rank, world)
        except Exception as e:
            ts = {'unavailable': f'{type(e).__name__}: {e}'}
        extras['testsplit'] = ts
    if rank == 0 and world == 1 and not args.skip_extras:
        try:
            extras.update(config_extras(eng, dev, not args.no_graph))
        except Exception as e:
            extras['l2cs_bs32_448'] = {'unavailable': f'{type(e).__name__}: {e}'}
        eng.set_option('time_kernels', 0)
        try:
            extras['preprocess'] = preprocess_extras(eng, dev, args.steps, not args.no_graph)
        except Exception as e:      # side measurement: never take the headline line down
            extras['preprocess'] = {'unavailable': f'{type(e).__name__}: {e}'}
        try:
            extras['png_decode'] = png_decode_extra(eng, sd, dev)
        except Exception as e:      # side measurement: never take the headline line down
            extras['png_decode'] = {'unavailable': f'{type(e).__name__}: {e}'}
        # the other precision modes on the same workload, for context: fp16 (fast) does NOT meet the 1e-3
        # (yaw,pitch) bar (~3e-3); fp16x3 and fp16c8 are the parity modes
        del eng
        torch.cuda.empty_cache()
        others = []
        for prec in [p for p in ('fp16x3', 'fp16c8', 'fp16') if p != args.precision]:
            fast = lib.Engine(sd, local_rank, prec)
            o2 = fast.forward(img, clip_length=T)
            fast.set_graph_mode(not args.no_graph)
            for _ in range(3):
                fast.forward_into(img, T, o2)
            torch.cuda.synchronize()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for _ in range(max(5, args.steps // 2)):
                fast.forward_into(img, T, o2)
            f1.record()
            torch.cuda.synchronize()
            fms = f0.elapsed_time(f1) / max(5, args.steps // 2)
            others.append({'precision': prec, 'value': CLIPS_PER_STEP / (fms * 1e-3), 'unit': 'clips/s', 'ms_per_step': fms})
            fast.close()
            del fast, o2
            torch.cuda.empty_cache()
        extras['other_precision'] = {'modes': others,
                                     'note': 'fp16 = single-fp16 operands (fast, ~3e-3 rad vs the fp32 oracle); fp16x3 = split-'
                                             'fp16 operands, 3 MMAs; fp16c8 = fp16 + e4m3 correction MMAs (both <= 1e-3 rad)'}
        v, n, el, cores = cpu_reference_throughput(args.cpu_budget)
        cpu_baseline = {'value': v, 'unit': 'clips/s', 'cores': cores, 'kind': 'port',
                        'sample': f'{n} clips of 7x3x224x224 in {el:.1f} s, one clip per forward, fp32 torch CPU, '
                                  f'{cores} threads (oracle port of the reference forward; mmcv not installable)'}

    if rank == 0:
        line = {'metric': METRIC, 'value': value, 'unit': 'clips/s', 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': args.precision, 'data': 'synthetic',
                'ms_per_step_over_ranks': rank_ms, 'ms_per_step_samples': repeats_ms,
                'config': {'workload': 'multiclue_gaze_r50 Gaze360-setting inference, bs=32 clips x 7 frames x 224x224 '
                                       'per GPU (BASELINE configs[1])', 'clips_per_step_per_gpu': CLIPS_PER_STEP,
                           'clip_length': T, 'height': H, 'width': W, 'weights': 'seeded random (reference key layout)',
                           'precision_mode': args.precision, 'cuda_graph': not args.no_graph,
                           'parallelism': f'{world} independent replicas over sharded clips, one all-gather of results',
                           'l2': 'per-step input (135 MB) and ~12 GB of activations exceed the 126 MB L2; no explicit flush'},
                'clocks': clocks,
                'e2e': {'value': e2e_value, 'unit': 'clips/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                        'steps': e2e_steps, 'api': 'mcg_submit_host/mcg_wait_host, 2 in flight (pinned host input, H2D every step, results read back)'},
                'gpu_launches': launches_per_step * args.steps,
                'roofline': roofline}
        if cpu_baseline:
            line['cpu_baseline'] = cpu_baseline
        line.update(extras)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
