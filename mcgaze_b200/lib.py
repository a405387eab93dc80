"""ctypes binding of the C ABI declared in include/mcgaze_b200.h.

There is NO CPU fallback: if the shared library is missing, or no CUDA device is visible,
every compute entry point raises.  torch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Optional, Sequence

HERE = os.path.dirname(os.path.abspath(__file__))
# MCG_LIB_PATH: A/B experiments against another build of the same C ABI (tools/ only)
LIB_PATH = os.environ.get('MCG_LIB_PATH') or os.path.join(HERE, 'libmcgaze_b200.so')

PRECISIONS = {'fp16x3': 0, 'fp16': 1, 'simt': 2, 'fp16c8': 3}

# every symbol include/mcgaze_b200.h declares
EXPORTS = ('mcg_create', 'mcg_destroy', 'mcg_forward', 'mcg_forward_host', 'mcg_submit_host', 'mcg_wait_host',
           'mcg_get_intermediate',
           'mcg_last_launch_count', 'mcg_range_report', 'mcg_last_umma_stats', 'mcg_last_umma_times', 'mcg_last_kernel_profile', 'mcg_set_graph_mode', 'mcg_set_option', 'mcg_debug_conv',
           'mcg_preprocess', 'mcg_png_parse', 'mcg_png_file_sizes', 'mcg_png_stage_files', 'mcg_png_decode', 'mcg_merge_clips', 'mcg_gaze_error', 'mcg_last_error', 'mcg_version')


class McgError(RuntimeError):
    pass


class mcg_tensor(ctypes.Structure):
    _fields_ = [('name', ctypes.c_char_p), ('data', ctypes.c_void_p), ('ndim', ctypes.c_int),
                ('shape', ctypes.c_int64 * 4)]


class mcg_frame(ctypes.Structure):
    _fields_ = [('src', ctypes.c_void_p), ('src_stride', ctypes.c_int64), ('src_h', ctypes.c_int32),
                ('src_w', ctypes.c_int32), ('crop_y', ctypes.c_int32), ('crop_x', ctypes.c_int32),
                ('crop_h', ctypes.c_int32), ('crop_w', ctypes.c_int32), ('dst_h', ctypes.c_int32),
                ('dst_w', ctypes.c_int32)]


class mcg_png_info(ctypes.Structure):
    _fields_ = [('width', ctypes.c_int32), ('height', ctypes.c_int32), ('bit_depth', ctypes.c_int32),
                ('color_type', ctypes.c_int32), ('interlace', ctypes.c_int32), ('channels', ctypes.c_int32),
                ('has_palette', ctypes.c_int32), ('supported', ctypes.c_int32), ('idat_bytes', ctypes.c_int64),
                ('palette', ctypes.c_uint8 * 768)]


class mcg_png_job(ctypes.Structure):
    _fields_ = [('zdata', ctypes.c_void_p), ('zbytes', ctypes.c_int64), ('width', ctypes.c_int32),
                ('height', ctypes.c_int32), ('color_type', ctypes.c_int32), ('reserved', ctypes.c_int32),
                ('palette', ctypes.c_void_p), ('scan', ctypes.c_void_p), ('dst', ctypes.c_void_p),
                ('dst_stride', ctypes.c_int64)]


_lib = None


def load_library() -> ctypes.CDLL:
    """Load libmcgaze_b200.so (built in-tree by `python -m mcgaze_b200.build`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise McgError(f'{LIB_PATH} not found: build it with `python -m mcgaze_b200.build` '
                       '(there is no CPU fallback)')
    lib = ctypes.CDLL(LIB_PATH)
    vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_float)
    lib.mcg_create.argtypes = [ctypes.POINTER(vp), ci, ctypes.POINTER(mcg_tensor), ci, ci]
    lib.mcg_destroy.argtypes = [vp]
    lib.mcg_forward.argtypes = [vp, vp, ci, ci, ci, ci, vp, vp, vp, vp, vp, vp]
    lib.mcg_forward_host.argtypes = [vp, vp, ci, ci, ci, ci, vp, vp, vp, vp, vp]
    lib.mcg_submit_host.argtypes = [vp, vp, ci, ci, ci, ci, vp, vp, ctypes.POINTER(ci)]
    lib.mcg_wait_host.argtypes = [vp, ci, vp, vp, vp]
    lib.mcg_get_intermediate.argtypes = [vp, ctypes.c_char_p, vp, ctypes.c_int64, ctypes.POINTER(ctypes.c_int64)]
    lib.mcg_last_launch_count.argtypes = [vp]
    lib.mcg_set_graph_mode.argtypes = [vp, ci]
    lib.mcg_last_umma_stats.argtypes = [vp, ctypes.POINTER(ctypes.c_double)]
    lib.mcg_last_umma_times.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ci]
    lib.mcg_last_kernel_profile.argtypes = [vp, ctypes.c_char_p, ci]
    lib.mcg_set_option.argtypes = [vp, ctypes.c_char_p, ci]
    lib.mcg_range_report.argtypes = [vp, ctypes.c_char_p, ci]
    lib.mcg_preprocess.argtypes = [ctypes.POINTER(mcg_frame), ci, cf, cf, ci, vp, ci, ci, vp]
    lib.mcg_png_parse.argtypes = [vp, ctypes.c_int64, ci, ctypes.POINTER(mcg_png_info), vp, ctypes.c_int64]
    lib.mcg_png_file_sizes.argtypes = [ctypes.POINTER(ctypes.c_char_p), ci, vp]
    lib.mcg_png_stage_files.argtypes = [ctypes.POINTER(ctypes.c_char_p), ci, ci, ci, vp, vp, vp, ctypes.POINTER(mcg_png_info), vp]
    lib.mcg_png_decode.argtypes = [ctypes.POINTER(mcg_png_job), ci, vp, vp]
    lib.mcg_gaze_error.argtypes = [vp, vp, vp, ci, ci, vp, vp]
    lib.mcg_merge_clips.argtypes = [vp, vp, vp, ci, ci, ci, vp, vp, vp]
    lib.mcg_debug_conv.argtypes = [ci, vp, ci, ci, ci, ci, vp, ci, ci, ci, ci, ci, vp, vp, ci, ci, ci, ci, ci, vp, vp]
    for name in EXPORTS:
        fn = getattr(lib, name)
        fn.restype = ctypes.c_char_p if name in ('mcg_last_error', 'mcg_version') else ci
    del cf
    _lib = lib
    return lib


def _check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load_library().mcg_last_error()
        raise McgError(f'{what} failed (code {rc}): {msg.decode() if msg else ""}')


_FRAME_DTYPE = None


def preprocess(frames, geometry, mean, std, to_rgb, out, stream: Optional[int] = None) -> None:
    """mcg_preprocess: `frames` = CUDA uint8 HWC tensors (decoded BGR frames; a list, or ONE [n, h, w, 3] tensor),
    `geometry` = per frame (crop_y, crop_x, crop_h, crop_w, dst_h, dst_w), `out` = CUDA fp32 [n, 3, Hp, Wp].
    Asynchronous on `stream` (default: torch's current stream)."""
    import numpy as np
    import torch
    global _FRAME_DTYPE
    if not torch.cuda.is_available():
        raise McgError('mcgaze_b200 needs a CUDA device (B200, sm_100a); there is no CPU path')
    lib = load_library()
    if _FRAME_DTYPE is None:
        _FRAME_DTYPE = np.dtype([('src', np.uint64), ('src_stride', np.int64), ('src_h', np.int32), ('src_w', np.int32),
                                 ('geo', np.int32, (6,))])
        assert _FRAME_DTYPE.itemsize == ctypes.sizeof(mcg_frame)
    n = len(frames)
    geo = np.asarray(geometry, dtype=np.int32)
    if n == 0 or geo.shape != (n, 6):
        raise McgError('preprocess: need one (crop_y, crop_x, crop_h, crop_w, dst_h, dst_w) per frame')
    if out.dtype != torch.float32 or not out.is_cuda or not out.is_contiguous() or out.dim() != 4 or out.shape[0] != n \
            or out.shape[1] != 3:
        raise McgError('preprocess: `out` must be a contiguous CUDA fp32 tensor [n, 3, Hp, Wp]')
    desc = np.zeros(n, dtype=_FRAME_DTYPE)
    bad = 'preprocess: frames must be CUDA uint8 [h, w, 3] tensors with packed pixels'
    if hasattr(frames, 'data_ptr'):               # one batched tensor: descriptors without a python loop
        f = frames
        if f.dtype != torch.uint8 or not f.is_cuda or f.dim() != 4 or f.shape[3] != 3 or f.stride(3) != 1 or f.stride(2) != 3:
            raise McgError(bad)
        desc['src'] = f.data_ptr() + np.arange(n, dtype=np.uint64) * np.uint64(f.stride(0))
        desc['src_stride'], desc['src_h'], desc['src_w'] = f.stride(1), f.shape[1], f.shape[2]
    else:
        for f in frames:
            if f.dtype != torch.uint8 or not f.is_cuda or f.dim() != 3 or f.shape[2] != 3 or f.stride(2) != 1 or f.stride(1) != 3:
                raise McgError(bad)
        desc['src'] = [f.data_ptr() for f in frames]
        desc['src_stride'] = [f.stride(0) for f in frames]
        desc['src_h'] = [f.shape[0] for f in frames]
        desc['src_w'] = [f.shape[1] for f in frames]
    desc['geo'] = geo
    c3 = ctypes.c_float * 3
    st = torch.cuda.current_stream().cuda_stream if stream is None else stream
    _check(lib.mcg_preprocess(ctypes.cast(desc.ctypes.data, ctypes.POINTER(mcg_frame)), n,
                              c3(*[float(v) for v in mean]), c3(*[float(v) for v in std]), 1 if to_rgb else 0,
                              out.data_ptr(), out.shape[2], out.shape[3], st), 'mcg_preprocess')


SCORER_VARIANTS = {'gaze360': 0, 'l2cs': 1}


def gaze_error_sums(pred, gt, lengths, variant: str = 'gaze360', stream: Optional[int] = None):
    """mcg_gaze_error, raw form: -> CUDA float64 [6] = {sum of (per-video mean error in degrees x frames), frames} for
    360 / front-180 / front-20.  The sums of disjoint sets of videos add up, so a sharded run needs one all-reduce of
    these six numbers.  `pred`, `gt` = CUDA fp32 [F, 3], `lengths` = frames per video."""
    import numpy as np
    import torch
    if not torch.cuda.is_available():
        raise McgError('mcgaze_b200 needs a CUDA device (B200, sm_100a); there is no CPU path')
    if variant not in SCORER_VARIANTS:
        raise McgError(f'unknown scorer variant {variant!r}; choose from {sorted(SCORER_VARIANTS)}')
    lib = load_library()
    lengths = np.asarray(lengths, dtype=np.int64)
    for t in (pred, gt):
        if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous() or t.dim() != 2 or t.shape[1] != 3 \
                or t.shape[0] != int(lengths.sum()):
            raise McgError('gaze_error: pred / gt must be contiguous CUDA fp32 [sum(lengths), 3] tensors')
    if len(lengths) == 0 or np.any(lengths <= 0):
        raise McgError('gaze_error: every video needs at least one frame')
    start = torch.from_numpy(np.concatenate([[0], np.cumsum(lengths)]).astype(np.int32)).to(pred.device)
    out = torch.empty(6, dtype=torch.float64, device=pred.device)
    st = torch.cuda.current_stream(pred.device).cuda_stream if stream is None else stream
    with torch.cuda.device(pred.device):
        _check(lib.mcg_gaze_error(pred.data_ptr(), gt.data_ptr(), start.data_ptr(), len(lengths), SCORER_VARIANTS[variant],
                                  out.data_ptr(), st), 'mcg_gaze_error')
    return out


def sums_to_mae(sums) -> Dict[str, float]:
    o = [float(v) for v in sums]
    res = {}
    for k, name in enumerate(('360', 'front90', 'front20')):
        res[f'mae_{name}'] = float(o[2 * k] / max(o[2 * k + 1], 1.0))
        res[f'frames_{name}'] = int(o[2 * k + 1])
    return res


def gaze_error(pred, gt, lengths, variant: str = 'gaze360', stream: Optional[int] = None) -> Dict[str, float]:
    """mcg_gaze_error: `pred`, `gt` = CUDA fp32 [F, 3] (frames of all videos concatenated), `lengths` = frames per
    video.  Returns the keys of mcgaze_b200.metric.gaze_error (mae_360 / mae_front90 / mae_front20 + frame counts)."""
    return sums_to_mae(gaze_error_sums(pred, gt, lengths, variant, stream).cpu().numpy())


def merge_clips(rows, clips_per_video, frames_per_video, clip_len: int = 7, stride: int = 4, stream: Optional[int] = None):
    """mcg_merge_clips: `rows` = CUDA fp32 [n_clips, clip_len, 27] (per clip and frame: boxes [3,4], scores [3], gaze
    [4,3]; clips of a video consecutive, short clips padded), `clips_per_video` / `frames_per_video` = per-video counts.
    -> (det CUDA fp32 [F, 3, 5], gaze CUDA fp32 [F, 4, 3]) with the reference's overlap merge applied."""
    import numpy as np
    import torch
    if not torch.cuda.is_available():
        raise McgError('mcgaze_b200 needs a CUDA device (B200, sm_100a); there is no CPU path')
    lib = load_library()
    cpv = np.asarray(clips_per_video, dtype=np.int64)
    fpv = np.asarray(frames_per_video, dtype=np.int64)
    if not rows.is_cuda or rows.dtype != torch.float32 or not rows.is_contiguous() or rows.dim() != 3 \
            or rows.shape[0] != int(cpv.sum()) or rows.shape[1] != clip_len or rows.shape[2] != 27:
        raise McgError('merge_clips: rows must be a contiguous CUDA fp32 tensor [sum(clips_per_video), clip_len, 27]')
    if len(cpv) == 0 or len(cpv) != len(fpv) or np.any(cpv <= 0) or np.any(fpv <= 0):
        raise McgError('merge_clips: every video needs at least one clip and one frame')
    dev = rows.device
    cs = torch.from_numpy(np.concatenate([[0], np.cumsum(cpv)]).astype(np.int32)).to(dev)
    fs = torch.from_numpy(np.concatenate([[0], np.cumsum(fpv)]).astype(np.int32)).to(dev)
    F = int(fpv.sum())
    det = torch.empty(F, 3, 5, dtype=torch.float32, device=dev)
    gaze = torch.empty(F, 4, 3, dtype=torch.float32, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream if stream is None else stream
    with torch.cuda.device(dev):
        _check(lib.mcg_merge_clips(rows.data_ptr(), cs.data_ptr(), fs.data_ptr(), len(cpv), clip_len, stride, det.data_ptr(),
                                   gaze.data_ptr(), st), 'mcg_merge_clips')
    return det, gaze


class Engine:
    """Owns one `mcg_handle`: packed weights + workspace on one device.

    `state_dict` is a reference-layout checkpoint dict (name -> CPU fp32 tensor / ndarray),
    exactly what `mmcv.runner.load_checkpoint` would hand to the reference's `MultiClueGaze`.
    """

    def __init__(self, state_dict: Dict[str, object], device: int = 0, precision: str = 'fp16c8'):
        import numpy as np
        import torch
        if not torch.cuda.is_available():
            raise McgError('mcgaze_b200 needs a CUDA device (B200, sm_100a); there is no CPU path')
        if precision not in PRECISIONS:
            raise McgError(f'unknown precision {precision!r}; choose from {sorted(PRECISIONS)}')
        self._lib = load_library()
        self.device = int(device)
        self.precision = precision
        keep, arr = [], []
        for k, v in state_dict.items():
            if hasattr(v, 'detach'):
                if not v.is_floating_point():
                    continue
                v = v.detach().to('cpu', torch.float32).contiguous().numpy()
            v = np.ascontiguousarray(v, dtype=np.float32)
            if v.ndim > 4:
                continue
            keep.append(v)
            t = mcg_tensor()
            t.name = k.encode()
            t.data = v.ctypes.data
            t.ndim = v.ndim
            for i in range(v.ndim):
                t.shape[i] = v.shape[i]
            arr.append(t)
        tensors = (mcg_tensor * len(arr))(*arr)
        h = ctypes.c_void_p()
        _check(self._lib.mcg_create(ctypes.byref(h), self.device, tensors, len(arr), PRECISIONS[precision]),
               'mcg_create')
        self._h = h
        del keep

    def close(self) -> None:
        if getattr(self, '_h', None):
            self._lib.mcg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ options
    def set_graph_mode(self, on: bool) -> None:
        _check(self._lib.mcg_set_graph_mode(self._h, int(bool(on))), 'mcg_set_graph_mode')

    def set_option(self, key: str, value: int) -> None:
        _check(self._lib.mcg_set_option(self._h, key.encode(), int(value)), 'mcg_set_option')

    @property
    def last_launch_count(self) -> int:
        return int(self._lib.mcg_last_launch_count(self._h))

    def umma_stats(self):
        """(launches, algorithmic FLOPs, summed device ms) of the tcgen05 GEMMs of the last eager forward."""
        out = (ctypes.c_double * 3)()
        _check(self._lib.mcg_last_umma_stats(self._h, out), 'mcg_last_umma_stats')
        return int(out[0]), float(out[1]), float(out[2])

    def umma_times(self):
        """Per-launch device times (ms) of the tcgen05 GEMMs of the last eager forward (option time_kernels)."""
        buf = (ctypes.c_double * 256)()
        n = self._lib.mcg_last_umma_times(self._h, buf, 256)
        if n < 0:
            _check(n, 'mcg_last_umma_times')
        return [float(buf[i]) for i in range(n)]

    def kernel_profile(self):
        """[(kernel name, ms)] for every launch of the last eager forward (option time_kernels)."""
        n = self._lib.mcg_last_kernel_profile(self._h, None, 0)
        if n < 0:
            _check(n, 'mcg_last_kernel_profile')
        buf = ctypes.create_string_buffer(n)
        _check(min(self._lib.mcg_last_kernel_profile(self._h, buf, n), 0), 'mcg_last_kernel_profile')
        return [(l.split('\t')[0], float(l.split('\t')[1])) for l in buf.value.decode().splitlines() if l]

    def range_report(self):
        """mcg_range_report: one dict per trunk / FPN activation of the last forward and per convolution weight
        (name, total, nonzero, over, under, nonfinite, maxabs, energy, energy_under)."""
        n = self._lib.mcg_range_report(self._h, None, 0)
        if n < 0:
            _check(n, 'mcg_range_report')
        buf = ctypes.create_string_buffer(n)
        _check(min(self._lib.mcg_range_report(self._h, buf, n), 0), 'mcg_range_report')
        keys = ('total', 'nonzero', 'over', 'under', 'nonfinite')
        rows = []
        for line in buf.value.decode().splitlines():
            f = line.split('\t')
            row = {'name': f[0]}
            row.update({k: int(v) for k, v in zip(keys, f[1:6])})
            row.update(maxabs=float(f[6]), energy=float(f[7]), energy_under=float(f[8]))
            rows.append(row)
        return rows

    def check_ranges(self, max_over_frac: float = 1e-6, max_under_energy: float = 0.25):
        """Refuse operands the fp16c8 corrections cannot represent: raises McgError when, in the last forward, a tensor
        holds non-finite values, more than `max_over_frac` of its elements saturate the e4m3 planes (|activation| > 448,
        |weight| > 28), or more than `max_under_energy` of an activation's energy sits in elements whose residue is an
        e4m3 subnormal (|v| < 2^-8: those are only fp16-accurate).  Returns the report otherwise.  The other parity mode
        (`fp16x3`) has no such window."""
        rows = self.range_report()
        if self.precision != 'fp16c8':
            return rows
        bad = []
        for r in rows:
            if r['nonfinite']:
                bad.append(f"{r['name']}: {r['nonfinite']} non-finite values")
            elif r['over'] > max_over_frac * max(r['total'], 1):
                bad.append(f"{r['name']}: {r['over']} of {r['total']} elements saturate e4m3 (max |v| = {r['maxabs']:.4g})")
            elif not r['name'].startswith('w:') and r['energy'] > 0 and r['energy_under'] > max_under_energy * r['energy']:
                bad.append(f"{r['name']}: {100 * r['energy_under'] / r['energy']:.1f} % of the energy below 2^-8 "
                           f"(max |v| = {r['maxabs']:.4g})")
        if bad:
            raise McgError('operands outside the fp16c8 correction window, use precision="fp16x3": ' + '; '.join(bad[:6]))
        return rows

    # ------------------------------------------------------------------ forward
    @staticmethod
    def _meta(arr, n: int, width: int):
        import numpy as np
        if arr is None:
            return None, None
        a = np.ascontiguousarray(np.asarray(arr, dtype=np.float32).reshape(n, width))
        return a, a.ctypes.data

    def _check_io(self, img, out=None) -> None:
        """Raw pointers cross the C ABI: refuse anything that is not a contiguous fp32 tensor on this engine's GPU."""
        import torch
        if not (img.is_cuda and img.dtype == torch.float32 and img.dim() == 4 and img.shape[1] == 3 and img.is_contiguous()):
            raise McgError('img must be a contiguous CUDA fp32 tensor [N, 3, H, W]')
        if img.device.index != self.device:
            raise McgError(f'img lives on cuda:{img.device.index}, the engine on cuda:{self.device}')
        if out is not None:
            N = img.shape[0]
            for k, shape in (('gaze', (N, 4, 3)), ('boxes', (N, 3, 4)), ('scores', (N, 3))):
                t = out[k]
                if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == shape
                        and t.device.index == self.device):
                    raise McgError(f"out['{k}'] must be a contiguous CUDA fp32 tensor {shape} on cuda:{self.device}")

    def forward(self, img, clip_length: Optional[int] = None, img_hw=None, scale_factor=None):
        """img: CUDA fp32 [N,3,H,W] (N = B*clip_length).  Returns CUDA tensors
        {'gaze' [N,4,3] (fused, face, eyes, head), 'boxes' [N,3,4], 'scores' [N,3]}; asynchronous on the
        current torch stream."""
        import torch
        assert img.is_cuda and img.dtype == torch.float32 and img.dim() == 4 and img.shape[1] == 3
        img = img.contiguous()
        self._check_io(img)
        N, _, H, W = img.shape
        T = N if clip_length is None else int(clip_length)
        assert N % T == 0
        gaze = torch.empty(N, 4, 3, device=img.device, dtype=torch.float32)
        boxes = torch.empty(N, 3, 4, device=img.device, dtype=torch.float32)
        scores = torch.empty(N, 3, device=img.device, dtype=torch.float32)
        k1, p1 = self._meta(img_hw, N, 2)
        k2, p2 = self._meta(scale_factor, N, 4)
        stream = torch.cuda.current_stream(img.device).cuda_stream
        _check(self._lib.mcg_forward(self._h, img.data_ptr(), N // T, T, H, W, p1, p2, gaze.data_ptr(),
                                     boxes.data_ptr(), scores.data_ptr(), ctypes.c_void_p(stream)), 'mcg_forward')
        del k1, k2
        return {'gaze': gaze, 'boxes': boxes, 'scores': scores}

    def forward_into(self, img, clip_length: int, out, img_hw=None, scale_factor=None):
        """Like forward() but writes into the preallocated dict `out` (stable pointers: lets the
        CUDA-graph replay path engage)."""
        import torch
        self._check_io(img, out)
        N, _, H, W = img.shape
        T = int(clip_length)
        if N % T:
            raise McgError(f'forward_into: {N} frames are not a whole number of {T}-frame clips')
        k1, p1 = self._meta(img_hw, N, 2)
        k2, p2 = self._meta(scale_factor, N, 4)
        stream = torch.cuda.current_stream(img.device).cuda_stream
        _check(self._lib.mcg_forward(self._h, img.data_ptr(), N // T, T, H, W, p1, p2, out['gaze'].data_ptr(),
                                     out['boxes'].data_ptr(), out['scores'].data_ptr(), ctypes.c_void_p(stream)),
               'mcg_forward')
        del k1, k2
        return out

    def forward_host(self, img, clip_length: Optional[int] = None, img_hw=None, scale_factor=None):
        """img: CPU fp32 [N,3,H,W] (pinned memory gives async copies).  Host-to-device copy, forward,
        device-to-host copy and a stream sync all happen inside the C call.  Returns CPU tensors."""
        import torch
        assert (not img.is_cuda) and img.dtype == torch.float32 and img.dim() == 4
        img = img.contiguous()
        N, _, H, W = img.shape
        T = N if clip_length is None else int(clip_length)
        gaze = torch.empty(N, 4, 3, dtype=torch.float32)
        boxes = torch.empty(N, 3, 4, dtype=torch.float32)
        scores = torch.empty(N, 3, dtype=torch.float32)
        k1, p1 = self._meta(img_hw, N, 2)
        k2, p2 = self._meta(scale_factor, N, 4)
        _check(self._lib.mcg_forward_host(self._h, img.data_ptr(), N // T, T, H, W, p1, p2, gaze.data_ptr(),
                                          boxes.data_ptr(), scores.data_ptr()), 'mcg_forward_host')
        del k1, k2
        return {'gaze': gaze, 'boxes': boxes, 'scores': scores}

    def submit_host(self, img, clip_length: Optional[int] = None, img_hw=None, scale_factor=None):
        """Pipelined host entry: enqueue H2D copy + forward + read-back, return a ticket for wait_host().
        Keep `img` (CPU fp32, ideally pinned) alive until the matching wait_host()."""
        import torch
        assert (not img.is_cuda) and img.dtype == torch.float32 and img.dim() == 4 and img.is_contiguous()
        N, _, H, W = img.shape
        T = N if clip_length is None else int(clip_length)
        k1, p1 = self._meta(img_hw, N, 2)
        k2, p2 = self._meta(scale_factor, N, 4)
        ticket = ctypes.c_int(-1)
        _check(self._lib.mcg_submit_host(self._h, img.data_ptr(), N // T, T, H, W, p1, p2, ctypes.byref(ticket)),
               'mcg_submit_host')
        del k1, k2
        return (ticket.value, N, img)

    def wait_host(self, ticket):
        import torch
        t, N, _keep = ticket
        gaze = torch.empty(N, 4, 3, dtype=torch.float32)
        boxes = torch.empty(N, 3, 4, dtype=torch.float32)
        scores = torch.empty(N, 3, dtype=torch.float32)
        _check(self._lib.mcg_wait_host(self._h, t, gaze.data_ptr(), boxes.data_ptr(), scores.data_ptr()),
               'mcg_wait_host')
        return {'gaze': gaze, 'boxes': boxes, 'scores': scores}

    def intermediate(self, name: str, max_elems: int = 1 << 28):
        import torch
        buf = torch.empty(max_elems, device=f'cuda:{self.device}', dtype=torch.float32)
        shape = (ctypes.c_int64 * 4)()
        _check(self._lib.mcg_get_intermediate(self._h, name.encode(), buf.data_ptr(), max_elems, shape),
               f'mcg_get_intermediate({name})')
        dims = [int(s) for s in shape if int(s) > 0]
        n = 1
        for d in dims:
            n *= d
        return buf[:n].reshape(dims).clone()


def debug_conv(engine: str, x, w, stride: int, pad: int, bias=None, res=None, res_mode: int = 0, relu: bool = False,
               force_im2col: bool = False, force_block_n: int = 0, out_mode: int = 0):
    """Kernel-level entry: x CUDA fp32 NCHW, w CUDA fp32 [Cout,Cin,R,S] -> CUDA fp32 NCHW (torch layout in/out;
    the NHWC / (r,s,c) repack the C ABI wants happens here)."""
    import torch
    lib = load_library()
    NB, C, H, W = x.shape
    Cout, _, R, S = w.shape
    P = (H + 2 * pad - R) // stride + 1
    Q = (W + 2 * pad - S) // stride + 1
    x_nhwc = x.permute(0, 2, 3, 1).contiguous()
    w_p = w.permute(0, 2, 3, 1).contiguous().reshape(Cout, R * S * C)
    res_nhwc = res.permute(0, 2, 3, 1).contiguous() if res is not None else None
    out = torch.empty(NB, P, Q, Cout, device=x.device, dtype=torch.float32)
    stream = torch.cuda.current_stream(x.device).cuda_stream
    rc = lib.mcg_debug_conv(PRECISIONS[engine], x_nhwc.data_ptr(), NB, H, W, C, w_p.data_ptr(), Cout, R, S, stride, pad,
                            bias.data_ptr() if bias is not None else None,
                            res_nhwc.data_ptr() if res_nhwc is not None else None, res_mode, int(relu),
                            int(force_im2col), force_block_n, int(out_mode), out.data_ptr(), ctypes.c_void_p(stream))
    _check(rc, 'mcg_debug_conv')
    return out.permute(0, 3, 1, 2).contiguous()
