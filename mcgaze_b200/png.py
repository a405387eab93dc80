"""PNG frames decoded on the GPU (the "decode" in front of crop / resize / normalise).

`GpuPngDecoder.decode(files)` returns, for every PNG file, the CUDA uint8 [h, w, 3] BGR tensor that
`LoadImageFromFile` (mmdet/datasets/pipelines/loading.py:58-69 -> mmcv.imfrombytes -> cv2.imdecode(IMREAD_COLOR)) would
have produced on the host - bit for bit - ready to be handed to `GpuTestPipeline.batch`.  The host only reads the files
and walks their chunk lists (`mcg_png_parse`: signature, IHDR, PLTE, IDAT payloads, chunk CRCs); the compressed bytes
cross PCIe in ONE copy per batch (about half of the decoded size) and `mcg_png_decode` inflates and reconstructs the
scanlines on the device, one warp per image.

Two phases, so that a loader thread can do the host part of batch k+1 while batch k runs:
    staged = decoder.stage(files)        # host: read + parse into one pinned block (thread safe, no CUDA calls but the
                                         #       pinned allocation)
    frames, status = decoder.launch(staged)   # H2D copy + kernels on the current stream; `status` is a CUDA int32 [n]
    decoder.check(staged, status)        # synchronises; raises on a corrupt stream

Images `mcg_png_decode` does not take (16-bit samples, 1/2/4-bit samples, Adam7 interlace) raise `UnsupportedPng` in
stage(); the callers in evaluate.py then decode THAT file with cv2 on the host, as the reference does for every file.
There is no CPU fallback for the decode itself: without the CUDA library / a GPU launch() raises.
"""
from __future__ import annotations

import ctypes
import os
from typing import Any, List, Optional, Sequence, Tuple, Union

import numpy as np

from . import lib

FileLike = Union[str, os.PathLike, bytes, bytearray, memoryview, np.ndarray]
_ALIGN = 16
_CHANNELS = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}
STATUS_NAMES = ('ok', 'bad zlib header', 'reserved block type', 'bad stored-block length', 'bad code lengths',
                'bad symbol', 'match distance before the start of the data', 'more data than the image holds',
                'compressed stream ends early', 'less data than the image holds', 'bad filter type', 'bad job descriptor',
                'Adler-32 of the decoded data does not match the stream trailer')


class UnsupportedPng(lib.McgError):
    """a file mcg_png_decode does not take: another image format, or a PNG with bit depth != 8 / Adam7 interlace"""


def _up(n: int) -> int:
    return (n + _ALIGN - 1) // _ALIGN * _ALIGN


class StagedPngs:
    """host side of one batch: the pinned block (zlib streams + palettes + room for the job table at `used`), per-image
    header fields and offsets"""
    __slots__ = ('block', 'used', 'infos', 'zoff', 'zlen', 'paloff', 'names')

    def __init__(self, block, used, infos, zoff, zlen, paloff, names):
        self.block, self.used, self.infos, self.zoff, self.zlen, self.paloff, self.names = block, used, infos, zoff, zlen, paloff, names

    def __len__(self) -> int:
        return len(self.infos)

    @property
    def shapes(self) -> List[Tuple[int, int]]:
        return [(h, w) for (w, h, _) in self.infos]


class GpuPngDecoder:
    def __init__(self, device: int = 0, check_crc: bool = True):
        self.device = int(device)
        self.check_crc = bool(check_crc)
        self._job_dtype = np.dtype([('zdata', np.uint64), ('zbytes', np.int64), ('width', np.int32), ('height', np.int32),
                                    ('color_type', np.int32), ('reserved', np.int32), ('palette', np.uint64),
                                    ('scan', np.uint64), ('dst', np.uint64), ('dst_stride', np.int64)])
        assert self._job_dtype.itemsize == ctypes.sizeof(lib.mcg_png_job)

    # ------------------------------------------------------------------------------------------------ host phase
    @staticmethod
    def _bytes(f: FileLike) -> np.ndarray:
        if isinstance(f, (str, os.PathLike)):
            return np.fromfile(f, dtype=np.uint8)
        if isinstance(f, np.ndarray):
            return np.ascontiguousarray(f, dtype=np.uint8).reshape(-1)
        return np.frombuffer(f, dtype=np.uint8)

    def _stage_paths(self, paths: Sequence[str], threads: int) -> StagedPngs:
        """all inputs are file names: sizes, layout, then ONE C call maps, parses and copies every file on its own worker
        threads (mcg_png_stage_files) - no per-file python work"""
        import torch
        so = lib.load_library()
        n = len(paths)
        cpaths = (ctypes.c_char_p * n)(*[os.fsencode(p) for p in paths])
        sizes = np.zeros(n, dtype=np.int64)
        lib._check(so.mcg_png_file_sizes(cpaths, n, sizes.ctypes.data), 'mcg_png_file_sizes')
        if (sizes < 0).any():
            raise FileNotFoundError(paths[int(np.nonzero(sizes < 0)[0][0])])
        slot = (sizes + _ALIGN - 1) // _ALIGN * _ALIGN
        offs = np.concatenate([[0], np.cumsum(slot + 768)]).astype(np.int64)
        jobs_off = (int(offs[-1]) + 63) // 64 * 64
        block = torch.empty(jobs_off + 64 * n, dtype=torch.uint8, pin_memory=torch.cuda.is_available())
        base = block.numpy()
        infos = (lib.mcg_png_info * n)()
        results = np.zeros(n, dtype=np.int32)
        rc = so.mcg_png_stage_files(cpaths, n, 1 if self.check_crc else 0, max(1, int(threads)), offs.ctypes.data, sizes.ctypes.data,
                                    base.ctypes.data, infos, results.ctypes.data)
        if rc != 0:
            msg = (so.mcg_last_error() or b'mcg_png_stage_files failed').decode()
            first = int(results[np.nonzero(results)[0][0]]) if results.any() else 0
            if first == 1:
                raise FileNotFoundError(msg)
            raise (UnsupportedPng if first == 3 or self._not_png(paths[int(np.nonzero(results)[0][0])]) else lib.McgError)(msg)
        arr = np.frombuffer(infos, dtype=np.uint8).reshape(n, ctypes.sizeof(lib.mcg_png_info))
        head = arr[:, :32].copy().view(np.int32)                     # width, height, bit_depth, color_type, ...
        zlen = arr[:, 32:40].copy().view(np.int64).reshape(n)
        paloff = np.full(n, -1, dtype=np.int64)
        for i in np.nonzero(head[:, 3] == 3)[0]:
            paloff[i] = int(offs[i] + slot[i])
            base[paloff[i]:paloff[i] + 768] = arr[i, 40:808]
        triples = [(int(w), int(h), int(c)) for w, h, c in zip(head[:, 0], head[:, 1], head[:, 3])]
        return StagedPngs(block, jobs_off, triples, offs[:-1].copy(), zlen, paloff, [str(p) for p in paths])

    @staticmethod
    def _not_png(path) -> bool:
        try:
            with open(path, 'rb') as fh:
                return fh.read(8) != b'\x89PNG\r\n\x1a\n'
        except OSError:
            return False

    def stage(self, files: Sequence[FileLike], pool=None) -> StagedPngs:
        """Read and parse `files` (paths or file images).  The IDAT payloads land in one pinned host block (pageable
        without a GPU).  `pool`: optional ThreadPoolExecutor (its size = the worker threads used)."""
        import torch
        so = lib.load_library()
        if len(files) and all(isinstance(f, (str, os.PathLike)) for f in files):
            return self._stage_paths([os.fspath(f) for f in files], getattr(pool, '_max_workers', 1) if pool is not None else 1)
        datas = [self._bytes(f) for f in files] if pool is None else list(pool.map(self._bytes, files))
        n = len(datas)
        if n == 0:
            raise lib.McgError('no files to decode')
        for i, d in enumerate(datas):
            if d.size < 8 or d[:8].tobytes() != b'\x89PNG\r\n\x1a\n':
                raise UnsupportedPng(f'{files[i] if isinstance(files[i], (str, os.PathLike)) else i}: not a PNG file')
        # a file's IDAT payload is shorter than the file: per-file slots of the file size (+ palette) always fit
        offs = np.zeros(n + 1, dtype=np.int64)
        for i, d in enumerate(datas):
            offs[i + 1] = offs[i] + _up(int(d.size)) + 768
        # ... and the job table behind them (64-byte descriptors, filled by launch() once the device addresses exist): the
        # kernels read it from HBM, so it travels in the same copy
        jobs_off = (int(offs[-1]) + 63) // 64 * 64
        block = torch.empty(jobs_off + 64 * n, dtype=torch.uint8, pin_memory=torch.cuda.is_available())
        base = block.numpy()
        names = [str(f) if isinstance(f, (str, os.PathLike)) else f'<image {i}>' for i, f in enumerate(files)]

        def one(i: int):
            d = datas[i]
            info = lib.mcg_png_info()
            rc = so.mcg_png_parse(d.ctypes.data, int(d.size), 1 if self.check_crc else 0, ctypes.byref(info),
                                  base.ctypes.data + int(offs[i]), int(d.size))
            if rc != 0:
                msg = so.mcg_last_error()
                raise lib.McgError(f'{names[i]}: {msg.decode() if msg else "mcg_png_parse failed"}')
            if not info.supported:
                raise UnsupportedPng(f'{names[i]}: bit depth {info.bit_depth}, interlace {info.interlace} is not decoded on the GPU')
            pal = -1
            if info.color_type == 3:
                pal = int(offs[i]) + _up(int(d.size))
                base[pal:pal + 768] = np.frombuffer(bytes(info.palette), dtype=np.uint8)
            return (int(info.width), int(info.height), int(info.color_type)), int(info.idat_bytes), pal

        res = [one(i) for i in range(n)] if pool is None else list(pool.map(one, range(n)))
        return StagedPngs(block, jobs_off, [r[0] for r in res], offs[:-1].copy(), np.array([r[1] for r in res], dtype=np.int64),
                          np.array([r[2] for r in res], dtype=np.int64), names)

    # ------------------------------------------------------------------------------------------------ device phase
    def launch(self, staged: StagedPngs, stream: Optional[int] = None, out: Any = None, copy_stream: Any = None):
        """-> (frames, status): `frames` = list of CUDA uint8 [h, w, 3] tensors (views of one allocation; or ONE
        [n, h, w, 3] tensor when all images have the same size), `status` = CUDA int32 [n] (0 = decoded), both valid in
        stream order.  `out`: optional CUDA uint8 [n, h, w, 3] tensor to decode into (images of one size).
        `copy_stream`: a torch.cuda.Stream for the H2D copy of the compressed bytes (the current stream then only waits
        for it, so the copy runs beside whatever the current stream is still computing)."""
        import torch
        if not torch.cuda.is_available():
            raise lib.McgError('mcgaze_b200 needs a CUDA device (B200, sm_100a); there is no CPU path')
        so = lib.load_library()
        dev = torch.device('cuda', self.device)
        n = len(staged)
        with torch.cuda.device(dev):
            if copy_stream is not None:      # from the copy stream's pool: a block this stream may overwrite at once
                with torch.cuda.stream(copy_stream):
                    zdev = torch.empty(staged.block.shape, dtype=torch.uint8, device=dev)
            else:
                zdev = torch.empty(staged.block.shape, dtype=torch.uint8, device=dev)
            w = np.array([i[0] for i in staged.infos], dtype=np.int64)
            h = np.array([i[1] for i in staged.infos], dtype=np.int64)
            ch = np.array([_CHANNELS[i[2]] for i in staged.infos], dtype=np.int64)
            scan_bytes = h * (1 + w * ch)
            scan_off = np.concatenate([[0], np.cumsum((scan_bytes + _ALIGN - 1) // _ALIGN * _ALIGN)])
            scan = torch.empty(int(scan_off[-1]), dtype=torch.uint8, device=dev)
            same = bool((w == w[0]).all() and (h == h[0]).all())
            dst_bytes = h * w * 3
            if out is not None:
                if not same or tuple(out.shape) != (n, int(h[0]), int(w[0]), 3) or out.dtype != torch.uint8 or not out.is_cuda \
                        or not out.is_contiguous():
                    raise lib.McgError('`out` must be a contiguous CUDA uint8 [n, h, w, 3] tensor and all images of that size')
                dst = out
                dst_off = np.arange(n + 1, dtype=np.int64) * int(dst_bytes[0])
            elif same:
                dst = torch.empty((n, int(h[0]), int(w[0]), 3), dtype=torch.uint8, device=dev)
                dst_off = np.arange(n + 1, dtype=np.int64) * int(dst_bytes[0])
            else:
                dst_off = np.concatenate([[0], np.cumsum((dst_bytes + _ALIGN - 1) // _ALIGN * _ALIGN)])
                dst = torch.empty(int(dst_off[-1]), dtype=torch.uint8, device=dev)
            status = torch.empty(n, dtype=torch.int32, device=dev)
            jobs = staged.block.numpy()[staged.used:staged.used + 64 * n].view(self._job_dtype)
            jobs[:] = 0
            zbase = zdev.data_ptr()
            jobs['zdata'] = zbase + staged.zoff.astype(np.uint64)
            jobs['zbytes'] = staged.zlen
            jobs['width'], jobs['height'] = w, h
            jobs['color_type'] = [i[2] for i in staged.infos]
            jobs['palette'] = np.where(staged.paloff >= 0, zbase + np.maximum(staged.paloff, 0), 0).astype(np.uint64)
            jobs['scan'] = scan.data_ptr() + scan_off[:-1].astype(np.uint64)
            jobs['dst'] = dst.data_ptr() + dst_off[:-1].astype(np.uint64)
            jobs['dst_stride'] = 3 * w
            if copy_stream is not None:                            # compressed bytes + palettes + job table: one copy
                cur = torch.cuda.current_stream()
                with torch.cuda.stream(copy_stream):
                    zdev.copy_(staged.block, non_blocking=True)
                    done = torch.cuda.Event()
                    done.record(copy_stream)
                cur.wait_event(done)
            else:
                zdev.copy_(staged.block, non_blocking=True)
            st = torch.cuda.current_stream().cuda_stream if stream is None else stream
            lib._check(so.mcg_png_decode(ctypes.cast(zbase + staged.used, ctypes.POINTER(lib.mcg_png_job)), n, status.data_ptr(), st),
                       'mcg_png_decode')
            cur = torch.cuda.current_stream()
            for t in (zdev, scan):           # freed at return: the caching allocator must not hand them out before the kernels ran
                t.record_stream(cur)
        if same:
            return dst, status
        frames = [dst[int(dst_off[i]):int(dst_off[i]) + int(dst_bytes[i])].view(int(h[i]), int(w[i]), 3) for i in range(n)]
        return frames, status

    @staticmethod
    def status_async(status):
        """-> (pinned int32 copy of `status`, event): the per-image results without stalling the stream"""
        import torch
        host = torch.empty(status.shape, dtype=status.dtype, pin_memory=True)
        host.copy_(status, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(status.device))
        return host, ev

    def check(self, staged: StagedPngs, status) -> None:
        """read the per-image status (synchronises when it is still on the device); raises McgError naming the first
        corrupt file"""
        st = status.cpu().numpy()
        bad = np.nonzero(st)[0]
        if bad.size:
            i = int(bad[0])
            code = int(st[i])
            what = STATUS_NAMES[code] if 0 <= code < len(STATUS_NAMES) else f'status {code}'
            raise lib.McgError(f'{staged.names[i]}: corrupt PNG data ({what}); {bad.size} of {st.size} images failed')

    def decode(self, files: Sequence[FileLike], pool=None):
        """stage + launch + check in one call -> list of CUDA uint8 [h, w, 3] BGR tensors (or one [n, h, w, 3] tensor)"""
        staged = self.stage(files, pool)
        frames, status = self.launch(staged)
        self.check(staged, status)
        return frames
