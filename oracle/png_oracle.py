"""CPU restatement of the PNG decode inside LoadImageFromFile - TEST INFRASTRUCTURE ONLY (the product path is
mcg_png_parse / mcg_png_decode in mcgaze_b200/csrc/png_decode.cu; nothing under mcgaze_b200/ imports this file).

What it restates: mmdet/datasets/pipelines/loading.py:58-69 reads the file and calls mmcv.imfrombytes(flag='color'),
i.e. cv2.imdecode(buf, IMREAD_COLOR).  mmcv and OpenCV are third-party dependencies that are not vendored in
/root/reference (requirements/runtime.txt: mmcv-full; opencv-python 4.x comes with it), so the algorithm restated here
is the published one: the PNG specification (ISO/IEC 15948) section 9 for the five scanline filters, RFC 1950 / 1951
for the zlib stream (delegated to python's zlib, itself the reference implementation), and OpenCV's PNG reader for the
conversion to 8-bit BGR (modules/imgcodecs/src/grfmt_png.cpp: png_set_strip_alpha, png_set_palette_to_rgb,
png_set_gray_to_rgb, png_set_bgr; no gamma handling).
Pinned by: cv2.imdecode itself, executed on the same files in tests/test_png.py (cv2 IS the reference's decoder), and
by the committed fixtures tests/golden/png_*.npz (oracle/gen_golden_png.py).

`encode()` is a test-vector generator: a minimal PNG writer with a chosen filter per row, zlib level / strategy
(stored, fixed-Huffman and dynamic blocks, long-distance matches) and IDAT chunk size.
"""
from __future__ import annotations

import struct
import zlib
from typing import Dict, Optional, Sequence

import numpy as np

CHANNELS = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}
SIGNATURE = b'\x89PNG\r\n\x1a\n'


def parse(data: bytes, check_crc: bool = True) -> Dict:
    """Chunk walk (PNG spec section 5): IHDR fields, PLTE, concatenated IDAT payload."""
    if data[:8] != SIGNATURE:
        raise ValueError('not a PNG file')
    pos, out, idat = 8, {}, []
    while pos + 12 <= len(data):
        n, typ = struct.unpack('>I4s', data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        if check_crc and zlib.crc32(typ + body) != struct.unpack('>I', data[pos + 8 + n:pos + 12 + n])[0]:
            raise ValueError('chunk CRC mismatch')
        if typ == b'IHDR':
            w, h, depth, ct, comp, flt, lace = struct.unpack('>IIBBBBB', body)
            out.update(width=w, height=h, bit_depth=depth, color_type=ct, interlace=lace)
        elif typ == b'PLTE':
            out['palette'] = np.frombuffer(body, np.uint8).reshape(-1, 3)
        elif typ == b'IDAT':
            idat.append(body)
        elif typ == b'IEND':
            break
        pos += 12 + n
    out['zdata'] = b''.join(idat)
    return out


def _paeth(a: int, b: int, c: int) -> int:
    p = a + b - c
    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
    return a if pa <= pb and pa <= pc else b if pb <= pc else c


def unfilter(raw: bytes, width: int, height: int, bpp: int) -> np.ndarray:
    """PNG spec 9.2: -> uint8 [height, width * bpp] reconstructed scanlines.  Arithmetic is modulo 256, bytes left of the
    first pixel and above the first row count as zero."""
    rb = width * bpp
    if len(raw) != height * (1 + rb):
        raise ValueError('scanline data has the wrong size')
    src = np.frombuffer(raw, np.uint8).reshape(height, 1 + rb)
    out = np.zeros((height, rb), np.uint8)
    prev = np.zeros(rb, np.int32)
    for r in range(height):
        ft, line = int(src[r, 0]), src[r, 1:].astype(np.int32)
        cur = np.zeros(rb, np.int32)
        if ft == 0:
            cur = line.copy()
        elif ft == 2:
            cur = (line + prev) & 255
        elif ft == 1:
            for c in range(bpp):          # running sum per channel
                cur[c::bpp] = np.cumsum(line[c::bpp]) & 255
        elif ft in (3, 4):
            for j in range(rb):
                a = int(cur[j - bpp]) if j >= bpp else 0
                b = int(prev[j])
                c = int(prev[j - bpp]) if j >= bpp else 0
                cur[j] = (int(line[j]) + ((a + b) >> 1 if ft == 3 else _paeth(a, b, c))) & 255
        else:
            raise ValueError(f'bad filter type {ft}')
        out[r] = cur
        prev = cur
    return out


def decode(data: bytes) -> np.ndarray:
    """-> what cv2.imdecode(data, cv2.IMREAD_COLOR) returns for an 8-bit non-interlaced PNG: uint8 [h, w, 3] BGR."""
    p = parse(data)
    if p['bit_depth'] != 8 or p['interlace'] != 0:
        raise NotImplementedError('only 8-bit non-interlaced PNG')
    ct, w, h = p['color_type'], p['width'], p['height']
    bpp = CHANNELS[ct]
    px = unfilter(zlib.decompress(p['zdata']), w, h, bpp).reshape(h, w, bpp)
    if ct in (0, 4):
        return np.repeat(px[:, :, :1], 3, axis=2)
    if ct == 3:
        pal = np.zeros((256, 3), np.uint8)
        pal[:len(p['palette'])] = p['palette']
        return pal[px[:, :, 0]][:, :, ::-1].copy()
    return px[:, :, 2::-1].copy()           # RGB(A) -> BGR, alpha dropped


# ------------------------------------------------------------------------------------------------ test-vector writer
def _chunk(typ: bytes, body: bytes) -> bytes:
    return struct.pack('>I', len(body)) + typ + body + struct.pack('>I', zlib.crc32(typ + body))


def filter_rows(px: np.ndarray, filters: Sequence[int]) -> bytes:
    """px: uint8 [h, w, bpp] -> filtered scanlines with filter type filters[r % len(filters)] on row r."""
    h, w, bpp = px.shape
    rb = w * bpp
    flat = px.reshape(h, rb).astype(np.int32)
    out = bytearray()
    for r in range(h):
        ft = int(filters[r % len(filters)])
        cur = flat[r]
        prev = flat[r - 1] if r else np.zeros(rb, np.int32)
        a = np.concatenate([np.zeros(bpp, np.int32), cur[:-bpp]]) if rb > bpp else np.zeros(rb, np.int32)
        c = np.concatenate([np.zeros(bpp, np.int32), prev[:-bpp]]) if rb > bpp else np.zeros(rb, np.int32)
        if ft == 0:
            pred = np.zeros(rb, np.int32)
        elif ft == 1:
            pred = a
        elif ft == 2:
            pred = prev
        elif ft == 3:
            pred = (a + prev) >> 1
        else:
            p = a + prev - c
            pa, pb, pc = np.abs(p - a), np.abs(p - prev), np.abs(p - c)
            pred = np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, prev, c))
        out.append(ft)
        out += ((cur - pred) & 255).astype(np.uint8).tobytes()
    return bytes(out)


def encode(px: np.ndarray, color_type: int, filters: Sequence[int] = (0,), level: int = 6, strategy: int = zlib.Z_DEFAULT_STRATEGY,
           palette: Optional[np.ndarray] = None, idat: int = 8192, wbits: int = 15, extra_chunks: Sequence[bytes] = ()) -> bytes:
    """px: uint8 [h, w, channels of color_type] in PNG sample order (RGB, not BGR)."""
    if px.ndim == 2:
        px = px[:, :, None]
    h, w, bpp = px.shape
    assert bpp == CHANNELS[color_type]
    co = zlib.compressobj(level, zlib.DEFLATED, wbits, 9, strategy)
    z = co.compress(filter_rows(px, filters)) + co.flush()
    out = SIGNATURE + _chunk(b'IHDR', struct.pack('>IIBBBBB', w, h, 8, color_type, 0, 0, 0))
    for e in extra_chunks:
        out += e
    if palette is not None:
        out += _chunk(b'PLTE', np.asarray(palette, np.uint8).tobytes())
    for k in range(0, len(z), idat):
        out += _chunk(b'IDAT', z[k:k + idat])
    return out + _chunk(b'IEND', b'')
