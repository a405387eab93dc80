"""PNG test files for tests/test_png.py (CPU: oracle / host build of the decoder core; GPU: mcg_png_decode), generated
in memory: what cv2.imwrite produces (the Gaze360 rawframes, tools/gaze360_img_reorganize.py:108), every colour type,
every scanline filter, every deflate block type, long codes, far matches, tiny and ragged sizes."""
import zlib

import cv2
import numpy as np

from oracle import png_oracle as P


def _natural(h, w, seed):
    """smooth gradients + texture + noise: compresses like a photograph (long literal runs, few matches)"""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.stack([128 + 90 * np.sin(x / 17 + c) * np.cos(y / 23 - c) + 12 * rng.standard_normal((h, w)) for c in range(3)], -1)
    return np.clip(img, 0, 255).astype(np.uint8)


def cases():
    rng = np.random.default_rng(7)
    out = {}
    nat = _natural(150, 131, 0)
    out['cv2_default'] = cv2.imencode('.png', nat)[1].tobytes()                       # Sub filter, Z_RLE, level 1
    out['cv2_level9'] = cv2.imencode('.png', nat, [cv2.IMWRITE_PNG_COMPRESSION, 9, cv2.IMWRITE_PNG_STRATEGY, 0])[1].tobytes()
    out['cv2_gray'] = cv2.imencode('.png', nat[:, :, 0])[1].tobytes()
    out['cv2_rgba'] = cv2.imencode('.png', np.dstack([nat, nat[:, :, :1]]), [cv2.IMWRITE_PNG_COMPRESSION, 3])[1].tobytes()
    out['cv2_448'] = cv2.imencode('.png', _natural(448, 448, 1))[1].tobytes()
    rnd = rng.integers(0, 256, (37, 29, 3), dtype=np.uint8)
    for ft in range(5):
        out[f'filter{ft}_rgb'] = P.encode(rnd, 2, (ft,))
    mix = _natural(70, 33, 2)
    out['mixed_rgb'] = P.encode(mix, 2, (4, 1, 3, 0, 2, 4, 4, 3))
    out['mixed_gray'] = P.encode(mix[:, :, 0], 0, (0, 1, 2, 3, 4))
    out['mixed_ga'] = P.encode(mix[:, :, :2], 4, (3, 4, 1))
    out['mixed_rgba'] = P.encode(np.dstack([mix, mix[:, :, 1]]), 6, (4, 3, 2, 1))
    pal = rng.integers(0, 256, (200, 3), dtype=np.uint8)
    out['palette'] = P.encode(rng.integers(0, 200, (41, 50), dtype=np.uint8), 3, (0, 4, 1), palette=pal)
    out['stored'] = P.encode(_natural(200, 180, 3), 2, (1,), level=0)                      # stored blocks, several of 65535
    out['fixed'] = P.encode(mix, 2, (1, 2), strategy=zlib.Z_FIXED)
    out['huffman_only'] = P.encode(mix, 2, (4,), strategy=zlib.Z_HUFFMAN_ONLY)
    tile = rng.integers(0, 256, (16, 40, 3), dtype=np.uint8)
    far = np.tile(tile, (14, 5, 1))                                                       # matches up to ~9.6 KB away, length 258
    out['far_matches'] = P.encode(far, 2, (0,), level=9)
    flat = np.zeros((64, 300, 3), np.uint8)
    flat[::7, ::11] = rng.integers(0, 256, flat[::7, ::11].shape, dtype=np.uint8)
    out['rle_long_codes'] = P.encode(flat, 2, (0, 2), level=9)                            # skewed statistics: codes > 10 bits
    skew = (rng.geometric(0.35, (120, 97, 3)) - 1).clip(0, 255).astype(np.uint8)
    out['long_codes'] = P.encode(skew, 2, (0,), strategy=zlib.Z_HUFFMAN_ONLY)
    out['small_idat'] = P.encode(mix, 2, (1,), idat=97)
    out['window512'] = P.encode(far, 2, (2,), level=9, wbits=9)
    out['ancillary'] = P.encode(mix, 2, (1,), extra_chunks=[P._chunk(b'gAMA', (45455).to_bytes(4, 'big')),
                                                             P._chunk(b'tEXt', b'Comment\x00hello')])
    out['one_pixel'] = P.encode(rng.integers(0, 256, (1, 1, 3), dtype=np.uint8), 2, (4,))
    out['one_row'] = P.encode(rng.integers(0, 256, (1, 77, 3), dtype=np.uint8), 2, (3,))
    out['one_col'] = P.encode(rng.integers(0, 256, (67, 1, 3), dtype=np.uint8), 2, (4, 2, 3))
    out['rows_32'] = P.encode(rng.integers(0, 256, (32, 5, 3), dtype=np.uint8), 2, (4,))
    out['rows_33'] = P.encode(rng.integers(0, 256, (33, 5, 4), dtype=np.uint8), 6, (4, 3))
    return out


def corrupt_cases():
    """name -> (file bytes, expected: 'parse' = mcg_png_parse rejects, 'status' = per-image status != 0)"""
    good = P.encode(_natural(40, 30, 5), 2, (1, 4))
    p = P.parse(good)
    z = p['zdata']

    def rebuild(zdata, rows=None):
        ihdr = P._chunk(b'IHDR', (30).to_bytes(4, 'big') + (rows or 40).to_bytes(4, 'big') + bytes([8, 2, 0, 0, 0]))
        return P.SIGNATURE + ihdr + P._chunk(b'IDAT', zdata) + P._chunk(b'IEND', b'')

    out = {}
    out['bad_signature'] = (b'\x89PNX' + good[4:], 'parse')
    out['bad_crc'] = (good[:45] + bytes([good[45] ^ 1]) + good[46:], 'parse')
    out['truncated_file'] = (good[:len(good) // 2], 'parse')
    out['truncated_stream'] = (rebuild(z[:len(z) // 2]), 'status')
    out['more_rows_than_data'] = (rebuild(z, rows=41), 'status')
    out['fewer_rows_than_data'] = (rebuild(z, rows=39), 'status')
    out['bad_zlib_header'] = (rebuild(bytes([0x79]) + z[1:]), 'status')
    raw = bytearray(zlib.decompress(z))
    raw[91 * 3] = 7                                                                          # filter byte of row 3
    out['bad_filter'] = (rebuild(zlib.compress(bytes(raw))), 'status')
    out['reserved_block_type'] = (rebuild(z[:2] + bytes([0x07]) + z[3:]), 'status')         # BFINAL=1, BTYPE=3
    return out
