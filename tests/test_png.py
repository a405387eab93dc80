"""PNG decode on the device (the "decode" step; mmdet/datasets/pipelines/loading.py:58-69).

CPU (`-m "not gpu"`): the oracle (oracle/png_oracle.py) against cv2.imdecode - the decoder the reference calls - and
against the committed fixtures; the host function mcg_png_parse; the decoder core of the CUDA library compiled for the
host (tests/png_host_sim.cpp = mcgaze_b200/csrc/png_core.cuh with MCG_PNG_HOST_SIM) against zlib and cv2.
GPU (`-m gpu`): mcg_png_decode through the C ABI, bit-exact against cv2.imdecode on every case; corrupt streams;
decoded frames through mcg_preprocess; the evaluation driver with decode='gpu' against decode='host'."""
import ctypes
import os
import subprocess
import sys
import zlib

import cv2
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import png_cases  # noqa: E402
from mcgaze_b200 import lib  # noqa: E402
from oracle import png_oracle as P  # noqa: E402

CASES = png_cases.cases()
CORRUPT = png_cases.corrupt_cases()


def _cv2(data: bytes) -> np.ndarray:
    img = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
    assert img is not None
    return img


@pytest.fixture(scope='module')
def sim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp('pngsim') / 'png_host_sim.so')
    subprocess.run(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-o', so, os.path.join(ROOT, 'tests', 'png_host_sim.cpp')],
                   check=True)
    s = ctypes.CDLL(so)
    s.sim_inflate.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_longlong, ctypes.POINTER(ctypes.c_longlong)]
    s.sim_unfilter.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong]
    return s


def _sim_decode(sim, data: bytes):
    """the two device stages on the host build -> (inflate status, unfilter status, BGR image)"""
    p = P.parse(data)
    w, h, ct = p['width'], p['height'], p['color_type']
    z = np.frombuffer(p['zdata'], np.uint8).copy()
    expected = h * (1 + w * P.CHANNELS[ct])
    scan = np.zeros(expected + 64, np.uint8)
    scan[expected:] = 0xA5                                   # canary behind the buffer
    produced = ctypes.c_longlong(0)
    st = sim.sim_inflate(z.ctypes.data, z.size, scan.ctypes.data, expected, ctypes.byref(produced))
    assert (scan[expected:] == 0xA5).all(), 'inflate wrote behind its output buffer'
    if st != 0 or produced.value != expected:
        return st or 9, None, None
    pal = np.zeros(768, np.uint8)
    if 'palette' in p:
        pal[:p['palette'].size] = p['palette'].reshape(-1)
    dst = np.zeros((h, w, 3), np.uint8)
    st2 = sim.sim_unfilter(scan.ctypes.data, w, h, ct, pal.ctypes.data, dst.ctypes.data, 3 * w)
    return 0, st2, dst


# ------------------------------------------------------------------------------------------------------ CPU
@pytest.mark.parametrize('name', sorted(CASES))
def test_oracle_matches_cv2(name):
    assert np.array_equal(P.decode(CASES[name]), _cv2(CASES[name]))


def test_oracle_matches_committed_fixtures():
    """files and cv2.imdecode outputs frozen by oracle/gen_golden_png.py (pins the oracle independently of the cv2 here)"""
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'png_decode.npz'))
    names = sorted(k[5:] for k in g.files if k.startswith('file_'))
    assert len(names) >= 8
    for n in names:
        assert np.array_equal(P.decode(g['file_' + n].tobytes()), g['bgr_' + n]), n


@pytest.mark.parametrize('name', sorted(CASES))
def test_host_build_of_the_decoder_core(sim, name):
    data = CASES[name]
    p = P.parse(data)
    # stage 1 against zlib itself
    raw = zlib.decompress(p['zdata'])
    out = np.zeros(len(raw), np.uint8)
    z = np.frombuffer(p['zdata'], np.uint8).copy()
    produced = ctypes.c_longlong(0)
    assert sim.sim_inflate(z.ctypes.data, z.size, out.ctypes.data, out.size, ctypes.byref(produced)) == 0
    assert produced.value == len(raw) and out.tobytes() == raw
    # a misaligned input pointer takes the byte-wise refill path first
    z1 = np.zeros(z.size + 1, np.uint8)
    z1[1:] = z
    out[:] = 0
    assert sim.sim_inflate(z1.ctypes.data + 1, z.size, out.ctypes.data, out.size, ctypes.byref(produced)) == 0
    assert out.tobytes() == raw
    # both stages against cv2
    st, st2, img = _sim_decode(sim, data)
    assert st == 0 and st2 == 0
    assert np.array_equal(img, _cv2(data))


def test_codes_longer_than_the_primary_table_take_the_canonical_path(sim):
    """decode_symbol's bit-by-bit path must be exercised: a Huffman-only stream of a geometric source spends more than
    the 10 index bits of the literal table on its rare symbols"""
    sim.sim_slow_symbols.restype = ctypes.c_longlong
    before = sim.sim_slow_symbols()
    st, st2, img = _sim_decode(sim, CASES['long_codes'])
    assert st == 0 and st2 == 0 and np.array_equal(img, _cv2(CASES['long_codes']))
    assert sim.sim_slow_symbols() - before > 10


@pytest.mark.parametrize('name', sorted(CASES))
def test_parse_matches_oracle(name):
    so = lib.load_library()
    data = np.frombuffer(CASES[name], np.uint8)
    p = P.parse(CASES[name])
    info = lib.mcg_png_info()
    assert so.mcg_png_parse(data.ctypes.data, data.size, 1, ctypes.byref(info), None, 0) == 0   # sizing call
    assert (info.width, info.height, info.bit_depth, info.color_type, info.interlace) == \
        (p['width'], p['height'], p['bit_depth'], p['color_type'], p['interlace'])
    assert info.idat_bytes == len(p['zdata']) and info.supported == 1 and info.channels == P.CHANNELS[p['color_type']]
    z = np.zeros(info.idat_bytes, np.uint8)
    assert so.mcg_png_parse(data.ctypes.data, data.size, 1, ctypes.byref(info), z.ctypes.data, z.size) == 0
    assert z.tobytes() == p['zdata']
    if 'palette' in p:
        assert bytes(info.palette)[:p['palette'].size] == p['palette'].tobytes() and info.has_palette == 1
    # too small a buffer is an error that still reports the size
    if info.idat_bytes > 1:
        small = np.zeros(info.idat_bytes - 1, np.uint8)
        assert so.mcg_png_parse(data.ctypes.data, data.size, 0, ctypes.byref(info), small.ctypes.data, small.size) != 0
        assert info.idat_bytes == len(p['zdata'])


def test_parse_rejects_and_flags():
    so = lib.load_library()
    info = lib.mcg_png_info()
    for name, (data, where) in CORRUPT.items():
        a = np.frombuffer(data, np.uint8)
        rc = so.mcg_png_parse(a.ctypes.data, a.size, 1, ctypes.byref(info), None, 0)
        assert (rc != 0) == (where == 'parse'), name
        if where == 'parse':
            assert cv2.imdecode(a, cv2.IMREAD_COLOR) is None or name == 'bad_crc'
    # the CRC is only looked at on request
    a = np.frombuffer(CORRUPT['bad_crc'][0], np.uint8)
    assert so.mcg_png_parse(a.ctypes.data, a.size, 0, ctypes.byref(info), None, 0) == 0
    # valid files the device decoder does not take are flagged, not rejected
    img16 = (np.arange(20 * 10 * 3, dtype=np.uint16) * 97).reshape(20, 10, 3)
    a = cv2.imencode('.png', img16)[1]
    assert so.mcg_png_parse(a.ctypes.data, a.size, 1, ctypes.byref(info), None, 0) == 0
    assert info.bit_depth == 16 and info.supported == 0


@pytest.mark.parametrize('name', sorted(n for n, (_, w) in CORRUPT.items() if w == 'status'))
def test_host_build_flags_corrupt_streams(sim, name):
    st, st2, _ = _sim_decode(sim, CORRUPT[name][0])
    assert st != 0 or st2 != 0


def test_host_build_flags_random_corruption(sim):
    """bit flips anywhere in the stream: flagged (a code error, a size mismatch, or the Adler-32 of the inflated bytes
    against the stream trailer) unless the decoded image is unchanged (padding bits); never a write outside the buffers"""
    rng = np.random.default_rng(0)
    p = P.parse(CASES['cv2_level9'])
    base = bytearray(p['zdata'])
    want = _cv2(CASES['cv2_level9'])
    ihdr = P._chunk(b'IHDR', p['width'].to_bytes(4, 'big') + p['height'].to_bytes(4, 'big') + bytes([8, 2, 0, 0, 0]))
    statuses = set()
    for _ in range(300):
        z = bytearray(base)
        for _ in range(int(rng.integers(1, 4))):
            z[int(rng.integers(2, len(z)))] ^= 1 << int(rng.integers(0, 8))
        st, st2, img = _sim_decode(sim, P.SIGNATURE + ihdr + P._chunk(b'IDAT', bytes(z)) + P._chunk(b'IEND', b''))
        assert st != 0 or st2 != 0 or np.array_equal(img, want)
        statuses.add(st)
    assert 12 in statuses          # MCG_PNG_BAD_CHECKSUM: a flipped literal decodes "fine" and only the checksum sees it


def test_staging_files_in_one_c_call_equals_staging_file_images(tmp_path):
    """mcg_png_file_sizes + mcg_png_stage_files (the loader-thread side of decode='gpu': read, chunk walk, copy into the
    staging block on worker threads inside the call) against the per-image mcg_png_parse path; and what they report for
    files the device decoder cannot take"""
    from concurrent.futures import ThreadPoolExecutor
    from mcgaze_b200.png import GpuPngDecoder, UnsupportedPng
    names = sorted(CASES)
    paths = []
    for n in names:
        (tmp_path / f'{n}.png').write_bytes(CASES[n])
        paths.append(str(tmp_path / f'{n}.png'))
    dec = GpuPngDecoder(check_crc=True)
    a = dec.stage([CASES[n] for n in names])
    with ThreadPoolExecutor(4) as pool:
        b = dec.stage(paths, pool)
    assert a.infos == b.infos and (a.zlen == b.zlen).all() and ((a.paloff >= 0) == (b.paloff >= 0)).all()
    for i, n in enumerate(names):
        za = a.block.numpy()[a.zoff[i]:a.zoff[i] + a.zlen[i]]
        zb = b.block.numpy()[b.zoff[i]:b.zoff[i] + b.zlen[i]]
        assert np.array_equal(za, zb), n
        if a.paloff[i] >= 0:
            assert np.array_equal(a.block.numpy()[a.paloff[i]:a.paloff[i] + 768], b.block.numpy()[b.paloff[i]:b.paloff[i] + 768])
    assert b.names == paths
    # not a PNG (a JPEG behind a .png name), a 16-bit PNG: the caller's cue to use the host loader; a missing file; a bad CRC
    assert cv2.imwrite(str(tmp_path / 'x.jpg'), np.zeros((10, 10, 3), np.uint8))
    os.replace(tmp_path / 'x.jpg', tmp_path / 'fake.png')
    with pytest.raises(UnsupportedPng, match='fake.png'):
        dec.stage([paths[0], str(tmp_path / 'fake.png')])
    assert cv2.imwrite(str(tmp_path / 's16.png'), (np.arange(600, dtype=np.uint16) * 97).reshape(20, 10, 3))
    with pytest.raises(UnsupportedPng, match='s16.png'):
        dec.stage([str(tmp_path / 's16.png')])
    with pytest.raises(FileNotFoundError):
        dec.stage([str(tmp_path / 'missing.png')])
    (tmp_path / 'crc.png').write_bytes(CORRUPT['bad_crc'][0])
    with pytest.raises(lib.McgError, match='crc.png'):
        dec.stage([str(tmp_path / 'crc.png')])
    assert len(GpuPngDecoder(check_crc=False).stage([str(tmp_path / 'crc.png')])) == 1


def test_decoder_needs_a_gpu_and_never_falls_back():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from mcgaze_b200.png import GpuPngDecoder
    dec = GpuPngDecoder()
    staged = dec.stage([CASES['cv2_default'], CASES['mixed_rgba']])        # the host phase works anywhere
    assert staged.shapes == [(150, 131), (70, 33)]
    with pytest.raises(lib.McgError):
        dec.launch(staged)
    src = open(os.path.join(ROOT, 'mcgaze_b200', 'png.py')).read()
    assert 'cv2' not in src.split('"""', 2)[2] and 'oracle' not in src.split('"""', 2)[2]


# ------------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_gpu_decode_matches_cv2_bit_exact():
    import torch
    from mcgaze_b200.png import GpuPngDecoder
    dec = GpuPngDecoder(0)
    names = sorted(CASES)
    frames = dec.decode([CASES[n] for n in names])
    assert len(frames) == len(names)
    for n, f in zip(names, frames):
        want = _cv2(CASES[n])
        assert tuple(f.shape) == want.shape and f.dtype == torch.uint8 and f.is_cuda, n
        assert np.array_equal(f.cpu().numpy(), want), n
    # images of one size come back as one [n, h, w, 3] block; a batch larger than one launch's descriptor table
    same = [CASES['cv2_default']] * 500
    block = dec.decode(same)
    assert tuple(block.shape) == (500, 150, 131, 3)
    want = torch.from_numpy(_cv2(CASES['cv2_default'])).cuda()
    assert bool((block == want[None]).all())


@pytest.mark.gpu
def test_gpu_decode_flags_corrupt_streams_and_keeps_the_rest():
    from mcgaze_b200.png import GpuPngDecoder
    dec = GpuPngDecoder(0)
    bad = sorted(n for n, (_, w) in CORRUPT.items() if w == 'status')
    files = [CASES['mixed_rgb']] + [CORRUPT[n][0] for n in bad] + [CASES['palette']]
    staged = dec.stage(files)
    frames, status = dec.launch(staged)
    st = status.cpu().numpy()
    assert st[0] == 0 and st[-1] == 0 and (st[1:-1] != 0).all(), dict(zip(bad, st[1:-1]))
    assert np.array_equal(frames[0].cpu().numpy(), _cv2(CASES['mixed_rgb']))
    assert np.array_equal(frames[-1].cpu().numpy(), _cv2(CASES['palette']))
    with pytest.raises(lib.McgError, match='corrupt PNG'):
        dec.check(staged, status)
    for name, (data, where) in CORRUPT.items():
        if where == 'parse':
            with pytest.raises(lib.McgError):
                dec.stage([data])


@pytest.mark.gpu
def test_gpu_decode_random_corruption_stays_inside_its_buffers():
    """200 bit-flipped streams next to intact neighbours: the neighbours still decode bit-exactly"""
    import torch
    from mcgaze_b200.png import GpuPngDecoder
    rng = np.random.default_rng(1)
    p = P.parse(CASES['cv2_level9'])
    base = bytearray(p['zdata'])
    ihdr = P._chunk(b'IHDR', p['width'].to_bytes(4, 'big') + p['height'].to_bytes(4, 'big') + bytes([8, 2, 0, 0, 0]))
    files = []
    for k in range(200):
        z = bytearray(base)
        for _ in range(int(rng.integers(1, 4))):
            z[int(rng.integers(2, len(z)))] ^= 1 << int(rng.integers(0, 8))
        files.append(P.SIGNATURE + ihdr + P._chunk(b'IDAT', bytes(z)) + P._chunk(b'IEND', b''))
        files.append(CASES['cv2_level9'])
    dec = GpuPngDecoder(0)
    staged = dec.stage(files)
    frames, status = dec.launch(staged)
    torch.cuda.synchronize()
    want = torch.from_numpy(_cv2(CASES['cv2_level9'])).cuda()
    st = status.cpu().numpy()
    assert (st[1::2] == 0).all() and bool((frames[1::2] == want[None]).all())
    # ... and every damaged stream is flagged unless its pixels are unchanged (Adler-32 verified on the device)
    same = (frames[0::2] == want[None]).flatten(1).all(1).cpu().numpy()
    assert ((st[0::2] != 0) | same).all() and (st[0::2] == 12).any()


@pytest.mark.gpu
def test_gpu_decoded_frames_feed_the_pipeline():
    """decode -> mcg_preprocess entirely on the device == cv2.imdecode on the host -> mcg_preprocess"""
    import torch
    from mcgaze_b200.compat import Config
    from mcgaze_b200.pipeline import GpuTestPipeline
    from mcgaze_b200.png import GpuPngDecoder
    cfg = Config.fromfile(os.path.join(ROOT, 'configs/multiclue_gaze/multiclue_gaze_r50_gaze360.py'))
    names = ['cv2_default', 'cv2_448', 'mixed_rgb', 'cv2_gray']
    rands = [0.1, 0.5, 0.9, 0.3]
    pipe = GpuTestPipeline(cfg.data.test.pipeline)
    host = pipe.batch([_cv2(CASES[n]) for n in names], rands=rands)
    dev = pipe.batch(GpuPngDecoder(0).decode([CASES[n] for n in names]), rands=rands)
    assert torch.equal(host['img'][0], dev['img'][0])
    assert [m['img_shape'] for m in host['img_metas'][0]] == [m['img_shape'] for m in dev['img_metas'][0]]


@pytest.mark.gpu
def test_gpu_evaluation_driver_with_device_decode_equals_host_decode(synthetic_sd, tmp_path):
    """run_clips on PNG files: decode='gpu' (files -> mcg_png_decode -> mcg_preprocess -> forward) gives the rows of
    decode='host' (cv2.imread in loader threads) exactly - same pixels in, same kernels after; a JPEG frame sends its
    batch through the host loader instead of failing."""
    import torch
    from mcgaze_b200 import evaluate as ev
    from mcgaze_b200.apis import init_detector
    from mcgaze_b200.pipeline import GpuTestPipeline
    model = init_detector(os.path.join(ROOT, 'configs/multiclue_gaze/multiclue_gaze_r50_gaze360.py'), None, 'cuda:0')
    model.load_state_dict(synthetic_sd)
    lengths = [7, 10, 4]
    videos = []
    for vi, L in enumerate(lengths):
        names = []
        size = (120, 100) if vi != 1 else (90, 128)            # one video with another frame size
        for t in range(L):
            name = f'v{vi:03d}/{t:05d}.png'
            os.makedirs(tmp_path / os.path.dirname(name), exist_ok=True)
            assert cv2.imwrite(str(tmp_path / name), png_cases._natural(size[0], size[1], 100 * vi + t))
            names.append(name)
        videos.append(dict(id=vi + 1, file_names=names))
    anno = dict(videos=videos, annotations=[])
    rows = {}
    for mode in ('host', 'gpu'):
        ds = ev.Gaze360ClipDataset(anno, img_prefix=str(tmp_path), decode=mode)
        pipe = GpuTestPipeline(model.cfg.data.test.pipeline, seed=3)
        rows[mode] = ev.single_gpu_test(model, ds, pipe, clips_per_batch=2, workers=2)
        assert ds.host_decoded_batches == 0
    for a, b in zip(rows['host'], rows['gpu']):
        assert np.array_equal(a, b)
    # a corrupt file is reported by name
    bad = tmp_path / 'v000' / '00003.png'
    data = bytearray(open(bad, 'rb').read())
    keep = bytes(data)
    p = P.parse(keep)
    z = bytearray(p['zdata'])
    z[len(z) // 2] ^= 0x10
    ihdr = P._chunk(b'IHDR', p['width'].to_bytes(4, 'big') + p['height'].to_bytes(4, 'big') + bytes([8, 2, 0, 0, 0]))
    open(bad, 'wb').write(P.SIGNATURE + ihdr + P._chunk(b'IDAT', bytes(z)) + P._chunk(b'IEND', b''))
    ds = ev.Gaze360ClipDataset(anno, img_prefix=str(tmp_path), decode='gpu')
    with pytest.raises(lib.McgError, match='00003.png'):      # a code error or the Adler-32 check, reported by file name
        ev.single_gpu_test(model, ds, GpuTestPipeline(model.cfg.data.test.pipeline, seed=3), clips_per_batch=2, workers=2)
    # a JPEG among the frames: that batch is decoded by cv2 on the host, as the reference does
    open(bad, 'wb').write(keep)
    jpg = tmp_path / 'v002' / '00001.png'
    cv2.imwrite(str(tmp_path / 'tmp.jpg'), png_cases._natural(120, 100, 55))
    os.replace(tmp_path / 'tmp.jpg', jpg)
    ds = ev.Gaze360ClipDataset(anno, img_prefix=str(tmp_path), decode='gpu')
    ev.single_gpu_test(model, ds, GpuTestPipeline(model.cfg.data.test.pipeline, seed=3), clips_per_batch=2, workers=2)
    assert ds.host_decoded_batches >= 1
    torch.cuda.synchronize()
