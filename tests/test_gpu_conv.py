"""GPU: kernel-level parity of the tcgen05/TMA implicit-GEMM convolution and of the CUDA-core
cross-check kernel against torch (fp64 reference on the same inputs), through the C ABI
(mcg_debug_conv)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

CASES = [
    # name, NB, C, H, W, Cout, k, stride, pad, res_mode, relu, force_im2col, block_n
    ('1x1_plain', 2, 64, 56, 56, 256, 1, 1, 0, 0, 1, 0, 0),
    ('1x1_tail_rows', 1, 256, 30, 30, 64, 1, 1, 0, 0, 1, 0, 0),       # M = 900: ragged last tile
    ('1x1_tiny_M', 1, 64, 3, 3, 64, 1, 1, 0, 0, 0, 0, 0),             # M = 9 < one tile
    ('1x1_residual', 2, 64, 28, 28, 256, 1, 1, 0, 1, 1, 0, 0),
    ('1x1_fpn_topdown', 2, 512, 14, 14, 256, 1, 1, 0, 2, 0, 0, 0),
    ('1x1_via_im2col', 2, 128, 28, 28, 128, 1, 1, 0, 0, 0, 1, 0),
    ('1x1_stride2', 2, 256, 56, 56, 512, 1, 2, 0, 0, 0, 0, 0),
    ('3x3_c64', 2, 64, 56, 56, 64, 3, 1, 1, 0, 1, 0, 0),
    ('3x3_c256', 1, 256, 14, 14, 256, 3, 1, 1, 0, 0, 0, 0),
    ('3x3_stride2', 2, 128, 56, 56, 128, 3, 2, 1, 0, 1, 0, 0),
    ('3x3_7x7map', 3, 512, 7, 7, 512, 3, 1, 1, 0, 1, 0, 0),
    ('3x3_nonsquare', 2, 64, 24, 40, 128, 3, 1, 1, 0, 1, 0, 0),
    ('block_n64', 2, 256, 28, 28, 256, 1, 1, 0, 0, 0, 0, 64),
    ('block_n128', 2, 256, 28, 28, 256, 1, 1, 0, 0, 0, 0, 128),
    ('bigK', 1, 2048, 7, 7, 512, 1, 1, 0, 0, 1, 0, 0),
    # several tiles per persistent CTA: both epilogue groups, all TMEM accumulator buffers and the
    # residual prefetch ring wrap around
    ('multi_1x1_residual', 8, 64, 56, 56, 256, 1, 1, 0, 1, 1, 0, 0),
    ('multi_1x1_residual_bn64', 8, 64, 56, 56, 256, 1, 1, 0, 1, 1, 0, 64),
    ('multi_1x1_plain_k128', 6, 128, 56, 56, 256, 1, 1, 0, 0, 1, 0, 0),
    ('multi_3x3_c64', 8, 64, 56, 56, 64, 3, 1, 1, 0, 1, 0, 0),
    ('multi_fpn_topdown', 8, 256, 28, 28, 256, 1, 1, 0, 2, 0, 0, 0),
    ('multi_3x3_c128_bn128', 8, 128, 28, 28, 256, 3, 1, 1, 0, 0, 0, 128),
    # nearest-2x top-down add with the TMA-prefetched source window (round 2): the FPN's own shapes at 224^2 (56 / 28 /
    # 14 wide) and 448^2 (112 wide), tiles crossing image rows and frames, a non-square map, and a map too wide for the
    # 128-row window (falls back to per-thread gathers)
    ('fpn_topdown_56', 5, 256, 56, 56, 256, 1, 1, 0, 2, 0, 0, 0),
    ('fpn_topdown_112', 1, 256, 112, 112, 256, 1, 1, 0, 2, 0, 0, 0),
    ('fpn_topdown_nonsquare', 3, 512, 24, 40, 256, 1, 1, 0, 2, 0, 0, 0),
    ('fpn_topdown_odd_frames', 3, 256, 14, 14, 256, 1, 1, 0, 2, 0, 0, 128),
    ('fpn_topdown_wide_gather', 1, 256, 8, 288, 256, 1, 1, 0, 2, 0, 0, 0),
]
# max |err| relative to max |ref|: split-fp16 x3 and the fp32 CUDA-core kernel are fp32-class,
# single fp16 carries 2^-11 operand rounding
# single fp16 carries 2^-11 operand rounding; fp16c8 corrects both roundings to ~4 more bits in e4m3
# (largest |activation| in these cases ~5: e4m3 lo8 / hi8 stay in their normal range)
TOL = {'simt': 5e-6, 'fp16x3': 2e-5, 'fp16c8': 2e-4, 'fp16': 1e-3}


def _run(engine, case, out_mode=0):
    from mcgaze_b200 import lib
    name, NB, C, H, W, Cout, k, stride, pad, res_mode, relu, fim, bn = case
    g = torch.Generator().manual_seed(len(name) * 7 + C)
    x = torch.randn(NB, C, H, W, generator=g).cuda()
    w = (torch.randn(Cout, C, k, k, generator=g) / (C * k * k) ** 0.5).cuda()
    b = torch.randn(Cout, generator=g).cuda()
    P, Q = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    res = None
    if res_mode == 1:
        res = torch.randn(NB, Cout, P, Q, generator=g).cuda()
    elif res_mode == 2:
        res = torch.randn(NB, Cout, P // 2, Q // 2, generator=g).cuda()
    ref = F.conv2d(x.double(), w.double(), b.double(), stride=stride, padding=pad)
    if res_mode == 1:
        ref = ref + res.double()
    elif res_mode == 2:
        ref = ref + F.interpolate(res.double(), size=(P, Q), mode='nearest')
    if relu:
        ref = ref.relu()
    out = lib.debug_conv(engine, x, w, stride, pad, bias=b, res=res, res_mode=res_mode, relu=bool(relu),
                         force_im2col=bool(fim), force_block_n=bn, out_mode=out_mode)
    torch.cuda.synchronize()
    assert not torch.isnan(out).any()
    if out_mode & 4:
        # the e4m3 copy of the output's hi plane: 4 significant bits, subnormal step 2^-9, saturates at 448
        err = (out.double() - ref.clamp(-448, 448)).abs()
        return (err / (0.0625 * ref.abs() + 2e-3)).max().item()
    return (out.double() - ref).abs().max().item() / ref.abs().max().item()


@pytest.mark.parametrize('engine', ['simt', 'fp16x3', 'fp16c8', 'fp16'])
@pytest.mark.parametrize('case', CASES, ids=[c[0] for c in CASES])
def test_conv_parity(engine, case):
    assert _run(engine, case) < TOL[engine]


@pytest.mark.parametrize('case', [CASES[0], CASES[3], CASES[4], CASES[8], CASES[15]], ids=lambda c: c[0])
def test_conv_fp16c8_emits_hi8_plane(case):
    """out_mode bit 2: the e4m3 copy of the output's hi plane (operand of the next layer's weight correction)."""
    assert _run('fp16c8', case, out_mode=4) <= 1.0


@pytest.mark.parametrize('engine', ['fp16x3', 'fp16c8', 'fp16'])
@pytest.mark.parametrize('case', [CASES[0], CASES[1], CASES[3], CASES[4], CASES[9]], ids=lambda c: c[0])
def test_conv_parity_fp32_epilogue(engine, case):
    """out_mode=1: the direct fp32 store epilogue the head's Linear layers use."""
    assert _run(engine, case, out_mode=1) < TOL[engine]


# CTA-pair variant (clusters of two CTAs, tcgen05 cta_group::2, 256-row tiles): block_n code 1000 + n
PAIR_CASES = [
    ('pair_3x3_c256', 4, 256, 28, 28, 256, 3, 1, 1, 0, 1, 0, 1000),          # M = 3136: 12.25 pair tiles (ragged)
    ('pair_3x3_c64', 8, 64, 56, 56, 64, 3, 1, 1, 0, 1, 0, 1000),
    ('pair_3x3_stride2', 2, 128, 56, 56, 128, 3, 2, 1, 0, 1, 0, 1000),
    ('pair_1x1_residual', 8, 64, 56, 56, 256, 1, 1, 0, 1, 1, 0, 1000),
    ('pair_1x1_residual_bn128', 8, 64, 56, 56, 256, 1, 1, 0, 1, 1, 0, 1128),
    ('pair_fpn_topdown', 8, 256, 28, 28, 256, 1, 1, 0, 2, 0, 0, 1000),
    ('pair_fpn_topdown_56', 5, 256, 56, 56, 256, 1, 1, 0, 2, 0, 0, 1000),
    ('pair_fpn_topdown_nonsquare', 3, 512, 24, 40, 256, 1, 1, 0, 2, 0, 0, 1128),
    ('pair_odd_tiles', 1, 256, 30, 30, 128, 1, 1, 0, 0, 1, 0, 1000),          # M = 900: 3.5 pair tiles, last half empty
    ('pair_tiny_M', 1, 64, 3, 3, 64, 1, 1, 0, 0, 0, 0, 1000),                  # M = 9: the second CTA has no rows
    ('pair_bigK', 2, 2048, 14, 14, 512, 1, 1, 0, 0, 1, 0, 1000),
    ('pair_many_tiles', 16, 128, 56, 56, 128, 3, 1, 1, 0, 1, 0, 1000),         # > 2 tiles per pair, all buffers wrap
]


@pytest.mark.parametrize('engine', ['fp16x3', 'fp16c8', 'fp16'])
@pytest.mark.parametrize('case', PAIR_CASES, ids=[c[0] for c in PAIR_CASES])
def test_conv_parity_cta_pairs(engine, case):
    assert _run(engine, case) < TOL[engine]


def test_tensor_core_and_cuda_core_kernels_agree():
    """x3 tcgen05 vs fp32 FFMA on identical split-fp16 operands: only summation order differs."""
    from mcgaze_b200 import lib
    g = torch.Generator().manual_seed(5)
    x = torch.randn(4, 128, 28, 28, generator=g).cuda()
    w = (torch.randn(128, 128, 3, 3, generator=g) / 34.0).cuda()
    a = lib.debug_conv('fp16x3', x, w, 1, 1)
    b = lib.debug_conv('simt', x, w, 1, 1)
    assert (a - b).abs().max().item() < 3e-5 * b.abs().max().item()


def test_unsupported_shape_is_an_error_not_a_fallback():
    from mcgaze_b200 import lib
    x = torch.randn(1, 24, 8, 8).cuda()        # C not a multiple of 64
    w = torch.randn(64, 24, 1, 1).cuda()
    with pytest.raises(lib.McgError):
        lib.debug_conv('fp16x3', x, w, 1, 0)
