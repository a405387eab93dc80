"""CPU-only numerics study for DESIGN.md section 9: could the two rounding corrections of the fp16c8 scheme run as
BLOCK-SCALED FP4 tensor-core MMAs (tcgen05 kind::mxf4 / mxf4nvf4: four times the fp16 MAC rate, so 1.5 instead of 2
fp16-MMA units per algorithmic MMA and 3 instead of 4 bytes per stored element)?

Every trunk / FPN convolution of the oracle (the stem keeps full precision, as in the engine) is replaced by the product
the tensor cores would form,

    A W  ~  A_hi W_hi  +  Qa(A_lo) Qw(W_hi)  +  Qa'(A_hi) Qw'(W_lo),      A = A_hi + A_lo, A_hi = fp16(A)

with the correction operands rounded by the scheme under test, and max |d(yaw, pitch)| over six clips is printed against
the fp64 oracle (the bar is 1e-3 rad):

    fp16     no corrections (the fast mode)
    c8       e4m3 operands with the engine's global power-of-two scales (common.cuh) - the shipped parity mode
    mxf4     e2m1 operands, one power-of-two (ue8m0) scale per 32 consecutive channels (kind::mxf4)
    nvf4     e2m1 operands, one e4m3 scale per 16 consecutive channels + a per-tensor fp32 scale (kind::mxf4nvf4)
    nvf4-fix the same with FIXED global power-of-two scales instead of the per-tensor one (what the engine would ship)
    c8/mxf4, mxf4/c8    one correction term in e4m3, the other in block-scaled fp4 (1.75 units)

usage: python tools/precision_fp4_study.py            (about two minutes on 8 cores)"""
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mcgaze_oracle as O  # noqa: E402

torch.set_num_threads(os.cpu_count() or 8)
GRID = torch.tensor([0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0])
MID = (GRID[1:] + GRID[:-1]) / 2


def e4m3(t):
    return t.clamp(-448.0, 448.0).to(torch.float8_e4m3fn).float()


def e2m1(t):
    """round |t| <= 6 to the e2m1 grid (ties to the even neighbour are irrelevant at this resolution)"""
    a = t.abs().clamp(max=6.0)
    return torch.sign(t) * GRID[torch.bucketize(a, MID)]


def fp4_block(t, block, scales):
    """t [.., C, ..] with the channel (K) dimension at dim 1: e2m1 values with one scale per `block` channels."""
    n, c = t.shape[0], t.shape[1]
    x = t.reshape(n, c // block, block, *t.shape[2:])
    amax = x.abs().amax(dim=2, keepdim=True).clamp(min=1e-30)
    if scales == 'pow2':                                   # ue8m0: the smallest power of two that maps the block into [-6, 6]
        s = torch.exp2(torch.ceil(torch.log2(amax / 6.0)))
    elif scales == 'e4m3':                                 # ue4m3 block scale relative to a per-tensor fp32 scale
        g = t.abs().max().clamp(min=1e-30) / (6.0 * 448.0)
        s = e4m3(amax / 6.0 / g).clamp(min=2.0 ** -9) * g
    else:                                                  # ue4m3 block scale times a FIXED global power of two (`scales` =
        g = 2.0 ** -scales                                 # its exponent): what an engine with compile-time scales would do,
        s = e4m3(amax / 6.0 / g).clamp(min=2.0 ** -9) * g  # the factor leaves through tcgen05.mma's scale-input-d like 2^15 does today
    return (e2m1(x / s) * s).reshape(t.shape)


def corr_c8(al, ah, wl, wh):
    """the shipped scheme: lo8 = e4m3(A_lo 2^11), hi8 = e4m3(A_hi), W hi8 = e4m3(W_hi 2^4), W lo8 = e4m3(W_lo 2^15)"""
    return (lambda: (e4m3(al * 2048.0) / 2048.0, e4m3(wh * 16.0) / 16.0)), (lambda: (e4m3(ah), e4m3(wl * 32768.0) / 32768.0))


def corr_fp4(block, scales):
    def make(al, ah, wl, wh):
        q = lambda t: fp4_block(t, block, scales)  # noqa: E731
        return (lambda: (q(al), q(wh))), (lambda: (q(ah), q(wl)))
    return make


def corr_nvf4_fixed(al, ah, wl, wh):
    """nvf4 with global power-of-two scale exponents chosen like common.cuh chooses the e4m3 ones: block scale x 2^e stays a
    normal e4m3 number for |activation| in [2^-6, 448] / |weight| in [2^-10, 28]"""
    qa_lo = lambda t: fp4_block(t, 16, 14)   # noqa: E731   A_lo <= 2^-11 * 448
    qa_hi = lambda t: fp4_block(t, 16, 2)    # noqa: E731
    qw_hi = lambda t: fp4_block(t, 16, 6)    # noqa: E731
    qw_lo = lambda t: fp4_block(t, 16, 17)   # noqa: E731
    return (lambda: (qa_lo(al), qw_hi(wh))), (lambda: (qa_hi(ah), qw_lo(wl)))


SCHEMES = {
    'fp16': None,
    'c8': (corr_c8, corr_c8),
    'mxf4': (corr_fp4(32, 'pow2'), corr_fp4(32, 'pow2')),
    'nvf4': (corr_fp4(16, 'e4m3'), corr_fp4(16, 'e4m3')),
    'nvf4-fix': (corr_nvf4_fixed, corr_nvf4_fixed),        # ... with fixed global power-of-two scale exponents
    'c8/mxf4': (corr_c8, corr_fp4(32, 'pow2')),            # activation-rounding term in e4m3, weight-rounding term in fp4
    'mxf4/c8': (corr_fp4(32, 'pow2'), corr_c8),
    'c8/nvf4': (corr_c8, corr_fp4(16, 'e4m3')),
    'nvf4/c8': (corr_fp4(16, 'e4m3'), corr_c8),
}


def make_conv(scheme):
    def conv(x, w, b, stride, pad, hk):
        if x.dtype != torch.float32 or x.shape[1] % 32 != 0:           # fp64 reference run / the stem
            return F.conv2d(x, w, b, stride=stride, padding=pad)
        ah, wh = x.half().float(), w.half().float()
        y = F.conv2d(ah, wh, b, stride=stride, padding=pad)
        if scheme is None:
            return y
        al, wl = x - ah, w - wh
        t1 = scheme[0](al, ah, wl, wh)[0]()
        t2 = scheme[1](al, ah, wl, wh)[1]()
        return y + F.conv2d(t1[0], t1[1], None, stride=stride, padding=pad) + F.conv2d(t2[0], t2[1], None, stride=stride, padding=pad)
    return conv


def main():
    sd = O.make_state_dict(0)
    sd64 = {k: v.double() for k, v in sd.items()}
    clips = [O.make_clip(s, 7) for s in (1, 2, 3, 4, 5, 6)]
    plain = O._conv
    ref = [O.forward(sd64, c.double()) for c in clips]
    keys = [k for k in ref[0] if 'gaze' in k]
    print(f'{"scheme":10s} {"max |d(yaw,pitch)| rad":>24s}   per clip')
    for name, scheme in SCHEMES.items():
        O._conv = make_conv(scheme)
        t0 = time.time()
        errs = []
        try:
            for c, r in zip(clips, ref):
                o = O.forward(sd, c)
                errs.append(max((O.vector_to_yaw_pitch(o[k].double()) - O.vector_to_yaw_pitch(r[k])).abs().max().item() for k in keys))
        finally:
            O._conv = plain
        print(f'{name:10s} {max(errs):24.3e}   {" ".join(f"{e:.2e}" for e in errs)}   ({time.time() - t0:.0f} s)', flush=True)


if __name__ == '__main__':
    main()
