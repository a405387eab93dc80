// Test-time image pipeline on the GPU: decoded uint8 BGR frames in HBM ->
// CenterCrop window -> cv2.resize(INTER_LINEAR) fixed-point bilinear -> BGR->RGB -> (x - mean) / std -> zero pad to the
// batch canvas -> fp32 NCHW, i.e. exactly the tensor the reference's test_pipeline + collate hands to the forward
// (configs/_base_/datasets/gaze360.py:27-36; mmdet/datasets/pipelines/transforms.py Resize :213-241,
// Normalize :739-754, Pad :665-681, CenterCrop :1022-1047; formatting.py:96 HWC -> CHW).
//
// Bit-exact by construction: the resize is OpenCV's integer algorithm (11-bit coefficients, int32 horizontal pass,
// ((b*(r>>4))>>16 ... +2)>>2 vertical pass), the coefficient tables follow OpenCV's double / float sequence with
// explicitly rounded operations (no FMA contraction), and the normalisation is a 3 x 256 table evaluated in double like
// cv2.subtract / cv2.multiply do for a float32 array and a double scalar.
//
// HBM-bound byte work: per frame it reads the crop window once (uint8) and writes 3 * Hp * Wp floats.  One CTA
// handles kRows output rows of one frame: it builds the frame's x tap table and the normalisation table in shared
// memory once, then every thread produces 4 consecutive pixels of a row and stores one float4 per colour plane
// (fully coalesced 512-byte segments per warp and plane); the two taps of a pixel (6 source bytes per row) come from
// one or two aligned 8-byte loads instead of 6 byte loads (the byte-load version was L1-issue bound: ncu 77 % l1tex).
// Frame descriptors travel as a __grid_constant__ kernel parameter (no device-side table to allocate or copy;
// asynchronous on the caller's stream, graph-capturable).
#include <algorithm>

#include "common.cuh"
#include "../../include/mcgaze_b200.h"

namespace mcg {

constexpr int kPpThreads = 256;
constexpr int kPpRows = 32;           // output rows per CTA (amortises the per-CTA tap / normalisation tables)
constexpr int kPpMaxW = 4096;         // widest canvas (x tap table in dynamic shared memory: 8 bytes per column)
constexpr int kPpChunk = 512;         // frames per launch (descriptor block: 20 KB of the 32 KB parameter space)

struct PpFrame {
  const uint8_t* src;                 // crop origin
  long long stride;                   // bytes per source row
  int sh, sw;                         // crop window size
  int dh, dw;                         // resized size
  float* out;                         // [3, Hp, Wp] canvas of this frame
};
struct PpBatch {
  PpFrame f[kPpChunk];
  double mean[3], stdinv[3];
  int Hp, Wp, to_rgb, n;
};
static_assert(sizeof(PpBatch) <= 32000, "descriptor block exceeds the kernel parameter space (32764 bytes, CUDA >= 12.1)");

// OpenCV resize(): scale = 1 / (dst / src) in double (NOT src / dst)
__device__ __forceinline__ double linear_scale(int src, int dst) {
  return __ddiv_rn(1.0, __ddiv_rn(static_cast<double>(dst), static_cast<double>(src)));
}
// f = float((d + 0.5) * scale - 0.5); s = floor(f); f -= s
__device__ __forceinline__ void linear_tap(int d, double scale, int& s, float& f) {
  const double p = __dsub_rn(__dmul_rn(__dadd_rn(static_cast<double>(d), 0.5), scale), 0.5);
  const float pf = __double2float_rn(p);
  const float fl = floorf(pf);
  s = static_cast<int>(fl);
  f = __fsub_rn(pf, fl);
}
// saturate_cast<short>(c * INTER_RESIZE_COEF_SCALE) for c = 1 - f and f (round half to even)
__device__ __forceinline__ void linear_coef(float f, int& c0, int& c1) {
  c1 = __float2int_rn(__fmul_rn(f, 2048.f));
  c0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
}
// `need` (3 or 6) bytes starting at an arbitrary address, as (bytes 0..3, bytes 4..5 + don't-care), from one or two
// ALIGNED 8-byte loads.  An aligned word that holds at least one valid byte never crosses an allocation (or page)
// boundary, so the bytes around the frame that come along are readable; they are never used.
__device__ __forceinline__ uint2 load_bytes(const uint8_t* p, int need) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const unsigned s = static_cast<unsigned>(a) & 7u;
  const uint2* w = reinterpret_cast<const uint2*>(a - s);
  const uint2 lo = __ldg(w);
  uint2 hi = make_uint2(0u, 0u);
  if (s + need > 8) hi = __ldg(w + 1);
  // 12 candidate bytes lo.x lo.y hi.x (hi.y is only reached for s >= 7 with need == 6: then through the second funnel)
  const bool up = s >= 4;
  const unsigned w0 = up ? lo.y : lo.x, w1 = up ? hi.x : lo.y, w2 = up ? hi.y : hi.x;
  const unsigned sh = (s & 3u) * 8u;
  return make_uint2(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh));
}
// horizontal pass of one source row for one pixel: per channel c, L_c * a0 + R_c * a1 with (a0 | a1 << 16) packed;
// bytes: v.x = L_B L_G L_R R_B, v.y = R_G R_R . .
__device__ __forceinline__ void hpass(uint2 v, unsigned a01, unsigned (&h)[3]) {
  const unsigned bg = __byte_perm(v.x, v.y, 0x4130);   // L_B R_B L_G R_G
  const unsigned r_ = __byte_perm(v.x, v.y, 0x0052);   // L_R R_R . .
  h[0] = __dp2a_lo(a01, bg, 0u);
  h[1] = __dp2a_hi(a01, bg, 0u);
  h[2] = __dp2a_lo(a01, r_, 0u);
}

__global__ void __launch_bounds__(kPpThreads, 6) preprocess_kernel(const __grid_constant__ PpBatch b) {
  extern __shared__ int s_dyn[];      // 2 * Wp ints
  int* s_x0 = s_dyn;                  // byte offset of the left tap within a source row
  int* s_xa = s_dyn + b.Wp;           // (a1 << 16) | a0 ; bit 31: left tap is the last pixel of the row (a1 == 0)
  __shared__ float s_lut[3][256];     // output plane p, pixel value v -> normalised float
  __shared__ unsigned s_y[kPpRows][2];    // per output row of this CTA: y0 | y1 << 16 (source rows), b0 | b1 << 16
  const PpFrame& fr = b.f[blockIdx.y];
  const int Wp = b.Wp, Hp = b.Hp;
  const int tid = threadIdx.x;
  const int y_base = blockIdx.x * kPpRows;

  if (y_base < fr.dh) {               // CTAs entirely below the frame only write zeros
    const double scale_x = linear_scale(fr.sw, fr.dw);
    for (int x = tid; x < fr.dw; x += kPpThreads) {
      int sx;
      float fx;
      linear_tap(x, scale_x, sx, fx);
      if (sx < 0) {
        fx = 0.f;
        sx = 0;
      }
      if (sx >= fr.sw - 1) {
        fx = 0.f;
        sx = fr.sw - 1;
      }
      int a0, a1;
      linear_coef(fx, a0, a1);
      s_x0[x] = sx * 3;
      s_xa[x] = (a1 << 16) | a0 | (sx >= fr.sw - 1 ? 0x80000000 : 0);
    }
    if (tid < kPpRows && y_base + tid < fr.dh) {
      int sy, b0, b1;
      float fy;
      linear_tap(y_base + tid, linear_scale(fr.sh, fr.dh), sy, fy);
      linear_coef(fy, b0, b1);
      s_y[tid][0] = static_cast<unsigned>(min(max(sy, 0), fr.sh - 1)) | (static_cast<unsigned>(min(max(sy + 1, 0), fr.sh - 1)) << 16);
      s_y[tid][1] = static_cast<unsigned>(b0) | (static_cast<unsigned>(b1) << 16);
    }
    for (int i = tid; i < 3 * 256; i += kPpThreads) {
      const int p = i >> 8, v = i & 255;
      // cv2.subtract / cv2.multiply (float32 array, double scalar): evaluate in double, round to float per op
      const float t = __double2float_rn(__dsub_rn(static_cast<double>(v), b.mean[p]));
      s_lut[p][v] = __double2float_rn(__dmul_rn(static_cast<double>(t), b.stdinv[p]));
    }
  }
  __syncthreads();

  const int groups = Wp >> 2;         // float4 groups per row
  // source channel feeding output plane p (Normalize(to_rgb=True): plane 0 = R = BGR channel 2)
  const int sh0 = b.to_rgb ? 16 : 0, sh2 = b.to_rgb ? 0 : 16;
  const size_t plane = static_cast<size_t>(Hp) * Wp;
  for (int it = tid; it < kPpRows * groups; it += kPpThreads) {
    const int row = it / groups;
    const int y = y_base + row;
    const int x4 = (it - row * groups) << 2;
    if (y >= Hp) break;
    float o[3][4];
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
      for (int j = 0; j < 4; ++j) o[p][j] = 0.f;   // Pad(size_divisor) / collate: zeros right of and below the frame
    if (y < fr.dh && x4 < fr.dw) {
      const unsigned yy = s_y[row][0], bb = s_y[row][1];
      const unsigned b0s = bb << 16, b1s = bb & 0xffff0000u;   // b0 << 16, b1 << 16
      const uint8_t* r0 = fr.src + static_cast<long long>(yy & 0xffff) * fr.stride;
      const uint8_t* r1 = fr.src + static_cast<long long>(yy >> 16) * fr.stride;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int x = x4 + j;
        if (x < fr.dw) {
          const int off0 = s_x0[x];
          const int xa = s_xa[x];
          const unsigned a01 = static_cast<unsigned>(xa) & 0x7fffffffu;
          const int need = xa < 0 ? 3 : 6;         // last pixel of the row: the right tap has weight 0
          unsigned h0[3], h1[3];
          hpass(load_bytes(r0 + off0, need), a01, h0);
          hpass(load_bytes(r1 + off0, need), a01, h1);
          unsigned v = 0;                          // the resized pixel, BGR in bytes 0..2
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            // ((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2, the products as high halves of
            // (b << 16) * (h >> 4): everything is non-negative and < 2^32
            const unsigned r = (__umulhi(b0s, h0[c] >> 4) + __umulhi(b1s, h1[c] >> 4) + 2u) >> 2;
            v |= min(r, 255u) << (8 * c);
          }
          o[0][j] = s_lut[0][(v >> sh0) & 255];
          o[1][j] = s_lut[1][(v >> 8) & 255];
          o[2][j] = s_lut[2][(v >> sh2) & 255];
        }
      }
    }
    float* dst = fr.out + static_cast<size_t>(y) * Wp + x4;
#pragma unroll
    for (int p = 0; p < 3; ++p)
      __stcs(reinterpret_cast<float4*>(dst + p * plane), make_float4(o[p][0], o[p][1], o[p][2], o[p][3]));
  }
}

// Host side of mcg_preprocess (include/mcgaze_b200.h); throws CudaError, the C wrapper in mcg_api.cu translates.
void preprocess_launch(const mcg_frame* frames, int n, const float* mean, const float* std_, int to_rgb, float* out,
                       int Hp, int Wp, cudaStream_t st, int* launches) {
  MCG_CHECK(frames != nullptr && n > 0 && mean != nullptr && std_ != nullptr && out != nullptr, "null argument");
  MCG_CHECK(Hp > 0 && Wp > 0 && Wp % 4 == 0 && Wp <= kPpMaxW, "canvas width must be a multiple of 4 and <= 4096");
  MCG_CHECK((reinterpret_cast<uintptr_t>(out) & 15) == 0, "output canvas must be 16-byte aligned");
  const size_t frame_elems = static_cast<size_t>(3) * Hp * Wp;
  int count = 0;
  for (int base = 0; base < n; base += kPpChunk) {
    PpBatch b;
    b.n = std::min(kPpChunk, n - base);
    b.Hp = Hp;
    b.Wp = Wp;
    b.to_rgb = to_rgb ? 1 : 0;
    for (int p = 0; p < 3; ++p) {
      // Normalize.__init__ keeps float32 mean / std (transforms.py:735-736); mmcv.imnormalize widens them to double.
      // Plane p takes its statistics in output (RGB) order.
      b.mean[p] = static_cast<double>(mean[p]);
      b.stdinv[p] = 1.0 / static_cast<double>(std_[p]);
    }
    for (int i = 0; i < b.n; ++i) {
      const mcg_frame& f = frames[base + i];
      MCG_CHECK(f.src != nullptr && f.src_h > 0 && f.src_w > 0 && f.src_stride >= 3LL * f.src_w, "bad source frame");
      MCG_CHECK(f.crop_x >= 0 && f.crop_y >= 0 && f.crop_w > 0 && f.crop_h > 0 && f.crop_x + f.crop_w <= f.src_w &&
                    f.crop_y + f.crop_h <= f.src_h,
                "crop window outside the source frame");
      MCG_CHECK(f.crop_h <= 65535 && f.crop_w <= 65535, "crop window larger than 65535 pixels");
      MCG_CHECK(f.dst_h > 0 && f.dst_w > 0 && f.dst_h <= Hp && f.dst_w <= Wp, "resized frame does not fit the canvas");
      PpFrame& d = b.f[i];
      d.src = f.src + static_cast<long long>(f.crop_y) * f.src_stride + 3LL * f.crop_x;
      d.stride = f.src_stride;
      d.sh = f.crop_h;
      d.sw = f.crop_w;
      d.dh = f.dst_h;
      d.dw = f.dst_w;
      d.out = out + static_cast<size_t>(base + i) * frame_elems;
    }
    dim3 grid(static_cast<unsigned>((Hp + kPpRows - 1) / kPpRows), static_cast<unsigned>(b.n));
    preprocess_kernel<<<grid, kPpThreads, 2 * static_cast<size_t>(Wp) * sizeof(int), st>>>(b);
    MCG_CUDA(cudaGetLastError());
    ++count;
  }
  if (launches) *launches = count;
}

}  // namespace mcg
