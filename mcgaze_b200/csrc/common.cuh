// Shared host/device definitions for the MCGaze B200 kernels.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdexcept>
#include <string>

namespace mcg {

// Split tensor: value = hi + lo.  hi is fp16.  The low part is stored either as fp16 (`lo`, mode
// fp16x3) or as e4m3 fp8 of (value - hi) * 2^kLo8Shift (`lo8`, mode fp16c8: 3 bytes / element and
// fp8 tensor-core correction terms); both null means single-fp16 ("fast") storage.
// Layout of every activation is NHWC (channels innermost), i.e. a row-major [pixels, C] matrix.
struct Planes {
  __half* hi = nullptr;
  __half* lo = nullptr;
  uint8_t* lo8 = nullptr;
  uint8_t* hi8 = nullptr;  // e4m3(hi): second fp8 operand of "T" layers (fp16c8 mode)
};

// fp16c8 scales (all powers of two, global constants).  |value - hi| <= 2^-11 |value|, so lo8 = e4m3(lo 2^11)
// stays below the e4m3 maximum (448) wherever hi8 = e4m3(hi) does, and is a normal e4m3 number for |value| >= 2^-6.
// Weights: hi8 = e4m3(W_hi 2^4), lo8 = e4m3(W_lo 2^15) (normal for |W| >= 2^-10, saturating at |W| >= 28).
// Both correction products then carry 2^15 = 2^kC8AccShift, the largest factor tcgen05.mma's scale-input-d removes.
constexpr int kLo8Shift = 11;
constexpr float kLo8Scale = 2048.f;          // 2^11
constexpr float kLo8InvScale = 1.f / 2048.f;
constexpr int kC8AccShift = 15;
constexpr int kW8HiShift = kC8AccShift - kLo8Shift;  // 4
constexpr int kW8LoShift = kC8AccShift;              // 15 (the activation's hi8 plane is unscaled)

__host__ __device__ inline uint8_t float_to_e4m3(float v) {
  return static_cast<uint8_t>(__nv_cvt_float_to_fp8(v, __NV_SATFINITE, __NV_E4M3));
}
__host__ __device__ inline float e4m3_to_float(uint8_t b) {
  const __half_raw hr = __nv_cvt_fp8_to_halfraw(static_cast<__nv_fp8_storage_t>(b), __NV_E4M3);
  return __half2float(__half(hr));
}

enum ResMode : int { RES_NONE = 0, RES_SAME = 1, RES_UP2X = 2 };

// Fused GEMM / conv epilogue:  y = acc + bias[n] (+ residual[row', n]) ; optional ReLU ;
// written either as split-fp16 planes or as fp32.  For RES_UP2X the residual lives on a
// grid of (P/2, Q/2) pixels per frame and is read with nearest-neighbour 2x upsampling
// (FPN top-down add, mmdet/models/necks/fpn.py:165-174).
struct Epilogue {
  const float* bias = nullptr;   // [N] or null
  const __half* res_hi = nullptr;
  const __half* res_lo = nullptr;
  const uint8_t* res_lo8 = nullptr;
  const float* res_f32 = nullptr;  // fp32 residual (head); used instead of res_hi/res_lo when set
  int res_mode = RES_NONE;
  int relu = 0;
  __half* out_hi = nullptr;
  __half* out_lo = nullptr;
  uint8_t* out_lo8 = nullptr;
  uint8_t* out_hi8 = nullptr;    // optional e4m3 copy of out_hi (fp16c8: input of a tensor-bound consumer)
  float* out_f32 = nullptr;      // if non-null, fp32 output instead of planes
  long long ldo = 0;             // output row stride (elements)
  long long ldr = 0;             // residual row stride (elements)
  int P = 0, Q = 0;              // output spatial size (RES_UP2X only)
};

// Geometry of the A operand.  kind 0: plain row-major [M, K] matrix with row stride lda.
// kind 1: implicit im2col over an NHWC tensor [NB, H, W, C]; K index = (r*S + s)*C + c.
struct AGeom {
  int kind = 0;
  long long lda = 0;
  int NB = 0, H = 0, W = 0, C = 0, R = 1, S = 1, stride = 1, pad = 0, P = 0, Q = 0;
};

__host__ __device__ inline long long res_row(const Epilogue& e, long long m) {
  if (e.res_mode != RES_UP2X) return m;
  long long pq = static_cast<long long>(e.P) * e.Q;
  long long n = m / pq;
  int rem = static_cast<int>(m - n * pq);
  int p = rem / e.Q, q = rem - p * e.Q;
  return (n * (e.P / 2) + (p >> 1)) * (e.Q / 2) + (q >> 1);
}

// first pixel of the source ROW that res_row(e, m) lies in: no pixel at or after m maps to an earlier source pixel, so
// it is the start of the window of source pixels a tile beginning at m needs (RES_UP2X prefetch)
__host__ __device__ inline long long res_row_window_start(const Epilogue& e, long long m) {
  long long pq = static_cast<long long>(e.P) * e.Q;
  long long n = m / pq;
  int rem = static_cast<int>(m - n * pq);
  int p = rem / e.Q;
  return (n * (e.P / 2) + (p >> 1)) * (e.Q / 2);
}

__device__ __forceinline__ void split_store(float v, __half* hi, __half* lo, long long idx, uint8_t* lo8 = nullptr,
                                            uint8_t* hi8 = nullptr) {
  __half h = __float2half_rn(v);
  hi[idx] = h;
  if (lo) lo[idx] = __float2half_rn(v - __half2float(h));
  if (lo8) lo8[idx] = float_to_e4m3((v - __half2float(h)) * kLo8Scale);
  if (hi8) hi8[idx] = float_to_e4m3(__half2float(h));
}

// 8 consecutive e4m3 values (one uint2) -> 8 floats (unscaled)
__device__ __forceinline__ void e4m3x8_to_float(const uint2& u, float (&f)[8]) {
  const uint16_t* p = reinterpret_cast<const uint16_t*>(&u);
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const __half2_raw hr = __nv_cvt_fp8x2_to_halfraw2(static_cast<__nv_fp8x2_storage_t>(p[t]), __NV_E4M3);
    const float2 v = __half22float2(__half2(hr));
    f[2 * t] = v.x;
    f[2 * t + 1] = v.y;
  }
}
// 8 floats -> 8 e4m3 (saturating), packed in a uint2
__device__ __forceinline__ uint2 float8_to_e4m3x8(const float (&f)[8]) {
  uint2 u;
  uint16_t* p = reinterpret_cast<uint16_t*>(&u);
#pragma unroll
  for (int t = 0; t < 4; ++t)
    p[t] = static_cast<uint16_t>(__nv_cvt_float2_to_fp8x2(make_float2(f[2 * t], f[2 * t + 1]), __NV_SATFINITE, __NV_E4M3));
  return u;
}

struct CudaError : std::runtime_error {
  explicit CudaError(const std::string& s) : std::runtime_error(s) {}
};

#define MCG_CUDA(call)                                                                               \
  do {                                                                                               \
    cudaError_t _e = (call);                                                                         \
    if (_e != cudaSuccess)                                                                           \
      throw ::mcg::CudaError(std::string(#call) + " failed: " + cudaGetErrorString(_e) + " at " +   \
                             __FILE__ + ":" + std::to_string(__LINE__));                            \
  } while (0)

#define MCG_CHECK(cond, msg)                                                                         \
  do {                                                                                               \
    if (!(cond))                                                                                     \
      throw ::mcg::CudaError(std::string("check failed: ") + #cond + " : " + (msg) + " at " +        \
                             __FILE__ + ":" + std::to_string(__LINE__));                            \
  } while (0)

}  // namespace mcg
