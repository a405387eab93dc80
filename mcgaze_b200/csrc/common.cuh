// Shared host/device definitions for the MCGaze B200 kernels.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdexcept>
#include <string>

namespace mcg {

// Split-fp16 tensor: value = hi + lo.  `lo == nullptr` means single-fp16 ("fast") storage.
// Layout of every activation is NHWC (channels innermost), i.e. a row-major [pixels, C] matrix.
struct Planes {
  __half* hi = nullptr;
  __half* lo = nullptr;
};

enum ResMode : int { RES_NONE = 0, RES_SAME = 1, RES_UP2X = 2 };

// Fused GEMM / conv epilogue:  y = acc + bias[n] (+ residual[row', n]) ; optional ReLU ;
// written either as split-fp16 planes or as fp32.  For RES_UP2X the residual lives on a
// grid of (P/2, Q/2) pixels per frame and is read with nearest-neighbour 2x upsampling
// (FPN top-down add, mmdet/models/necks/fpn.py:165-174).
struct Epilogue {
  const float* bias = nullptr;   // [N] or null
  const __half* res_hi = nullptr;
  const __half* res_lo = nullptr;
  const float* res_f32 = nullptr;  // fp32 residual (head); used instead of res_hi/res_lo when set
  int res_mode = RES_NONE;
  int relu = 0;
  __half* out_hi = nullptr;
  __half* out_lo = nullptr;
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

__device__ __forceinline__ void split_store(float v, __half* hi, __half* lo, long long idx) {
  __half h = __float2half_rn(v);
  hi[idx] = h;
  if (lo) lo[idx] = __float2half_rn(v - __half2float(h));
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
