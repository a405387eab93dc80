// fp32 CUDA-core GEMM / implicit-conv kernel with the same operand formats and fused epilogue
// as the tcgen05 kernel.  It is (a) the on-device cross-check for the tensor-core kernel,
// (b) the path for shapes tcgen05 tiles do not cover (N not a multiple of 64: the 1/3/4-wide
// classifier / regressor / gaze outputs), and (c) the "simt" precision mode (pure fp32 FMA).
#pragma once
#include "common.cuh"

namespace mcg {

struct SimtParams {
  long long M = 0;
  int N = 0, K = 0;
  AGeom a;
  // A operand: fp32 matrix (kind 0 only) or split-fp16 planes (kind 0 or 1)
  const float* a_f32 = nullptr;
  const __half* a_hi = nullptr;
  const __half* a_lo = nullptr;
  const uint8_t* a_lo8 = nullptr;
  // W operand [N, K] row-major: fp32, or split-fp16 planes
  const float* w_f32 = nullptr;
  const __half* w_hi = nullptr;
  const __half* w_lo = nullptr;
  Epilogue ep;
};

constexpr int kSimtBM = 64, kSimtBN = 64, kSimtBK = 16;

__device__ __forceinline__ void epilogue_store(const Epilogue& ep, long long m, int n, float v, long long rrow) {
  if (ep.bias) v += ep.bias[n];
  if (ep.res_mode != RES_NONE) {
    const long long ri = rrow * ep.ldr + n;
    if (ep.res_f32) {
      v += ep.res_f32[ri];
    } else {
      v += __half2float(ep.res_hi[ri]);
      if (ep.res_lo) v += __half2float(ep.res_lo[ri]);
      if (ep.res_lo8) v += e4m3_to_float(ep.res_lo8[ri]) * kLo8InvScale;
    }
  }
  if (ep.relu) v = fmaxf(v, 0.f);
  const long long oi = m * ep.ldo + n;
  if (ep.out_f32)
    ep.out_f32[oi] = v;
  else
    split_store(v, ep.out_hi, ep.out_lo, oi, ep.out_lo8, ep.out_hi8);
}

__global__ void __launch_bounds__(256) simt_gemm_kernel(const SimtParams p) {
  __shared__ float As[kSimtBK][kSimtBM + 4];
  __shared__ float Ws[kSimtBK][kSimtBN + 4];
  const int tid = threadIdx.x;
  const long long m0 = static_cast<long long>(blockIdx.x) * kSimtBM;
  const int n0 = blockIdx.y * kSimtBN;

  // loader mapping: 64 rows x 4 groups of 4 consecutive k
  const int lrow = tid >> 2;
  const int lk = (tid & 3) * 4;
  const long long am = m0 + lrow;
  const bool a_ok = am < p.M;
  int img_n = 0, base_h = 0, base_w = 0;
  if (p.a.kind == 1 && a_ok) {
    const long long pq = static_cast<long long>(p.a.P) * p.a.Q;
    img_n = static_cast<int>(am / pq);
    const int rem = static_cast<int>(am - img_n * pq);
    const int pp = rem / p.a.Q;
    base_h = pp * p.a.stride - p.a.pad;
    base_w = (rem - pp * p.a.Q) * p.a.stride - p.a.pad;
  }
  const int wn = n0 + lrow;
  const bool w_ok = wn < p.N;

  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4] = {};

  for (int k0 = 0; k0 < p.K; k0 += kSimtBK) {
    const int k = k0 + lk;
    float av[4] = {0.f, 0.f, 0.f, 0.f};
    if (a_ok && k < p.K) {
      long long idx = -1;
      if (p.a.kind == 0) {
        idx = am * p.a.lda + k;
      } else {
        const int tap = k / p.a.C;
        const int c = k - tap * p.a.C;
        const int r = tap / p.a.S;
        const int s = tap - r * p.a.S;
        const int h = base_h + r, w = base_w + s;
        if (h >= 0 && h < p.a.H && w >= 0 && w < p.a.W)
          idx = ((static_cast<long long>(img_n) * p.a.H + h) * p.a.W + w) * p.a.C + c;
      }
      if (idx >= 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (k + j < p.K) {
            if (p.a_f32)
              av[j] = p.a_f32[idx + j];
            else
              av[j] = __half2float(p.a_hi[idx + j]) + (p.a_lo ? __half2float(p.a_lo[idx + j]) : 0.f) +
                      (p.a_lo8 ? e4m3_to_float(p.a_lo8[idx + j]) * kLo8InvScale : 0.f);
          }
        }
      }
    }
    float wv[4] = {0.f, 0.f, 0.f, 0.f};
    if (w_ok && k < p.K) {
      const long long widx = static_cast<long long>(wn) * p.K + k;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (k + j < p.K) {
          if (p.w_f32)
            wv[j] = p.w_f32[widx + j];
          else
            wv[j] = __half2float(p.w_hi[widx + j]) + (p.w_lo ? __half2float(p.w_lo[widx + j]) : 0.f);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      As[lk + j][lrow] = av[j];
      Ws[lk + j][lrow] = wv[j];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kSimtBK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 w4 = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
    const long long rrow = p.ep.res_mode != RES_NONE ? res_row(p.ep, m) : 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < p.N) epilogue_store(p.ep, m, n, acc[i][j], rrow);
    }
  }
}

inline void launch_simt_gemm(const SimtParams& p, cudaStream_t stream) {
  dim3 grid(static_cast<unsigned>((p.M + kSimtBM - 1) / kSimtBM), static_cast<unsigned>((p.N + kSimtBN - 1) / kSimtBN));
  simt_gemm_kernel<<<grid, 256, 0, stream>>>(p);
  MCG_CUDA(cudaGetLastError());
}

}  // namespace mcg
