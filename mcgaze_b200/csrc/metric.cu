// Gaze360 scorer on the GPU: the reference's tools/calculate_mae_gaze360.py:gaze_error
// (:110-188) with smooth_filter (:16-29), compute_angular_error (:77-94) and compute_yaw_angular (:69-74) as one fused
// kernel over per-frame predictions that are already on the device (the outputs of mcg_forward after the overlap merge).
//
// One CTA per video.  Per frame: temporal smoothing of the prediction with its neighbours (alpha = 0.6, re-normalised;
// videos of one frame are left alone), angle to the normalised ground truth, yaw of the ground truth -> the three
// categories 360 / front-180 (|yaw| <= 90 deg) / front-20.  Per video and category the MEAN angle in degrees is weighted
// by its frame count (:157-158 and below), summed over the videos in double with atomics.  fp32 arithmetic in the
// reference's operation order (explicitly rounded products: no FMA contraction); the reduction order inside a video
// differs from torch's, which moves the result by ~1e-6 relative.
#include <math.h>

#include "common.cuh"
#include "../../include/mcgaze_b200.h"

namespace mcg {

constexpr int kGeThreads = 128;

__device__ __forceinline__ void load3(const float* p, long long i, float (&v)[3]) {
  v[0] = p[3 * i];
  v[1] = p[3 * i + 1];
  v[2] = p[3 * i + 2];
}

__global__ void __launch_bounds__(kGeThreads) gaze_error_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                               const int* __restrict__ video_start, double* out) {
  const int v = blockIdx.x;
  const long long s = video_start[v];
  const int L = video_start[v + 1] - video_start[v];
  const float alpha = 0.6f, beta = 0.4f;   // python's (1 - 0.6) rounds to 0.4f when it meets a float32 tensor
  float sum[3] = {0.f, 0.f, 0.f};
  int cnt[3] = {0, 0, 0};
  for (int t = threadIdx.x; t < L; t += kGeThreads) {
    float p[3];
    load3(pred, s + t, p);
    if (L >= 2) {                          // smooth_filter
      float nb[3];
      if (t == 0) {
        load3(pred, s + 1, nb);
      } else if (t == L - 1) {
        load3(pred, s + L - 2, nb);
      } else {
        float a[3], b[3];
        load3(pred, s + t - 1, a);
        load3(pred, s + t + 1, b);
#pragma unroll
        for (int k = 0; k < 3; ++k) nb[k] = __fadd_rn(a[k], b[k]);
      }
      const bool mid = t != 0 && t != L - 1;
      float o[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        float add = __fmul_rn(beta, nb[k]);
        if (mid) add = add / 2.f;          // (1 - alpha) * (prev + next) / 2
        o[k] = __fadd_rn(__fmul_rn(alpha, p[k]), add);
      }
      const float n = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(o[0], o[0]), __fmul_rn(o[1], o[1])), __fmul_rn(o[2], o[2])));
#pragma unroll
      for (int k = 0; k < 3; ++k) p[k] = o[k] / n;
    }
    float g[3];
    load3(gt, s + t, g);
    const float gn = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(g[0], g[0]), __fmul_rn(g[1], g[1])), __fmul_rn(g[2], g[2])));
#pragma unroll
    for (int k = 0; k < 3; ++k) g[k] = g[k] / gn;
    const float dot = __fadd_rn(__fadd_rn(__fmul_rn(g[0], p[0]), __fmul_rn(g[1], p[1])), __fmul_rn(g[2], p[2]));
    const float err = acosf(dot);          // NaN when rounding pushes the dot product past 1, exactly like torch.acos
    const float yaw = 180.f * fabsf(atan2f(g[0], -g[2])) / 3.14159265358979323846f;
    sum[0] += err;
    cnt[0] += 1;
    if (yaw <= 90.f) {
      sum[1] += err;
      cnt[1] += 1;
    }
    if (yaw <= 20.f) {
      sum[2] += err;
      cnt[2] += 1;
    }
  }
  __shared__ float s_sum[3][kGeThreads / 32];
  __shared__ int s_cnt[3][kGeThreads / 32];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sum[c] += __shfl_xor_sync(0xffffffffu, sum[c], o);
      cnt[c] += __shfl_xor_sync(0xffffffffu, cnt[c], o);
    }
    if ((threadIdx.x & 31) == 0) {
      s_sum[c][threadIdx.x >> 5] = sum[c];
      s_cnt[c][threadIdx.x >> 5] = cnt[c];
    }
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    const int c = threadIdx.x;
    float ts = 0.f;
    int tc = 0;
    for (int w = 0; w < kGeThreads / 32; ++w) {
      ts += s_sum[c][w];
      tc += s_cnt[c][w];
    }
    if (tc > 0) {
      const float mean_deg = 180.f * (ts / static_cast<float>(tc)) / 3.14159265358979323846f;   // 180 * mean / pi
      atomicAdd(out + 2 * c, static_cast<double>(mean_deg) * tc);
      atomicAdd(out + 2 * c + 1, static_cast<double>(tc));
    }
  }
}

// Host side of mcg_gaze_error; throws CudaError, the C wrapper in mcg_api.cu translates.
void gaze_error_launch(const float* pred, const float* gt, const int32_t* video_start, int n_videos, double* out,
                       cudaStream_t st) {
  MCG_CHECK(pred != nullptr && gt != nullptr && video_start != nullptr && out != nullptr && n_videos > 0, "null argument");
  MCG_CUDA(cudaMemsetAsync(out, 0, 6 * sizeof(double), st));
  gaze_error_kernel<<<n_videos, kGeThreads, 0, st>>>(pred, gt, video_start, out);
  MCG_CUDA(cudaGetLastError());
}

}  // namespace mcg
