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

// variant 0 = tools/calculate_mae_gaze360.py, 1 = tools/calculate_mae_l2cs.py: the front-20 class of the l2cs scorer
// also needs |pitch(gt)| <= 20 deg (:139); everything else (incl. the smoothing, which the l2cs script applies despite
// its comment, :124) is identical.  Its ground truth lives at annotations[3 * video] (:110): a host-side indexing matter.
__global__ void __launch_bounds__(kGeThreads) gaze_error_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                               const int* __restrict__ video_start, int variant,
                                                               double* out) {
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
    bool front20 = yaw <= 20.f;
    if (variant == 1) front20 = front20 && 180.f * fabsf(asinf(g[1])) / 3.14159265358979323846f <= 20.f;
    if (front20) {
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
void gaze_error_launch(const float* pred, const float* gt, const int32_t* video_start, int n_videos, int variant,
                       double* out, cudaStream_t st) {
  MCG_CHECK(pred != nullptr && gt != nullptr && video_start != nullptr && out != nullptr && n_videos > 0, "null argument");
  MCG_CHECK(variant == MCG_SCORER_GAZE360 || variant == MCG_SCORER_L2CS, "unknown scorer variant");
  MCG_CUDA(cudaMemsetAsync(out, 0, 6 * sizeof(double), st));
  gaze_error_kernel<<<n_videos, kGeThreads, 0, st>>>(pred, gt, video_start, variant, out);
  MCG_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------------------------
// Overlap merge of the per-clip results of a video on the device (the reference's
// tools/test_gaze360_gaze.py:129-201 does it clip by clip with torch ops on the host side of the loop).
// The reference appends clip after clip: frames a clip adds are copied (boxes zeroed where that clip's score < 0.5),
// frames it shares with what is already there are averaged with it - box coordinates zeroed when the running (already
// averaged) score or the clip's own score is < 0.5, scores and gaze vectors averaged without re-normalisation.  Every
// frame is covered by at most three clips (regular windows i with stride*i <= f < stride*i + clip_len, plus the
// right-aligned last window), and the sequential update of a frame only involves the clips that cover it, in clip
// order: one thread folds one frame.
//   rows       [n_clips_total, clip_len, 27]   per clip and frame: boxes [3,4], scores [3], gaze [4,3]
//   clip_start [n_videos + 1]                  first clip of every video in `rows`
//   frame_start[n_videos + 1]                  first frame of every video in the outputs
//   det [F,3,5] (x1,y1,x2,y2,score)   gaze [F,4,3]
// ---------------------------------------------------------------------------------------------------------------
constexpr int kRowFloats = 27;

__global__ void __launch_bounds__(128) merge_clips_kernel(const float* __restrict__ rows, const int* __restrict__ clip_start,
                                                         const int* __restrict__ frame_start, int n_videos, int clip_len,
                                                         int stride, float* __restrict__ det, float* __restrict__ gaze) {
  const int v = blockIdx.x;
  const int f0 = frame_start[v], L = frame_start[v + 1] - frame_start[v];
  const int c0 = clip_start[v], nclips = clip_start[v + 1] - clip_start[v];
  for (int f = threadIdx.x; f < L; f += blockDim.x) {
    float box[12], sc[3], gz[12];
    bool have = false;
    auto fold = [&](int clip, int t) {
      const float* r = rows + (static_cast<long long>(c0 + clip) * clip_len + t) * kRowFloats;
      float cb[12], cs[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        cs[c] = r[12 + c];
#pragma unroll
        for (int k = 0; k < 4; ++k) cb[c * 4 + k] = cs[c] < 0.5f ? 0.f : r[c * 4 + k];   // :135-141 / :188-194
      }
      if (!have) {
#pragma unroll
        for (int k = 0; k < 12; ++k) {
          box[k] = cb[k];
          gz[k] = r[15 + k];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) sc[c] = cs[c];
        have = true;
        return;
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const bool bad = sc[c] < 0.5f || cs[c] < 0.5f;                                 // :170-174
#pragma unroll
        for (int k = 0; k < 4; ++k) box[c * 4 + k] = bad ? 0.f : (box[c * 4 + k] + cb[c * 4 + k]) / 2.f;
        sc[c] = (sc[c] + cs[c]) / 2.f;
      }
#pragma unroll
      for (int k = 0; k < 12; ++k) gz[k] = (gz[k] + r[15 + k]) / 2.f;                 // :182-183, not re-normalised
    };
    if (nclips == 1) {
      fold(0, f);
    } else {
      // regular windows 0 .. nclips - 2 start at stride * i; the last one is right-aligned
      int lo = f - (clip_len - 1);
      lo = lo <= 0 ? 0 : (lo + stride - 1) / stride;
      int hi = f / stride;
      if (hi > nclips - 2) hi = nclips - 2;
      for (int i = lo; i <= hi; ++i) fold(i, f - stride * i);
      if (f >= L - clip_len) fold(nclips - 1, f - (L - clip_len));
    }
    float* d = det + static_cast<long long>(f0 + f) * 15;
    float* g = gaze + static_cast<long long>(f0 + f) * 12;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int k = 0; k < 4; ++k) d[c * 5 + k] = box[c * 4 + k];
      d[c * 5 + 4] = sc[c];
    }
#pragma unroll
    for (int k = 0; k < 12; ++k) g[k] = gz[k];
  }
}

void merge_clips_launch(const float* rows, const int32_t* clip_start, const int32_t* frame_start, int n_videos, int clip_len,
                        int stride, float* det, float* gaze, cudaStream_t st) {
  MCG_CHECK(rows && clip_start && frame_start && det && gaze && n_videos > 0, "null argument");
  MCG_CHECK(clip_len >= 1 && stride >= 1 && stride <= clip_len, "need 1 <= stride <= clip_len");
  merge_clips_kernel<<<n_videos, 128, 0, st>>>(rows, clip_start, frame_start, n_videos, clip_len, stride, det, gaze);
  MCG_CUDA(cudaGetLastError());
}

}  // namespace mcg
