// tcgen05 / TMA / TMEM implicit-GEMM kernel for sm_100a.
//
//   D[m, n] = sum_k A[m, k] * W[n, k]   (+ fused epilogue, see common.cuh)
//
// A is either a plain row-major [M, K] split-fp16 matrix (1x1 stride-1 convolutions, the head's
// Linear layers) or the implicit im2col view of an NHWC activation (3x3 / strided convolutions),
// fetched by TMA in im2col mode.  W is the [N, K] K-major packed weight (BN folded).
//
// Precision: operands are split-fp16 (value = hi + lo).  kTerms == 1 issues one MMA per k-step
// (hi*hi, "fast" mode); kTerms == 3 issues lo*hi + hi*lo + hi*hi into the same fp32 TMEM
// accumulator (~22-bit operand mantissa: fp32-equivalent products, the mode that meets the
// 1e-3 (yaw,pitch) parity bar against the fp32 oracle).
// kTerms == 2 ("fp16c8") keeps hi*hi in fp16 and moves the two rounding corrections to e4m3 tensor-core
// MMAs (kind::f8f6f4, K = 32 per instruction: twice the fp16 rate), all into ONE accumulator:
//     A*W ~= A_hi*W_hi + 2^-15 [ e4m3(A_lo 2^11) * e4m3(W_hi 2^4) + e4m3(A_hi) * e4m3(W_lo 2^15) ]
// Each tile runs two passes over K through the same smem ring: pass 1 streams the e4m3 tiles (A lo8 | A hi8,
// W hi8 | W lo8; same bytes per stage as the fp16 tiles) and accumulates the corrections scaled by 2^15; the
// first fp16 MMA of pass 2 rescales the accumulator (tcgen05.mma scale-input-d = 15: D = A*B + D 2^-15) and
// pass 2 adds hi*hi.  8 MMA units per 64-wide k-block instead of 12, a single accumulator (block_n up to 256,
// double buffered), the plain epilogue.  Activations are stored as fp16 hi + e4m3 lo8 + e4m3 hi8 (e4m3(hi), the
// operand of the weight-rounding correction).
//
// Structure (persistent, one CTA per SM, 384 threads):
//   warp 0    : TMA producer  (one elected lane; A tile 128 x 64, W tile block_n x 64, SWIZZLE_128B)
//   warp 1    : MMA issuer    (one lane issues tcgen05.mma M=128, N=block_n, K=16; commit -> mbarrier)
//   warp 2    : TMEM allocator (512 columns = 2 accumulator buffers of 256 fp32 columns, or 4 of 128)
//   warps 4-11: two independent epilogue groups of four warps.  Group g drains the accumulators of the
//               CTA's tiles g, g+2, g+4, ... so two tiles are in the epilogue at once: with one warp per
//               SM sub-partition the epilogue of the HBM-bound layers was issue/latency bound.  Inside a
//               group each warp owns 32 accumulator rows (its TMEM lane quarter) and the group walks the
//               tile in 32-column chunks: tcgen05.ld -> + bias -> + residual -> ReLU -> split to fp16
//               hi/lo -> 64B-swizzled smem staging -> one TMA store per plane ([128 x 32] box, bulk
//               async group, double buffered).  A same-shape residual (bottleneck identity) is
//               TMA-prefetched kResBufs chunks ahead into smem, continuously across tile boundaries.
//               fp32 outputs (head) and the FPN's 2x-upsampled residual use direct vector accesses.
// Pipelines: smem full/empty ring (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue groups), per-group
// residual mbarriers.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"

namespace mcg {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // 64 fp16 = 128 B = one swizzle row
constexpr int kUmmaK = 16;
constexpr int kGemmThreads = 384;
constexpr int kEpiGroups = 2;
constexpr int kTmemCols = 512;
constexpr int kMaxStages = 8;
constexpr int kMaxAcc = 4;
constexpr int kATileBytes = kBlockM * kBlockK * 2;  // 16 KB
constexpr int kSmemBarrierBytes = 1024;
constexpr int kMaxDynSmem = 232448;  // 227 KB: the sm_100 opt-in limit per block
constexpr int kEpiChunk = 32;        // columns per epilogue chunk (64 B of fp16 per row)
constexpr int kEpiPlaneBytes = kBlockM * kEpiChunk * 2;  // [128 rows x 32 cols] fp16 staging tile = 8 KB
constexpr int kMaxResBufs = 4;       // residual chunks in flight per epilogue group (UmmaParams::res_bufs: 2..4)
// Attribution experiments (env MCG_DEBUG_FLAGS -> UmmaParams::dbg) are compiled in only with -DMCG_KERNEL_DEBUG=1:
// the issue loops are single-warp, latency-bound code and every runtime flag test costs them cycles.
#ifndef MCG_KERNEL_DEBUG
#define MCG_KERNEL_DEBUG 0
#endif
constexpr bool kDbg = MCG_KERNEL_DEBUG != 0;

struct UmmaParams {
  int M = 0, N = 0, K = 0;
  int block_n = 0;
  int num_stages = 0;
  int num_kb = 0;
  int m_tiles = 0, n_tiles = 0;
  int cblocks = 1;  // C / 64 for im2col
  int k_split = 1;  // split-K factor: tile t covers k-blocks [ks*num_kb/k_split, (ks+1)*num_kb/k_split) and writes
                    // its fp32 partial sums to out_f32 + ks * split_stride (bias / residual / ReLU only in slice 0
                    // resp. never: the reduction kernel applies them)
  long long split_stride = 0;
  int num_acc = 2;   // TMEM accumulator buffers (2 x 256 columns or 4 x 128)
  int acc_cols = 256;
  int out_tma = 0;  // planes output through smem staging + TMA store
  int out_sets = 1; // staging sets per epilogue group for the TMA-store epilogue (2 = double buffered)
  int res_tma = 0;  // residual planes prefetched by TMA: RES_SAME chunks, or (res_up) the window of the coarser level
  int res_up = 0;   // res_tma with RES_UP2X: the [128 x 32] box starts at the source pixel of the tile's first row; all
                    // source pixels of a 128-row tile lie within 128 consecutive source rows (checked on the host)
  int res_bufs = 2; // ... this many [128 x 32] chunks ahead per epilogue group
  int out_fmt = 0;  // planes written by the epilogue: 0 hi, 1 hi + fp16 lo, 2 hi + e4m3 lo8 (+ e4m3 hi8 if out_hi8)
  int out_hi8 = 0;
  int res_fmt = 0;  // planes of the residual: 0 hi, 1 hi + fp16 lo, 2 hi + e4m3 lo8
  int dbg = 0;      // attribution experiments only (env MCG_DEBUG_FLAGS): 1 no stores, 2 no epilogue math,
                    // 4 no A loads, 8 no W loads, 16 no MMA issue, 32 no residual loads.  Results are garbage.
  // K-concatenated second A operand: k-blocks [kb2_begin, num_kb) read the tensor described by `a2` (maps a2_*)
  // instead of `a`; W is the [N, K1 + K2] concatenation.  D = A1 W1^T + A2 W2^T in one accumulator: a bottleneck's
  // conv3 and its downsample branch (resnet.py:286-295) as ONE GEMM, the identity never round-trips through HBM.
  int reverse = 0;    // walk the tile list from its end: consecutive launches alternate, so a launch starts on the tiles
                      // its producer wrote last (still in L2) instead of the ones written first (long evicted)
  int kb2_begin = 0;  // 0 = single A operand
  AGeom a;
  AGeom a2;
  Epilogue ep;
};

// kTerms == 2: a_lo = A lo8, w_lo = W lo8 (e4m3 maps).  o_lo / r_lo are fp16 or e4m3 maps by out_fmt / res_fmt.
struct UmmaMaps {
  CUtensorMap a_hi, a_lo, w_hi, w_lo, o_hi, o_lo, r_hi, r_lo;
  CUtensorMap a_hi8, w_hi8, o_hi8;
  CUtensorMap a2_hi, a2_lo, a2_hi8;  // second A operand (UmmaParams::kb2_begin); a2_lo = fp16 lo or e4m3 lo8 like a_lo
};

// shared-memory bytes of one pipeline stage / one epilogue staging set for a precision mode
// kTerms == 2: one stage holds EITHER the fp16 tiles of a k-block (pass 2: A_hi 16K, W_hi) OR its e4m3 tiles
// (pass 1: A_lo8 8K | A_hi8 8K, W_hi8 | W_lo8) - the same bytes as in single-fp16 mode
__host__ __device__ constexpr int a_stage_bytes(int terms) { return terms == 3 ? 2 * kATileBytes : kATileBytes; }
__host__ __device__ constexpr int w_stage_bytes(int terms, int bn) { return (terms == 3 ? 2 : 1) * bn * kBlockK * 2; }
// one epilogue staging set / residual slot: fp16 hi chunk (+ fp16 lo chunk | + e4m3 lo8 chunk + e4m3 hi8 chunk)
__host__ __device__ constexpr int epi_set_bytes(int fmt) { return (fmt == 0 ? 1 : 2) * kEpiPlaneBytes; }
constexpr int kEpiLo8Off = kEpiPlaneBytes;                       // format 2: [128 x 32 B] e4m3 lo8 chunk
constexpr int kEpiHi8Off = kEpiPlaneBytes + kEpiPlaneBytes / 2;  //              [128 x 32 B] e4m3 hi8 chunk

// v[0..15] += e4m3x16(u) * 2^-13   (the lo8 plane of a residual)
__device__ __forceinline__ void add_lo8x16(float* v, const uint4& u) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
  const __half2 k = __float2half2_rn(kLo8InvScale);
#pragma unroll
  for (int t = 0; t < 4; ++t) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const __half2_raw hr = __nv_cvt_fp8x2_to_halfraw2(static_cast<__nv_fp8x2_storage_t>((w[t] >> (16 * h)) & 0xffffu), __NV_E4M3);
      const __half2 s2 = __hmul2(__half2(hr), k);
      ptx::add_half2(v[4 * t + 2 * h], v[4 * t + 2 * h + 1], *reinterpret_cast<const uint32_t*>(&s2));
    }
  }
}
// e4m3x2 of the rounding residue {(a - low(h2)) 2^13, (b - high(h2)) 2^13}
__device__ __forceinline__ uint32_t residue_e4m3x2(float a, float b, uint32_t h2) {
  float da, db;
  ptx::residue2(a, b, h2, da, db);
  return static_cast<uint32_t>(__nv_cvt_float2_to_fp8x2(make_float2(da * kLo8Scale, db * kLo8Scale), __NV_SATFINITE, __NV_E4M3));
}
__device__ __forceinline__ uint32_t half2_to_e4m3x2(uint32_t h2) {
  __half2_raw hr;
  hr.x = static_cast<unsigned short>(h2 & 0xffffu);
  hr.y = static_cast<unsigned short>(h2 >> 16);
  return static_cast<uint32_t>(__nv_cvt_halfraw2_to_fp8x2(hr, __NV_SATFINITE, __NV_E4M3));
}

// byte offset of logical 16-byte chunk j (0/1) of row `row` in a 32B-swizzled [rows][32 B] tile (e4m3 chunks)
__device__ __forceinline__ uint32_t sw32_off(int row, int j) {
  return static_cast<uint32_t>(row * 32 + ((j ^ ((row >> 2) & 1)) << 4));
}
// byte offset of logical 16-byte chunk j of row `row` in a 64B-swizzled [rows][64 B] tile
__device__ __forceinline__ uint32_t sw64_off(int row, int j) {
  return static_cast<uint32_t>(row * 64 + ((j ^ ((row >> 1) & 3)) << 4));
}

// kPair: the kernel runs as clusters of two CTAs (one SM pair) that share each tile of 256 rows x block_n
// columns through tcgen05 cta_group::2: every CTA loads its own 128 rows of A and HALF of the W rows, the leader
// CTA issues one M = 256 MMA over both, each CTA drains its own 128 accumulator rows.  Halves the W bytes every
// SM pulls from L2 per FLOP (the 3x3 convolutions are L2 -> shared-memory bound with 128 x 256 tiles).
// kKbs: 64-wide k-blocks per pipeline stage (compile time, so the per-stage loops unroll): layers with little work
// per k-block (narrow tiles) amortise the barrier handshake and issue-loop overhead of a stage over 2-3 k-blocks.
template <int kTerms, bool kPair, int kKbs>
__global__ void __launch_bounds__(kGemmThreads, 1)
umma_gemm_kernel(const __grid_constant__ UmmaMaps tm, const UmmaParams p) {
  // storage format of the activation planes this instantiation reads / writes: 0 hi, 1 hi + fp16 lo,
  // 2 hi + e4m3 lo8 (+ e4m3 hi8)
  constexpr int kFmt = kTerms == 1 ? 0 : kTerms == 3 ? 1 : 2;
  constexpr int kSet = epi_set_bytes(kFmt);   // bytes of one epilogue staging set (hi chunk [+ lo / lo8 + hi8 chunk])
  constexpr int kRSet = kSet;                 // bytes of one residual prefetch slot
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  // programmatic dependent launch: the next GEMM's CTAs may take this SM as soon as this CTA exits and run their
  // prologue (barrier init, TMEM allocation, descriptor prefetch) under the tail of this grid
  ptx::grid_dep_launch_dependents();

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tfull_bar = empty_bar + kMaxStages;
  uint64_t* tempty_bar = tfull_bar + kMaxAcc;
  uint64_t* res_bar = tempty_bar + kMaxAcc;  // [kEpiGroups][kResBufs]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(res_bar + kEpiGroups * kMaxResBufs);
  uint8_t* stage_base = smem + kSmemBarrierBytes;

  // rows of W this CTA keeps in shared memory (a CTA pair splits the block_n rows)
  const int w_rows = kPair ? p.block_n / 2 : p.block_n;
  const uint32_t cta_rank = kPair ? ptx::cluster_ctarank() : 0u;
  const bool leader_cta = cta_rank == 0;
  const int w_tile_bytes = w_rows * kBlockK * 2;
  const int a_bytes = a_stage_bytes(kTerms);
  const int sub_bytes = a_bytes + w_stage_bytes(kTerms, w_rows);  // one k-block: A tile(s) + W tile(s)
  const int stage_bytes = kKbs * sub_bytes;
  // kTerms == 2, pass 1 (e4m3 tiles) re-uses the stage: [A_lo8 8K | A_hi8 8K] [W_hi8 | W_lo8]
  const int w8_tile_bytes = w_rows * kBlockK;
  const int off_a_hi8 = kATileBytes / 2;
  const int off_w_lo8 = a_bytes + w8_tile_bytes;
  // k-block iterations per tile: kTerms == 2 runs the K range twice (e4m3 pass, then fp16 pass)
  const int passes = kTerms == 2 ? 2 : 1;
  uint8_t* obuf_base = stage_base + static_cast<size_t>(p.num_stages) * stage_bytes;        // [groups][out_sets][kSet]
  uint8_t* rbuf_base = obuf_base + (p.out_tma ? kEpiGroups * p.out_sets * kSet : 0);       // [groups][kResBufs][kSet]
  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int mn_tiles = p.m_tiles * p.n_tiles;
  const int num_tiles = mn_tiles * p.k_split;

  if (warp_idx == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm.a_hi);
    ptx::prefetch_tmap(&tm.w_hi);
    if (kTerms != 1) {
      ptx::prefetch_tmap(&tm.a_lo);
      ptx::prefetch_tmap(&tm.w_lo);
    }
    if (kTerms == 2) {
      ptx::prefetch_tmap(&tm.w_hi8);
      ptx::prefetch_tmap(&tm.a_hi8);
    }
    if (p.out_tma) ptx::prefetch_tmap(&tm.o_hi);
    if (p.res_tma) ptx::prefetch_tmap(&tm.r_hi);
    if (p.kb2_begin) {
      ptx::prefetch_tmap(&tm.a2_hi);
      if (kTerms != 1) ptx::prefetch_tmap(&tm.a2_lo);
      if (kTerms == 2) ptx::prefetch_tmap(&tm.a2_hi8);
    }
  }
  if (warp_idx == 1 && lane == 0) {
    for (int i = 0; i < p.num_stages; ++i) {
      ptx::mbar_init(&full_bar[i], kPair ? 2 : 1);  // pair: one arrive per CTA, on the leader's barrier
      ptx::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < kMaxAcc; ++i) {
      ptx::mbar_init(&tfull_bar[i], 1);
      // one arrive per warp of the owning epilogue group (pair: of both CTAs, on the leader's barrier)
      ptx::mbar_init(&tempty_bar[i], kPair ? 8 : 4);
    }
    for (int i = 0; i < kEpiGroups * kMaxResBufs; ++i) ptx::mbar_init(&res_bar[i], 1);
    ptx::fence_mbar_init();
  }
  if (warp_idx == 2) {
    if (kPair) {
      ptx::tmem_alloc_pair(tmem_ptr_smem, kTmemCols);
      ptx::tmem_relinquish_pair();
    } else {
      ptx::tmem_alloc(tmem_ptr_smem, kTmemCols);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (kPair) ptx::cluster_sync();  // the peer's barriers are initialised before anything arrives on them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // everything below reads or writes tensors of earlier kernels
  ptx::grid_dep_wait();
  // persistent tile walk: a CTA pair walks the tile list together
  const int walker = kPair ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int walkers = kPair ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);

  if (warp_idx == 0) {
    // ===================== TMA producer =====================
    // The whole warp runs the loop (warp-uniform control flow keeps coordinates and smem addresses in
    // uniform registers, which the TMA instructions read directly); one elected lane issues.
    int stage = 0;
    uint32_t phase = 0;
    long long dbg_wait = 0;
    const long long dbg_p0 = clock64();
    const uint32_t full_a = ptx::smem_u32(full_bar), empty_a = ptx::smem_u32(empty_bar);
    const uint32_t ring_a = ptx::smem_u32(stage_base);
    for (int tile_i = walker; tile_i < num_tiles; tile_i += walkers) {
      const int tile = p.reverse ? num_tiles - 1 - tile_i : tile_i;
      const int ks = tile / mn_tiles;
      const int mn = tile - ks * mn_tiles;
      const int m_tile = mn / p.n_tiles;
      const int n_tile = mn - m_tile * p.n_tiles;
      const int kb_begin = static_cast<int>(static_cast<long long>(ks) * p.num_kb / p.k_split);
      const int kb_end = static_cast<int>(static_cast<long long>(ks + 1) * p.num_kb / p.k_split);
      // pair: m_tile counts 256-row tiles, this CTA owns rows [128 rank, 128 rank + 128) and W rows [w_rows rank, ...)
      const long long m0 = (static_cast<long long>(m_tile) * (kPair ? 2 : 1) + cta_rank) * kBlockM;
      const int n0 = n_tile * p.block_n + static_cast<int>(cta_rank) * (kPair ? w_rows : 0);
      int img_n = 0, base_h = 0, base_w = 0;
      if (p.a.kind == 1) {
        const long long pq = static_cast<long long>(p.a.P) * p.a.Q;
        img_n = static_cast<int>(m0 / pq);
        const int rem = static_cast<int>(m0 - img_n * pq);
        const int p0 = rem / p.a.Q;
        const int q0 = rem - p0 * p.a.Q;
        base_h = p0 * p.a.stride - p.a.pad;
        base_w = q0 * p.a.stride - p.a.pad;
      }
      // second A operand (1x1 convolution, plain or strided im2col view on the same output grid)
      int img_n2 = 0, base_h2 = 0, base_w2 = 0;
      if (p.kb2_begin && p.a2.kind == 1) {
        const long long pq = static_cast<long long>(p.a2.P) * p.a2.Q;
        img_n2 = static_cast<int>(m0 / pq);
        const int rem = static_cast<int>(m0 - img_n2 * pq);
        const int p0 = rem / p.a2.Q;
        const int q0 = rem - p0 * p.a2.Q;
        base_h2 = p0 * p.a2.stride - p.a2.pad;
        base_w2 = q0 * p.a2.stride - p.a2.pad;
      }
      for (int pass = 0; pass < passes; ++pass) {
        const bool f8 = kTerms == 2 && pass == 0;
        int tap = kb_begin / p.cblocks, cb = kb_begin - tap * p.cblocks;
        int tr = tap / p.a.S, tsx = tap - tr * p.a.S;
        for (int kb = kb_begin; kb < kb_end; kb += kKbs) {
          if (kDbg && (p.dbg & 128)) {
            const long long w0 = clock64();
            ptx::mbar_wait_a(empty_a + stage * 8, phase ^ 1u);
            dbg_wait += clock64() - w0;
          } else {
            ptx::mbar_wait_a(empty_a + stage * 8, phase ^ 1u);
          }
          if (ptx::elect_one()) {
            const uint32_t s0 = ring_a + static_cast<uint32_t>(stage * stage_bytes);
            // pair: both CTAs' loads signal the LEADER's full barrier (it counts the bytes of both)
            const uint32_t fb = kPair ? ptx::mapa(full_a + stage * 8, 0) : full_a + stage * 8;
            uint32_t tx = static_cast<uint32_t>(stage_bytes);
            if (kDbg && (p.dbg & 4)) tx -= static_cast<uint32_t>(kKbs * a_bytes);
            if (kDbg && (p.dbg & 8)) tx -= static_cast<uint32_t>(kKbs * (sub_bytes - a_bytes));
            if (kPair) {
              if (leader_cta)
                ptx::mbar_arrive_expect_tx_a(full_a + stage * 8, 2 * tx);
              else
                ptx::mbar_arrive_cluster_a(fb);
            } else if (tx == 0) {
              ptx::mbar_arrive_a(fb);
            } else {
              ptx::mbar_arrive_expect_tx_a(fb, tx);
            }
            auto tma_2d = [&](uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
              if (kPair)
                ptx::tma_load_2d_pair_a(dst, m, bar, c0, c1);
              else
                ptx::tma_load_2d_a(dst, m, bar, c0, c1);
            };
            auto tma_im2col = [&](uint32_t dst, const CUtensorMap* m, uint32_t bar, int c, int w, int h, int n, uint16_t ow,
                                  uint16_t oh) {
              if (kPair)
                ptx::tma_load_im2col_4d_pair_a(dst, m, bar, c, w, h, n, ow, oh);
              else
                ptx::tma_load_im2col_4d_a(dst, m, bar, c, w, h, n, ow, oh);
            };
            // first / second A tile of a k-block: (hi, lo) fp16 planes, or the (lo8, hi8) e4m3 planes in pass 1
            const CUtensorMap* ma0 = f8 ? &tm.a_lo : &tm.a_hi;
            const CUtensorMap* ma1 = f8 ? &tm.a_hi8 : &tm.a_lo;
            const bool two_a = kTerms == 3 || f8;
            int cb2 = cb, tsx2 = tsx, tr2 = tr;  // filter tap / channel block of the k-block being loaded
#pragma unroll
            for (int sub = 0; sub < kKbs; ++sub) {
              const uint32_t s = s0 + static_cast<uint32_t>(sub * sub_bytes);
              const uint32_t sW = s + a_bytes;
              const uint32_t sA1 = s + (f8 ? off_a_hi8 : kATileBytes);
              const int kbk = kb + sub;
              if (kDbg && (p.dbg & 4)) {
              } else if (p.kb2_begin && kbk >= p.kb2_begin) {
                // k-blocks of the second A operand (a 1x1 convolution: one filter tap, channel block kbk - kb2_begin)
                const CUtensorMap* mb0 = f8 ? &tm.a2_lo : &tm.a2_hi;
                const CUtensorMap* mb1 = f8 ? &tm.a2_hi8 : &tm.a2_lo;
                const int c0 = (kbk - p.kb2_begin) * kBlockK;
                if (p.a2.kind == 1) {
                  tma_im2col(s, mb0, fb, c0, base_w2, base_h2, img_n2, 0, 0);
                  if (two_a) tma_im2col(sA1, mb1, fb, c0, base_w2, base_h2, img_n2, 0, 0);
                } else {
                  tma_2d(s, mb0, fb, c0, static_cast<int>(m0));
                  if (two_a) tma_2d(sA1, mb1, fb, c0, static_cast<int>(m0));
                }
              } else if (p.a.kind == 1) {
                tma_im2col(s, ma0, fb, cb2 * kBlockK, base_w, base_h, img_n, static_cast<uint16_t>(tsx2),
                           static_cast<uint16_t>(tr2));
                if (two_a)
                  tma_im2col(sA1, ma1, fb, cb2 * kBlockK, base_w, base_h, img_n, static_cast<uint16_t>(tsx2),
                             static_cast<uint16_t>(tr2));
                if (kKbs > 1 && ++cb2 == p.cblocks) {
                  cb2 = 0;
                  if (++tsx2 == p.a.S) {
                    tsx2 = 0;
                    ++tr2;
                  }
                }
              } else {
                tma_2d(s, ma0, fb, kbk * kBlockK, static_cast<int>(m0));
                if (two_a) tma_2d(sA1, ma1, fb, kbk * kBlockK, static_cast<int>(m0));
              }
              if (!(kDbg && (p.dbg & 8))) {
                if (f8) {
                  tma_2d(sW, &tm.w_hi8, fb, kbk * kBlockK, n0);
                  tma_2d(s + off_w_lo8, &tm.w_lo, fb, kbk * kBlockK, n0);
                } else {
                  tma_2d(sW, &tm.w_hi, fb, kbk * kBlockK, n0);
                  if (kTerms == 3) tma_2d(sW + w_tile_bytes, &tm.w_lo, fb, kbk * kBlockK, n0);
                }
              }
            }
          }
          __syncwarp();
          // next filter tap / channel block (im2col view: kb = tap * cblocks + cb, tap = r * S + s)
#pragma unroll
          for (int sub = 0; sub < kKbs; ++sub) {
            if (++cb == p.cblocks) {
              cb = 0;
              if (++tsx == p.a.S) {
                tsx = 0;
                ++tr;
              }
            }
          }
          if (++stage == p.num_stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
    if ((kDbg && (p.dbg & 128)) && blockIdx.x == 0 && lane == 0)
      printf("prod M=%d N=%d K=%d: %lld cycles, %lld waiting for a free stage\n", p.M, p.N, p.K, clock64() - dbg_p0, dbg_wait);
  } else if (warp_idx == 1) {
    // ===================== MMA issuer =====================
    // Warp-uniform loop, one elected lane issues each tcgen05.mma / commit (the shared-memory descriptors are
    // then computed on the uniform datapath; a lane-0-only loop costs ~5 R2UR moves per MMA and made the issue
    // rate, not the tensor pipe, the limit).
    const uint32_t idesc = ptx::make_idesc_f16_f32(kPair ? 2 * kBlockM : kBlockM, p.block_n);
    const uint32_t stage_base_u32 = ptx::smem_u32(stage_base);
    auto mma_f16 = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t acc) {
      if (kPair)
        ptx::umma_f16_pair(d, a, b, idesc, acc);
      else
        ptx::umma_f16(d, a, b, idesc, acc);
    };
    auto mma_f8 = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t acc) {
      if (kPair)
        ptx::umma_f8_pair(d, a, b, idesc, acc);
      else
        ptx::umma_f8(d, a, b, idesc, acc);
    };
    auto commit = [&](uint32_t bar) {
      if (kPair)
        ptx::umma_commit_pair_a(bar);
      else
        ptx::umma_commit_a(bar);
    };
    const uint32_t full_a = ptx::smem_u32(full_bar), empty_a = ptx::smem_u32(empty_bar);
    const uint32_t tfull_a = ptx::smem_u32(tfull_bar), tempty_a = ptx::smem_u32(tempty_bar);
    int stage = 0;
    uint32_t phase = 0;
    int local = 0;
    long long dbg_c0 = 0, dbg_wfull = 0, dbg_wacc = 0;
    unsigned long long dbg_t0 = 0;
    if ((kDbg && (p.dbg & 128)) && blockIdx.x == 0 && lane == 0) {
      dbg_c0 = clock64();
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_t0));
    }
    for (int tile_i = walker; tile_i < num_tiles && leader_cta; tile_i += walkers, ++local) {
      const int tile = p.reverse ? num_tiles - 1 - tile_i : tile_i;
      const int acc = local % p.num_acc;
      const uint32_t acc_phase = static_cast<uint32_t>(local / p.num_acc) & 1u;
      if (kDbg && (p.dbg & 128)) {
        const long long w0 = clock64();
        ptx::mbar_wait_a(tempty_a + acc * 8, acc_phase ^ 1u);
        dbg_wacc += clock64() - w0;
      } else {
        ptx::mbar_wait_a(tempty_a + acc * 8, acc_phase ^ 1u);
      }
      ptx::tc_fence_after();
      const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * p.acc_cols);
      const int ks = tile / mn_tiles;
      const int kb_begin = static_cast<int>(static_cast<long long>(ks) * p.num_kb / p.k_split);
      const int kb_end = static_cast<int>(static_cast<long long>(ks + 1) * p.num_kb / p.k_split);
      for (int pass = 0; pass < passes; ++pass) {
        const bool f8 = kTerms == 2 && pass == 0;
        for (int kb = kb_begin; kb < kb_end; kb += kKbs) {
          if (kDbg && (p.dbg & 128)) {
            const long long w0 = clock64();
            ptx::mbar_wait_a(full_a + stage * 8, phase);
            dbg_wfull += clock64() - w0;
          } else {
            ptx::mbar_wait_a(full_a + stage * 8, phase);
          }
          ptx::tc_fence_after();
#pragma unroll
          for (int sub = 0; sub < kKbs; ++sub) {
            if (kDbg && (p.dbg & 16)) break;
            const uint32_t s = stage_base_u32 + static_cast<uint32_t>(stage * stage_bytes + sub * sub_bytes);
            const uint32_t first = (sub > 0 || kb > kb_begin) ? 1u : 0u;
            // descriptors of the k = 0 slice; slice j advances the 16-byte-granular start address by 2 (32 B)
            if (f8) {
              // pass 1, e4m3 corrections x 2^15 (K = 32 per MMA): lo8_a * hi8_w + hi8_a * lo8_w
              const uint64_t dA_lo8 = ptx::make_sw64_kmajor_desc(s);
              const uint64_t dA_hi8 = ptx::make_sw64_kmajor_desc(s + off_a_hi8);
              const uint64_t dW_hi8 = ptx::make_sw64_kmajor_desc(s + a_bytes);
              const uint64_t dW_lo8 = ptx::make_sw64_kmajor_desc(s + off_w_lo8);
#pragma unroll
              for (int j = 0; j < kBlockK / 32; ++j) {
                mma_f8(tmem_d, dA_lo8 + 2 * j, dW_hi8 + 2 * j, (first || j > 0) ? 1u : 0u);
                mma_f8(tmem_d, dA_hi8 + 2 * j, dW_lo8 + 2 * j, 1u);
              }
            } else {
              const uint64_t dA_hi = ptx::make_sw128_kmajor_desc(s);
              const uint64_t dA_lo = ptx::make_sw128_kmajor_desc(s + kATileBytes);
              const uint64_t dW_hi = ptx::make_sw128_kmajor_desc(s + a_bytes);
              const uint64_t dW_lo = ptx::make_sw128_kmajor_desc(s + a_bytes + w_tile_bytes);
#pragma unroll
              for (int j = 0; j < kBlockK / kUmmaK; ++j) {
                uint32_t accum = (first || j > 0) ? 1u : 0u;
                if (kTerms == 3) {
                  mma_f16(tmem_d, dA_lo + 2 * j, dW_hi + 2 * j, accum);
                  mma_f16(tmem_d, dA_hi + 2 * j, dW_lo + 2 * j, 1u);
                  accum = 1u;
                }
                if (kTerms == 2 && j == 0 && !first) {
                  // first fp16 MMA of the tile: D = A*B + D 2^-15 brings the e4m3 corrections to scale
                  if (kPair)
                    ptx::umma_f16_rescale_d_pair<kC8AccShift>(tmem_d, dA_hi, dW_hi, idesc);
                  else
                    ptx::umma_f16_rescale_d<kC8AccShift>(tmem_d, dA_hi, dW_hi, idesc);
                } else {
                  mma_f16(tmem_d, dA_hi + 2 * j, dW_hi + 2 * j, accum);
                }
              }
            }
          }
          commit(empty_a + stage * 8);  // smem slot reusable once these MMAs retire
          if (kb + kKbs >= kb_end && pass == passes - 1) commit(tfull_a + acc * 8);
          if (++stage == p.num_stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
    if ((kDbg && (p.dbg & 128)) && blockIdx.x == 0 && lane == 0) {
      // effective SM clock while this launch ran (attribution experiments only)
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      const long long c1 = clock64();
      printf("umma M=%d N=%d K=%d terms=%d bn=%d stages=%d: %lld cycles in %llu ns = %.0f MHz; MMA warp waited %lld for operands, %lld for an accumulator\n",
             p.M, p.N, p.K, kTerms, p.block_n, p.num_stages, c1 - dbg_c0, t1 - dbg_t0, 1e3 * double(c1 - dbg_c0) / double(t1 - dbg_t0),
             dbg_wfull, dbg_wacc);
    }
  } else if (warp_idx >= 4) {
    // ===================== epilogue (two groups of four warps) =====================
    const int ew = warp_idx - 4;
    const int grp = ew >> 2;
    const int quarter = warp_idx & 3;     // TMEM lane quarter this warp may access (= warp id % 4)
    const int row = quarter * 32 + lane;  // row of the tile == TMEM lane
    const bool leader = (ew & 3) == 0 && lane == 0;
    const int bar_id = 1 + grp;           // named barrier of this group (128 threads)
    const Epilogue& ep = p.ep;
    const int nchunks = p.block_n / kEpiChunk;
    const uint32_t osets = static_cast<uint32_t>(p.out_sets);
    uint8_t* obuf = obuf_base + grp * p.out_sets * kSet;
    const uint32_t res_bufs = static_cast<uint32_t>(p.res_bufs);
    uint8_t* rbuf = rbuf_base + grp * p.res_bufs * kRSet;
    uint64_t* rbar = res_bar + grp * kMaxResBufs;
    // residual chunk stream of this group's tiles, prefetched kResBufs chunks ahead across tile boundaries
    int ri_local = grp, ri_c = 0;
    uint32_t r_issued = 0, r_consumed = 0;
    auto res_issue = [&]() {
      const int t_i = walker + ri_local * walkers;
      if (t_i >= num_tiles) return;
      const int t = p.reverse ? num_tiles - 1 - t_i : t_i;
      const uint32_t b = r_issued % res_bufs;
      if (leader && (kDbg && (p.dbg & 32))) {
        ptx::mbar_arrive(&rbar[b]);
      } else if (leader) {
        const int mn_i = t % mn_tiles;
        const int mtile = mn_i / p.n_tiles;
        const int nt = mn_i - mtile * p.n_tiles;
        const int mt = mtile * (kPair ? 2 : 1) + static_cast<int>(cta_rank);  // 128-row tile of this CTA
        uint8_t* dst = rbuf + b * kRSet;
        ptx::fence_proxy_async();
        ptx::mbar_arrive_expect_tx(&rbar[b], static_cast<uint32_t>(kFmt == 2   ? kEpiPlaneBytes + kEpiPlaneBytes / 2
                                                                   : kFmt == 1 ? 2 * kEpiPlaneBytes
                                                                                    : kEpiPlaneBytes));
        const int r0 = p.res_up ? static_cast<int>(res_row_window_start(ep, static_cast<long long>(mt) * kBlockM)) : mt * kBlockM;
        ptx::tma_load_2d(dst, &tm.r_hi, &rbar[b], nt * p.block_n + ri_c * kEpiChunk, r0);
        if (kFmt != 0)  // fp16 lo chunk or e4m3 lo8 chunk
          ptx::tma_load_2d(dst + kEpiPlaneBytes, &tm.r_lo, &rbar[b], nt * p.block_n + ri_c * kEpiChunk, r0);
      }
      ++r_issued;
      if (++ri_c == nchunks) {
        ri_c = 0;
        ri_local += kEpiGroups;
      }
    };
    if (p.res_tma) {
      for (int i = 0; i < p.res_bufs; ++i) res_issue();
    }
    uint32_t ostores = 0;  // chunks handed to TMA so far (staging set = ostores % osets)
    for (int local = grp;; local += kEpiGroups) {
      const int tile_i = walker + local * walkers;
      if (tile_i >= num_tiles) break;
      const int tile = p.reverse ? num_tiles - 1 - tile_i : tile_i;
      const int ks = tile / mn_tiles;
      const int mn = tile - ks * mn_tiles;
      const int m_tile = mn / p.n_tiles;
      const int n_tile = mn - m_tile * p.n_tiles;
      const int acc = local % p.num_acc;
      const uint32_t acc_phase = static_cast<uint32_t>(local / p.num_acc) & 1u;
      const int m_tile_cta = m_tile * (kPair ? 2 : 1) + static_cast<int>(cta_rank);  // 128-row tile of this CTA
      const long long m = static_cast<long long>(m_tile_cta) * kBlockM + row;
      const bool valid = m < p.M;
      const int n_base = n_tile * p.block_n;
      ptx::mbar_wait(&tfull_bar[acc], acc_phase);
      ptx::tc_fence_after();
      const long long rrow = (valid && ep.res_mode != RES_NONE && !p.res_tma) ? res_row(ep, m) : 0;
      // row of this thread's residual inside a TMA-prefetched slot: its own row, or (nearest-2x upsampling) the offset
      // of its source pixel from the tile's first source pixel
      int srow = row;
      if (p.res_up) {
        const long long first = res_row_window_start(ep, static_cast<long long>(m_tile_cta) * kBlockM);
        srow = valid ? static_cast<int>(res_row(ep, m) - first) : 0;
      }
      const uint32_t taddr0 =
          tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * p.acc_cols);
      // gathered residual (FPN nearest-2x top-down add): per-thread vector loads, software-pipelined
      // one chunk ahead in registers so the L2 latency overlaps the previous chunk's math
      const bool res_direct = valid && ep.res_mode != RES_NONE && !p.res_tma && ep.res_f32 == nullptr;
      uint4 rpre[8];
      auto load_direct = [&](int n) {
        const uint4* rh = reinterpret_cast<const uint4*>(ep.res_hi + rrow * ep.ldr + n);
#pragma unroll
        for (int j = 0; j < 4; ++j) rpre[j] = __ldg(rh + j);
        if (kFmt == 1) {
          const uint4* rl = reinterpret_cast<const uint4*>(ep.res_lo + rrow * ep.ldr + n);
#pragma unroll
          for (int j = 0; j < 4; ++j) rpre[4 + j] = __ldg(rl + j);
        }
        if (kFmt == 2) {
          const uint4* rl = reinterpret_cast<const uint4*>(ep.res_lo8 + rrow * ep.ldr + n);
#pragma unroll
          for (int j = 0; j < 2; ++j) rpre[4 + j] = __ldg(rl + j);
        }
      };
      if (res_direct) load_direct(n_base);
#pragma unroll 1
      for (int c = 0; c < nchunks; ++c) {
        const int n = n_base + c * kEpiChunk;
        uint32_t r[32];
        ptx::tmem_ld_32x32(taddr0 + static_cast<uint32_t>(c * kEpiChunk), r);
        uint4 rnow[8];
        if (res_direct) {
#pragma unroll
          for (int j = 0; j < 8; ++j) rnow[j] = rpre[j];
          if (c + 1 < nchunks) load_direct(n + kEpiChunk);
        }
        const uint8_t* rcur = nullptr;
        if (p.res_tma) {
          const uint32_t b = r_consumed % res_bufs;
          ptx::mbar_wait(&rbar[b], (r_consumed / res_bufs) & 1u);
          rcur = rbuf + b * kRSet;
          ++r_consumed;
        }
        ptx::tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (!(kDbg && (p.dbg & 2))) {
          if (ep.bias && p.k_split == 1) {
            const float4* b4 = reinterpret_cast<const float4*>(ep.bias + n);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = __ldg(b4 + j);
              v[4 * j + 0] += b.x;
              v[4 * j + 1] += b.y;
              v[4 * j + 2] += b.z;
              v[4 * j + 3] += b.w;
            }
          }
          if (rcur) {
            if (kFmt == 2) {
              const uint8_t* r8 = rcur + kEpiLo8Off;
              add_lo8x16(v, *reinterpret_cast<const uint4*>(r8 + sw32_off(srow, 0)));
              add_lo8x16(v + 16, *reinterpret_cast<const uint4*>(r8 + sw32_off(srow, 1)));
            }
#pragma unroll
            for (int pl = 0; pl < 2; ++pl) {
              if (pl == 1 && kFmt != 1) break;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint4 u = *reinterpret_cast<const uint4*>(rcur + pl * kEpiPlaneBytes + sw64_off(srow, j));
                ptx::add_half2(v[8 * j + 0], v[8 * j + 1], u.x);
                ptx::add_half2(v[8 * j + 2], v[8 * j + 3], u.y);
                ptx::add_half2(v[8 * j + 4], v[8 * j + 5], u.z);
                ptx::add_half2(v[8 * j + 6], v[8 * j + 7], u.w);
              }
            }
          } else if (valid && ep.res_mode != RES_NONE && !p.res_tma) {
            if (ep.res_f32) {
              const float4* rf = reinterpret_cast<const float4*>(ep.res_f32 + rrow * ep.ldr + n);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 f = __ldg(rf + j);
                v[4 * j + 0] += f.x;
                v[4 * j + 1] += f.y;
                v[4 * j + 2] += f.z;
                v[4 * j + 3] += f.w;
              }
            } else {
              if (kFmt == 2) {
                add_lo8x16(v, rnow[4]);
                add_lo8x16(v + 16, rnow[5]);
              }
              const int npl = kFmt == 1 ? 2 : 1;
#pragma unroll
              for (int pl = 0; pl < 2; ++pl) {
                if (pl < npl) {
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const uint4 u = rnow[pl * 4 + j];
                    ptx::add_half2(v[8 * j + 0], v[8 * j + 1], u.x);
                    ptx::add_half2(v[8 * j + 2], v[8 * j + 3], u.y);
                    ptx::add_half2(v[8 * j + 4], v[8 * j + 5], u.z);
                    ptx::add_half2(v[8 * j + 6], v[8 * j + 7], u.w);
                  }
                }
              }
            }
          }
          if (ep.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          }
        }
        if (p.out_tma) {
          uint8_t* ob = obuf + (ostores % osets) * kSet;
          if (osets == 1) {
            // single staging set: the store issued for the previous chunk must have read it out
            if (leader) ptx::tma_store_wait_read<0>();
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
          }
          if (!(kDbg && (p.dbg & 2))) {
            uint2 l8[4], h8[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 uh, ul;
              uh.x = ptx::pack_half2(v[8 * j + 0], v[8 * j + 1]);
              uh.y = ptx::pack_half2(v[8 * j + 2], v[8 * j + 3]);
              uh.z = ptx::pack_half2(v[8 * j + 4], v[8 * j + 5]);
              uh.w = ptx::pack_half2(v[8 * j + 6], v[8 * j + 7]);
              *reinterpret_cast<uint4*>(ob + sw64_off(row, j)) = uh;
              if (kFmt == 1) {
                ul.x = ptx::residue_half2(v[8 * j + 0], v[8 * j + 1], uh.x);
                ul.y = ptx::residue_half2(v[8 * j + 2], v[8 * j + 3], uh.y);
                ul.z = ptx::residue_half2(v[8 * j + 4], v[8 * j + 5], uh.z);
                ul.w = ptx::residue_half2(v[8 * j + 6], v[8 * j + 7], uh.w);
                *reinterpret_cast<uint4*>(ob + kEpiPlaneBytes + sw64_off(row, j)) = ul;
              }
              if (kFmt == 2) {
                l8[j].x = residue_e4m3x2(v[8 * j + 0], v[8 * j + 1], uh.x) | (residue_e4m3x2(v[8 * j + 2], v[8 * j + 3], uh.y) << 16);
                l8[j].y = residue_e4m3x2(v[8 * j + 4], v[8 * j + 5], uh.z) | (residue_e4m3x2(v[8 * j + 6], v[8 * j + 7], uh.w) << 16);
                if (p.out_hi8) {
                  h8[j].x = half2_to_e4m3x2(uh.x) | (half2_to_e4m3x2(uh.y) << 16);
                  h8[j].y = half2_to_e4m3x2(uh.z) | (half2_to_e4m3x2(uh.w) << 16);
                }
              }
            }
            if (kFmt == 2) {
              // e4m3 chunks: [128 rows x 32 B], 32B-swizzled (16-byte writes of consecutive rows hit distinct banks)
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                *reinterpret_cast<uint4*>(ob + kEpiLo8Off + sw32_off(row, j)) =
                    make_uint4(l8[2 * j].x, l8[2 * j].y, l8[2 * j + 1].x, l8[2 * j + 1].y);
                if (p.out_hi8)
                  *reinterpret_cast<uint4*>(ob + kEpiHi8Off + sw32_off(row, j)) =
                      make_uint4(h8[2 * j].x, h8[2 * j].y, h8[2 * j + 1].x, h8[2 * j + 1].y);
              }
            }
          }
          // all 128 threads have consumed the residual slot and filled the staging set; with two sets the
          // store of the previous chunk (other set) must have been read out before the next chunk refills it
          ptx::fence_proxy_async();
          if (osets == 2 && leader) ptx::tma_store_wait_read<0>();
          asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
          if (p.res_tma) res_issue();
          if (leader) {
            if (!(kDbg && (p.dbg & 1))) {
              const int m0 = m_tile_cta * kBlockM;
              ptx::tma_store_2d(&tm.o_hi, ob, n, m0);
              if (kFmt != 0) ptx::tma_store_2d(&tm.o_lo, ob + kEpiPlaneBytes, n, m0);  // fp16 lo / e4m3 lo8
              if (p.out_hi8) ptx::tma_store_2d(&tm.o_hi8, ob + kEpiHi8Off, n, m0);
            }
            ptx::tma_store_commit();
          }
          ++ostores;
        } else if (valid && !(kDbg && (p.dbg & 2))) {
          if (ep.out_f32) {
            float4* o = reinterpret_cast<float4*>(ep.out_f32 + ks * p.split_stride + m * ep.ldo + n);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else {
            uint4* oh = reinterpret_cast<uint4*>(ep.out_hi + m * ep.ldo + n);
            uint4* ol = ep.out_lo ? reinterpret_cast<uint4*>(ep.out_lo + m * ep.ldo + n) : nullptr;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 uh, ul;
              uh.x = ptx::pack_half2(v[8 * j + 0], v[8 * j + 1]);
              uh.y = ptx::pack_half2(v[8 * j + 2], v[8 * j + 3]);
              uh.z = ptx::pack_half2(v[8 * j + 4], v[8 * j + 5]);
              uh.w = ptx::pack_half2(v[8 * j + 6], v[8 * j + 7]);
              oh[j] = uh;
              if (ol) {
                ul.x = ptx::residue_half2(v[8 * j + 0], v[8 * j + 1], uh.x);
                ul.y = ptx::residue_half2(v[8 * j + 2], v[8 * j + 3], uh.y);
                ul.z = ptx::residue_half2(v[8 * j + 4], v[8 * j + 5], uh.z);
                ul.w = ptx::residue_half2(v[8 * j + 6], v[8 * j + 7], uh.w);
                ol[j] = ul;
              }
              if (ep.out_lo8) {
                uint2 l8;
                l8.x = residue_e4m3x2(v[8 * j + 0], v[8 * j + 1], uh.x) | (residue_e4m3x2(v[8 * j + 2], v[8 * j + 3], uh.y) << 16);
                l8.y = residue_e4m3x2(v[8 * j + 4], v[8 * j + 5], uh.z) | (residue_e4m3x2(v[8 * j + 6], v[8 * j + 7], uh.w) << 16);
                *reinterpret_cast<uint2*>(ep.out_lo8 + m * ep.ldo + n + 8 * j) = l8;
              }
              if (ep.out_hi8) {
                uint2 h8;
                h8.x = half2_to_e4m3x2(uh.x) | (half2_to_e4m3x2(uh.y) << 16);
                h8.y = half2_to_e4m3x2(uh.z) | (half2_to_e4m3x2(uh.w) << 16);
                *reinterpret_cast<uint2*>(ep.out_hi8 + m * ep.ldo + n + 8 * j) = h8;
              }
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kPair && !leader_cta)
          ptx::mbar_arrive_cluster_a(ptx::mapa(ptx::smem_u32(&tempty_bar[acc]), 0));  // the leader's MMA warp waits
        else
          ptx::mbar_arrive(&tempty_bar[acc]);
      }
    }
    // smem must stay valid until every bulk store has been read out
    if (p.out_tma && leader) ptx::tma_store_wait_all<0>();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (kPair) ptx::cluster_sync();  // no CTA leaves while its peer may still arrive on its barriers / read its smem
  if (warp_idx == 2) {
    ptx::tc_fence_after();
    if (kPair)
      ptx::tmem_dealloc_pair(tmem_base, kTmemCols);
    else
      ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------
// host side: tensor maps + launch plan
// ---------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*PFN_encodeIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct DriverApi {
  PFN_encodeTiled encodeTiled = nullptr;
  PFN_encodeIm2col encodeIm2col = nullptr;
  static const DriverApi& get() {
    static DriverApi api = [] {
      DriverApi a;
      cudaDriverEntryPointQueryResult qr;
      void* fn = nullptr;
      MCG_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
      MCG_CHECK(fn != nullptr && qr == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled unavailable");
      a.encodeTiled = reinterpret_cast<PFN_encodeTiled>(fn);
      fn = nullptr;
      MCG_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &qr));
      MCG_CHECK(fn != nullptr && qr == cudaDriverEntryPointSuccess, "cuTensorMapEncodeIm2col unavailable");
      a.encodeIm2col = reinterpret_cast<PFN_encodeIm2col>(fn);
      return a;
    }();
    return api;
  }
};

// 2-D row-major fp16 matrix [rows, cols] with row pitch ld; box = box_cols x box_rows
inline CUtensorMap make_tmap_2d(const __half* base, long long rows, long long cols, long long ld, int box_rows,
                                int box_cols = kBlockK, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  CUtensorMap m;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = DriverApi::get().encodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), dims,
                                            strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MCG_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed, code " + std::to_string(static_cast<int>(r)));
  return m;
}

// uint8 (e4m3) variants: one byte per element
inline CUtensorMap make_tmap_2d_u8(const uint8_t* base, long long rows, long long cols, long long ld, int box_rows,
                                   int box_cols, CUtensorMapSwizzle swz) {
  CUtensorMap m;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld)};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = DriverApi::get().encodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(base), dims,
                                            strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MCG_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(u8) failed, code " + std::to_string(static_cast<int>(r)));
  return m;
}

inline CUtensorMap make_tmap_im2col_u8(const uint8_t* base, const AGeom& g) {
  CUtensorMap m;
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(g.C), static_cast<cuuint64_t>(g.W), static_cast<cuuint64_t>(g.H),
                        static_cast<cuuint64_t>(g.NB)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(g.C), static_cast<cuuint64_t>(g.W) * g.C,
                           static_cast<cuuint64_t>(g.H) * g.W * g.C};
  int lower[2] = {-g.pad, -g.pad};
  int upper[2] = {g.pad - (g.S - 1), g.pad - (g.R - 1)};
  cuuint32_t estr[4] = {1, static_cast<cuuint32_t>(g.stride), static_cast<cuuint32_t>(g.stride), 1};
  CUresult r = DriverApi::get().encodeIm2col(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<uint8_t*>(base), dims,
                                             strides, lower, upper, kBlockK, kBlockM, estr,
                                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MCG_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeIm2col(u8) failed, code " + std::to_string(static_cast<int>(r)));
  return m;
}

inline CUtensorMap make_tmap_im2col(const __half* base, const AGeom& g) {
  CUtensorMap m;
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(g.C), static_cast<cuuint64_t>(g.W), static_cast<cuuint64_t>(g.H),
                        static_cast<cuuint64_t>(g.NB)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(g.C) * 2, static_cast<cuuint64_t>(g.W) * g.C * 2,
                           static_cast<cuuint64_t>(g.H) * g.W * g.C * 2};
  int lower[2] = {-g.pad, -g.pad};                            // (W, H)
  int upper[2] = {g.pad - (g.S - 1), g.pad - (g.R - 1)};
  cuuint32_t estr[4] = {1, static_cast<cuuint32_t>(g.stride), static_cast<cuuint32_t>(g.stride), 1};
  CUresult r = DriverApi::get().encodeIm2col(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half*>(base), dims,
                                             strides, lower, upper, kBlockK, kBlockM, estr,
                                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MCG_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeIm2col failed, code " + std::to_string(static_cast<int>(r)));
  return m;
}

struct UmmaPlan {
  UmmaMaps tm;
  UmmaParams p;
  int pair = 0;  // 1: clusters of two CTAs, tcgen05 cta_group::2 (tiles of 256 rows)
  int kbs = 1;   // k-blocks per pipeline stage (kernel template parameter)
  int terms = 3;
  int grid = 0;
  int smem = 0;
};

inline bool umma_supported(long long M, int N, int K, const AGeom& a) {
  if (N % 64 != 0 || K % kBlockK != 0 || M <= 0) return false;
  if (a.kind == 1 && (a.C % kBlockK != 0)) return false;
  if (a.kind == 0 && (a.lda % 16 != 0)) return false;
  return true;
}

// tuning knobs for GPU experiments (environment, read once): MCG_TUNE_RES_BN forces block_n of the
// residual (bottleneck conv3) layers, MCG_TUNE_OUT_SETS the number of staging sets per epilogue group
inline int tune_env(const char* name) {
  const char* e = std::getenv(name);
  return e ? std::atoi(e) : 0;
}

// A planes: for kind 0 the [M,K] matrix, for kind 1 the NHWC activation.  W planes: [N,K].
// terms == 2 (fp16c8): A = hi + lo8 + hi8, W = hi + hi8 + lo8 (scales: common.cuh).
// The planes the epilogue writes / the residual carries follow from the pointers set in `ep`.
// Optional second A operand (A2, a2): K = K1 + K2 with K1 = K - a2 channels; see UmmaParams::kb2_begin.
inline UmmaPlan make_umma_plan(int terms, Planes A, const AGeom& a, Planes W, long long M, int N, int K,
                               const Epilogue& ep, int num_sms, int force_block_n = 0, int k_split = 1,
                               long long split_stride = 0, int pair = 0, const Planes* A2 = nullptr,
                               const AGeom* a2 = nullptr) {
  MCG_CHECK(umma_supported(M, N, K, a), "shape not supported by the tcgen05 GEMM");
  if (A2) {
    MCG_CHECK(a2 != nullptr && a2->R == 1 && a2->S == 1 && a2->pad == 0 && a2->C % kBlockK == 0 && a2->C < K && k_split == 1 &&
                  a.kind == 0,
              "the K-concatenated second operand must be a 1x1 convolution beside a plain first operand");
    MCG_CHECK(terms != 3 || A2->lo, "3-term GEMM needs lo planes");
    MCG_CHECK(terms != 2 || (A2->lo8 && A2->hi8), "fp16c8 GEMM needs the e4m3 planes of the second operand");
  }
  MCG_CHECK(!pair || k_split == 1, "CTA-pair tiles do not combine with split-K");
  MCG_CHECK(terms >= 1 && terms <= 3, "the tcgen05 GEMM runs 1, 2 (fp16 + e4m3 corrections) or 3 MMA terms per k-step");
  MCG_CHECK(terms != 3 || (A.lo && W.lo), "3-term GEMM needs lo planes");
  MCG_CHECK(terms != 2 || (A.lo8 && A.hi8 && W.hi8 && W.lo8), "fp16c8 GEMM needs the e4m3 lo8 / hi8 planes of both operands");
  static const int tune_res_bn = tune_env("MCG_TUNE_RES_BN");
  static const int tune_out_sets = tune_env("MCG_TUNE_OUT_SETS");
  static const int tune_bn = tune_env("MCG_TUNE_BN");
  static const int dbg_flags = tune_env("MCG_DEBUG_FLAGS");
  UmmaPlan pl;
  pl.terms = terms;
  pl.pair = pair ? 1 : 0;
  UmmaParams& p = pl.p;
  p.M = static_cast<int>(M);
  p.N = N;
  p.K = K;
  p.dbg = dbg_flags;
  p.out_fmt = ep.out_lo ? 1 : ep.out_lo8 ? 2 : 0;
  p.res_fmt = ep.res_lo ? 1 : ep.res_lo8 ? 2 : 0;
  MCG_CHECK(!(ep.out_lo && ep.out_lo8) && !(ep.res_lo && ep.res_lo8), "a tensor has either an fp16 or an e4m3 low plane");
  MCG_CHECK(ep.out_hi8 == nullptr || p.out_fmt == 2, "the hi8 plane belongs to the hi + lo8 storage format");
  // the kernel instantiation fixes the storage format of the planes it reads and writes
  const int fmt = terms == 1 ? 0 : terms == 3 ? 1 : 2;
  MCG_CHECK(ep.out_f32 != nullptr || p.out_fmt == fmt, "output planes do not match the precision mode's storage format");
  MCG_CHECK(ep.res_mode == RES_NONE || ep.res_f32 != nullptr || p.res_fmt == fmt,
            "residual planes do not match the precision mode's storage format");
  // epilogue staging (per epilogue group): TMA-store staging when the output is fp16 planes, plus residual
  // prefetch buffers for the same-shape residual
  p.out_tma = (ep.out_f32 == nullptr && ep.ldo % 16 == 0) ? 1 : 0;
  p.res_tma = (p.out_tma && ep.res_mode == RES_SAME && ep.res_f32 == nullptr && ep.res_hi != nullptr && ep.ldr % 16 == 0) ? 1 : 0;
  // nearest-2x upsampled residual (FPN top-down add): the source pixels of 128 consecutive output pixels form a window
  // of at most (rows touched / 2 + 1) * Q / 2 consecutive source pixels; when that is <= 128 for every tile the window
  // is TMA-prefetched like a same-shape residual instead of gathered with per-thread loads
  static const int tune_no_res_up = tune_env("MCG_TUNE_NO_RES_UP");
  if (p.out_tma && !tune_no_res_up && ep.res_mode == RES_UP2X && ep.res_f32 == nullptr && ep.res_hi != nullptr && ep.ldr % 16 == 0 &&
      ep.P % 2 == 0 && ep.Q % 2 == 0) {
    // exact check over every distinct tile offset inside a frame (the geometry repeats per frame)
    const long long pq = static_cast<long long>(ep.P) * ep.Q;
    long long worst = 0;
    for (long long t = 0; t * kBlockM < M && t < pq; ++t) {
      const long long m0 = t * kBlockM;
      const long long first = res_row_window_start(ep, m0);
      for (long long m = m0; m < m0 + kBlockM && m < M; ++m) worst = std::max(worst, res_row(ep, m) - first);
      if (worst >= kBlockM) break;
    }
    if (worst < kBlockM) {
      p.res_tma = 1;
      p.res_up = 1;
    }
  }
  p.out_hi8 = (p.out_tma && ep.out_hi8 != nullptr) ? 1 : 0;
  const int set_bytes = epi_set_bytes(p.out_fmt);
  // residual chunks in flight per epilogue group (more chunks cost pipeline stages: measured slower)
  static const int tune_res_bufs = tune_env("MCG_TUNE_RES_BUFS");
  p.res_bufs = tune_res_bufs >= 1 && tune_res_bufs <= kMaxResBufs ? tune_res_bufs : 2;  // measured: 2 > 3 > 4 (ring depth matters more)
  // ... as long as two stages of the widest usable tile (block_n >= 128 where N allows) still fit
  while (p.res_tma && p.res_bufs > 2) {
    const int wide = force_block_n ? force_block_n : (N % 128 == 0 ? 128 : 64);
    const int sb = a_stage_bytes(terms) + w_stage_bytes(terms, pair ? wide / 2 : wide);
    const int left = kMaxDynSmem - 1024 - kSmemBarrierBytes - kEpiGroups * p.res_bufs * epi_set_bytes(p.res_fmt) -
                     kEpiGroups * set_bytes;
    if (left / sb >= 2) break;
    --p.res_bufs;
  }
  const int res_bytes = p.res_tma ? kEpiGroups * p.res_bufs * epi_set_bytes(p.res_fmt) : 0;
  const int fixed = 1024 + kSmemBarrierBytes + res_bytes;
  auto ring_budget = [&](int out_sets) { return kMaxDynSmem - fixed - (p.out_tma ? kEpiGroups * out_sets * set_bytes : 0); };
  // residual (bottleneck conv3) layers are HBM-bound with short K loops: 2 stages are enough there
  const int min_stages = p.res_tma ? 2 : 3;
  int bn = force_block_n;
  if (bn == 0 && p.res_tma && tune_res_bn > 0 && N % tune_res_bn == 0) bn = tune_res_bn;
  if (bn == 0 && tune_bn > 0 && N % tune_bn == 0) bn = tune_bn;
  if (bn == 0) {
    // Tile width by a wave model: persistent workers (SMs, or SM pairs) each take ceil(tiles / workers) tiles of
    // cost ~ block_n (MMA cycles and W bytes per 128 rows); narrow tiles pay more A traffic / issue overhead per
    // FLOP, N = 64 MMAs run at 2/3 efficiency.  Widest tile wins ties; it must leave min_stages pipeline stages.
    static const int tune_wave = tune_env("MCG_TUNE_NO_WAVE_MODEL");
    const long long workers = pair ? num_sms / 2 : num_sms;
    const long long mt = (M + (pair ? 2 : 1) * kBlockM - 1) / ((pair ? 2 : 1) * kBlockM);
    double best = 0.0;
    const int cands[3] = {256, 128, 64};
    for (int c : cands) {
      if (N % c) continue;
      const int sb = a_stage_bytes(terms) + w_stage_bytes(terms, pair ? c / 2 : c);
      if (!(ring_budget(1) / sb >= min_stages || c == 64)) continue;
      const long long tiles = mt * (N / c);
      const long long waves = (tiles + workers - 1) / workers;
      const double unit = c == 256 ? 256.0 : c == 128 ? 128.0 * 1.25 : 64.0 * 1.9;  // measured: layer3 / layer4 A-B
      const double cost = static_cast<double>(waves) * unit;
      if (bn == 0 || cost < best * 0.97) {
        bn = c;
        best = cost;
      }
      if (tune_wave) break;  // previous rule: the widest tile that fits
    }
  }
  MCG_CHECK(bn > 0 && N % bn == 0 && bn % 64 == 0 && bn <= 256, "bad block_n");
  p.block_n = bn;
  const int w_rows = pair ? bn / 2 : bn;  // W rows per CTA
  const int sub_bytes = a_stage_bytes(terms) + w_stage_bytes(terms, w_rows);
  // k-blocks per stage (1, 2 or 3): as many as divide the tile's k-blocks, keep a stage within 64 KB and the ring
  // at >= 3 stages beside double-buffered epilogue staging
  static const int tune_kbs = tune_env("MCG_TUNE_KBS");
  int kbs = 1;
  if (k_split == 1 && tune_kbs != 1) {
    const int nkb = K / kBlockK;
    for (int c = 3; c >= 2; --c) {
      if (nkb % c == 0 && c * sub_bytes <= 64 * 1024 && ring_budget(1) / (c * sub_bytes) >= 3) {
        kbs = c;
        break;
      }
    }
  }
  pl.kbs = kbs;
  const int stage_bytes = kbs * sub_bytes;
  // double-buffer the staging when that does not cost a needed pipeline stage
  p.out_sets = 1;
  if (p.out_tma && ring_budget(2) / stage_bytes >= min_stages) p.out_sets = 2;
  if (p.out_tma && tune_out_sets > 0 && ring_budget(tune_out_sets) / stage_bytes >= 2) p.out_sets = tune_out_sets > 2 ? 2 : tune_out_sets;
  p.num_stages = ring_budget(p.out_sets) / stage_bytes;
  if (p.num_stages > kMaxStages) p.num_stages = kMaxStages;
  static const int tune_stages = tune_env("MCG_TUNE_STAGES");
  if (tune_stages >= 2 && p.num_stages > tune_stages) p.num_stages = tune_stages;
  MCG_CHECK(p.num_stages >= 2, "not enough shared memory for 2 stages");
  p.acc_cols = bn > 128 ? 256 : 128;
  p.num_acc = kTmemCols / p.acc_cols;
  p.num_kb = K / kBlockK;
  const int tile_m = pair ? 2 * kBlockM : kBlockM;
  p.m_tiles = static_cast<int>((M + tile_m - 1) / tile_m);
  p.n_tiles = N / bn;
  p.a = a;
  p.cblocks = a.kind == 1 ? a.C / kBlockK : 1;
  if (a.kind == 1) MCG_CHECK(K == a.R * a.S * a.C, "im2col K mismatch");
  if (A2) {
    p.a2 = *a2;
    p.kb2_begin = (K - a2->C) / kBlockK;
    MCG_CHECK(p.kb2_begin >= 1, "first operand needs at least one k-block");
  }
  p.ep = ep;
  if (k_split > 1) {
    MCG_CHECK(ep.out_f32 != nullptr && ep.res_mode == RES_NONE && !ep.relu && k_split <= p.num_kb && terms != 2,
              "split-K needs a plain fp32 output (bias / activation are applied by the reduction)");
    p.k_split = k_split;
    p.split_stride = split_stride;
  }
  pl.smem = fixed + (p.out_tma ? kEpiGroups * p.out_sets * set_bytes : 0) + p.num_stages * stage_bytes;
  MCG_CHECK(pl.smem <= kMaxDynSmem, "shared memory plan exceeds the 227 KB limit");
  const long long tiles = static_cast<long long>(p.m_tiles) * p.n_tiles * p.k_split;
  if (pair) {
    const long long pairs = num_sms / 2;
    pl.grid = 2 * static_cast<int>(tiles < pairs ? tiles : pairs);
  } else {
    pl.grid = static_cast<int>(tiles < num_sms ? tiles : num_sms);
  }
  UmmaMaps& tm = pl.tm;
  const int K1 = A2 ? K - a2->C : K;  // columns of the first A operand
  if (a.kind == 1) {
    tm.a_hi = make_tmap_im2col(A.hi, a);
    tm.a_lo = terms == 3 ? make_tmap_im2col(A.lo, a) : tm.a_hi;
  } else {
    tm.a_hi = make_tmap_2d(A.hi, M, K1, a.lda, kBlockM);
    tm.a_lo = terms == 3 ? make_tmap_2d(A.lo, M, K1, a.lda, kBlockM) : tm.a_hi;
  }
  tm.w_hi = make_tmap_2d(W.hi, N, K, K, w_rows);
  tm.w_lo = terms == 3 ? make_tmap_2d(W.lo, N, K, K, w_rows) : tm.w_hi;
  tm.o_hi = tm.o_lo = tm.r_hi = tm.r_lo = tm.w_hi;  // placeholders when unused
  tm.a_hi8 = tm.w_hi8 = tm.o_hi8 = tm.w_hi;
  const CUtensorMapSwizzle sw64 = CU_TENSOR_MAP_SWIZZLE_64B;
  if (terms == 2) {
    tm.a_lo = a.kind == 1 ? make_tmap_im2col_u8(A.lo8, a) : make_tmap_2d_u8(A.lo8, M, K1, a.lda, kBlockM, kBlockK, sw64);
    tm.a_hi8 = a.kind == 1 ? make_tmap_im2col_u8(A.hi8, a) : make_tmap_2d_u8(A.hi8, M, K1, a.lda, kBlockM, kBlockK, sw64);
    tm.w_hi8 = make_tmap_2d_u8(W.hi8, N, K, K, w_rows, kBlockK, sw64);
    tm.w_lo = make_tmap_2d_u8(W.lo8, N, K, K, w_rows, kBlockK, sw64);
  }
  tm.a2_hi = tm.a2_lo = tm.a2_hi8 = tm.a_hi;
  if (A2) {
    const int C2 = a2->C;
    if (a2->kind == 1) {
      tm.a2_hi = make_tmap_im2col(A2->hi, *a2);
      if (terms == 3) tm.a2_lo = make_tmap_im2col(A2->lo, *a2);
      if (terms == 2) {
        tm.a2_lo = make_tmap_im2col_u8(A2->lo8, *a2);
        tm.a2_hi8 = make_tmap_im2col_u8(A2->hi8, *a2);
      }
    } else {
      tm.a2_hi = make_tmap_2d(A2->hi, M, C2, a2->lda, kBlockM);
      if (terms == 3) tm.a2_lo = make_tmap_2d(A2->lo, M, C2, a2->lda, kBlockM);
      if (terms == 2) {
        tm.a2_lo = make_tmap_2d_u8(A2->lo8, M, C2, a2->lda, kBlockM, kBlockK, sw64);
        tm.a2_hi8 = make_tmap_2d_u8(A2->hi8, M, C2, a2->lda, kBlockM, kBlockK, sw64);
      }
    }
  }
  if (p.out_tma) {
    tm.o_hi = make_tmap_2d(ep.out_hi, M, N, ep.ldo, kBlockM, kEpiChunk, sw64);
    if (p.out_fmt == 1) tm.o_lo = make_tmap_2d(ep.out_lo, M, N, ep.ldo, kBlockM, kEpiChunk, sw64);
    if (p.out_fmt == 2) tm.o_lo = make_tmap_2d_u8(ep.out_lo8, M, N, ep.ldo, kBlockM, kEpiChunk, CU_TENSOR_MAP_SWIZZLE_32B);
    if (p.out_hi8) tm.o_hi8 = make_tmap_2d_u8(ep.out_hi8, M, N, ep.ldo, kBlockM, kEpiChunk, CU_TENSOR_MAP_SWIZZLE_32B);
  }
  if (p.res_tma) {
    const long long RM = p.res_up ? M / 4 : M;   // rows of the residual tensor
    tm.r_hi = make_tmap_2d(ep.res_hi, RM, N, ep.ldr, kBlockM, kEpiChunk, sw64);
    if (p.res_fmt == 1) tm.r_lo = make_tmap_2d(ep.res_lo, RM, N, ep.ldr, kBlockM, kEpiChunk, sw64);
    if (p.res_fmt == 2) tm.r_lo = make_tmap_2d_u8(ep.res_lo8, RM, N, ep.ldr, kBlockM, kEpiChunk, CU_TENSOR_MAP_SWIZZLE_32B);
  }
  return pl;
}

template <int kTerms, bool kPair, int kKbs>
inline void umma_launch_one(const cudaLaunchConfig_t& cfg, const UmmaPlan& pl) {
  static bool attr_set = false;
  if (!attr_set) {
    MCG_CUDA(cudaFuncSetAttribute(umma_gemm_kernel<kTerms, kPair, kKbs>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  kMaxDynSmem));
    attr_set = true;
  }
  MCG_CUDA(cudaLaunchKernelEx(&cfg, umma_gemm_kernel<kTerms, kPair, kKbs>, pl.tm, pl.p));
}
template <int kTerms, bool kPair>
inline void umma_launch_kbs(const cudaLaunchConfig_t& cfg, const UmmaPlan& pl) {
  if (pl.kbs == 3)
    umma_launch_one<kTerms, kPair, 3>(cfg, pl);
  else if (pl.kbs == 2)
    umma_launch_one<kTerms, kPair, 2>(cfg, pl);
  else
    umma_launch_one<kTerms, kPair, 1>(cfg, pl);
}
template <int kTerms>
inline void umma_launch_pair(const cudaLaunchConfig_t& cfg, const UmmaPlan& pl) {
  if (pl.pair)
    umma_launch_kbs<kTerms, true>(cfg, pl);
  else
    umma_launch_kbs<kTerms, false>(cfg, pl);
}

inline void launch_umma(const UmmaPlan& pl, cudaStream_t stream) {
  static const int tune_pdl = tune_env("MCG_TUNE_NO_PDL");
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(pl.grid);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = pl.smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pl.pair) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (!tune_pdl) {
    // may start while the previous kernel of the stream drains (the kernel waits with griddepcontrol.wait)
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  if (pl.terms == 3)
    umma_launch_pair<3>(cfg, pl);
  else if (pl.terms == 2)
    umma_launch_pair<2>(cfg, pl);
  else
    umma_launch_pair<1>(cfg, pl);
}

}  // namespace mcg
