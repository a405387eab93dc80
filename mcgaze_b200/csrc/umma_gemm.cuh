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
// 1e-3 (yaw,pitch) parity bar against the fp32 oracle).  kTerms == 2 ("fp16lo8") keeps hi*hi and
// hi*lo_w in fp16 but stores the activations' low part as e4m3 (x 2^13) and adds lo8_a * hi8_w as an
// fp8 MMA (kind::f8f6f4, half the cycles) into a SECOND TMEM accumulator that the epilogue scales
// and adds: 3 bytes per activation element instead of 4 and 2.5 instead of 3 MMA units per k-step.
//
// Structure (persistent, one CTA per SM, 256 threads):
//   warp 0   : TMA producer  (one elected lane; A tile 128 x 64, W tile block_n x 64, SWIZZLE_128B)
//   warp 1   : MMA issuer    (one lane issues tcgen05.mma M=128, N=block_n, K=16; commit -> mbarrier)
//   warp 2   : TMEM allocator (512 columns = 2 accumulator buffers of up to 256 fp32 columns)
//   warps 4-7: epilogue.  Each warp owns 32 accumulator rows (its TMEM lane quarter) and walks the
//              tile in 32-column chunks: tcgen05.ld -> + bias -> + residual -> ReLU -> split to
//              fp16 hi/lo -> 64B-swizzled smem staging -> TMA store (bulk async group), double
//              buffered so the store of chunk c overlaps the math of chunk c+1.  A same-shape
//              residual (bottleneck identity) is TMA-prefetched one chunk ahead into smem by the
//              same warp, so neither stores nor residual reads are issued as per-thread strided
//              global accesses.  fp32 outputs (head) and the FPN's 2x-upsampled residual use
//              direct vector loads/stores.
// Pipelines: smem full/empty ring (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue), per-warp
// residual mbarriers, so the epilogue of tile i overlaps the main loop of tile i+1.
#pragma once
#include <cmath>
#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"

namespace mcg {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // 64 fp16 = 128 B = one swizzle row
constexpr int kUmmaK = 16;
constexpr int kGemmThreads = 256;
constexpr int kTmemCols = 512;
constexpr int kMaxStages = 8;
constexpr int kATileBytes = kBlockM * kBlockK * 2;  // 16 KB
constexpr int kSmemBarrierBytes = 1024;
constexpr int kMaxDynSmem = 232448;  // 227 KB: the sm_100 opt-in limit per block
constexpr int kEpiChunk = 64;        // columns per epilogue chunk (128 B of fp16 per row)
constexpr int kEpiHiBytes = kBlockM * kEpiChunk * 2;  // [128 rows x 64 cols] fp16 staging tile = 16 KB
constexpr int kResBufs = 2;          // residual chunks (128 x 64) in flight, prefetched across tile boundaries

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
  int out_tma = 0;  // planes output through smem staging + TMA store
  int out_sets = 1; // staging sets for the TMA-store epilogue (2 = double buffered)
  int res_tma = 0;  // RES_SAME residual planes prefetched by TMA
  float corr_scale = 0.f;  // fp16lo8: factor of the fp8 correction accumulator = 2^-(13 + weight shift)
  int dbg = 0;             // attribution experiments only (env MCG_DEBUG_FLAGS): 1 no stores, 2 no epilogue math,
                           // 4 no A loads, 8 no W loads, 16 no MMA issue, 32 no residual loads.  Results are garbage.
  AGeom a;
  Epilogue ep;
};

struct UmmaMaps {
  CUtensorMap a_hi, a_lo, w_hi, w_lo, o_hi, o_lo, r_hi, r_lo;
  CUtensorMap a_lo8, w_hi8, o_lo8, r_lo8;  // fp16lo8 mode (uint8 tensors)
};

// shared-memory bytes of one pipeline stage / one epilogue staging set for a precision mode
__host__ __device__ constexpr int a_stage_bytes(int terms) { return terms == 3 ? 2 * kATileBytes : (terms == 2 ? kATileBytes + kATileBytes / 2 : kATileBytes); }
__host__ __device__ constexpr int w_stage_bytes(int terms, int bn) {
  return terms == 3 ? 2 * bn * kBlockK * 2 : (terms == 2 ? 2 * bn * kBlockK * 2 + bn * kBlockK : bn * kBlockK * 2);
}
// one epilogue staging set / residual slot: fp16 hi tile (+ fp16 lo tile | + e4m3 lo8 tile of half the size)
__host__ __device__ constexpr int epi_set_bytes(int terms) {
  return terms == 3 ? 2 * kEpiHiBytes : (terms == 2 ? kEpiHiBytes + kEpiHiBytes / 2 : kEpiHiBytes);
}

// byte offset of logical 16-byte chunk j of row `row` in a 64B-swizzled [rows][64 B] tile
__device__ __forceinline__ uint32_t sw64_off(int row, int j) {
  return static_cast<uint32_t>(row * 64 + ((j ^ ((row >> 1) & 3)) << 4));
}
// same for a 128B-swizzled [rows][128 B] tile (8 chunks per row)
__device__ __forceinline__ uint32_t sw128_off(int row, int j) {
  return static_cast<uint32_t>(row * 128 + ((j ^ (row & 7)) << 4));
}

template <int kTerms>
__global__ void __launch_bounds__(kGemmThreads, 1)
umma_gemm_kernel(const __grid_constant__ UmmaMaps tm, const UmmaParams p) {
  constexpr int kSet = epi_set_bytes(kTerms);  // bytes of one epilogue staging set (hi tile [+ lo / lo8 tile])
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tfull_bar = empty_bar + kMaxStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* res_bar = tempty_bar + 2;  // [kResBufs]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(res_bar + kResBufs);
  uint8_t* stage_base = smem + kSmemBarrierBytes;

  const int w_tile_bytes = p.block_n * kBlockK * 2;
  const int a_bytes = a_stage_bytes(kTerms);
  const int stage_bytes = a_bytes + w_stage_bytes(kTerms, p.block_n);
  uint8_t* obuf_base = stage_base + static_cast<size_t>(p.num_stages) * stage_bytes;  // [out_sets][kSet]
  uint8_t* rbuf_base = obuf_base + (p.out_tma ? p.out_sets * kSet : 0);               // [kResBufs][kSet]
  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int mn_tiles = p.m_tiles * p.n_tiles;
  const int num_tiles = mn_tiles * p.k_split;

  if (warp_idx == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm.a_hi);
    ptx::prefetch_tmap(&tm.w_hi);
    if (kTerms == 3) {
      ptx::prefetch_tmap(&tm.a_lo);
      ptx::prefetch_tmap(&tm.w_lo);
    }
    if (kTerms == 2) {
      ptx::prefetch_tmap(&tm.a_lo8);
      ptx::prefetch_tmap(&tm.w_lo);
      ptx::prefetch_tmap(&tm.w_hi8);
    }
    if (p.out_tma) ptx::prefetch_tmap(&tm.o_hi);
    if (p.res_tma) ptx::prefetch_tmap(&tm.r_hi);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int i = 0; i < p.num_stages; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tfull_bar[i], 1);
      ptx::mbar_init(&tempty_bar[i], 4);  // one arrive per epilogue warp
    }
    for (int i = 0; i < kResBufs; ++i) ptx::mbar_init(&res_bar[i], 1);
    ptx::fence_mbar_init();
  }
  if (warp_idx == 2) {
    ptx::tmem_alloc(tmem_ptr_smem, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp_idx == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int ks = tile / mn_tiles;
        const int mn = tile - ks * mn_tiles;
        const int m_tile = mn / p.n_tiles;
        const int n_tile = mn - m_tile * p.n_tiles;
        const int kb_begin = static_cast<int>(static_cast<long long>(ks) * p.num_kb / p.k_split);
        const int kb_end = static_cast<int>(static_cast<long long>(ks + 1) * p.num_kb / p.k_split);
        const long long m0 = static_cast<long long>(m_tile) * kBlockM;
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
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* s = stage_base + static_cast<size_t>(stage) * stage_bytes;
          uint8_t* sA_hi = s;
          uint8_t* sA_lo = s + kATileBytes;  // fp16 lo tile (x3) or e4m3 lo8 tile (fp16lo8)
          uint8_t* sW_hi = s + a_bytes;
          uint8_t* sW_lo = sW_hi + w_tile_bytes;
          uint8_t* sW_hi8 = sW_lo + w_tile_bytes;
          uint32_t tx = static_cast<uint32_t>(stage_bytes);
          if (p.dbg & 4) tx -= static_cast<uint32_t>(a_bytes);
          if (p.dbg & 8) tx -= static_cast<uint32_t>(stage_bytes - a_bytes);
          if (tx == 0) {
            ptx::mbar_arrive(&full_bar[stage]);
          } else {
            ptx::mbar_arrive_expect_tx(&full_bar[stage], tx);
          }
          if (p.dbg & 4) {
          } else if (p.a.kind == 1) {
            const int tap = kb / p.cblocks;
            const int cb = kb - tap * p.cblocks;
            const int r = tap / p.a.S;
            const int sx = tap - r * p.a.S;
            ptx::tma_load_im2col_4d(sA_hi, &tm.a_hi, &full_bar[stage], cb * kBlockK, base_w, base_h, img_n,
                                    static_cast<uint16_t>(sx), static_cast<uint16_t>(r));
            if (kTerms == 3)
              ptx::tma_load_im2col_4d(sA_lo, &tm.a_lo, &full_bar[stage], cb * kBlockK, base_w, base_h, img_n,
                                      static_cast<uint16_t>(sx), static_cast<uint16_t>(r));
            if (kTerms == 2)
              ptx::tma_load_im2col_4d(sA_lo, &tm.a_lo8, &full_bar[stage], cb * kBlockK, base_w, base_h, img_n,
                                      static_cast<uint16_t>(sx), static_cast<uint16_t>(r));
          } else {
            ptx::tma_load_2d(sA_hi, &tm.a_hi, &full_bar[stage], kb * kBlockK, static_cast<int>(m0));
            if (kTerms == 3) ptx::tma_load_2d(sA_lo, &tm.a_lo, &full_bar[stage], kb * kBlockK, static_cast<int>(m0));
            if (kTerms == 2) ptx::tma_load_2d(sA_lo, &tm.a_lo8, &full_bar[stage], kb * kBlockK, static_cast<int>(m0));
          }
          if (!(p.dbg & 8)) {
            ptx::tma_load_2d(sW_hi, &tm.w_hi, &full_bar[stage], kb * kBlockK, n_tile * p.block_n);
            if (kTerms >= 2) ptx::tma_load_2d(sW_lo, &tm.w_lo, &full_bar[stage], kb * kBlockK, n_tile * p.block_n);
            if (kTerms == 2) ptx::tma_load_2d(sW_hi8, &tm.w_hi8, &full_bar[stage], kb * kBlockK, n_tile * p.block_n);
          }
          if (++stage == p.num_stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = ptx::make_idesc_f16_f32(kBlockM, p.block_n);
    int stage = 0;
    uint32_t phase = 0;
    int local = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
      ptx::tc_fence_after();
      const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * 256);
      const int ks = tile / mn_tiles;
      const int kb_begin = static_cast<int>(static_cast<long long>(ks) * p.num_kb / p.k_split);
      const int kb_end = static_cast<int>(static_cast<long long>(ks + 1) * p.num_kb / p.k_split);
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        if (lane == 0 && (p.dbg & 16)) {
          ptx::umma_commit(&empty_bar[stage]);
          if (kb == kb_end - 1) ptx::umma_commit(&tfull_bar[acc]);
        } else if (lane == 0) {
          const uint32_t s = ptx::smem_u32(stage_base + static_cast<size_t>(stage) * stage_bytes);
          const uint32_t aA_hi = s;
          const uint32_t aA_lo = s + kATileBytes;
          const uint32_t aW_hi = s + a_bytes;
          const uint32_t aW_lo = aW_hi + w_tile_bytes;
          const uint32_t aW_hi8 = aW_lo + w_tile_bytes;
#pragma unroll
          for (int j = 0; j < kBlockK / kUmmaK; ++j) {
            const uint32_t koff = j * kUmmaK * 2;  // bytes inside the 128 B swizzle row
            const uint64_t dA_hi = ptx::make_sw128_kmajor_desc(aA_hi + koff);
            const uint64_t dW_hi = ptx::make_sw128_kmajor_desc(aW_hi + koff);
            uint32_t accum = (kb > kb_begin || j > 0) ? 1u : 0u;
            if (kTerms == 3) {
              const uint64_t dA_lo = ptx::make_sw128_kmajor_desc(aA_lo + koff);
              const uint64_t dW_lo = ptx::make_sw128_kmajor_desc(aW_lo + koff);
              ptx::umma_f16(tmem_d, dA_lo, dW_hi, idesc, accum);
              ptx::umma_f16(tmem_d, dA_hi, dW_lo, idesc, 1u);
              accum = 1u;
            }
            if (kTerms == 2) {
              const uint64_t dW_lo = ptx::make_sw128_kmajor_desc(aW_lo + koff);
              ptx::umma_f16(tmem_d, dA_hi, dW_lo, idesc, accum);
              accum = 1u;
            }
            ptx::umma_f16(tmem_d, dA_hi, dW_hi, idesc, accum);
          }
          if (kTerms == 2) {
            // fp8 correction lo8_a * hi8_w into the second accumulator (columns +128), K = 32 per MMA
#pragma unroll
            for (int j = 0; j < kBlockK / 32; ++j) {
              const uint64_t dA8 = ptx::make_sw64_kmajor_desc(aA_lo + j * 32);
              const uint64_t dW8 = ptx::make_sw64_kmajor_desc(aW_hi8 + j * 32);
              ptx::umma_f8(tmem_d + 128u, dA8, dW8, idesc, (kb > kb_begin || j > 0) ? 1u : 0u);
            }
          }
          ptx::umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
          if (kb == kb_end - 1) ptx::umma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        if (++stage == p.num_stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp_idx >= 4) {
    // ===================== epilogue =====================
    // The four warps work in lock step on one 64-column chunk of the 128-row tile at a time so that
    // ONE thread can move the whole [128 x 64] chunk with a single TMA op per plane (the per-SM TMA
    // unit is shared with the main loop's loads: many small boxes starve it).
    const int quarter = warp_idx & 3;
    const int row = quarter * 32 + lane;  // row of the tile == TMEM lane
    const bool leader = threadIdx.x == 128;
    const Epilogue& ep = p.ep;
    const int nchunks = p.block_n / kEpiChunk;
    constexpr uint32_t kResTx = kTerms == 3 ? 2 * kEpiHiBytes : (kTerms == 2 ? kEpiHiBytes + kEpiHiBytes / 2 : kEpiHiBytes);
    // residual chunk stream, prefetched kResBufs chunks ahead across tile boundaries by the leader
    int ri_tile = blockIdx.x, ri_c = 0;
    uint32_t r_issued = 0, r_consumed = 0;
    auto res_issue = [&]() {
      if (ri_tile >= num_tiles) return;
      const int mn_i = ri_tile % mn_tiles;
      const int mt = mn_i / p.n_tiles;
      const int nt = mn_i - mt * p.n_tiles;
      const uint32_t b = r_issued % kResBufs;
      if (leader && (p.dbg & 32)) {
        ptx::mbar_arrive(&res_bar[b]);
      } else if (leader) {
        uint8_t* dst = rbuf_base + b * kSet;
        ptx::fence_proxy_async();
        ptx::mbar_arrive_expect_tx(&res_bar[b], kResTx);
        ptx::tma_load_2d(dst, &tm.r_hi, &res_bar[b], nt * p.block_n + ri_c * kEpiChunk, mt * kBlockM);
        if (kTerms == 3) ptx::tma_load_2d(dst + kEpiHiBytes, &tm.r_lo, &res_bar[b], nt * p.block_n + ri_c * kEpiChunk, mt * kBlockM);
        if (kTerms == 2) ptx::tma_load_2d(dst + kEpiHiBytes, &tm.r_lo8, &res_bar[b], nt * p.block_n + ri_c * kEpiChunk, mt * kBlockM);
      }
      ++r_issued;
      if (++ri_c == nchunks) {
        ri_c = 0;
        ri_tile += gridDim.x;
      }
    };
    if (p.res_tma) {
#pragma unroll
      for (int i = 0; i < kResBufs; ++i) res_issue();
    }
    uint32_t ostores = 0;  // chunks handed to TMA so far (staging set = ostores % osets)
    const uint32_t osets = static_cast<uint32_t>(p.out_sets);
    int local = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
      const int ks = tile / mn_tiles;
      const int mn = tile - ks * mn_tiles;
      const int m_tile = mn / p.n_tiles;
      const int n_tile = mn - m_tile * p.n_tiles;
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      const long long m = static_cast<long long>(m_tile) * kBlockM + row;
      const bool valid = m < p.M;
      const int n_base = n_tile * p.block_n;
      ptx::mbar_wait(&tfull_bar[acc], acc_phase);
      ptx::tc_fence_after();
      const long long rrow = (valid && ep.res_mode != RES_NONE && !p.res_tma) ? res_row(ep, m) : 0;
      const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * 256);
      // gathered residual (FPN nearest-2x top-down add): per-thread vector loads, software-pipelined
      // one 32-column half ahead in registers so the L2 latency overlaps the previous half's math
      const bool res_direct = valid && ep.res_mode != RES_NONE && !p.res_tma && ep.res_f32 == nullptr;
      uint4 rpre[8];
      auto load_direct = [&](int n) {
        const uint4* rh = reinterpret_cast<const uint4*>(ep.res_hi + rrow * ep.ldr + n);
#pragma unroll
        for (int j = 0; j < 4; ++j) rpre[j] = __ldg(rh + j);
        if (ep.res_lo) {
          const uint4* rl = reinterpret_cast<const uint4*>(ep.res_lo + rrow * ep.ldr + n);
#pragma unroll
          for (int j = 0; j < 4; ++j) rpre[4 + j] = __ldg(rl + j);
        }
        if (ep.res_lo8) {
          const uint4* rl = reinterpret_cast<const uint4*>(ep.res_lo8 + rrow * ep.ldr + n);
#pragma unroll
          for (int j = 0; j < 2; ++j) rpre[4 + j] = __ldg(rl + j);
        }
      };
      if (res_direct) load_direct(n_base);
      for (int c = 0; c < nchunks; ++c) {
        const uint8_t* rcur = nullptr;
        if (p.res_tma) {
          const uint32_t b = r_consumed % kResBufs;
          ptx::mbar_wait(&res_bar[b], (r_consumed / kResBufs) & 1u);
          rcur = rbuf_base + b * kSet;
          ++r_consumed;
        }
        uint8_t* ob = obuf_base + (ostores % osets) * kSet;
        if (p.out_tma) {
          // the staging set was handed to TMA `osets` chunks ago: wait until that store has read it
          if (leader) {
            if (osets == 2) ptx::tma_store_wait_read<1>(); else ptx::tma_store_wait_read<0>();
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
#pragma unroll 1
        for (int hf = 0; hf < ((p.dbg & 2) ? 0 : 2); ++hf) {
          const int n = n_base + c * kEpiChunk + hf * 32;
          uint4 rnow[8];
          if (res_direct) {
#pragma unroll
            for (int j = 0; j < 8; ++j) rnow[j] = rpre[j];
            if (hf == 0 || c + 1 < nchunks) load_direct(n + 32);
          }
          uint32_t r[32];
          float v[32];
          ptx::tmem_ld_32x32(taddr0 + static_cast<uint32_t>(c * kEpiChunk + hf * 32), r);
          if (kTerms == 2) {
            uint32_t r2[32];
            ptx::tmem_ld_32x32(taddr0 + 128u + static_cast<uint32_t>(c * kEpiChunk + hf * 32), r2);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaf(__uint_as_float(r2[j]), p.corr_scale, __uint_as_float(r[j]));
          } else {
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          }
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
            if (kTerms == 2) {  // e4m3 low part: [128 rows][64 B], 64B swizzle; this half = 16-byte chunks 2hf, 2hf+1
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                const uint4 u = *reinterpret_cast<const uint4*>(rcur + kEpiHiBytes + sw64_off(row, hf * 2 + j));
                float l8[8];
                e4m3x8_to_float(make_uint2(u.x, u.y), l8);
#pragma unroll
                for (int t = 0; t < 8; ++t) v[16 * j + t] = fmaf(l8[t], kLo8InvScale, v[16 * j + t]);
                e4m3x8_to_float(make_uint2(u.z, u.w), l8);
#pragma unroll
                for (int t = 0; t < 8; ++t) v[16 * j + 8 + t] = fmaf(l8[t], kLo8InvScale, v[16 * j + 8 + t]);
              }
            }
#pragma unroll
            for (int pl = 0; pl < (kTerms == 3 ? 2 : 1); ++pl) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint4 u = *reinterpret_cast<const uint4*>(rcur + pl * kEpiHiBytes + sw128_off(row, hf * 4 + j));
                const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                  const float2 f = __half22float2(h2[t]);
                  v[8 * j + 2 * t] += f.x;
                  v[8 * j + 2 * t + 1] += f.y;
                }
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
              if (ep.res_lo8) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                  const uint4 u = rnow[4 + j];
                  float l8[8];
                  e4m3x8_to_float(make_uint2(u.x, u.y), l8);
#pragma unroll
                  for (int t = 0; t < 8; ++t) v[16 * j + t] = fmaf(l8[t], kLo8InvScale, v[16 * j + t]);
                  e4m3x8_to_float(make_uint2(u.z, u.w), l8);
#pragma unroll
                  for (int t = 0; t < 8; ++t) v[16 * j + 8 + t] = fmaf(l8[t], kLo8InvScale, v[16 * j + 8 + t]);
                }
              }
              const int npl = ep.res_lo ? 2 : 1;
#pragma unroll
              for (int pl = 0; pl < 2; ++pl) {
                if (pl < npl) {
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const uint4 u = rnow[pl * 4 + j];
                    const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                      const float2 f = __half22float2(h2[t]);
                      v[8 * j + 2 * t] += f.x;
                      v[8 * j + 2 * t + 1] += f.y;
                    }
                  }
                }
              }
            }
          }
          if (ep.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          if (p.out_tma) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 uh, ul;
              __half2* hh = reinterpret_cast<__half2*>(&uh);
              __half2* hl = reinterpret_cast<__half2*>(&ul);
              float rs[8];
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const float a = v[8 * j + 2 * t], b = v[8 * j + 2 * t + 1];
                const __half2 h = __floats2half2_rn(a, b);
                const float2 hfv = __half22float2(h);
                hh[t] = h;
                hl[t] = __floats2half2_rn(a - hfv.x, b - hfv.y);
                rs[2 * t] = (a - hfv.x) * kLo8Scale;
                rs[2 * t + 1] = (b - hfv.y) * kLo8Scale;
              }
              *reinterpret_cast<uint4*>(ob + sw128_off(row, hf * 4 + j)) = uh;
              if (kTerms == 3) *reinterpret_cast<uint4*>(ob + kEpiHiBytes + sw128_off(row, hf * 4 + j)) = ul;
              if (kTerms == 2)  // low part as e4m3 of (v - hi) * 2^13: 8 bytes per 8 columns
                *reinterpret_cast<uint2*>(ob + kEpiHiBytes + sw64_off(row, hf * 2 + (j >> 1)) + (j & 1) * 8) =
                    float8_to_e4m3x8(rs);
            }
          } else if (valid) {
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
                __half2* hh = reinterpret_cast<__half2*>(&uh);
                __half2* hl = reinterpret_cast<__half2*>(&ul);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                  const float a = v[8 * j + 2 * t], b = v[8 * j + 2 * t + 1];
                  const __half2 h = __floats2half2_rn(a, b);
                  const float2 hfv = __half22float2(h);
                  hh[t] = h;
                  hl[t] = __floats2half2_rn(a - hfv.x, b - hfv.y);
                }
                oh[j] = uh;
                if (ol) ol[j] = ul;
              }
            }
          }
        }
        if (p.res_tma || p.out_tma) {
          // all 128 threads have consumed the residual slot and filled the staging set
          ptx::fence_proxy_async();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (p.res_tma) res_issue();
          if (p.out_tma) {
            if (leader && !(p.dbg & 1)) {
              const int n = n_base + c * kEpiChunk;
              const int m0 = m_tile * kBlockM;
              ptx::tma_store_2d(&tm.o_hi, ob, n, m0);
              if (kTerms == 3) ptx::tma_store_2d(&tm.o_lo, ob + kEpiHiBytes, n, m0);
              if (kTerms == 2) ptx::tma_store_2d(&tm.o_lo8, ob + kEpiHiBytes, n, m0);
              ptx::tma_store_commit();
            } else if (leader) {
              ptx::tma_store_commit();
            }
            ++ostores;
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);
    }
    // smem must stay valid until every bulk store has been read out
    if (p.out_tma && leader) ptx::tma_store_wait_all<0>();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp_idx == 2) {
    ptx::tc_fence_after();
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

// A planes: for kind 0 the [M,K] matrix, for kind 1 the NHWC activation.  W planes: [N,K].
// W.lo8 (terms == 2) holds e4m3(W_hi * 2^w_shift); the fp8 accumulator is scaled by 2^-(13 + w_shift).
inline UmmaPlan make_umma_plan(int terms, Planes A, const AGeom& a, Planes W, long long M, int N, int K,
                               const Epilogue& ep, int num_sms, int force_block_n = 0, int k_split = 1,
                               long long split_stride = 0, int w_shift = 0) {
  MCG_CHECK(umma_supported(M, N, K, a), "shape not supported by the tcgen05 GEMM");
  MCG_CHECK(terms != 3 || (A.lo && W.lo), "3-term GEMM needs lo planes");
  MCG_CHECK(terms != 2 || (A.lo8 && W.lo && W.lo8), "fp16lo8 GEMM needs an e4m3 activation low plane and fp16 lo + e4m3 hi weights");
  UmmaPlan pl;
  pl.terms = terms;
  UmmaParams& p = pl.p;
  p.M = static_cast<int>(M);
  p.N = N;
  p.K = K;
  // epilogue staging (per CTA): TMA-store staging when the output is fp16 planes, plus residual
  // prefetch buffers for the same-shape residual
  const bool out_lo_ok = terms == 1 || (terms == 3 && ep.out_lo != nullptr) || (terms == 2 && ep.out_lo8 != nullptr);
  const bool res_lo_ok = terms == 1 || (terms == 3 && ep.res_lo != nullptr) || (terms == 2 && ep.res_lo8 != nullptr);
  p.out_tma = (ep.out_f32 == nullptr && ep.ldo % 16 == 0 && out_lo_ok) ? 1 : 0;
  MCG_CHECK(ep.out_f32 != nullptr || terms != 2 || p.out_tma, "fp16lo8 planes output needs the TMA-store epilogue");
  p.res_tma = (p.out_tma && ep.res_mode == RES_SAME && ep.res_f32 == nullptr && ep.res_hi != nullptr && ep.ldr % 16 == 0 &&
               res_lo_ok)
                  ? 1
                  : 0;
  // staging: double buffered when two ring stages still fit beside it (the HBM-bound layers), else single
  const int set_bytes = epi_set_bytes(terms);
  const int res_bytes = p.res_tma ? kResBufs * set_bytes : 0;
  const int min_ring = 2 * (a_stage_bytes(terms) + w_stage_bytes(terms, 128 < N ? 128 : 64));
  p.out_sets = 1;
  if (p.out_tma && p.res_tma && kMaxDynSmem - 1024 - kSmemBarrierBytes - res_bytes - 2 * set_bytes >= min_ring) p.out_sets = 2;
  const int epi_bytes = (p.out_tma ? p.out_sets * set_bytes : 0) + res_bytes;
  if (terms == 2) p.corr_scale = std::ldexp(1.0f, -(kLo8Shift + w_shift));
  const int ring_budget = kMaxDynSmem - 1024 - kSmemBarrierBytes - epi_bytes;
  int bn = force_block_n;
  if (bn == 0) {
    // largest tile that divides N and still leaves >= 3 pipeline stages
    const int cands[3] = {256, 128, 64};
    for (int c : cands) {
      if (N % c) continue;
      if (terms == 2 && c > 128) continue;  // two accumulators (main + fp8 correction) share a 256-column buffer
      const int sb = a_stage_bytes(terms) + w_stage_bytes(terms, c);
      // residual (bottleneck conv3) layers are HBM-bound with short K loops: 2 stages are enough there
      if (ring_budget / sb >= (p.res_tma ? 2 : 3) || c == 64) {
        bn = c;
        break;
      }
    }
  }
  MCG_CHECK(bn > 0 && N % bn == 0, "bad block_n");
  p.block_n = bn;
  MCG_CHECK(terms != 2 || bn <= 128, "fp16lo8 needs block_n <= 128");
  const int stage_bytes = a_stage_bytes(terms) + w_stage_bytes(terms, bn);
  p.num_stages = ring_budget / stage_bytes;
  if (p.num_stages > kMaxStages) p.num_stages = kMaxStages;
  MCG_CHECK(p.num_stages >= 2, "not enough shared memory for 2 stages");
  p.num_kb = K / kBlockK;
  p.m_tiles = static_cast<int>((M + kBlockM - 1) / kBlockM);
  p.n_tiles = N / bn;
  p.a = a;
  p.cblocks = a.kind == 1 ? a.C / kBlockK : 1;
  if (a.kind == 1) MCG_CHECK(K == a.R * a.S * a.C, "im2col K mismatch");
  p.ep = ep;
  {
    static const int dbg_flags = [] {
      const char* e = std::getenv("MCG_DEBUG_FLAGS");
      return e ? std::atoi(e) : 0;
    }();
    p.dbg = dbg_flags;
  }
  if (k_split > 1) {
    MCG_CHECK(ep.out_f32 != nullptr && ep.res_mode == RES_NONE && !ep.relu && k_split <= p.num_kb,
              "split-K needs a plain fp32 output (bias / activation are applied by the reduction)");
    p.k_split = k_split;
    p.split_stride = split_stride;
  }
  pl.smem = 1024 + kSmemBarrierBytes + p.num_stages * stage_bytes + epi_bytes;
  MCG_CHECK(pl.smem <= kMaxDynSmem, "shared memory plan exceeds the 227 KB limit");
  const long long tiles = static_cast<long long>(p.m_tiles) * p.n_tiles * p.k_split;
  pl.grid = static_cast<int>(tiles < num_sms ? tiles : num_sms);
  UmmaMaps& tm = pl.tm;
  if (a.kind == 1) {
    tm.a_hi = make_tmap_im2col(A.hi, a);
    tm.a_lo = terms == 3 ? make_tmap_im2col(A.lo, a) : tm.a_hi;
  } else {
    tm.a_hi = make_tmap_2d(A.hi, M, K, a.lda, kBlockM);
    tm.a_lo = terms == 3 ? make_tmap_2d(A.lo, M, K, a.lda, kBlockM) : tm.a_hi;
  }
  tm.w_hi = make_tmap_2d(W.hi, N, K, K, bn);
  tm.w_lo = terms == 3 ? make_tmap_2d(W.lo, N, K, K, bn) : tm.w_hi;
  tm.o_hi = tm.o_lo = tm.r_hi = tm.r_lo = tm.w_hi;  // placeholders when unused
  tm.a_lo8 = tm.w_hi8 = tm.o_lo8 = tm.r_lo8 = tm.w_hi;
  if (terms == 2) {
    tm.a_lo8 = a.kind == 1 ? make_tmap_im2col_u8(A.lo8, a)
                           : make_tmap_2d_u8(A.lo8, M, K, a.lda, kBlockM, kBlockK, CU_TENSOR_MAP_SWIZZLE_64B);
    tm.w_lo = make_tmap_2d(W.lo, N, K, K, bn);
    tm.w_hi8 = make_tmap_2d_u8(W.lo8, N, K, K, bn, kBlockK, CU_TENSOR_MAP_SWIZZLE_64B);
    if (p.out_tma) tm.o_lo8 = make_tmap_2d_u8(ep.out_lo8, M, N, ep.ldo, kBlockM, kEpiChunk, CU_TENSOR_MAP_SWIZZLE_64B);
    if (p.res_tma) tm.r_lo8 = make_tmap_2d_u8(ep.res_lo8, M, N, ep.ldr, kBlockM, kEpiChunk, CU_TENSOR_MAP_SWIZZLE_64B);
  }
  if (p.out_tma) {
    tm.o_hi = make_tmap_2d(ep.out_hi, M, N, ep.ldo, kBlockM, kEpiChunk, CU_TENSOR_MAP_SWIZZLE_128B);
    tm.o_lo = terms == 3 ? make_tmap_2d(ep.out_lo, M, N, ep.ldo, kBlockM, kEpiChunk, CU_TENSOR_MAP_SWIZZLE_128B) : tm.o_hi;
  }
  if (p.res_tma) {
    tm.r_hi = make_tmap_2d(ep.res_hi, M, N, ep.ldr, kBlockM, kEpiChunk, CU_TENSOR_MAP_SWIZZLE_128B);
    tm.r_lo = terms == 3 ? make_tmap_2d(ep.res_lo, M, N, ep.ldr, kBlockM, kEpiChunk, CU_TENSOR_MAP_SWIZZLE_128B) : tm.r_hi;
  }
  return pl;
}

inline void umma_set_attrs() {
  static bool done = false;
  if (done) return;
  MCG_CUDA(cudaFuncSetAttribute(umma_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
  MCG_CUDA(cudaFuncSetAttribute(umma_gemm_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
  MCG_CUDA(cudaFuncSetAttribute(umma_gemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
  done = true;
}

inline void launch_umma(const UmmaPlan& pl, cudaStream_t stream) {
  umma_set_attrs();
  if (pl.terms == 3)
    umma_gemm_kernel<3><<<pl.grid, kGemmThreads, pl.smem, stream>>>(pl.tm, pl.p);
  else if (pl.terms == 2)
    umma_gemm_kernel<2><<<pl.grid, kGemmThreads, pl.smem, stream>>>(pl.tm, pl.p);
  else
    umma_gemm_kernel<1><<<pl.grid, kGemmThreads, pl.smem, stream>>>(pl.tm, pl.p);
  MCG_CUDA(cudaGetLastError());
}

}  // namespace mcg
