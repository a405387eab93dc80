// Fused ResNet stem for sm_100a: 7x7/2 convolution (+ folded BN) + ReLU + 3x3/2 max-pool in ONE
// tcgen05 kernel (mmdet/models/backbones/resnet.py:599-611, 636-639).
//
// The unfused path (stem_im2col_kernel -> umma_gemm_kernel -> maxpool3x3s2_kernel) writes and re-reads
// a [pixels, 192] im2col matrix and the 112^2 x 64 stem map through HBM (2.9 GB per 224-frame step in
// split-fp16).  Here the im2col tile is built in shared memory straight from the fp32 NCHW input,
// multiplied on the tensor cores, and pooled out of TMEM: HBM traffic = input + pooled output only.
//
// Work unit = (frame n, pooled row i, column group g of 60 pooled columns): the three stem rows
// 2i-1, 2i, 2i+1 are three M-tiles (accumulators) of 128 rows x 64 channels.  Tile row m maps to stem
// column q(m) = 120 g - 1 + 30 (m / 32) + (m % 32): every warp-sized slice of 32 rows overlaps the next
// by two columns so that each 3-wide pooling window lies inside one warp of the epilogue (window of
// odd lane l = lanes l-1, l, l+1; 15 pooled columns per warp, no cross-warp exchange).
//
// Units are ordered (frame, column group, pooled row) and every CTA takes a CONTIGUOUS range of them, so the unit after
// (n, g, i) is (n, g, i + 1) and its first stem row 2i+1 is the previous unit's last one: that accumulator stays in
// TMEM ("carry") and only two of the three stem rows are built and multiplied per unit (the first unit of a CTA's
// range computes all three).  TMEM regions of 64 columns: row 2i -> A[local & 1], row 2i+1 -> C[local % 3], row 2i-1 =
// C[(local + 2) % 3] (the previous unit's C); the accumulator-empty barrier of unit local - 2 frees A[local & 1] and
// C[local % 3] (last read, as a carry, by that unit's epilogue).
//
// Roles (448 threads, persistent, one CTA per SM):
//   warp 0     : MMA issuer (tcgen05.mma M=128, N=64, K=16; kTerms = 3 issues lo*hi + hi*lo + hi*hi)
//   warp 1     : TMEM allocator, loads the packed stem weights [64 x 192] once by TMA
//   warps 2-9  : builders - cp.async the 11 input rows of the next unit into a double-buffered fp32
//                staging area, build the current unit's A k-tiles (SWIZZLE_128B K-major) as split-fp16
//   warps 10-13: epilogue - tcgen05.ld the three accumulators, max, + bias, ReLU, horizontal 3-max by
//                warp shuffles, split to fp16 hi/lo, 16-byte stores of the pooled row
#pragma once
#include "common.cuh"
#include "ptx.cuh"
#include "umma_gemm.cuh"

namespace mcg {

constexpr int kSfThreads = 448;
constexpr int kSfBuilderThreads = 256;
constexpr int kSfStages = 3;          // A ring: one stage = one [128 x 64] k-tile (hi [+ lo])
constexpr int kSfInRows = 11;         // input rows feeding three stem rows
constexpr int kSfInCols = 256;        // staged input columns per unit (249 needed)
constexpr int kSfInStride = 268;      // staging row pitch in floats: rows (c, r) that one warp instruction reads
                                      // land in different banks (pitch 256 -> 5-way conflicts, 268 -> 2-3-way)
constexpr int kSfK = 176;             // 21 chunks (r, c) x 8 columns = 168 (stem_k_index), rounded up to a multiple of 16
constexpr int kSfPoolCols = 60;       // pooled columns per unit
constexpr int kSfStageBufBytes = 3 * kSfInRows * kSfInStride * 4;  // 35376
constexpr int kSfWTileBytes = 64 * 64 * 2;                       // one [64 x 64] weight k-tile plane

struct StemFusedParams {
  const float* img = nullptr;  // [NB, 3, H, W] fp32
  const float* bias = nullptr; // [64] folded BN shift
  __half* out_hi = nullptr;    // [NB, H/4, W/4, 64]
  __half* out_lo = nullptr;
  uint8_t* out_lo8 = nullptr;  // fp16c8 storage: e4m3((v - hi) 2^11) instead of the fp16 lo plane ...
  uint8_t* out_hi8 = nullptr;  // ... and the e4m3 copy of hi
  int NB = 0, H = 0, W = 0;
  int P = 0, Q = 0;            // stem map
  int PP = 0, QQ = 0;          // pooled map
  int groups = 1;              // column groups per pooled row
  int units = 0;               // NB * PP * groups
};

struct StemFusedMaps {
  CUtensorMap w_hi, w_lo;
};

__host__ __device__ constexpr int sf_a_stage_bytes(int terms) { return (terms == 3 ? 2 : 1) * kATileBytes; }
__host__ __device__ constexpr int sf_w_bytes(int terms) { return (terms == 3 ? 2 : 1) * 3 * kSfWTileBytes; }
__host__ __device__ constexpr int sf_smem_bytes(int terms) {
  return 1024 + 1024 + sf_w_bytes(terms) + kSfStages * sf_a_stage_bytes(terms) + 2 * kSfStageBufBytes;
}

__device__ __forceinline__ void cp_async16_zfill(void* dst, const void* src, bool valid) {
  const uint32_t n = valid ? 16u : 0u;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(ptx::smem_u32(dst)), "l"(src), "r"(n) : "memory");
}

// TMEM column of stem row d (0: 2i-1 carried or computed, 1: 2i, 2: 2i+1) of the CTA's local-th unit
__device__ __forceinline__ uint32_t sf_tmem_col(int local, int d) {
  return d == 1 ? static_cast<uint32_t>((local & 1) * 64)
                : static_cast<uint32_t>(128 + ((d == 2 ? local : local + 2) % 3) * 64);
}

template <int kTerms>
__global__ void __launch_bounds__(kSfThreads, 1)
stem_fused_kernel(const __grid_constant__ StemFusedMaps tm, const StemFusedParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  ptx::grid_dep_launch_dependents();  // the first GEMM's prologue may run under this kernel's tail
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);  // [kSfStages] builders -> MMA
  uint64_t* empty_bar = full_bar + kSfStages;              // MMA -> builders
  uint64_t* tfull_bar = empty_bar + kSfStages;             // [2] MMA -> epilogue
  uint64_t* tempty_bar = tfull_bar + 2;                    // [2] epilogue -> MMA
  uint64_t* w_bar = tempty_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(w_bar + 1);
  uint8_t* w_smem = smem + 1024;
  uint8_t* a_ring = w_smem + sf_w_bytes(kTerms);
  float* stage_in = reinterpret_cast<float*>(a_ring + kSfStages * sf_a_stage_bytes(kTerms));

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // contiguous unit range of this CTA; unit = (n * groups + g) * PP + i
  const int units_per_cta = (p.units + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int u_begin = static_cast<int>(blockIdx.x) * units_per_cta;
  const int u_end = min(u_begin + units_per_cta, p.units);

  if (warp_idx == 0 && lane == 0) {
    for (int i = 0; i < kSfStages; ++i) {
      ptx::mbar_init(&full_bar[i], kSfBuilderThreads / 32);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tfull_bar[i], 1);
      ptx::mbar_init(&tempty_bar[i], 4);
    }
    ptx::mbar_init(w_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp_idx == 1) {
    ptx::tmem_alloc(tmem_ptr_smem, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp_idx == 1) {
    if (lane == 0) {
      // packed stem weights [64, 192] K-major: three [64 x 64] boxes per plane, resident for the whole kernel
      ptx::prefetch_tmap(&tm.w_hi);
      ptx::mbar_arrive_expect_tx(w_bar, static_cast<uint32_t>(sf_w_bytes(kTerms)));
      for (int kt = 0; kt < 3; ++kt) {
        ptx::tma_load_2d(w_smem + kt * kSfWTileBytes, &tm.w_hi, w_bar, kt * 64, 0);
        if (kTerms == 3) ptx::tma_load_2d(w_smem + (3 + kt) * kSfWTileBytes, &tm.w_lo, w_bar, kt * 64, 0);
      }
    }
  } else if (warp_idx == 0) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = ptx::make_idesc_f16_f32(128, 64);
    ptx::mbar_wait(w_bar, 0);
    int stage = 0;
    uint32_t phase = 0;
    int local = 0;
    for (int unit = u_begin; unit < u_end; ++unit, ++local) {
      const int ab = local & 1;
      const uint32_t ab_phase = static_cast<uint32_t>(local >> 1) & 1u;
      const int i = unit % p.PP;
      ptx::mbar_wait(&tempty_bar[ab], ab_phase ^ 1u);
      ptx::tc_fence_after();
      const int d_first = (i == 0 || local > 0) ? 1 : 0;   // top padding row / carried from the previous unit
      for (int d = d_first; d < 3; ++d) {
        const uint32_t tmem_d = tmem_base + sf_tmem_col(local, d);
        for (int kt = 0; kt < 3; ++kt) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          {
            // warp-uniform: descriptors live in uniform registers, umma_* elect the issuing lane
            const uint32_t aA_hi = ptx::smem_u32(a_ring + stage * sf_a_stage_bytes(kTerms));
            const uint32_t aA_lo = aA_hi + kATileBytes;
            const uint32_t aW_hi = ptx::smem_u32(w_smem + kt * kSfWTileBytes);
            const uint32_t aW_lo = aW_hi + 3 * kSfWTileBytes;
            const uint64_t dA_hi = ptx::make_sw128_kmajor_desc(aA_hi);
            const uint64_t dW_hi = ptx::make_sw128_kmajor_desc(aW_hi);
            const uint64_t dA_lo = ptx::make_sw128_kmajor_desc(aA_lo);
            const uint64_t dW_lo = ptx::make_sw128_kmajor_desc(aW_lo);
            const int ksteps = kt == 2 ? (kSfK - 128) / kUmmaK : 4;
            for (int j = 0; j < ksteps; ++j) {
              uint32_t accum = (kt > 0 || j > 0) ? 1u : 0u;
              if (kTerms == 3) {
                ptx::umma_f16(tmem_d, dA_lo + 2 * j, dW_hi + 2 * j, idesc, accum);
                ptx::umma_f16(tmem_d, dA_hi + 2 * j, dW_lo + 2 * j, idesc, 1u);
                accum = 1u;
              }
              ptx::umma_f16(tmem_d, dA_hi + 2 * j, dW_hi + 2 * j, idesc, accum);
            }
            ptx::umma_commit(&empty_bar[stage]);
            if (d == 2 && kt == 2) ptx::umma_commit(&tfull_bar[ab]);
          }
          if (++stage == kSfStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp_idx < 10) {
    // ===================== builders =====================
    const int bt = threadIdx.x - 64;  // 0..255
    const int bw = bt >> 5;           // builder warp = 16-byte chunk (8 consecutive k) inside a k-tile row
    // carried: the unit's first stem row comes from TMEM, its two top input rows are not read
    auto prefetch = [&](int unit, int buf, bool carried) {
      const int i = unit % p.PP;
      const int t = unit / p.PP;
      const int g = t % p.groups;
      const int n = t / p.groups;
      const int w0 = 240 * g - 8;
      float* dst = stage_in + buf * (kSfStageBufBytes / 4);
      // thread = (16-byte column chunk vc, row phase rr0): 9 unrolled copies (3 channels x rows rr0, rr0+4, rr0+8) with
      // compile-time (c, j) -- the generic v -> (row, chunk) -> (c, rr) index arithmetic was a quarter of the builder
      // warps' instructions
      static_assert(kSfInCols / 4 == 64 && kSfBuilderThreads == 256 && kSfInRows <= 12, "prefetch thread mapping");
      const int vc = bt & 63;
      const int rr0 = bt >> 6;
      const int w = w0 + 4 * vc;
      const bool w_ok = w >= 0 && w < p.W;
      const int h0 = 4 * i - 5;
      const float* img_n = p.img + static_cast<long long>(n) * 3 * p.H * p.W;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int rr = rr0 + 4 * j;
          if (rr < kSfInRows && !(carried && rr < 2)) {
            const int h = h0 + rr;
            const bool ok = w_ok && h >= 0 && h < p.H;
            const float* src = ok ? img_n + (static_cast<long long>(c) * p.H + h) * p.W + w : p.img;
            cp_async16_zfill(dst + (c * kSfInRows + rr) * kSfInStride + 4 * vc, src, ok);
          }
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int stage = 0;
    uint32_t phase = 0;
    int local = 0;
    if (u_begin < u_end) prefetch(u_begin, 0, false);
    for (int unit = u_begin; unit < u_end; ++unit, ++local) {
      const int buf = local & 1;
      const int i = unit % p.PP;
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      asm volatile("bar.sync 1, 256;" ::: "memory");  // staging of this unit complete; previous unit fully built
      if (unit + 1 < u_end) prefetch(unit + 1, buf ^ 1, true);
      const float* sin = stage_in + buf * (kSfStageBufBytes / 4);
      const int d_first = (i == 0 || local > 0) ? 1 : 0;
      for (int d = d_first; d < 3; ++d) {
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {
          uint8_t* sA = a_ring + stage * sf_a_stage_bytes(kTerms);
          // chunk gc = kt*8 + bw = r*3 + c holds input columns 2q-3 .. 2q+3 (+ one zero-weight pad) of input row
          // (c, 2d + r): the 8 floats around them are four aligned 8-byte loads, contiguous across the warp's
          // 32 consecutive stem columns (conflict-free); chunk 21 pads K to a multiple of 16 and must be zero
          const int gc = kt * 8 + bw;
          const int r = gc / 3, c = gc - 3 * r;
          float v[4][8];
          {
            // all 16 loads of the four 32-row slices first, then the conversions and stores: the compiler cannot move a
            // shared-memory load above the previous slice's shared-memory stores (possible aliasing), which left every
            // F2FP waiting on its own LDS (ncu: short_scoreboard was the builders' top stall).  The loads read the
            // staging area only, so they are issued before waiting for the ring stage to drain.
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              if (gc < 21) {
                const float2* src =
                    reinterpret_cast<const float2*>(sin + (c * kSfInRows + 2 * d + r) * kSfInStride + 60 * it + 2 * lane + 2);
                const float2 a0 = src[0], a1 = src[1], a2 = src[2], a3 = src[3];
                v[it][0] = a0.y;  // s = 0: staged column 60 it + 2 lane + 3
                v[it][1] = a1.x;
                v[it][2] = a1.y;
                v[it][3] = a2.x;
                v[it][4] = a2.y;
                v[it][5] = a3.x;
                v[it][6] = a3.y;
                v[it][7] = 0.f;   // s = 7: zero weight
              } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[it][e] = 0.f;
              }
            }
          }
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
          if (gc < kSfK / 8) {
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              const int m = it * 32 + lane;
              uint4 uh;
              uh.x = ptx::pack_half2(v[it][0], v[it][1]);
              uh.y = ptx::pack_half2(v[it][2], v[it][3]);
              uh.z = ptx::pack_half2(v[it][4], v[it][5]);
              uh.w = ptx::pack_half2(v[it][6], v[it][7]);
              const uint32_t so = static_cast<uint32_t>(m * 128 + ((bw ^ (m & 7)) << 4));
              *reinterpret_cast<uint4*>(sA + so) = uh;
              if (kTerms == 3) {
                uint4 ul;
                ul.x = ptx::residue_half2(v[it][0], v[it][1], uh.x);
                ul.y = ptx::residue_half2(v[it][2], v[it][3], uh.y);
                ul.z = ptx::residue_half2(v[it][4], v[it][5], uh.z);
                ul.w = ptx::residue_half2(v[it][6], v[it][7], uh.w);
                *reinterpret_cast<uint4*>(sA + kATileBytes + so) = ul;
              }
            }
          }
          ptx::fence_proxy_async();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&full_bar[stage]);
          if (++stage == kSfStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else {
    // ===================== epilogue =====================
    const int quarter = warp_idx & 3;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    int local = 0;
    for (int unit = u_begin; unit < u_end; ++unit, ++local) {
      const int ab = local & 1;
      const uint32_t ab_phase = static_cast<uint32_t>(local >> 1) & 1u;
      const int i = unit % p.PP;
      const int t = unit / p.PP;
      const int g = t % p.groups;
      const int n = t / p.groups;
      const int q = 120 * g - 1 + 30 * quarter + lane;  // stem column of this lane's row
      const bool q_ok = q >= 0 && q < p.Q;
      const int j = kSfPoolCols * g + 15 * quarter + (lane >> 1);  // pooled column owned by odd lanes < 31
      const bool owner = (lane & 1) && lane < 31 && j < p.QQ;
      ptx::mbar_wait(&tfull_bar[ab], ab_phase);
      ptx::tc_fence_after();
      const int d_first = i == 0 ? 1 : 0;   // row 2i-1 (computed by this unit or carried) exists unless it is the padding row
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        float v[32];
        {
          uint32_t r[32];
          ptx::tmem_ld_32x32(tmem_base + lane_addr + sf_tmem_col(local, 2) + static_cast<uint32_t>(half * 32), r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 32; ++c) v[c] = __uint_as_float(r[c]);
        }
        for (int d = d_first; d < 2; ++d) {
          uint32_t r[32];
          ptx::tmem_ld_32x32(tmem_base + lane_addr + sf_tmem_col(local, d) + static_cast<uint32_t>(half * 32), r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 32; ++c) v[c] = fmaxf(v[c], __uint_as_float(r[c]));
        }
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + half * 32);
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const float4 b = __ldg(b4 + c4);
          v[4 * c4 + 0] = fmaxf(v[4 * c4 + 0] + b.x, 0.f);
          v[4 * c4 + 1] = fmaxf(v[4 * c4 + 1] + b.y, 0.f);
          v[4 * c4 + 2] = fmaxf(v[4 * c4 + 2] + b.z, 0.f);
          v[4 * c4 + 3] = fmaxf(v[4 * c4 + 3] + b.w, 0.f);
        }
        // out-of-range stem columns are max-pool padding: 0 is neutral after the ReLU
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const float x = q_ok ? v[c] : 0.f;
          const float l = __shfl_up_sync(0xffffffffu, x, 1);
          const float rgt = __shfl_down_sync(0xffffffffu, x, 1);
          v[c] = fmaxf(x, fmaxf(l, rgt));
        }
        if (owner) {
          const long long o = ((static_cast<long long>(n) * p.PP + i) * p.QQ + j) * 64 + half * 32;
          uint4* oh = reinterpret_cast<uint4*>(p.out_hi + o);
          uint4* ol = p.out_lo ? reinterpret_cast<uint4*>(p.out_lo + o) : nullptr;
#pragma unroll
          for (int c8 = 0; c8 < 4; ++c8) {
            uint4 uh, ul;
            uh.x = ptx::pack_half2(v[8 * c8 + 0], v[8 * c8 + 1]);
            uh.y = ptx::pack_half2(v[8 * c8 + 2], v[8 * c8 + 3]);
            uh.z = ptx::pack_half2(v[8 * c8 + 4], v[8 * c8 + 5]);
            uh.w = ptx::pack_half2(v[8 * c8 + 6], v[8 * c8 + 7]);
            oh[c8] = uh;
            if (ol) {
              ul.x = ptx::residue_half2(v[8 * c8 + 0], v[8 * c8 + 1], uh.x);
              ul.y = ptx::residue_half2(v[8 * c8 + 2], v[8 * c8 + 3], uh.y);
              ul.z = ptx::residue_half2(v[8 * c8 + 4], v[8 * c8 + 5], uh.z);
              ul.w = ptx::residue_half2(v[8 * c8 + 6], v[8 * c8 + 7], uh.w);
              ol[c8] = ul;
            }
            if (p.out_lo8) {
              uint2 l8;
              l8.x = residue_e4m3x2(v[8 * c8 + 0], v[8 * c8 + 1], uh.x) | (residue_e4m3x2(v[8 * c8 + 2], v[8 * c8 + 3], uh.y) << 16);
              l8.y = residue_e4m3x2(v[8 * c8 + 4], v[8 * c8 + 5], uh.z) | (residue_e4m3x2(v[8 * c8 + 6], v[8 * c8 + 7], uh.w) << 16);
              *reinterpret_cast<uint2*>(p.out_lo8 + o + 8 * c8) = l8;
            }
            if (p.out_hi8) {
              uint2 h8;
              h8.x = half2_to_e4m3x2(uh.x) | (half2_to_e4m3x2(uh.y) << 16);
              h8.y = half2_to_e4m3x2(uh.z) | (half2_to_e4m3x2(uh.w) << 16);
              *reinterpret_cast<uint2*>(p.out_hi8 + o + 8 * c8) = h8;
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty_bar[ab]);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

struct StemFusedPlan {
  StemFusedMaps tm;
  StemFusedParams p;
  int terms = 3;
  int grid = 0;
  int smem = 0;
};

inline bool stem_fused_supported(int H, int W) { return H % 4 == 0 && W % 8 == 0 && H >= 8 && W >= 8; }

// W: packed stem weight planes [64, 192] (k = (r*7+s)*3 + c, zero padded)
inline StemFusedPlan make_stem_fused_plan(int terms, const float* img, int NB, int H, int W, Planes Wt, const float* bias,
                                          Planes out, int num_sms) {
  MCG_CHECK(stem_fused_supported(H, W), "fused stem needs H % 4 == 0 and W % 8 == 0");
  MCG_CHECK(terms == 1 || (terms == 3 && Wt.lo && (out.lo || out.lo8)), "fused stem: 3-term mode needs lo planes");
  StemFusedPlan pl;
  pl.terms = terms;
  StemFusedParams& p = pl.p;
  p.img = img;
  p.bias = bias;
  p.out_hi = out.hi;
  p.out_lo = terms == 3 ? out.lo : nullptr;
  p.out_lo8 = terms == 3 ? out.lo8 : nullptr;
  p.out_hi8 = terms == 3 ? out.hi8 : nullptr;
  p.NB = NB;
  p.H = H;
  p.W = W;
  p.P = H / 2;
  p.Q = W / 2;
  p.PP = H / 4;
  p.QQ = W / 4;
  p.groups = (p.QQ + kSfPoolCols - 1) / kSfPoolCols;
  p.units = NB * p.PP * p.groups;
  pl.tm.w_hi = make_tmap_2d(Wt.hi, 64, 192, 192, 64);
  pl.tm.w_lo = terms == 3 ? make_tmap_2d(Wt.lo, 64, 192, 192, 64) : pl.tm.w_hi;
  pl.smem = sf_smem_bytes(terms);
  pl.grid = p.units < num_sms ? p.units : num_sms;
  return pl;
}

inline void launch_stem_fused(const StemFusedPlan& pl, cudaStream_t stream) {
  static bool attrs = false;
  if (!attrs) {
    MCG_CUDA(cudaFuncSetAttribute(stem_fused_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, sf_smem_bytes(1)));
    MCG_CUDA(cudaFuncSetAttribute(stem_fused_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, sf_smem_bytes(3)));
    attrs = true;
  }
  if (pl.terms == 3)
    stem_fused_kernel<3><<<pl.grid, kSfThreads, pl.smem, stream>>>(pl.tm, pl.p);
  else
    stem_fused_kernel<1><<<pl.grid, kSfThreads, pl.smem, stream>>>(pl.tm, pl.p);
  MCG_CUDA(cudaGetLastError());
}

}  // namespace mcg
