// Fused bottleneck tail for sm_100a: conv2 (3x3, stride 1) -> BN -> ReLU -> conv3 (1x1) -> BN -> + identity -> ReLU
// (mmdet/models/backbones/resnet.py:277-302 for the blocks without a downsample branch) as ONE persistent tcgen05
// kernel.  The intermediate t2 = relu(conv2(t1)) never leaves the SM: the epilogue of conv2 writes it into shared
// memory in the K-major swizzled layout the tensor core reads operands from, and conv3 multiplies it from there.
//
// fp16c8 precision mode only (see umma_gemm.cuh for the scheme): every GEMM runs an e4m3 correction pass and an fp16
// pass into one fp32 TMEM accumulator, so t2 lives on chip as three operand planes per 64-channel k-block:
//   hi  [128 x 64] fp16, 128B-swizzled (16 KB) | lo8 [128 x 64] e4m3, 64B-swizzled (8 KB) | hi8 the same (8 KB)
//
// Per CTA (or CTA pair: 256 rows, cta_group::2) and tile of 128 output pixels:
//   C2(i):  acc1[i % nbuf1] = im2col(t1) x W2         (TMA im2col A tiles + W2 tiles through the smem ring)
//   E1(i):  acc1 -> + b2 -> ReLU -> hi / lo8 / hi8 -> t2 operand planes in shared memory           (warps 4-7)
//   C3(i):  for each 128-wide slice j of the N3 = 4 x planes outputs: acc3[q % 2] = t2 x W3[j]      (W3 through the ring)
//   E3(q):  acc3 -> + b3 -> + identity (TMA-prefetched) -> ReLU -> planes -> staging -> TMA store   (warps 8-11 / 12-15)
// The issue order is software-pipelined, C2(0) C2(1) C3(0) C2(2) C3(1) ..., so the tensor core runs conv2 of the next
// tile while E1 converts the current one; the single t2 buffer is handed back by a tcgen05.commit after C3(i).
//
// Roles (512 threads): warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-7 E1, warps 8-15 two E3
// groups.  TMEM: nbuf1 x N1 columns for acc1 + 2 x 128 for acc3 (<= 512).
#pragma once
#include "common.cuh"
#include "ptx.cuh"
#include "umma_gemm.cuh"

namespace mcg {

constexpr int kBfThreads = 512;
constexpr int kBfN3Tile = 128;                       // conv3 output slice per accumulator
constexpr int kBfT2KbBytes = 2 * kATileBytes;        // hi 16 KB + lo8 8 KB + hi8 8 KB per 64-channel k-block of t2
constexpr int kBfMaxStages = 8;
constexpr int kBfSet = 2 * kEpiPlaneBytes;           // E3 staging set: hi 8 KB + lo8 4 KB + hi8 4 KB per [128 x 32] chunk
constexpr int kBfRSet = kEpiPlaneBytes + kEpiPlaneBytes / 2;   // residual slot: hi 8 KB + lo8 4 KB

struct BneckParams {
  int M = 0;         // output pixels
  int N1 = 0;        // planes (conv2 in = out channels): 64 / 128 / 256
  int N3 = 0;        // 4 x planes
  int m_tiles = 0;   // tiles of 128 (256 for a pair) rows
  int num_stages = 0;
  int stage_bytes = 0;
  int nbuf1 = 2;     // acc1 buffers
  int out_sets = 1;
  int res_slots = 2; // identity chunks in flight per E3 group (1 or 2)
  int reverse = 0;   // walk the tiles from the end (see UmmaParams::reverse)
  int out_hi8 = 1;
  AGeom a;           // im2col geometry of conv2 (3x3, stride 1, pad 1)
  const float* bias2 = nullptr;
  const float* bias3 = nullptr;
};

struct BneckMaps {
  CUtensorMap a_hi, a_lo8, a_hi8;       // t1 (im2col views)
  CUtensorMap w2_hi, w2_hi8, w2_lo8;    // [N1, 9 N1]
  CUtensorMap w3_hi, w3_hi8, w3_lo8;    // [N3, N1]
  CUtensorMap r_hi, r_lo8;              // identity [M, N3]
  CUtensorMap o_hi, o_lo8, o_hi8;       // block output [M, N3]
};

template <bool kPair>
__global__ void __launch_bounds__(kBfThreads, 1) bneck_tail_kernel(const __grid_constant__ BneckMaps tm, const BneckParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  ptx::grid_dep_launch_dependents();

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty_bar = full_bar + kBfMaxStages;
  uint64_t* a1full = empty_bar + kBfMaxStages;   // [2]
  uint64_t* a1empty = a1full + 2;                // [2]
  uint64_t* a3full = a1empty + 2;                // [2]
  uint64_t* a3empty = a3full + 2;                // [2]
  uint64_t* t2full = a3empty + 2;                // [1]
  uint64_t* t2empty = t2full + 1;                // [1]
  uint64_t* res_bar = t2empty + 1;               // [2 groups][2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(res_bar + 4);
  const int cb = p.N1 / kBlockK;                 // 64-channel k-blocks of t1 / t2
  uint8_t* t2_base = smem + kSmemBarrierBytes;                      // [cb][hi | lo8 | hi8]
  uint8_t* stage_base = t2_base + cb * kBfT2KbBytes;
  uint8_t* obuf_base = stage_base + static_cast<size_t>(p.num_stages) * p.stage_bytes;   // [2 groups][out_sets][kBfSet]
  uint8_t* rbuf_base = obuf_base + 2 * p.out_sets * kBfSet;                               // [2 groups][2][kBfRSet]

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = kPair ? ptx::cluster_ctarank() : 0u;
  const bool leader_cta = cta_rank == 0;
  const int w2_rows = kPair ? p.N1 / 2 : p.N1;             // rows of W2 / W3 slices this CTA stages
  const int w3_rows = kPair ? kBfN3Tile / 2 : kBfN3Tile;
  const int nt3 = p.N3 / kBfN3Tile;
  const int kb2 = 9 * cb;                                   // k-blocks of conv2 per pass
  const int walker = kPair ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int walkers = kPair ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int n_local = walker < p.m_tiles ? (p.m_tiles - walker + walkers - 1) / walkers : 0;
  const int rows_per_tile = kPair ? 2 * kBlockM : kBlockM;
  // i-th tile of this walker
  auto tile_of = [&](int i) {
    const int t = walker + i * walkers;
    return p.reverse ? p.m_tiles - 1 - t : t;
  };

  if (warp_idx == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm.a_hi);
    ptx::prefetch_tmap(&tm.a_lo8);
    ptx::prefetch_tmap(&tm.a_hi8);
    ptx::prefetch_tmap(&tm.w2_hi);
    ptx::prefetch_tmap(&tm.w2_hi8);
    ptx::prefetch_tmap(&tm.w2_lo8);
    ptx::prefetch_tmap(&tm.w3_hi);
    ptx::prefetch_tmap(&tm.w3_hi8);
    ptx::prefetch_tmap(&tm.w3_lo8);
    ptx::prefetch_tmap(&tm.r_hi);
    ptx::prefetch_tmap(&tm.o_hi);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int i = 0; i < p.num_stages; ++i) {
      ptx::mbar_init(&full_bar[i], kPair ? 2 : 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&a1full[i], 1);
      ptx::mbar_init(&a1empty[i], kPair ? 8 : 4);   // one arrive per E1 warp (of both CTAs, on the leader's barrier)
      ptx::mbar_init(&a3full[i], 1);
      ptx::mbar_init(&a3empty[i], kPair ? 8 : 4);
    }
    ptx::mbar_init(t2full, kPair ? 8 : 4);
    ptx::mbar_init(t2empty, 1);
    for (int i = 0; i < 4; ++i) ptx::mbar_init(&res_bar[i], 1);
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
  if (kPair) ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t acc3_col = static_cast<uint32_t>(p.nbuf1 * p.N1);
  ptx::grid_dep_wait();

  if (warp_idx == 0) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t full_a = ptx::smem_u32(full_bar), empty_a = ptx::smem_u32(empty_bar);
    const uint32_t ring_a = ptx::smem_u32(stage_base);
    auto tma_2d = [&](uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
      if (kPair)
        ptx::tma_load_2d_pair_a(dst, m, bar, c0, c1);
      else
        ptx::tma_load_2d_a(dst, m, bar, c0, c1);
    };
    auto tma_im2col = [&](uint32_t dst, const CUtensorMap* m, uint32_t bar, int c, int w, int h, int n, uint16_t ow, uint16_t oh) {
      if (kPair)
        ptx::tma_load_im2col_4d_pair_a(dst, m, bar, c, w, h, n, ow, oh);
      else
        ptx::tma_load_im2col_4d_a(dst, m, bar, c, w, h, n, ow, oh);
    };
    // arm the stage's full barrier for `tx` bytes landing in THIS CTA (the leader counts both CTAs' bytes)
    auto arm = [&](uint32_t fb, uint32_t tx) {
      if (kPair) {
        if (leader_cta)
          ptx::mbar_arrive_expect_tx_a(full_a + stage * 8, 2 * tx);
        else
          ptx::mbar_arrive_cluster_a(fb);
      } else {
        ptx::mbar_arrive_expect_tx_a(fb, tx);
      }
    };
    auto advance = [&]() {
      if (++stage == p.num_stages) {
        stage = 0;
        phase ^= 1u;
      }
    };
    auto load_c2 = [&](int tile) {
      const long long m0 = (static_cast<long long>(tile) * (kPair ? 2 : 1) + cta_rank) * kBlockM;
      const int n0 = static_cast<int>(cta_rank) * (kPair ? w2_rows : 0);
      const long long pq = static_cast<long long>(p.a.P) * p.a.Q;
      const int img_n = static_cast<int>(m0 / pq);
      const int rem = static_cast<int>(m0 - img_n * pq);
      const int p0 = rem / p.a.Q;
      const int q0 = rem - p0 * p.a.Q;
      const int base_h = p0 - 1, base_w = q0 - 1;    // stride 1, pad 1
      for (int pass = 0; pass < 2; ++pass) {
        int tap = 0, c = 0;
        for (int kb = 0; kb < kb2; ++kb) {
          ptx::mbar_wait_a(empty_a + stage * 8, phase ^ 1u);
          if (ptx::elect_one()) {
            const uint32_t s = ring_a + static_cast<uint32_t>(stage * p.stage_bytes);
            const uint32_t fb = kPair ? ptx::mapa(full_a + stage * 8, 0) : full_a + stage * 8;
            arm(fb, static_cast<uint32_t>(kATileBytes + w2_rows * kBlockK * 2));
            const uint16_t tr = static_cast<uint16_t>(tap / 3), ts = static_cast<uint16_t>(tap % 3);
            if (pass == 0) {
              tma_im2col(s, &tm.a_lo8, fb, c * kBlockK, base_w, base_h, img_n, ts, tr);
              tma_im2col(s + kATileBytes / 2, &tm.a_hi8, fb, c * kBlockK, base_w, base_h, img_n, ts, tr);
              tma_2d(s + kATileBytes, &tm.w2_hi8, fb, kb * kBlockK, n0);
              tma_2d(s + kATileBytes + w2_rows * kBlockK, &tm.w2_lo8, fb, kb * kBlockK, n0);
            } else {
              tma_im2col(s, &tm.a_hi, fb, c * kBlockK, base_w, base_h, img_n, ts, tr);
              tma_2d(s + kATileBytes, &tm.w2_hi, fb, kb * kBlockK, n0);
            }
          }
          __syncwarp();
          if (++c == cb) {
            c = 0;
            ++tap;
          }
          advance();
        }
      }
    };
    auto load_c3 = [&](int /*tile*/) {
      for (int j = 0; j < nt3; ++j) {
        const int n0 = j * kBfN3Tile + static_cast<int>(cta_rank) * (kPair ? w3_rows : 0);
        for (int pass = 0; pass < 2; ++pass) {
          for (int kb = 0; kb < cb; ++kb) {
            ptx::mbar_wait_a(empty_a + stage * 8, phase ^ 1u);
            if (ptx::elect_one()) {
              const uint32_t s = ring_a + static_cast<uint32_t>(stage * p.stage_bytes);
              const uint32_t fb = kPair ? ptx::mapa(full_a + stage * 8, 0) : full_a + stage * 8;
              arm(fb, static_cast<uint32_t>(w3_rows * kBlockK * 2));
              if (pass == 0) {
                tma_2d(s, &tm.w3_hi8, fb, kb * kBlockK, n0);
                tma_2d(s + w3_rows * kBlockK, &tm.w3_lo8, fb, kb * kBlockK, n0);
              } else {
                tma_2d(s, &tm.w3_hi, fb, kb * kBlockK, n0);
              }
            }
            __syncwarp();
            advance();
          }
        }
      }
    };
    for (int i = 0; i <= n_local; ++i) {
      if (i < n_local) load_c2(tile_of(i));
      if (i >= 1) load_c3(tile_of(i - 1));
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer (leader CTA) =====================
    const uint32_t idesc2 = ptx::make_idesc_f16_f32(kPair ? 2 * kBlockM : kBlockM, p.N1);
    const uint32_t idesc3 = ptx::make_idesc_f16_f32(kPair ? 2 * kBlockM : kBlockM, kBfN3Tile);
    const uint32_t stage_u32 = ptx::smem_u32(stage_base);
    const uint32_t t2_u32 = ptx::smem_u32(t2_base);
    const uint32_t full_a = ptx::smem_u32(full_bar), empty_a = ptx::smem_u32(empty_bar);
    auto mma_f16 = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
      if (kPair)
        ptx::umma_f16_pair(d, a, b, idesc, acc);
      else
        ptx::umma_f16(d, a, b, idesc, acc);
    };
    auto mma_f8 = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
      if (kPair)
        ptx::umma_f8_pair(d, a, b, idesc, acc);
      else
        ptx::umma_f8(d, a, b, idesc, acc);
    };
    auto rescale = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
      if (kPair)
        ptx::umma_f16_rescale_d_pair<kC8AccShift>(d, a, b, idesc);
      else
        ptx::umma_f16_rescale_d<kC8AccShift>(d, a, b, idesc);
    };
    auto commit = [&](uint32_t bar) {
      if (kPair)
        ptx::umma_commit_pair_a(bar);
      else
        ptx::umma_commit_a(bar);
    };
    int stage = 0;
    uint32_t phase = 0;
    auto advance = [&]() {
      if (++stage == p.num_stages) {
        stage = 0;
        phase ^= 1u;
      }
    };
    auto mma_c2 = [&](int i) {
      const int b = p.nbuf1 == 2 ? (i & 1) : 0;
      const uint32_t use = static_cast<uint32_t>(p.nbuf1 == 2 ? (i >> 1) : i);
      ptx::mbar_wait_a(ptx::smem_u32(&a1empty[b]), (use & 1u) ^ 1u);
      ptx::tc_fence_after();
      const uint32_t d = tmem_base + static_cast<uint32_t>(b * p.N1);
      for (int pass = 0; pass < 2; ++pass) {
        for (int kb = 0; kb < kb2; ++kb) {
          ptx::mbar_wait_a(full_a + stage * 8, phase);
          ptx::tc_fence_after();
          const uint32_t s = stage_u32 + static_cast<uint32_t>(stage * p.stage_bytes);
          if (pass == 0) {
            const uint64_t dA_lo8 = ptx::make_sw64_kmajor_desc(s);
            const uint64_t dA_hi8 = ptx::make_sw64_kmajor_desc(s + kATileBytes / 2);
            const uint64_t dW_hi8 = ptx::make_sw64_kmajor_desc(s + kATileBytes);
            const uint64_t dW_lo8 = ptx::make_sw64_kmajor_desc(s + kATileBytes + w2_rows * kBlockK);
#pragma unroll
            for (int j = 0; j < kBlockK / 32; ++j) {
              mma_f8(d, dA_lo8 + 2 * j, dW_hi8 + 2 * j, idesc2, (kb > 0 || j > 0) ? 1u : 0u);
              mma_f8(d, dA_hi8 + 2 * j, dW_lo8 + 2 * j, idesc2, 1u);
            }
          } else {
            const uint64_t dA_hi = ptx::make_sw128_kmajor_desc(s);
            const uint64_t dW_hi = ptx::make_sw128_kmajor_desc(s + kATileBytes);
#pragma unroll
            for (int j = 0; j < kBlockK / kUmmaK; ++j) {
              if (kb == 0 && j == 0)
                rescale(d, dA_hi, dW_hi, idesc2);   // D = A*B + D 2^-15: brings the e4m3 corrections to scale
              else
                mma_f16(d, dA_hi + 2 * j, dW_hi + 2 * j, idesc2, 1u);
            }
          }
          commit(empty_a + stage * 8);
          if (pass == 1 && kb == kb2 - 1) commit(ptx::smem_u32(&a1full[b]));
          advance();
        }
      }
    };
    auto mma_c3 = [&](int i) {
      ptx::mbar_wait_a(ptx::smem_u32(t2full), static_cast<uint32_t>(i) & 1u);
      ptx::tc_fence_after();
      for (int j = 0; j < nt3; ++j) {
        const int q = i * nt3 + j;
        const int g = q & 1;
        ptx::mbar_wait_a(ptx::smem_u32(&a3empty[g]), (static_cast<uint32_t>(q >> 1) & 1u) ^ 1u);
        ptx::tc_fence_after();
        const uint32_t d = tmem_base + acc3_col + static_cast<uint32_t>(g * kBfN3Tile);
        for (int pass = 0; pass < 2; ++pass) {
          for (int kb = 0; kb < cb; ++kb) {
            ptx::mbar_wait_a(full_a + stage * 8, phase);
            ptx::tc_fence_after();
            const uint32_t s = stage_u32 + static_cast<uint32_t>(stage * p.stage_bytes);
            const uint32_t a = t2_u32 + static_cast<uint32_t>(kb * kBfT2KbBytes);
            if (pass == 0) {
              const uint64_t dA_lo8 = ptx::make_sw64_kmajor_desc(a + kATileBytes);
              const uint64_t dA_hi8 = ptx::make_sw64_kmajor_desc(a + kATileBytes + kATileBytes / 2);
              const uint64_t dW_hi8 = ptx::make_sw64_kmajor_desc(s);
              const uint64_t dW_lo8 = ptx::make_sw64_kmajor_desc(s + w3_rows * kBlockK);
#pragma unroll
              for (int jj = 0; jj < kBlockK / 32; ++jj) {
                mma_f8(d, dA_lo8 + 2 * jj, dW_hi8 + 2 * jj, idesc3, (kb > 0 || jj > 0) ? 1u : 0u);
                mma_f8(d, dA_hi8 + 2 * jj, dW_lo8 + 2 * jj, idesc3, 1u);
              }
            } else {
              const uint64_t dA_hi = ptx::make_sw128_kmajor_desc(a);
              const uint64_t dW_hi = ptx::make_sw128_kmajor_desc(s);
#pragma unroll
              for (int jj = 0; jj < kBlockK / kUmmaK; ++jj) {
                if (kb == 0 && jj == 0)
                  rescale(d, dA_hi, dW_hi, idesc3);
                else
                  mma_f16(d, dA_hi + 2 * jj, dW_hi + 2 * jj, idesc3, 1u);
              }
            }
            commit(empty_a + stage * 8);
            if (pass == 1 && kb == cb - 1) commit(ptx::smem_u32(&a3full[g]));
            advance();
          }
        }
      }
      commit(ptx::smem_u32(t2empty));   // every MMA that reads this tile's t2 has retired -> E1 may overwrite it
    };
    if (leader_cta) {
      for (int i = 0; i <= n_local; ++i) {
        if (i < n_local) mma_c2(i);
        if (i >= 1) mma_c3(i - 1);
      }
    }
  } else if (warp_idx >= 4 && warp_idx < 8) {
    // ===================== E1: acc1 -> t2 operand planes in shared memory =====================
    const int quarter = warp_idx & 3;
    const int row = quarter * 32 + lane;
    const int nchunks = p.N1 / kEpiChunk;
    for (int i = 0; i < n_local; ++i) {
      const int b = p.nbuf1 == 2 ? (i & 1) : 0;
      const uint32_t use = static_cast<uint32_t>(p.nbuf1 == 2 ? (i >> 1) : i);
      ptx::mbar_wait(&a1full[b], use & 1u);
      ptx::mbar_wait(t2empty, (static_cast<uint32_t>(i) & 1u) ^ 1u);   // conv3 of the previous tile is done with t2
      ptx::tc_fence_after();
      const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(b * p.N1);
#pragma unroll 1
      for (int c = 0; c < nchunks; ++c) {
        uint32_t r[32];
        ptx::tmem_ld_32x32(taddr0 + static_cast<uint32_t>(c * kEpiChunk), r);
        ptx::tmem_ld_wait();
        float v[32];
        const float4* b4 = reinterpret_cast<const float4*>(p.bias2 + c * kEpiChunk);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bb = __ldg(b4 + j);
          v[4 * j + 0] = fmaxf(__uint_as_float(r[4 * j + 0]) + bb.x, 0.f);
          v[4 * j + 1] = fmaxf(__uint_as_float(r[4 * j + 1]) + bb.y, 0.f);
          v[4 * j + 2] = fmaxf(__uint_as_float(r[4 * j + 2]) + bb.z, 0.f);
          v[4 * j + 3] = fmaxf(__uint_as_float(r[4 * j + 3]) + bb.w, 0.f);
        }
        // k-block of t2 this chunk belongs to, and its half (32 of the 64 channels)
        uint8_t* kbp = t2_base + (c >> 1) * kBfT2KbBytes;
        const int half = c & 1;
        uint2 l8[4], h8[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 uh;
          uh.x = ptx::pack_half2(v[8 * j + 0], v[8 * j + 1]);
          uh.y = ptx::pack_half2(v[8 * j + 2], v[8 * j + 3]);
          uh.z = ptx::pack_half2(v[8 * j + 4], v[8 * j + 5]);
          uh.w = ptx::pack_half2(v[8 * j + 6], v[8 * j + 7]);
          // hi plane: 128-byte rows, 16-byte chunk index XOR (row & 7) (SWIZZLE_128B, K-major)
          const int ch = half * 4 + j;
          *reinterpret_cast<uint4*>(kbp + row * 128 + ((ch ^ (row & 7)) << 4)) = uh;
          l8[j].x = residue_e4m3x2(v[8 * j + 0], v[8 * j + 1], uh.x) | (residue_e4m3x2(v[8 * j + 2], v[8 * j + 3], uh.y) << 16);
          l8[j].y = residue_e4m3x2(v[8 * j + 4], v[8 * j + 5], uh.z) | (residue_e4m3x2(v[8 * j + 6], v[8 * j + 7], uh.w) << 16);
          h8[j].x = half2_to_e4m3x2(uh.x) | (half2_to_e4m3x2(uh.y) << 16);
          h8[j].y = half2_to_e4m3x2(uh.z) | (half2_to_e4m3x2(uh.w) << 16);
        }
        // e4m3 planes: 64-byte rows, 64B-swizzled; this chunk fills 16-byte chunks 2 half, 2 half + 1 of the row
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          *reinterpret_cast<uint4*>(kbp + kATileBytes + sw64_off(row, half * 2 + j)) =
              make_uint4(l8[2 * j].x, l8[2 * j].y, l8[2 * j + 1].x, l8[2 * j + 1].y);
          *reinterpret_cast<uint4*>(kbp + kATileBytes + kATileBytes / 2 + sw64_off(row, half * 2 + j)) =
              make_uint4(h8[2 * j].x, h8[2 * j].y, h8[2 * j + 1].x, h8[2 * j + 1].y);
        }
      }
      // t2 is read by the tensor core through the async proxy
      ptx::fence_proxy_async();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kPair && !leader_cta) {
          ptx::mbar_arrive_cluster_a(ptx::mapa(ptx::smem_u32(&a1empty[b]), 0));
          ptx::mbar_arrive_cluster_a(ptx::mapa(ptx::smem_u32(t2full), 0));
        } else {
          ptx::mbar_arrive(&a1empty[b]);
          ptx::mbar_arrive(t2full);
        }
      }
    }
  } else if (warp_idx >= 8) {
    // ===================== E3: acc3 -> + b3 + identity -> ReLU -> planes -> TMA store =====================
    const int grp = (warp_idx - 8) >> 2;
    const int quarter = warp_idx & 3;
    const int row = quarter * 32 + lane;
    const bool leader = (warp_idx & 3) == 0 && lane == 0;
    const int bar_id = 1 + grp;
    constexpr int nchunks = kBfN3Tile / kEpiChunk;   // 4
    const uint32_t osets = static_cast<uint32_t>(p.out_sets);
    uint8_t* obuf = obuf_base + grp * p.out_sets * kBfSet;
    uint8_t* rbuf = rbuf_base + grp * p.res_slots * kBfRSet;
    const uint32_t rslots = static_cast<uint32_t>(p.res_slots);
    uint64_t* rbar = res_bar + grp * 2;
    const int total_q = n_local * nt3;
    // residual chunk stream of this group's tasks, prefetched two chunks ahead
    int ri_q = grp, ri_c = 0;
    uint32_t r_issued = 0, r_consumed = 0;
    auto res_issue = [&]() {
      if (ri_q >= total_q) return;
      const uint32_t b = r_issued % rslots;
      if (leader) {
        const int i = ri_q / nt3, j = ri_q - i * nt3;
        const int tile = tile_of(i);
        const int mt = tile * (kPair ? 2 : 1) + static_cast<int>(cta_rank);
        uint8_t* dst = rbuf + b * kBfRSet;
        ptx::fence_proxy_async();
        ptx::mbar_arrive_expect_tx(&rbar[b], static_cast<uint32_t>(kEpiPlaneBytes + kEpiPlaneBytes / 2));
        ptx::tma_load_2d(dst, &tm.r_hi, &rbar[b], j * kBfN3Tile + ri_c * kEpiChunk, mt * kBlockM);
        ptx::tma_load_2d(dst + kEpiPlaneBytes, &tm.r_lo8, &rbar[b], j * kBfN3Tile + ri_c * kEpiChunk, mt * kBlockM);
      }
      ++r_issued;
      if (++ri_c == nchunks) {
        ri_c = 0;
        ri_q += 2;
      }
    };
    for (int k = 0; k < p.res_slots; ++k) res_issue();
    uint32_t ostores = 0;
    int n = 0;
    for (int q = grp; q < total_q; q += 2, ++n) {
      const int i = q / nt3, j = q - i * nt3;
      const int tile = tile_of(i);
      const int m_tile_cta = tile * (kPair ? 2 : 1) + static_cast<int>(cta_rank);
      ptx::mbar_wait(&a3full[grp], static_cast<uint32_t>(n) & 1u);
      ptx::tc_fence_after();
      const uint32_t taddr0 =
          tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc3_col + static_cast<uint32_t>(grp * kBfN3Tile);
#pragma unroll 1
      for (int c = 0; c < nchunks; ++c) {
        const int ncol = j * kBfN3Tile + c * kEpiChunk;
        uint32_t r[32];
        ptx::tmem_ld_32x32(taddr0 + static_cast<uint32_t>(c * kEpiChunk), r);
        const uint32_t rb = r_consumed % rslots;
        ptx::mbar_wait(&rbar[rb], (r_consumed / rslots) & 1u);
        const uint8_t* rcur = rbuf + rb * kBfRSet;
        ++r_consumed;
        ptx::tmem_ld_wait();
        float v[32];
        const float4* b4 = reinterpret_cast<const float4*>(p.bias3 + ncol);
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const float4 bb = __ldg(b4 + jj);
          v[4 * jj + 0] = __uint_as_float(r[4 * jj + 0]) + bb.x;
          v[4 * jj + 1] = __uint_as_float(r[4 * jj + 1]) + bb.y;
          v[4 * jj + 2] = __uint_as_float(r[4 * jj + 2]) + bb.z;
          v[4 * jj + 3] = __uint_as_float(r[4 * jj + 3]) + bb.w;
        }
        {
          const uint8_t* r8 = rcur + kEpiLo8Off;
          add_lo8x16(v, *reinterpret_cast<const uint4*>(r8 + sw32_off(row, 0)));
          add_lo8x16(v + 16, *reinterpret_cast<const uint4*>(r8 + sw32_off(row, 1)));
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const uint4 u = *reinterpret_cast<const uint4*>(rcur + sw64_off(row, jj));
            ptx::add_half2(v[8 * jj + 0], v[8 * jj + 1], u.x);
            ptx::add_half2(v[8 * jj + 2], v[8 * jj + 3], u.y);
            ptx::add_half2(v[8 * jj + 4], v[8 * jj + 5], u.z);
            ptx::add_half2(v[8 * jj + 6], v[8 * jj + 7], u.w);
          }
        }
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) v[jj] = fmaxf(v[jj], 0.f);
        uint8_t* ob = obuf + (ostores % osets) * kBfSet;
        if (osets == 1) {
          if (leader) ptx::tma_store_wait_read<0>();
          asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        }
        uint2 l8[4], h8[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          uint4 uh;
          uh.x = ptx::pack_half2(v[8 * jj + 0], v[8 * jj + 1]);
          uh.y = ptx::pack_half2(v[8 * jj + 2], v[8 * jj + 3]);
          uh.z = ptx::pack_half2(v[8 * jj + 4], v[8 * jj + 5]);
          uh.w = ptx::pack_half2(v[8 * jj + 6], v[8 * jj + 7]);
          *reinterpret_cast<uint4*>(ob + sw64_off(row, jj)) = uh;
          l8[jj].x = residue_e4m3x2(v[8 * jj + 0], v[8 * jj + 1], uh.x) | (residue_e4m3x2(v[8 * jj + 2], v[8 * jj + 3], uh.y) << 16);
          l8[jj].y = residue_e4m3x2(v[8 * jj + 4], v[8 * jj + 5], uh.z) | (residue_e4m3x2(v[8 * jj + 6], v[8 * jj + 7], uh.w) << 16);
          if (p.out_hi8) {
            h8[jj].x = half2_to_e4m3x2(uh.x) | (half2_to_e4m3x2(uh.y) << 16);
            h8[jj].y = half2_to_e4m3x2(uh.z) | (half2_to_e4m3x2(uh.w) << 16);
          }
        }
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          *reinterpret_cast<uint4*>(ob + kEpiLo8Off + sw32_off(row, jj)) =
              make_uint4(l8[2 * jj].x, l8[2 * jj].y, l8[2 * jj + 1].x, l8[2 * jj + 1].y);
          if (p.out_hi8)
            *reinterpret_cast<uint4*>(ob + kEpiHi8Off + sw32_off(row, jj)) =
                make_uint4(h8[2 * jj].x, h8[2 * jj].y, h8[2 * jj + 1].x, h8[2 * jj + 1].y);
        }
        ptx::fence_proxy_async();
        if (osets == 2 && leader) ptx::tma_store_wait_read<0>();
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        res_issue();
        if (leader) {
          const int m0 = m_tile_cta * kBlockM;
          ptx::tma_store_2d(&tm.o_hi, ob, ncol, m0);
          ptx::tma_store_2d(&tm.o_lo8, ob + kEpiPlaneBytes, ncol, m0);
          if (p.out_hi8) ptx::tma_store_2d(&tm.o_hi8, ob + kEpiHi8Off, ncol, m0);
          ptx::tma_store_commit();
        }
        ++ostores;
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kPair && !leader_cta)
          ptx::mbar_arrive_cluster_a(ptx::mapa(ptx::smem_u32(&a3empty[grp]), 0));
        else
          ptx::mbar_arrive(&a3empty[grp]);
      }
    }
    if (leader) ptx::tma_store_wait_all<0>();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (kPair) ptx::cluster_sync();
  if (warp_idx == 2) {
    ptx::tc_fence_after();
    if (kPair)
      ptx::tmem_dealloc_pair(tmem_base, kTmemCols);
    else
      ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
  (void)rows_per_tile;
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
struct BneckPlan {
  BneckMaps tm;
  BneckParams p;
  int pair = 0;
  int grid = 0;
  int smem = 0;
  double flops = 0.0;   // algorithmic FLOPs of both convolutions
};

// Whether the fused tail exists for a bottleneck with `planes` mid channels on a [NB, H, W] map (stride-1 conv2, identity
// residual).  planes = 256 (layer3) needs 128 KB for t2 alone: not with the TMA-prefetched identity (see make plan).
inline bool bneck_supported(int planes, int stride, bool has_downsample) {
  return stride == 1 && !has_downsample && (planes == 64 || planes == 128);
}

inline BneckPlan make_bneck_plan(const Planes& t1, int NB, int H, int W, int planes, const Planes& W2, const float* bias2,
                                 const Planes& W3, const float* bias3, const Planes& res, const Planes& out, int num_sms,
                                 int pair) {
  MCG_CHECK(planes == 64 || planes == 128, "fused bottleneck tail: planes must be 64 or 128");
  MCG_CHECK(t1.hi && t1.lo8 && t1.hi8 && W2.hi && W2.hi8 && W2.lo8 && W3.hi && W3.hi8 && W3.lo8 && res.hi && res.lo8 && out.hi &&
                out.lo8,
            "fused bottleneck tail runs in the fp16c8 mode (hi + e4m3 planes of every operand)");
  BneckPlan pl;
  BneckParams& p = pl.p;
  pl.pair = pair ? 1 : 0;
  const long long M = static_cast<long long>(NB) * H * W;
  p.M = static_cast<int>(M);
  p.N1 = planes;
  p.N3 = 4 * planes;
  const int tile_m = pair ? 2 * kBlockM : kBlockM;
  p.m_tiles = static_cast<int>((M + tile_m - 1) / tile_m);
  p.a.kind = 1;
  p.a.NB = NB;
  p.a.H = H;
  p.a.W = W;
  p.a.C = planes;
  p.a.R = p.a.S = 3;
  p.a.stride = 1;
  p.a.pad = 1;
  p.a.P = H;
  p.a.Q = W;
  p.bias2 = bias2;
  p.bias3 = bias3;
  p.out_hi8 = out.hi8 != nullptr ? 1 : 0;
  p.nbuf1 = 2;
  MCG_CHECK(p.nbuf1 * p.N1 + 2 * kBfN3Tile <= kTmemCols, "TMEM plan");
  const int w2_rows = pair ? planes / 2 : planes;
  const int w3_rows = pair ? kBfN3Tile / 2 : kBfN3Tile;
  const int c2_stage = kATileBytes + w2_rows * kBlockK * 2;
  const int c3_stage = w3_rows * kBlockK * 2;
  p.stage_bytes = ((c2_stage > c3_stage ? c2_stage : c3_stage) + 1023) & ~1023;
  static const int tune_rslots = tune_env("MCG_TUNE_BF_RES_SLOTS");
  // measured (same box, graph replay): 1 slot 9.49 ms per step, 2 slots 9.64 - the 24 KB are worth more as a ring stage
  p.res_slots = tune_rslots == 1 || tune_rslots == 2 ? tune_rslots : 1;
  const int fixed = 1024 + kSmemBarrierBytes + (planes / kBlockK) * kBfT2KbBytes + 2 * p.res_slots * kBfRSet;   // + residual slots
  // staging sets per E3 group: a second set only when it does not cost the ring its 4th stage (conv2 is L2-bound: the
  // bytes in flight matter more than the store overlap); env MCG_TUNE_BF_OUT_SETS forces it
  static const int tune_sets = tune_env("MCG_TUNE_BF_OUT_SETS");
  p.out_sets = 2;
  int ring = kMaxDynSmem - fixed - 2 * p.out_sets * kBfSet;
  if (ring / p.stage_bytes < 4 || tune_sets == 1) {
    p.out_sets = 1;
    ring = kMaxDynSmem - fixed - 2 * p.out_sets * kBfSet;
  }
  if (tune_sets == 2) {
    p.out_sets = 2;
    ring = kMaxDynSmem - fixed - 2 * p.out_sets * kBfSet;
  }
  p.num_stages = ring / p.stage_bytes;
  if (p.num_stages > kBfMaxStages) p.num_stages = kBfMaxStages;
  MCG_CHECK(p.num_stages >= 2, "fused bottleneck tail: not enough shared memory for 2 pipeline stages");
  pl.smem = fixed + 2 * p.out_sets * kBfSet + p.num_stages * p.stage_bytes;
  MCG_CHECK(pl.smem <= kMaxDynSmem, "fused bottleneck tail: shared memory plan exceeds the 227 KB limit");
  if (pair) {
    const long long pairs = num_sms / 2;
    pl.grid = 2 * static_cast<int>(p.m_tiles < pairs ? p.m_tiles : pairs);
  } else {
    pl.grid = static_cast<int>(p.m_tiles < num_sms ? p.m_tiles : num_sms);
  }
  pl.flops = 2.0 * static_cast<double>(M) * planes * (9.0 * planes) + 2.0 * static_cast<double>(M) * (4.0 * planes) * planes;
  const CUtensorMapSwizzle sw64 = CU_TENSOR_MAP_SWIZZLE_64B, sw32 = CU_TENSOR_MAP_SWIZZLE_32B;
  BneckMaps& tm = pl.tm;
  tm.a_hi = make_tmap_im2col(t1.hi, p.a);
  tm.a_lo8 = make_tmap_im2col_u8(t1.lo8, p.a);
  tm.a_hi8 = make_tmap_im2col_u8(t1.hi8, p.a);
  const int K2 = 9 * planes;
  tm.w2_hi = make_tmap_2d(W2.hi, planes, K2, K2, w2_rows);
  tm.w2_hi8 = make_tmap_2d_u8(W2.hi8, planes, K2, K2, w2_rows, kBlockK, sw64);
  tm.w2_lo8 = make_tmap_2d_u8(W2.lo8, planes, K2, K2, w2_rows, kBlockK, sw64);
  tm.w3_hi = make_tmap_2d(W3.hi, p.N3, planes, planes, w3_rows);
  tm.w3_hi8 = make_tmap_2d_u8(W3.hi8, p.N3, planes, planes, w3_rows, kBlockK, sw64);
  tm.w3_lo8 = make_tmap_2d_u8(W3.lo8, p.N3, planes, planes, w3_rows, kBlockK, sw64);
  tm.r_hi = make_tmap_2d(res.hi, M, p.N3, p.N3, kBlockM, kEpiChunk, sw64);
  tm.r_lo8 = make_tmap_2d_u8(res.lo8, M, p.N3, p.N3, kBlockM, kEpiChunk, sw32);
  tm.o_hi = make_tmap_2d(out.hi, M, p.N3, p.N3, kBlockM, kEpiChunk, sw64);
  tm.o_lo8 = make_tmap_2d_u8(out.lo8, M, p.N3, p.N3, kBlockM, kEpiChunk, sw32);
  tm.o_hi8 = p.out_hi8 ? make_tmap_2d_u8(out.hi8, M, p.N3, p.N3, kBlockM, kEpiChunk, sw32) : tm.o_lo8;
  return pl;
}

inline void launch_bneck(const BneckPlan& pl, cudaStream_t stream) {
  static bool attr_set[2] = {false, false};
  if (!attr_set[pl.pair]) {
    if (pl.pair)
      MCG_CUDA(cudaFuncSetAttribute(bneck_tail_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
    else
      MCG_CUDA(cudaFuncSetAttribute(bneck_tail_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
    attr_set[pl.pair] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(pl.grid);
  cfg.blockDim = dim3(kBfThreads);
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
  attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[na].val.programmaticStreamSerializationAllowed = 1;
  ++na;
  cfg.attrs = attr;
  cfg.numAttrs = na;
  if (pl.pair)
    MCG_CUDA(cudaLaunchKernelEx(&cfg, bneck_tail_kernel<true>, pl.tm, pl.p));
  else
    MCG_CUDA(cudaLaunchKernelEx(&cfg, bneck_tail_kernel<false>, pl.tm, pl.p));
}

}  // namespace mcg
