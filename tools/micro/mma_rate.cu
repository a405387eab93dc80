// Micro-benchmark: rate of tcgen05.mma by kind / tile / smem swizzle, random operands in smem.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I mcgaze_b200/csrc tools/micro/mma_rate.cu -o tools/micro/mma_rate.bin
#include <cstdio>
#include <vector>
#include "ptx.cuh"
using namespace mcg;

// MODE 0: f16 SW128 (K=16)          1: e4m3 SW64 (K=32)         2: e4m3 SW128 (K=32)
//      3: 4 x f16 -> acc0, then 4 x e4m3 SW64 -> acc1 (fp16c8 "T" k-block as first built)
//      4: 4 x f16 -> acc0, then 4 x e4m3 SW128 -> acc1
//      5: 12 x f16 -> acc0 (fp16x3 k-block)
template <int MODE>
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = raw + ((1024u - (raw & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t rng = 0x9e3779b9u * (threadIdx.x + 1) + blockIdx.x;
  for (int i = threadIdx.x; i < 100 * 1024 / 4; i += blockDim.x) {
    rng = rng * 1664525u + 1013904223u;
    uint32_t v = rng ^ (rng >> 13);
    v &= (MODE == 1 || MODE == 2) ? 0xb7b7b7b7u : 0xb7ffb7ffu;  // finite, |v| < 2 in both views
    reinterpret_cast<uint32_t*>(smem_raw)[i] = v;
  }
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(&tmem_ptr, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  if (warp == 1 && lane == 0) {
    const uint32_t idesc = ptx::make_idesc_f16_f32(128, N);
    const uint32_t a = base, b = base + 32768;
    uint64_t da[4], db[4], ea[4], eb[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      da[j] = ptx::make_sw128_kmajor_desc(a + j * 32);
      db[j] = ptx::make_sw128_kmajor_desc(b + j * 32);
      ea[j] = MODE == 4 ? ptx::make_sw128_kmajor_desc(a + 16384 + j * 32) : ptx::make_sw64_kmajor_desc(a + 16384 + (j >> 1) * 8192 + (j & 1) * 32);
      eb[j] = MODE == 4 ? ptx::make_sw128_kmajor_desc(b + 32768 + j * 32) : ptx::make_sw64_kmajor_desc(b + 32768 + (j >> 1) * 16384 + (j & 1) * 32);
    }
    long long t0 = clock64();
    int mmas = 0;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
      if (MODE == 0 || MODE == 3 || MODE == 4 || MODE == 5) {
#pragma unroll
        for (int j = 0; j < 4; ++j) ptx::umma_f16(tmem, da[j], db[j], idesc, 1u);
      }
      if (MODE == 5) {
#pragma unroll
        for (int j = 0; j < 4; ++j) ptx::umma_f16(tmem, da[j], db[3 - j], idesc, 1u);
#pragma unroll
        for (int j = 0; j < 4; ++j) ptx::umma_f16(tmem, da[3 - j], db[j], idesc, 1u);
      }
      if (MODE == 1) {
#pragma unroll
        for (int j = 0; j < 4; ++j) ptx::umma_f8(tmem, ea[j], eb[j], idesc, 1u);
      }
      if (MODE == 2) {
#pragma unroll
        for (int j = 0; j < 4; ++j) ptx::umma_f8(tmem, da[j], db[j], idesc, 1u);
      }
      if (MODE == 3 || MODE == 4) {
#pragma unroll
        for (int j = 0; j < 4; ++j) ptx::umma_f8(tmem + 256, ea[j], eb[j], idesc, 1u);
      }
    }
    mmas = iters * (MODE == 5 ? 12 : (MODE == 3 || MODE == 4) ? 8 : 4);
    ptx::umma_commit(&bar);
    ptx::mbar_wait(&bar, 0);
    long long t1 = clock64();
    out[blockIdx.x] = (t1 - t0) * 1000 / mmas;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

template <int MODE>
void run(const char* name, long long* d) {
  cudaFuncSetAttribute(rate_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
  for (int N : {64, 128, 256}) {
    if (N == 256 && (MODE == 3 || MODE == 4)) continue;
    rate_kernel<MODE><<<148, 128, 128 * 1024>>>(N, 4000, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    std::vector<long long> h(148);
    cudaMemcpy(h.data(), d, 148 * 8, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (auto v : h) mx = v > mx ? v : mx;
    printf("N %3d %-52s %.1f cycles / MMA\n", N, name, double(mx) / 1000);
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * 8);
  run<0>("f16 sw128", d);
  run<1>("e4m3 sw64", d);
  run<2>("e4m3 sw128", d);
  run<5>("f16 x 12 (fp16x3 k-block)", d);
  run<3>("4 f16 + 4 e4m3 sw64, 2 accumulators", d);
  run<4>("4 f16 + 4 e4m3 sw128, 2 accumulators", d);
  return 0;
}
