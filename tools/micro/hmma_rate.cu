// Micro-benchmark: issue rate of the legacy warp-level mma.sync.m16n8k16 (fp16 x fp16 -> fp32) on sm_100a.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/micro/hmma_rate.cu -o tools/micro/hmma_rate.bin
#include <cstdio>
#include <cstdint>
__global__ void k(int iters, long long* out, float* sink) {
  uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 0x3c003c00u, 0x38003800u}, b[2] = {0x3c003c00u, 0x34003400u};
  float c[8][4] = {};
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  long long t1 = clock64();
  float s = 0;
  for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}
int main() {
  long long* d; float* s;
  cudaMalloc(&d, 148 * 8); cudaMalloc(&s, 148 * 1024 * 4);
  for (int warps : {1, 4, 8, 16}) {
    const int iters = 2000;
    k<<<148, warps * 32>>>(iters, d, s);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
    const double per_sm = double(mx) / (double(iters) * 8 * warps);
    printf("%2d warps/SM: %.2f cycles per m16n8k16 per SM (%.0f dense fp16 FLOP/clk/SM)\n", warps, per_sm, 4096.0 / per_sm);
  }
  return 0;
}
