// Microbenchmark: issue rate of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on one SM sub-partition.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate ffma2_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k_ffma(float* out, int iters, long long* clk) {
  float a[16];
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
  const float m = 1.0001f, c = 0.5f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], m, c);
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
__global__ void k_ffma2(float* out, int iters, long long* clk) {
  uint64_t a[8];
  for (int i = 0; i < 8; ++i) {
    float2 v = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i);
    a[i] = *reinterpret_cast<uint64_t*>(&v);
  }
  float2 mv = make_float2(1.0001f, 1.0001f), cv = make_float2(0.5f, 0.5f);
  const uint64_t m = *reinterpret_cast<uint64_t*>(&mv), c = *reinterpret_cast<uint64_t*>(&cv);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(m), "l"(c));
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) {
    float2 v = *reinterpret_cast<float2*>(&a[i]);
    s += v.x + v.y;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
int main() {
  float* out;
  long long* clk;
  cudaMalloc(&out, 1 << 20);
  cudaMalloc(&clk, 8 * 148);
  const int iters = 4000;
  for (int threads : {128, 256, 512, 1024}) {
    long long h;
    k_ffma<<<148, threads>>>(out, iters, clk);
    cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    double fma_per_clk = (double)iters * 16 * threads / (double)h;
    k_ffma2<<<148, threads>>>(out, iters, clk);
    long long h2;
    cudaMemcpy(&h2, clk, 8, cudaMemcpyDeviceToHost);
    double fma2_per_clk = (double)iters * 16 * threads / (double)h2;  // same number of scalar FMAs (8 packed x 2)
    printf("threads/SM=%4d: FFMA %.1f fp32-FMA/clk/SM   FFMA2 %.1f fp32-FMA/clk/SM\n", threads, fma_per_clk, fma2_per_clk);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
