// Microbenchmark: tcgen05.ld throughput per SM (bytes/clk) vs warps per CTA and load width.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_read_bw tmem_read_bw.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int W>
__device__ __forceinline__ void ld(uint32_t addr, uint32_t* v);
template <>
__device__ __forceinline__ void ld<32>(uint32_t addr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(addr));
}
template <>
__device__ __forceinline__ void ld<8>(uint32_t addr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(addr));
}

// mode 0: one wait per load (dependent); mode 1: 4 loads in flight before a wait
template <int W, int INFLIGHT>
__global__ void __launch_bounds__(512) k(int iters, long long* out, uint32_t* sink) {
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tbase)));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = tbase + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  uint32_t v[INFLIGHT][W];
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int f = 0; f < INFLIGHT; ++f) ld<W>(base + (uint32_t)(((it * INFLIGHT + f) * W) & 511 & ~(W - 1)), v[f]);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int f = 0; f < INFLIGHT; ++f)
#pragma unroll
      for (int j = 0; j < W; ++j) acc ^= v[f][j];
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase));
}

template <int W, int INFLIGHT>
void run(int warps, const char* name) {
  long long* out;
  uint32_t* sink;
  cudaMalloc(&out, 148 * 8);
  cudaMalloc(&sink, 4);
  const int iters = 2000;
  k<W, INFLIGHT><<<148, warps * 32>>>(iters, out, sink);
  k<W, INFLIGHT><<<148, warps * 32>>>(iters, out, sink);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  const double bytes = (double)iters * INFLIGHT * W * 32 * 4 * warps;
  printf("%-28s warps=%2d  %8lld cyc  %7.1f B/clk/SM  (%s)\n", name, warps, h[0], bytes / (double)h[0], cudaGetErrorString(e));
  cudaFree(out);
  cudaFree(sink);
}

int main() {
  for (int warps : {4, 8, 16}) {
    run<32, 1>(warps, "32x32b.x32, 1 in flight");
    run<32, 2>(warps, "32x32b.x32, 2 in flight");
    run<8, 1>(warps, "32x32b.x8, 1 in flight");
    run<8, 4>(warps, "32x32b.x8, 4 in flight");
  }
  return 0;
}
