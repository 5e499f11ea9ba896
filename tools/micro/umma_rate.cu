// Microbenchmark: tcgen05.mma (kind::f16, bf16, M = 128, K = 16) back-to-back issue rate per SM as a function of
//   * N (64 / 128 / 256),
//   * where A comes from: shared memory (SS) or tensor memory (TS),
//   * the shared-memory operand layout: K-major no-swizzle ("interleaved" 8x16B core matrices) or K-major SWIZZLE_128B.
// One CTA per SM, one thread issues `iters` MMAs accumulating into the same TMEM tile and commits once.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// layout: 0 = no swizzle (LBO = 128 B between K-adjacent cores, SBO = 256 B between 8-row groups for a K = 16 operand)
//         1 = SWIZZLE_128B (8 rows x 128 B atoms, SBO = 1024 B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, int layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  if (layout == 0) {
    d |= (uint64_t)(128u >> 4) << 16;
    d |= (uint64_t)(256u >> 4) << 32;
  } else {
    d |= (uint64_t)(1u) << 16;
    d |= (uint64_t)(1024u >> 4) << 32;
    d |= (uint64_t)2 << 61;
  }
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}

template <int TS>
__global__ void __launch_bounds__(128) k(int iters, int n, int layout, int spread, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tbase;
  __shared__ __align__(8) unsigned long long bar;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (warp == 0) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tbase)));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  long long t0 = 0, t1 = 0;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(n);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 64 * 1024;
    const uint32_t d_tmem = tbase;
    // descriptors of 8 K-steps precomputed: the loop body is 8 MMAs and a branch (what an unrolled K loop issues)
    uint64_t ad[8], bd[8];
    uint32_t at[8];
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      const int k2 = ks % spread;
      const uint32_t step = layout == 0 ? (uint32_t)k2 * 4096u : (uint32_t)(k2 & 3) * 32u + (uint32_t)(k2 >> 2) * 16384u;
      ad[ks] = make_desc(a0 + step, layout);
      bd[ks] = make_desc(b0 + (layout == 0 ? (uint32_t)k2 * 8192u : step), layout);
      at[ks] = d_tmem + 256u + (uint32_t)k2 * 8u;
    }
    t0 = clock64();
    for (int it = 0; it < iters; it += 8) {
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        if (TS) {
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
              "r"(at[ks]), "l"(bd[ks]), "r"(idesc), "r"(1));
        } else {
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
              "l"(ad[ks]), "l"(bd[ks]), "r"(idesc), "r"(1));
        }
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile(
        "{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(
            smem_u32(&bar))
        : "memory");
    t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase));
}

int main() {
  long long* out;
  cudaMalloc(&out, 148 * sizeof(long long));
  const int iters = 2048;
  cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("# tcgen05.mma kind::f16 M=128 K=16: cycles per MMA (max over 148 CTAs), ideal = N/2\n");
  for (int ts = 0; ts < 2; ++ts)
    for (int layout = 0; layout < 2; ++layout)
      for (int n : {64, 128, 256})
        for (int spread : {1, 8}) {
          for (int rep = 0; rep < 2; ++rep) {
            if (ts) k<1><<<148, 128, 200 * 1024>>>(iters, n, layout, spread, out);
            else k<0><<<148, 128, 200 * 1024>>>(iters, n, layout, spread, out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          }
          long long h[148];
          cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
          long long mx = 0;
          for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
          printf("A from %s, B %s, N=%3d, %d distinct K-steps: %7.1f cycles/MMA  (ideal %d)  operand bytes/clk %.0f\n",
                 ts ? "TMEM" : "smem", layout ? "SWIZZLE_128B" : "no-swizzle  ", n, spread, (double)mx / iters, n / 2,
                 ((ts ? 0 : 4096) + n * 32.0) / ((double)mx / iters));
        }
  return 0;
}
