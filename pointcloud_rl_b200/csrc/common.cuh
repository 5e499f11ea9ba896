// Shared device/host helpers for libpcrl (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pcrl.h"

namespace pcrl {

void set_error(const char* fmt, ...);

#define PCRL_CHECK_ARG(cond)                                                        \
  do {                                                                              \
    if (!(cond)) {                                                                  \
      pcrl::set_error("%s:%d: invalid argument: %s", __FILE__, __LINE__, #cond);    \
      return PCRL_EINVAL;                                                           \
    }                                                                               \
  } while (0)

#define PCRL_CHECK_CUDA(expr)                                                                     \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      pcrl::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));      \
      return PCRL_ECUDA;                                                                          \
    }                                                                                             \
  } while (0)

#define PCRL_CHECK_LAUNCH() PCRL_CHECK_CUDA(cudaGetLastError())

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t align_up(int64_t a, int64_t b) { return cdiv(a, b) * b; }

int sm_count();

// Per-device context (pcrl_create / pcrl_destroy): everything the library keeps between calls lives here, one per
// CUDA device, created on first use or explicitly through the handle API: the SM count, and the internal side stream
// + fork/join events pcrl_pointnet_bwd runs its weight-gradient GEMMs on.  Nothing is process-global any more.
struct DeviceCtx {
  int device = -1;
  int sms = 0;
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};
DeviceCtx* device_ctx();  // context of the CURRENT device (lazily created); nullptr + set_error on failure

// TF32 requests that ended on the FFMA kernel because an operand was not TMA-addressable (16-byte base, pitch % 4).
// pcrl_tf32_fallbacks() reads the count; with pcrl_set_strict_tf32(1) such a call fails instead of falling back.
void note_tf32_fallback();
bool strict_tf32();

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), counter-based: no state, graph-replay safe because the
// per-update counter lives in device memory.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
// uniform in [0,1) with 24 random bits
__device__ __forceinline__ float u01(uint32_t x) { return (x >> 8) * (1.0f / 16777216.0f); }
// standard normal pair (Box-Muller)
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
  float u1 = ((a >> 8) + 1) * (1.0f / 16777216.0f);  // (0,1]
  float u2 = u01(b);
  float r = sqrtf(-2.0f * logf(u1));
  float s, c;
  sincospif(2.0f * u2, &s, &c);
  return make_float2(r * c, r * s);
}

// fp32 -> nearest TF32 (10-bit mantissa, ties away), kept in an fp32 container: the TF32 tensor-core path ignores the low
// 13 mantissa bits of its operands, i.e. TRUNCATES; producers that feed it round instead (half the error, no bias)
__device__ __forceinline__ float round_tf32(float v) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  return __uint_as_float(u);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------------------------------------
// Generic CUDA-core SGEMM used by the fp32 parity path, the compacted backward and the MLP heads:
//   C[i][j] (+)= sum_l A(i,l) * B(l,j),  A(i,l) = A[i*a_si + l*a_sl],  B(l,j) = B[l*b_sl + j*b_sj]
// m_dev (optional) bounds the rows actually present (compacted sets sized on the device).
// ---------------------------------------------------------------------------------------------
struct GemmArgs {
  const float* A;
  int64_t a_si, a_sl;
  const float* B;
  int64_t b_sl, b_sj;
  float* C;
  int64_t ldc;
  int M, N, K;
  const float* bias;  // per column j, may be null
  int relu;           // bit 0: ReLU; bit 1 (FFMA kernel only): round the stored value to TF32 (it feeds a TF32 GEMM)
  int accumulate;     // atomicAdd into C (required when split_k > 1)
  int split_k;
  const int* m_dev;   // if set, rows i >= *m_dev are skipped (M rows)
  const int* k_dev;   // if set, contraction index l >= *k_dev is skipped (K)
  const float* mask;  // if set: C[i][j] = 0 where mask[i*ldmask + j] <= 0 (ReLU backward fused into the dgrad GEMM)
  int64_t ldmask;
  int bn_hint;        // tensor-core path only: force the output tile width (0 = auto)
};
int launch_sgemm(const GemmArgs& g, cudaStream_t st);

}  // namespace pcrl
