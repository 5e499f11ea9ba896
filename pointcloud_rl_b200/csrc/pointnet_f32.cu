// Exact-fp32 PointNet per-point MLP on CUDA cores: the parity path of the forward, and the compacted
// sparse backward through the saved argmax.  Layer by layer over generic kernels (sgemm.cu,
// rowwise.cu); the fused tensor-core forward is pointnet_tc.cu.
#include "common.cuh"
#include "rowwise.cuh"

namespace pcrl {
namespace tc {
bool recompute_supported(int c1, int c2, int c3);
int recompute_active_tc(const void* xh, const void* wpack, const int32_t* src, const int* count_dev, int capacity,
                        int c1, int c2, int c3, float ln_eps, void* xha_scratch, float* h0, float* xhat1, float* rstd1,
                        float* h1, float* xhat2, float* rstd2, const float* m1, const float* m2, cudaStream_t st);
}

// max over the N real points of each cloud, ties -> smallest index (torch.max semantics,
// pointnet.py:151).  h [rows, NP, c3] post-ReLU.  One thread per (cloud, channel): reads are
// coalesced across channels.
__global__ void __launch_bounds__(256) maxpool_points_kernel(const float* __restrict__ h, int rows, int N, int NP,
                                                             int c3, float* __restrict__ pooled,
                                                             int32_t* __restrict__ argmax) {
  const int r = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= c3 || r >= rows) return;
  const float* p = h + (int64_t)r * NP * c3 + c;
  float best = p[0];
  int bi = 0;
  for (int n = 1; n < N; ++n) {
    float v = p[(int64_t)n * c3];
    if (v > best) {
      best = v;
      bi = n;
    }
  }
  pooled[(int64_t)r * c3 + c] = best;
  if (argmax) argmax[(int64_t)r * c3 + c] = bi;
}

// LayerNorm + ReLU + max over the points, fused (the last layer of the unfused fp32 / TF32 chains): y [rows, NP, c3] is the
// raw layer-2 GEMM output; nothing is written back except the packed per-(cloud, channel) maximum, so the 2 x c3 floats
// per point of a separate LN pass + max-pool pass never move.  Same arithmetic, in the same order, as ln_rows_kernel
// followed by maxpool_points_kernel (a warp per point, lane l owns channels l, l+32, ...).
// grid (slices, rows): block (r, s) takes points n = s*warps + w, step slices*warps; partial maxima meet in a 64-bit
// atomicMax on (float bits << 32 | ~index): values are >= 0 after the ReLU, so bit order is value order and ties
// resolve to the smallest index (torch.max semantics).  keys must be zero on entry; finalize_keys_kernel unpacks.
template <int VPL>
__global__ void __launch_bounds__(256)
ln_relu_maxpool_kernel(const float* __restrict__ y, int N, int NP, int c3, const float* __restrict__ g,
                       const float* __restrict__ b, float eps, unsigned long long* __restrict__ keys) {
  const int r = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int warps = blockDim.x >> 5;
  const float inv_d = 1.0f / (float)c3;
  float gv[VPL], bv[VPL], best[VPL];
  int bi[VPL];
#pragma unroll
  for (int q = 0; q < VPL; ++q) {
    gv[q] = g[lane + 32 * q];
    bv[q] = b[lane + 32 * q];
    best[q] = -1.f;
    bi[q] = 0;
  }
  for (int n = blockIdx.x * warps + warp; n < N; n += gridDim.x * warps) {
    const float* xr = y + ((int64_t)r * NP + n) * c3;
    float x[VPL];
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < VPL; ++q) {
      x[q] = xr[lane + 32 * q];
      s += x[q];
    }
    const float mean = warp_sum(s) * inv_d;
    float v = 0.f;
#pragma unroll
    for (int q = 0; q < VPL; ++q) {
      const float d = x[q] - mean;
      v = fmaf(d, d, v);
    }
    const float rstd = 1.0f / sqrtf(warp_sum(v) * inv_d + eps);
#pragma unroll
    for (int q = 0; q < VPL; ++q) {
      const float xh = (x[q] - mean) * rstd;
      const float o = fmaxf(fmaf(xh, gv[q], bv[q]), 0.f);
      if (o > best[q]) {  // points ascend within a warp: strict > keeps the smallest index
        best[q] = o;
        bi[q] = n;
      }
    }
  }
#pragma unroll
  for (int q = 0; q < VPL; ++q)
    if (best[q] >= 0.f)
      atomicMax(keys + (int64_t)r * c3 + lane + 32 * q,
                ((unsigned long long)__float_as_uint(best[q]) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)bi[q]));
}

__global__ void finalize_keys_kernel(unsigned long long* __restrict__ keys, int64_t n, float* __restrict__ pooled,
                                     int32_t* __restrict__ argmax) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long k = keys[i];
  pooled[i] = __uint_as_float((uint32_t)(k >> 32));
  if (argmax) argmax[i] = (int32_t)(0xFFFFFFFFu - (uint32_t)k);
}

// ---- sparse backward helpers -----------------------------------------------------------------

// A point carries gradient when it won the max for some channel (r, c) with pooled > 0 (ReLU passes) and
// dpooled != 0.  slot[r*NP+n] = compacted index (or -1); src[a] = r*NP+n; clouds in order, points ascending
// (deterministic layout); sizes stay on the device.
// Fused compaction, two launches instead of memset + five kernels.
// (1) one block per cloud: clear the cloud's flags, mark the points that carry gradient, count them.
__global__ void __launch_bounds__(256) mark_count_kernel(const float* __restrict__ pooled, const int32_t* __restrict__ argmax,
                                                         const float* __restrict__ dpooled, int NP, int c3,
                                                         int32_t* __restrict__ flag, int32_t* __restrict__ counts) {
  const int r = blockIdx.x;
  int32_t* f = flag + (int64_t)r * NP;
  for (int n = threadIdx.x; n < NP; n += blockDim.x) f[n] = 0;
  __syncthreads();
  for (int c = threadIdx.x; c < c3; c += blockDim.x) {
    const int64_t e = (int64_t)r * c3 + c;
    if (pooled[e] > 0.f && dpooled[e] != 0.f) f[argmax[e]] = 1;
  }
  __syncthreads();
  int s = 0;
  for (int n = threadIdx.x; n < NP; n += blockDim.x) s += f[n];
  __shared__ int red[256];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[r] = red[0];
}
// (2) one block per cloud: its base offset is the sum of the previous clouds' counts (R is at most a few thousand),
// then slots in ascending point order, and the rows are gathered right away: xa (fp32 staged points) and, in fast
// mode, the 32-byte bf16 tile rows of the fused kernel's operand image.
__global__ void __launch_bounds__(256) assign_gather_kernel(const int32_t* __restrict__ flag, const int32_t* __restrict__ counts,
                                                            int R, int NP, int capacity, int CP, const float* __restrict__ xf,
                                                            const char* __restrict__ xh, int32_t* __restrict__ slot,
                                                            int32_t* __restrict__ src, int32_t* __restrict__ total,
                                                            float* __restrict__ xa, char* __restrict__ xha) {
  const int r = blockIdx.x;
  __shared__ int red[256];
  __shared__ int warp_tot[8];
  __shared__ int base;
  int s = 0;
  for (int i = threadIdx.x; i < r; i += blockDim.x) s += counts[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    base = red[0];
    if (r == R - 1) *total = min(red[0] + counts[r], capacity);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int n0 = 0; n0 < NP; n0 += 256) {
    const int n = n0 + threadIdx.x;
    const int fl = (n < NP) ? flag[(int64_t)r * NP + n] : 0;
    const unsigned m = __ballot_sync(0xffffffffu, fl != 0);
    if (lane == 0) warp_tot[warp] = __popc(m);
    __syncthreads();
    int pre = 0;
    for (int w = 0; w < warp; ++w) pre += warp_tot[w];
    int tot = 0;
    for (int w = 0; w < 8; ++w) tot += warp_tot[w];
    if (n < NP) {
      int a = -1;
      if (fl) {
        a = base + pre + __popc(m & ((1u << lane) - 1));
        if (a < capacity) {
          const int pidx = r * NP + n;
          src[a] = pidx;
          const float4* sx = reinterpret_cast<const float4*>(xf + (int64_t)pidx * CP);
          float4* dx = reinterpret_cast<float4*>(xa + (int64_t)a * CP);
          for (int q = 0; q < CP / 4; ++q) dx[q] = sx[q];
          if (xh) {  // two 16-byte core-matrix rows of the 128x16 bf16 tile image
            const char* sp_ = xh + (int64_t)(pidx >> 7) * 4096 + ((pidx & 127) >> 3) * 256 + (pidx & 7) * 16;
            char* dp = xha + (int64_t)(a >> 7) * 4096 + ((a & 127) >> 3) * 256 + (a & 7) * 16;
            *reinterpret_cast<uint4*>(dp) = *reinterpret_cast<const uint4*>(sp_);
            *reinterpret_cast<uint4*>(dp + 128) = *reinterpret_cast<const uint4*>(sp_ + 128);
          }
        } else {
          a = -1;
        }
      }
      slot[(int64_t)r * NP + n] = a;
    }
    __syncthreads();
    if (threadIdx.x == 0) base += tot;
    __syncthreads();
  }
}

// dh2[slot(r, argmax[r,c]), c] = dpooled[r,c]  (dh2 zeroed beforehand)
__global__ void scatter_dpool_kernel(const float* __restrict__ pooled, const int32_t* __restrict__ argmax,
                                     const float* __restrict__ dpooled, const int32_t* __restrict__ slot, int R, int NP,
                                     int c3, float* __restrict__ dh2) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)R * c3) return;
  if (pooled[e] > 0.f && dpooled[e] != 0.f) {
    int r = (int)(e / c3), c = (int)(e % c3);
    int a = slot[(int64_t)r * NP + argmax[e]];
    if (a >= 0) dh2[(int64_t)a * c3 + c] = dpooled[e];
  }
}

// Layer-2 LayerNorm backward when dout is the max-pool scatter (fast mode).  dout[a][c] = dpooled[r][c] only where point
// a won channel c of cloud r, and there xhat2[a][c] = (pooled[r][c] - b2[c]) / g2[c] is known from the forward.  Pass 1
// accumulates the two row means LN backward needs and the (sparse) parameter gradients; the dense part of dy2 is then
// written by the recompute kernel itself and pass 2 adds the sparse term rstd2 * g2 * dout.
__global__ void ln2_sparse_stats_kernel(const float* __restrict__ pooled, const int32_t* __restrict__ argmax,
                                        const float* __restrict__ dpooled, const int32_t* __restrict__ slot,
                                        const float* __restrict__ g2, const float* __restrict__ be2, int R, int NP, int c3,
                                        float* __restrict__ m1, float* __restrict__ m2, float* __restrict__ dg2,
                                        float* __restrict__ dbe2) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)R * c3) return;
  const float pv = pooled[e], g = dpooled[e];
  if (!(pv > 0.f) || g == 0.f) return;
  const int r = (int)(e / c3), c = (int)(e % c3);
  const int a = slot[(int64_t)r * NP + argmax[e]];
  if (a < 0) return;
  const float gam = g2[c];
  const float xh = gam != 0.f ? (pv - be2[c]) / gam : 0.f;
  const float t = gam * g / (float)c3;
  atomicAdd(m1 + a, t);
  atomicAdd(m2 + a, t * xh);
  atomicAdd(dg2 + c, g * xh);
  atomicAdd(dbe2 + c, g);
}
__global__ void ln2_sparse_fix_kernel(const float* __restrict__ pooled, const int32_t* __restrict__ argmax,
                                      const float* __restrict__ dpooled, const int32_t* __restrict__ slot,
                                      const float* __restrict__ g2, const float* __restrict__ rstd2, int R, int NP, int c3,
                                      float* __restrict__ dy2) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)R * c3) return;
  const float g = dpooled[e];
  if (!(pooled[e] > 0.f) || g == 0.f) return;
  const int r = (int)(e / c3), c = (int)(e % c3);
  const int a = slot[(int64_t)r * NP + argmax[e]];
  if (a < 0) return;
  dy2[(int64_t)a * c3 + c] += rstd2[a] * g2[c] * g;  // (a, c) pairs are unique: no atomics
}

__global__ void relu_bwd_rows_kernel(float* __restrict__ dy, const float* __restrict__ y, int width,
                                     const int* __restrict__ count) {
  int64_t n = (int64_t)(*count) * width;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride)
    if (!(y[i] > 0.f)) dy[i] = 0.f;
}

// ---- layer 0 of the compacted backward: K = C <= 16 is far too thin for a GEMM tile -------------------------
// h0[a, n] = relu(b0[n] + sum_c xa[a, c] w0[n, c]).  CTA = 16 rows x c1 channels; W0 row of the thread in registers.
template <int CP>
__global__ void __launch_bounds__(256) layer0_fwd_kernel(const float* __restrict__ xa, const float* __restrict__ w0,
                                                         const float* __restrict__ b0, int C, int c1,
                                                         const int* __restrict__ count, float* __restrict__ h0) {
  const int rows = *count;
  const int a0 = blockIdx.x * 16;
  if (a0 >= rows) return;
  const int n = threadIdx.x % 128, phase = threadIdx.x / 128;
  for (int nn = n; nn < c1; nn += 128) {
    float wr[CP];
#pragma unroll
    for (int c = 0; c < CP; ++c) wr[c] = c < C ? __ldg(w0 + nn * C + c) : 0.f;
    const float bias = __ldg(b0 + nn);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int a = a0 + phase + 2 * i;
      if (a >= rows) break;
      const float4* xr = reinterpret_cast<const float4*>(xa + (int64_t)a * CP);
      float acc = bias;
#pragma unroll
      for (int q = 0; q < CP / 4; ++q) {
        const float4 x = __ldg(xr + q);
        acc = fmaf(x.x, wr[4 * q], acc);
        acc = fmaf(x.y, wr[4 * q + 1], acc);
        acc = fmaf(x.z, wr[4 * q + 2], acc);
        acc = fmaf(x.w, wr[4 * q + 3], acc);
      }
      h0[(int64_t)a * c1 + nn] = fmaxf(acc, 0.f);
    }
  }
}
// dw0[n, c] += sum_a d0[a, n] xa[a, c];  db0[n] += sum_a d0[a, n]   (d0 already ReLU-masked).
// CTA = 256 rows: thread (n, phase) accumulates every other row of its chunk, one atomic per output per CTA.
template <int CP>
__global__ void __launch_bounds__(256) layer0_wgrad_kernel(const float* __restrict__ d0, const float* __restrict__ xa,
                                                           int C, int c1, const int* __restrict__ count,
                                                           float* __restrict__ dw0, float* __restrict__ db0) {
  const int rows = *count;
  const int a0 = blockIdx.x * 256;
  if (a0 >= rows) return;
  const int a1 = min(rows, a0 + 256);
  const int n = threadIdx.x % 128, phase = threadIdx.x / 128;
  for (int nn = n; nn < c1; nn += 128) {
    float acc[CP], accb = 0.f;
#pragma unroll
    for (int c = 0; c < CP; ++c) acc[c] = 0.f;
#pragma unroll 4
    for (int a = a0 + phase; a < a1; a += 2) {
      const float g = __ldg(d0 + (int64_t)a * c1 + nn);
      const float4* xr = reinterpret_cast<const float4*>(xa + (int64_t)a * CP);
#pragma unroll
      for (int q = 0; q < CP / 4; ++q) {
        const float4 x = __ldg(xr + q);
        acc[4 * q] = fmaf(g, x.x, acc[4 * q]);
        acc[4 * q + 1] = fmaf(g, x.y, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(g, x.z, acc[4 * q + 2]);
        acc[4 * q + 3] = fmaf(g, x.w, acc[4 * q + 3]);
      }
      accb += g;
    }
#pragma unroll
    for (int c = 0; c < CP; ++c)
      if (c < C && acc[c] != 0.f) atomicAdd(dw0 + nn * C + c, acc[c]);
    if (accb != 0.f) atomicAdd(db0 + nn, accb);
  }
}

static int gemm_nt(const float* A, int lda, const float* W, int ldw, const float* bias, int relu, float* C, int ldc,
                   int M, int K, int Nout, const int* m_dev, cudaStream_t st, int tf32 = 0) {
  // C[M,Nout] = act(A[M,K] W[Nout,K]^T + bias)
  GemmArgs g{};
  g.A = A; g.a_si = lda; g.a_sl = 1;
  g.B = W; g.b_sl = 1; g.b_sj = ldw;
  g.C = C; g.ldc = ldc;
  g.M = M; g.N = Nout; g.K = K;
  g.bias = bias; g.relu = relu; g.split_k = 1; g.m_dev = m_dev;
  return launch_gemm(g, tf32, st);
}
static int gemm_nn(const float* A, int lda, const float* W, int ldw, float* C, int ldc, int M, int K, int Nout,
                   const int* m_dev, cudaStream_t st, int tf32 = 0, const float* relu_mask = nullptr) {
  // C[M,Nout] = (A[M,K] W[K,Nout]) * (relu_mask > 0)
  GemmArgs g{};
  g.A = A; g.a_si = lda; g.a_sl = 1;
  g.B = W; g.b_sl = ldw; g.b_sj = 1;
  g.C = C; g.ldc = ldc;
  g.M = M; g.N = Nout; g.K = K;
  g.split_k = 1; g.m_dev = m_dev;
  g.mask = relu_mask; g.ldmask = ldc;
  return launch_gemm(g, tf32, st);
}
static int gemm_tn_acc(const float* A, int lda, const float* Bm, int ldb, float* C, int ldc, int Mrows, int Na, int Nb,
                       const int* rows_dev, int expected_rows, cudaStream_t st, int tf32 = 0) {
  // C[Na,Nb] += A[Mrows,Na]^T Bm[Mrows,Nb]   (contraction over rows, split across CTAs)
  GemmArgs g{};
  g.A = A; g.a_si = 1; g.a_sl = lda;
  g.B = Bm; g.b_sl = ldb; g.b_sj = 1;
  g.C = C; g.ldc = ldc;
  g.M = Na; g.N = Nb; g.K = Mrows;
  g.accumulate = 1; g.k_dev = rows_dev;
  const int tile = tf32 ? 128 : 64;
  int64_t tiles = cdiv(Na, tile) * cdiv(Nb, tile);
  g.split_k = (int)std::max<int64_t>(1, std::min<int64_t>(cdiv(4 * sm_count(), tiles), cdiv(expected_rows, 128)));
  if (tf32) {
    // full-width tiles (every operand panel streamed once per M tile) and just enough K-splits to fill the SMs
    g.bn_hint = Nb >= 128 ? 128 : (Nb > 32 ? 64 : 32);
    const int64_t t128 = cdiv(Na, 128) * cdiv(Nb, g.bn_hint);
    g.split_k = (int)std::max<int64_t>(1, std::min<int64_t>(cdiv(sm_count(), t128), cdiv(expected_rows, 256)));
  }
  return launch_gemm(g, tf32, st);
}

}  // namespace pcrl

using namespace pcrl;

extern "C" {

int64_t pcrl_pointnet_fwd_f32_workspace(int clouds, int NP, int c1, int c2, int c3) {
  return (int64_t)clouds * NP * (c1 + c2 + c3) * sizeof(float);
}
int64_t pcrl_pointnet_fwd_tf32_workspace(int clouds, int NP, int c1, int c2, int c3) {
  return pcrl_pointnet_fwd_f32_workspace(clouds, NP, c1, c2, c3) + align_up((int64_t)(c2 * c1 + c3 * c2) * 4, 256);
}

static int pointnet_fwd_chain(const float* xf, int R, int N, int NP, int CP, int C, const float* w0, const float* b0,
                              const float* w1, const float* g1, const float* be1, const float* w2, const float* g2,
                              const float* be2, int c1, int c2, int c3, float ln_eps, float* pooled, int32_t* argmax,
                              void* workspace, int64_t workspace_bytes, int tf32, void* stream);

int pcrl_pointnet_fwd_f32(const float* xf, int R, int N, int NP, int CP, int C, const float* w0, const float* b0,
                          const float* w1, const float* g1, const float* be1, const float* w2, const float* g2,
                          const float* be2, int c1, int c2, int c3, float ln_eps, float* pooled, int32_t* argmax,
                          void* workspace, int64_t workspace_bytes, void* stream) {
  return pointnet_fwd_chain(xf, R, N, NP, CP, C, w0, b0, w1, g1, be1, w2, g2, be2, c1, c2, c3, ln_eps, pooled, argmax,
                            workspace, workspace_bytes, 0, stream);
}

int pcrl_pointnet_fwd_tf32(const float* xf, int R, int N, int NP, int CP, int C, const float* w0, const float* b0,
                           const float* w1, const float* g1, const float* be1, const float* w2, const float* g2,
                           const float* be2, int c1, int c2, int c3, float ln_eps, float* pooled, int32_t* argmax,
                           void* workspace, int64_t workspace_bytes, void* stream) {
  if (c1 % 4 || c2 % 4 || c1 < 8 || c2 < 8) {
    set_error("pcrl_pointnet_fwd_tf32: widths (%d,%d,%d) need c1, c2 multiples of 4 (TMA row pitch) and >= 8", c1, c2, c3);
    return PCRL_EUNSUPPORTED;
  }
  return pointnet_fwd_chain(xf, R, N, NP, CP, C, w0, b0, w1, g1, be1, w2, g2, be2, c1, c2, c3, ln_eps, pooled, argmax,
                            workspace, workspace_bytes, 1, stream);
}

__global__ void round_tf32_kernel(const float* __restrict__ src, float* __restrict__ dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = round_tf32(src[i]);
}

// tf32 = 1: layers 1 and 2 (99.3 % of the FLOPs) run on the TF32 tcgen05 GEMM straight from the fp32 activations; layer 0
// (K = C <= 12), LayerNorm statistics, the max / argmax stay exact fp32, and the 64-bit (value, ~index) keys are untruncated.
static int pointnet_fwd_chain(const float* xf, int R, int N, int NP, int CP, int C, const float* w0, const float* b0,
                              const float* w1, const float* g1, const float* be1, const float* w2, const float* g2,
                              const float* be2, int c1, int c2, int c3, float ln_eps, float* pooled, int32_t* argmax,
                              void* workspace, int64_t workspace_bytes, int tf32, void* stream) {
  PCRL_CHECK_ARG(xf && w0 && b0 && w1 && g1 && be1 && w2 && g2 && be2 && pooled && workspace);
  PCRL_CHECK_ARG(R >= 0 && N > 0 && NP >= N && NP % 128 == 0 && C <= CP);
  cudaStream_t st = as_stream(stream);
  const int64_t per_cloud = (int64_t)NP * (c1 + c2 + c3) * sizeof(float);
  // TF32 tier: the tensor core truncates its operands to 10 mantissa bits; feeding it values already ROUNDED to TF32
  // halves the error and removes its bias.  Activations are rounded by their producers (relu flag bit 1), the two
  // weight matrices into copies at the end of the workspace.
  const int64_t wbytes = tf32 ? align_up((int64_t)(c2 * c1 + c3 * c2) * 4, 256) : 0;
  PCRL_CHECK_ARG(workspace_bytes > wbytes);
  const int chunk = (int)std::min<int64_t>(R, (workspace_bytes - wbytes) / per_cloud);
  PCRL_CHECK_ARG(chunk >= 1 || R == 0);
  const int rnd = tf32 ? 2 : 0;
  if (tf32 && R > 0) {
    float* w1r = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + workspace_bytes - wbytes);
    float* w2r = w1r + c2 * c1;
    round_tf32_kernel<<<(unsigned)cdiv(c2 * c1, 256), 256, 0, st>>>(w1, w1r, c2 * c1);
    round_tf32_kernel<<<(unsigned)cdiv(c3 * c2, 256), 256, 0, st>>>(w2, w2r, c3 * c2);
    PCRL_CHECK_LAUNCH();
    w1 = w1r;
    w2 = w2r;
  }
  for (int r0 = 0; r0 < R; r0 += chunk) {
    const int rows = std::min(chunk, R - r0);
    const int P = rows * NP;
    float* h0 = reinterpret_cast<float*>(workspace);
    float* h1 = h0 + (int64_t)P * c1;
    float* h2 = h1 + (int64_t)P * c2;
    const float* x = xf + (int64_t)r0 * NP * CP;
    int rc;
    // h0 = relu(x W0^T + b0)                     (conv0 + ReLU; no norm: ignore_first_ln)
    if ((rc = gemm_nt(x, CP, w0, C, b0, 1 | rnd, h0, c1, P, C, c1, nullptr, st))) return rc;
    // h1 = relu(LN(h0 W1^T))
    if ((rc = gemm_nt(h0, c1, w1, c1, nullptr, 0, h1, c2, P, c1, c2, nullptr, st, tf32))) return rc;
    if ((rc = launch_ln_rows(h1, c2, g1, be1, h1, c2, nullptr, nullptr, P, c2, ln_eps, 1 | rnd, nullptr, st))) return rc;
    // h2 = relu(LN(h1 W2^T))
    if ((rc = gemm_nt(h1, c2, w2, c2, nullptr, 0, h2, c3, P, c2, c3, nullptr, st, tf32))) return rc;
    if (c3 % 32 == 0 && c3 <= 1024 && (int64_t)rows * c3 * 8 <= (int64_t)P * c1 * 4) {
      // fused LayerNorm + ReLU + max-pool; the packed maxima live in the h0 region (dead since layer 1's GEMM)
      auto* keys = reinterpret_cast<unsigned long long*>(h0);
      PCRL_CHECK_CUDA(cudaMemsetAsync(keys, 0, (size_t)rows * c3 * 8, st));
      const int slices = (int)std::max<int64_t>(1, std::min<int64_t>(cdiv(N, 8), cdiv(2 * sm_count(), rows)));
      dim3 grid((unsigned)slices, (unsigned)rows);
      switch (c3 / 32) {
        case 4: ln_relu_maxpool_kernel<4><<<grid, 256, 0, st>>>(h2, N, NP, c3, g2, be2, ln_eps, keys); break;
        case 8: ln_relu_maxpool_kernel<8><<<grid, 256, 0, st>>>(h2, N, NP, c3, g2, be2, ln_eps, keys); break;
        case 16: ln_relu_maxpool_kernel<16><<<grid, 256, 0, st>>>(h2, N, NP, c3, g2, be2, ln_eps, keys); break;
        case 32: ln_relu_maxpool_kernel<32><<<grid, 256, 0, st>>>(h2, N, NP, c3, g2, be2, ln_eps, keys); break;
        default: keys = nullptr;
      }
      if (keys) {
        PCRL_CHECK_LAUNCH();
        finalize_keys_kernel<<<(unsigned)cdiv((int64_t)rows * c3, 256), 256, 0, st>>>(
            keys, (int64_t)rows * c3, pooled + (int64_t)r0 * c3, argmax ? argmax + (int64_t)r0 * c3 : nullptr);
        PCRL_CHECK_LAUNCH();
        continue;
      }
    }
    if ((rc = launch_ln_rows(h2, c3, g2, be2, h2, c3, nullptr, nullptr, P, c3, ln_eps, 1, nullptr, st))) return rc;
    dim3 grid((unsigned)cdiv(c3, 256), (unsigned)rows);
    maxpool_points_kernel<<<grid, 256, 0, st>>>(h2, rows, N, NP, c3, pooled + (int64_t)r0 * c3,
                                                argmax ? argmax + (int64_t)r0 * c3 : nullptr);
    PCRL_CHECK_LAUNCH();
  }
  return PCRL_OK;
}

// workspace layout of the backward (A = capacity = R*c3 active points at most)
struct BwdWs {
  int32_t *flag, *slot, *src, *counts, *offsets, *total;
  float *xa, *h0, *y1hat, *rstd1, *h1, *y2hat, *rstd2, *d2, *d1, *d0, *m12;
  int64_t bytes;
};
static BwdWs carve_bwd(void* base, int R, int NP, int c1, int c2, int c3, int CP) {
  BwdWs w{};
  char* p = reinterpret_cast<char*>(base);
  const int64_t A = (int64_t)R * c3;
  auto take = [&](int64_t bytes) {
    char* q = p;
    p += align_up(bytes, 256);
    return q;
  };
  w.flag = (int32_t*)take((int64_t)R * NP * 4);
  w.slot = (int32_t*)take((int64_t)R * NP * 4);
  w.src = (int32_t*)take(A * 4);
  w.counts = (int32_t*)take((int64_t)R * 4);
  w.offsets = (int32_t*)take((int64_t)R * 4);
  w.total = (int32_t*)take(256);
  w.xa = (float*)take(A * CP * 4);
  w.h0 = (float*)take(A * c1 * 4);
  w.y1hat = (float*)take(A * c2 * 4);
  w.rstd1 = (float*)take(A * 4);
  w.h1 = (float*)take(A * c2 * 4);
  w.y2hat = (float*)take(A * c3 * 4);
  w.rstd2 = (float*)take(A * 4);
  w.d2 = (float*)take(A * c3 * 4);
  w.d1 = (float*)take(A * c2 * 4);
  w.d0 = (float*)take(A * c1 * 4);
  w.m12 = (float*)take(A * 2 * 4);  // fast mode: row means of the sparse layer-2 LN backward
  w.bytes = p - reinterpret_cast<char*>(base);
  return w;
}

int64_t pcrl_pointnet_bwd_workspace(int R, int NP, int c1, int c2, int c3, int CP) {
  return carve_bwd(nullptr, R, NP, c1, c2, c3, CP).bytes;
}

int pcrl_pointnet_bwd(const float* xf, int R, int N, int NP, int CP, int C, const float* pooled,
                      const int32_t* argmax, const float* dpooled, const float* w0, const float* b0, const float* w1,
                      const float* g1, const float* be1, const float* w2, const float* g2, const float* be2, int c1,
                      int c2, int c3, float ln_eps, float* dw0, float* db0, float* dw1, float* dg1, float* dbe1,
                      float* dw2, float* dg2, float* dbe2, void* workspace, int64_t workspace_bytes, int tf32,
                      const void* xh, const void* wpack, void* stream) {
  PCRL_CHECK_ARG(xf && pooled && argmax && dpooled && workspace && dw0 && db0 && dw1 && dg1 && dbe1 && dw2 && dg2 && dbe2);
  PCRL_CHECK_ARG(R >= 0 && NP % 128 == 0 && NP >= N && C <= CP && R <= 1024 * 1024);
  if (R == 0) return PCRL_OK;
  cudaStream_t st = as_stream(stream);
  BwdWs w = carve_bwd(workspace, R, NP, c1, c2, c3, CP);
  PCRL_CHECK_ARG(w.bytes <= workspace_bytes);
  const int A = R * c3;  // capacity
  const int expect = std::max(1, A * 3 / 4);
  int rc;

  // 1. which points carry gradient; compact them and gather their staged rows
  // the fused recompute (first-generation tcgen05 kernel in dump mode) covers widths up to (256, 256, 256); wider
  // PointNets recompute their compacted rows on the GEMM chain below (TF32 tcgen05 GEMMs when tf32 != 0)
  const bool fast = xh && wpack && tc::recompute_supported(c1, c2, c3);
  mark_count_kernel<<<R, 256, 0, st>>>(pooled, argmax, dpooled, NP, c3, w.flag, w.counts);
  PCRL_CHECK_LAUNCH();
  // (fast mode: w.d0 doubles as the gathered bf16 tile scratch, it is only written at the very end of the backward)
  assign_gather_kernel<<<R, 256, 0, st>>>(w.flag, w.counts, R, NP, A, CP, xf, fast ? (const char*)xh : nullptr, w.slot,
                                          w.src, w.total, w.xa, (char*)w.d0);
  PCRL_CHECK_LAUNCH();

  // 2. recompute the forward of the active points, keeping what LN backward needs
  float* dy2 = w.d2;
  if (fast) {
    // fast mode: the same fused tcgen05 kernel that produced the argmax, in dump mode (w.d0 doubles as the
    // gathered bf16 tile scratch: it is only written at the very end of the backward).  dout2 is the max-pool
    // scatter, so the layer-2 LN backward collapses: row means from the sparse entries first, the dense part of dy2
    // written by the recompute kernel in place of xhat2, the sparse term added afterwards (see ln2_sparse_*).
    float *m1 = w.m12, *m2 = w.m12 + A;
    PCRL_CHECK_CUDA(cudaMemsetAsync(w.m12, 0, (int64_t)A * 2 * 4, st));
    const unsigned gpairs = (unsigned)cdiv((int64_t)R * c3, 256);
    ln2_sparse_stats_kernel<<<gpairs, 256, 0, st>>>(pooled, argmax, dpooled, w.slot, g2, be2, R, NP, c3, m1, m2, dg2, dbe2);
    PCRL_CHECK_LAUNCH();
    dy2 = w.y2hat;
    if ((rc = tc::recompute_active_tc(xh, wpack, /*src=*/nullptr, w.total, A, c1, c2, c3, ln_eps, w.d0, w.h0, w.y1hat, w.rstd1,
                                      w.h1, dy2, w.rstd2, m1, m2, st)))
      return rc;
    ln2_sparse_fix_kernel<<<gpairs, 256, 0, st>>>(pooled, argmax, dpooled, w.slot, g2, w.rstd2, R, NP, c3, dy2);
    PCRL_CHECK_LAUNCH();
  } else {
  {
    const unsigned grid = (unsigned)cdiv(A, 16);
    if (CP == 8) layer0_fwd_kernel<8><<<grid, 256, 0, st>>>(w.xa, w0, b0, C, c1, w.total, w.h0);
    else layer0_fwd_kernel<16><<<grid, 256, 0, st>>>(w.xa, w0, b0, C, c1, w.total, w.h0);
    PCRL_CHECK_LAUNCH();
  }
  if ((rc = gemm_nt(w.h0, c1, w1, c1, nullptr, 0, w.h1, c2, A, c1, c2, w.total, st, tf32))) return rc;
  if ((rc = launch_ln_rows(w.h1, c2, g1, be1, w.h1, c2, w.y1hat, w.rstd1, A, c2, ln_eps, 1, w.total, st))) return rc;
  if ((rc = gemm_nt(w.h1, c2, w2, c2, nullptr, 0, w.d2, c3, A, c2, c3, w.total, st, tf32))) return rc;
  // (the post-LN activation of layer 2 itself is not needed: only xhat2 / rstd2)
  if ((rc = launch_ln_rows(w.d2, c3, g2, be2, w.d2, c3, w.y2hat, w.rstd2, A, c3, ln_eps, 1, w.total, st))) return rc;

  // 3. dL/dh2: zero except the argmax entries (ReLU mask holds there: pooled > 0)
  PCRL_CHECK_CUDA(cudaMemsetAsync(w.d2, 0, (int64_t)A * c3 * 4, st));
  scatter_dpool_kernel<<<(unsigned)cdiv((int64_t)R * c3, 256), 256, 0, st>>>(pooled, argmax, dpooled, w.slot, R, NP,
                                                                              c3, w.d2);
  PCRL_CHECK_LAUNCH();
  }

  // The weight-gradient GEMMs only feed the optimizer, the data-gradient chain feeds the next layer: run the
  // wgrads on an internal side stream (event fork/join, graph-capturable) so they overlap the dgrad chain.
  DeviceCtx* ctx = device_ctx();  // the side stream and its events belong to the current device's context
  if (!ctx) return PCRL_ECUDA;
  cudaStream_t side = ctx->side;
  cudaEvent_t ev_fork = ctx->ev_fork, ev_join = ctx->ev_join;
  // 4. layer 2 backward: LN -> dW2 (side), dh1 (main)
  if (!fast && (rc = launch_ln_rows_bwd(w.d2, c3, w.y2hat, w.rstd2, g2, dg2, dbe2, w.d2, c3, A, c3, w.total, st))) return rc;
  PCRL_CHECK_CUDA(cudaEventRecord(ev_fork, st));
  PCRL_CHECK_CUDA(cudaStreamWaitEvent(side, ev_fork, 0));
  if (tf32) {  // contraction tails [count, round_up(count, 32)) must be zero for the TMA-fed wgrad GEMMs
    if ((rc = launch_zero_tail(dy2, c3, w.total, A, side))) return rc;
    if ((rc = launch_zero_tail(w.h1, c2, w.total, A, side))) return rc;
    if ((rc = launch_zero_tail(w.h0, c1, w.total, A, side))) return rc;
  }
  if ((rc = gemm_tn_acc(dy2, c3, w.h1, c2, dw2, c2, A, c3, c2, w.total, expect, side, tf32))) return rc;
  if ((rc = gemm_nn(dy2, c3, w2, c2, w.d1, c2, A, c3, c2, w.total, st, tf32, w.h1))) return rc;  // ReLU bwd fused
  // 5. layer 1 backward: LN -> dW1 (side), dh0 (main)
  if ((rc = launch_ln_rows_bwd(w.d1, c2, w.y1hat, w.rstd1, g1, dg1, dbe1, w.d1, c2, A, c2, w.total, st))) return rc;
  PCRL_CHECK_CUDA(cudaEventRecord(ev_fork, st));
  PCRL_CHECK_CUDA(cudaStreamWaitEvent(side, ev_fork, 0));
  if (tf32 && (rc = launch_zero_tail(w.d1, c2, w.total, A, side))) return rc;
  if ((rc = gemm_tn_acc(w.d1, c2, w.h0, c1, dw1, c1, A, c2, c1, w.total, expect, side, tf32))) return rc;
  PCRL_CHECK_CUDA(cudaEventRecord(ev_join, side));
  if ((rc = gemm_nn(w.d1, c2, w1, c1, w.d0, c1, A, c2, c1, w.total, st, tf32, w.h0))) return rc;  // ReLU bwd fused
  // 6. layer 0 backward: dW0 [c1,C] += d0^T xa[:, :C];  db0 += colsum(d0)
  if (tf32) {
    // fast mode: one more split-K tensor-core GEMM (N = C padded to a 32-column tile) + a column sum
    if ((rc = launch_zero_tail(w.d0, c1, w.total, A, st))) return rc;
    if ((rc = launch_zero_tail(w.xa, CP, w.total, A, st))) return rc;
    if ((rc = gemm_tn_acc(w.d0, c1, w.xa, CP, dw0, C, A, c1, C, w.total, expect, st, 1))) return rc;
    if ((rc = launch_colsum(w.d0, c1, A, c1, w.total, db0, st))) return rc;
  } else {  // parity mode: exact fp32 on the raw coordinates
    const unsigned grid = (unsigned)cdiv(A, 256);
    if (CP == 8) layer0_wgrad_kernel<8><<<grid, 256, 0, st>>>(w.d0, w.xa, C, c1, w.total, dw0, db0);
    else layer0_wgrad_kernel<16><<<grid, 256, 0, st>>>(w.d0, w.xa, C, c1, w.total, dw0, db0);
    PCRL_CHECK_LAUNCH();
  }
  PCRL_CHECK_CUDA(cudaStreamWaitEvent(st, ev_join, 0));  // join the wgrad stream
  return PCRL_OK;
}

}  // extern "C"
