// Dense layers, LayerNorm, column copies: the small row-wise pieces around the GEMMs.
#include "common.cuh"
#include "rowwise.cuh"
#include "tc_gemm.cuh"

namespace pcrl {

// colsum: out[j] += sum_i x[i*ld + j]   (bias gradients).  rows_dev optional.
__global__ void colsum_kernel(const float* __restrict__ x, int64_t ld, int M, int N, const int* rows_dev,
                              float* __restrict__ out) {
  const int Mr = rows_dev ? min(M, *rows_dev) : M;
  const int j = blockIdx.x * 32 + (threadIdx.x & 31);
  const int rows_per_block = (Mr + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(Mr, r0 + rows_per_block);
  float s = 0.f;
  if (j < N)
    for (int i = r0 + (threadIdx.x >> 5); i < r1; i += 8) s += x[(int64_t)i * ld + j];
  __shared__ float red[8][33];
  red[threadIdx.x >> 5][threadIdx.x & 31] = s;
  __syncthreads();
  if (threadIdx.x < 32 && j < N) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    if (t != 0.f) atomicAdd(out + j, t);
  }
}

int launch_colsum(const float* x, int64_t ld, int M, int N, const int* rows_dev, float* out, cudaStream_t st) {
  if (M == 0 || N == 0) return PCRL_OK;
  int gy = (int)std::min<int64_t>(cdiv(M, 64), 256);
  dim3 grid((unsigned)cdiv(N, 32), (unsigned)gy);
  colsum_kernel<<<grid, 256, 0, st>>>(x, ld, M, N, rows_dev, out);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

int launch_gemm(const GemmArgs& g, int tf32, cudaStream_t st) {
  if (tf32 && (g.a_sl == 1 || g.a_si == 1) && (g.b_sl == 1 || g.b_sj == 1)) {
    tcg::TcGemmArgs t{};
    // prefer the K-major reading when both strides are 1 (degenerate K == 1 / M == 1 shapes)
    t.a_mn = (g.a_sl == 1) ? 0 : 1;
    t.A = g.A; t.lda = t.a_mn ? g.a_sl : g.a_si;
    t.b_mn = (g.b_sl == 1) ? 0 : 1;
    t.B = g.B; t.ldb = t.b_mn ? g.b_sl : g.b_sj;
    t.bias = g.bias; t.C = g.C; t.ldc = g.ldc; t.M = g.M; t.N = g.N; t.K = g.K; t.relu = g.relu & 1;
    t.split_k = g.split_k;
    t.mode = g.accumulate ? (g.split_k > 1 ? 2 : 1) : 0;
    t.k_dev = g.k_dev; t.m_dev = g.m_dev; t.mask = g.mask; t.ldmask = g.ldmask; t.bn_hint = g.bn_hint;
    if (g.K >= 8 && tcg::tc_gemm_supported(t)) return tcg::launch_tc_gemm(t, st);
    if (g.K >= 8) {
      // a shape the tensor path is meant for, but an operand is not TMA-addressable (16-byte base, pitch % 4 floats)
      note_tf32_fallback();
      if (strict_tf32()) {
        set_error("tf32 GEMM M=%d N=%d K=%d: operand not TMA-addressable (16-byte base, row pitch %% 4 == 0) and "
                  "pcrl_set_strict_tf32 is on: refusing the FFMA fallback", g.M, g.N, g.K);
        return PCRL_EUNSUPPORTED;
      }
    }
  }
  return launch_sgemm(g, st);
}

__global__ void zero_tail_kernel(float* __restrict__ x, int width, const int* __restrict__ count_dev, int cap) {
  const int c = *count_dev;
  const int end = min((c + 31) & ~31, cap);
  const int64_t n = (int64_t)(end - c) * width;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    x[(int64_t)c * width + i] = 0.f;
}

int launch_zero_tail(float* x, int width, const int* count_dev, int capacity_rows, cudaStream_t st) {
  zero_tail_kernel<<<8, 256, 0, st>>>(x, width, count_dev, capacity_rows);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

// ---- single-output layers (the Q heads' last Linear, Nout == 1): matrix-vector kernels instead of a GEMM tile
// y[m] = x[m,:] . w + b   -- one warp per row
__global__ void __launch_bounds__(256) gemv_fwd_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ w,
                                                       const float* __restrict__ b, float* __restrict__ y, int64_t ldy,
                                                       int M, int K, int relu) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* xr = x + (int64_t)row * ldx;
  float s = 0.f;
  for (int k = lane; k < K; k += 32) s = fmaf(xr[k], __ldg(w + k), s);
  s = warp_sum(s);
  if (lane == 0) {
    s += b ? b[0] : 0.f;
    y[(int64_t)row * ldy] = relu ? fmaxf(s, 0.f) : s;
  }
}
// dw[k] += sum_m dy[m] x[m,k];  db += sum_m dy[m];  dx[m,k] = dy[m] w[k] (masked by relu_mask > 0 if given)
__global__ void __launch_bounds__(256) gemv_bwd_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ w,
                                                       const float* __restrict__ dy, int64_t lddy, float* __restrict__ dw,
                                                       float* __restrict__ db, float* __restrict__ dx, int64_t lddx,
                                                       const float* __restrict__ mask, int64_t ldmask, int M, int K) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int rows_per = (M + gridDim.y - 1) / gridDim.y;
  const int m0 = blockIdx.y * rows_per, m1 = min(M, m0 + rows_per);
  const float wk = (k < K) ? w[k] : 0.f;
  float acc = 0.f, accb = 0.f;
#pragma unroll 4
  for (int m = m0; m < m1; ++m) {
    const float g = __ldg(dy + (int64_t)m * lddy);
    if (k < K) {
      if (dw) acc = fmaf(g, __ldg(x + (int64_t)m * ldx + k), acc);
      if (dx) {
        float v = g * wk;
        if (mask && !(__ldg(mask + (int64_t)m * ldmask + k) > 0.f)) v = 0.f;
        dx[(int64_t)m * lddx + k] = v;
      }
    }
    accb += g;
  }
  if (k < K && dw) atomicAdd(dw + k, acc);
  if (db && k == 0) atomicAdd(db, accb);
}

__global__ void relu_bwd_kernel(float* __restrict__ dy, const float* __restrict__ y, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride)
    if (!(y[i] > 0.f)) dy[i] = 0.f;
}

__global__ void add_cols_kernel(const float* __restrict__ a, int lda, const float* __restrict__ b, int ldb,
                                float* __restrict__ out, int ldo, int M, int width) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)M * width) return;
  int m = (int)(e / width), j = (int)(e % width);
  out[(int64_t)m * ldo + j] = a[(int64_t)m * lda + j] + b[(int64_t)m * ldb + j];
}

__global__ void copy_cols_kernel(const float* __restrict__ src, int lds, int row_div, int row_mul,
                                 float* __restrict__ dst, int ldd, int dst_off, int M, int width) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)M * width) return;
  int m = (int)(e / width), j = (int)(e % width);
  int sm = (m / row_div) * row_mul;
  dst[(int64_t)m * ldd + dst_off + j] = src[(int64_t)sm * lds + j];
}

}  // namespace pcrl

using namespace pcrl;

extern "C" {

int pcrl_linear_fwd(const float* x, int ldx, const float* w, const float* b, float* y, int ldy, int M, int K,
                    int Nout, int relu, int tf32, void* stream) {
  PCRL_CHECK_ARG(x && w && y && M >= 0 && K > 0 && Nout > 0 && ldx >= K && ldy >= Nout);
  if (Nout == 1 && M > 0) {
    gemv_fwd_kernel<<<(unsigned)cdiv((int64_t)M * 32, 256), 256, 0, as_stream(stream)>>>(x, ldx, w, b, y, ldy, M, K, relu);
    PCRL_CHECK_LAUNCH();
    return PCRL_OK;
  }
  GemmArgs g{};
  g.A = x; g.a_si = ldx; g.a_sl = 1;
  g.B = w; g.b_sl = 1; g.b_sj = K;
  g.C = y; g.ldc = ldy;
  g.M = M; g.N = Nout; g.K = K;
  g.bias = b; g.relu = relu; g.accumulate = 0; g.split_k = 1;
  return launch_gemm(g, tf32, as_stream(stream));
}

int pcrl_linear_bwd(const float* x, int ldx, const float* w, const float* dy, int lddy, float* dw, float* db,
                    float* dx, int lddx, const float* relu_mask, int ld_mask, int M, int K, int Nout, int tf32,
                    void* stream) {
  PCRL_CHECK_ARG(x && w && dy && M >= 0 && K > 0 && Nout > 0);
  cudaStream_t st = as_stream(stream);
  if (Nout == 1 && M > 0) {
    dim3 grid((unsigned)cdiv(K, 256), (unsigned)std::min<int64_t>(cdiv(M, 8), 128));
    gemv_bwd_kernel<<<grid, 256, 0, st>>>(x, ldx, w, dy, lddy, dw, db, dx, lddx, relu_mask, ld_mask, M, K);
    PCRL_CHECK_LAUNCH();
    return PCRL_OK;
  }
  // dw[n][k] += sum_m dy[m][n] * x[m][k]
  GemmArgs g{};
  g.A = dy; g.a_si = 1; g.a_sl = lddy;
  g.B = x; g.b_sl = ldx; g.b_sj = 1;
  g.C = dw; g.ldc = K;
  g.M = Nout; g.N = K; g.K = M;
  g.accumulate = 1;
  // enough CTAs to fill the machine: tiles * split_k >= ~2 waves
  const int tile = tf32 ? 128 : 64;
  int64_t tiles = cdiv(Nout, tile) * cdiv(K, tile);
  int split = (int)std::max<int64_t>(1, std::min<int64_t>(cdiv(2 * sm_count(), tiles), cdiv(M, tile)));
  if (tf32) split = (M >= 8192) ? (int)std::min<int64_t>(32, cdiv(M, 2048)) : 1;  // few atomics per address; narrow tiles fill the SMs
  g.split_k = split;
  int rc = 0;
  if (dw && (rc = launch_gemm(g, tf32, st))) return rc;
  if (db) {
    rc = launch_colsum(dy, lddy, M, Nout, nullptr, db, st);
    if (rc) return rc;
  }
  if (dx) {
    // dx[m][k] = sum_n dy[m][n] * w[n][k]
    GemmArgs h{};
    h.A = dy; h.a_si = lddy; h.a_sl = 1;
    h.B = w; h.b_sl = K; h.b_sj = 1;
    h.C = dx; h.ldc = lddx;
    h.M = M; h.N = K; h.K = Nout;
    h.split_k = 1;
    h.mask = relu_mask; h.ldmask = ld_mask;  // dx *= (x_post_activation > 0), fused
    rc = launch_gemm(h, tf32, st);
    if (rc) return rc;
  }
  return PCRL_OK;
}

int pcrl_relu_bwd(float* dy, const float* y, int64_t n, void* stream) {
  if (n == 0) return PCRL_OK;
  int blocks = (int)std::min<int64_t>(cdiv(n, 256), 148 * 8);
  relu_bwd_kernel<<<blocks, 256, 0, as_stream(stream)>>>(dy, y, n);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

int pcrl_add_cols(const float* a, int lda, const float* b, int ldb, float* out, int ldo, int M, int width,
                  void* stream) {
  PCRL_CHECK_ARG(a && b && out && M >= 0 && width >= 0);
  int64_t n = (int64_t)M * width;
  if (n == 0) return PCRL_OK;
  add_cols_kernel<<<(unsigned)cdiv(n, 256), 256, 0, as_stream(stream)>>>(a, lda, b, ldb, out, ldo, M, width);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

int pcrl_layernorm_fwd(const float* x, const float* g, const float* b, float* y, int ldy, float* xhat, float* rstd,
                       int M, int D, float eps, void* stream) {
  PCRL_CHECK_ARG(x && g && b && y && M >= 0 && D > 0 && D <= 1024);
  return launch_ln_rows(x, D, g, b, y, ldy, xhat, rstd, M, D, eps, /*relu=*/0, nullptr, as_stream(stream));
}

int pcrl_layernorm_bwd(const float* dy, int lddy, const float* xhat, const float* rstd, const float* g, float* dg,
                       float* db, float* dx, int M, int D, void* stream) {
  PCRL_CHECK_ARG(dy && xhat && rstd && g && dx && M >= 0 && D > 0 && D <= 1024);
  return launch_ln_rows_bwd(dy, lddy, xhat, rstd, g, dg, db, dx, /*lddx=*/lddy, M, D, nullptr, as_stream(stream));
}

int pcrl_copy_cols(const float* src, int lds, int src_row_div, int src_row_mul, float* dst, int ldd, int dst_off,
                   int M, int width, void* stream) {
  PCRL_CHECK_ARG(src && dst && src_row_div >= 1 && src_row_mul >= 1 && width >= 0);
  int64_t n = (int64_t)M * width;
  if (n == 0) return PCRL_OK;
  copy_cols_kernel<<<(unsigned)cdiv(n, 256), 256, 0, as_stream(stream)>>>(src, lds, src_row_div, src_row_mul, dst,
                                                                            ldd, dst_off, M, width);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

}  // extern "C"
