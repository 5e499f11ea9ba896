// LayerNorm over the last dimension, one warp per row.  Used for LN1d over channels per point
// (nn_layer.py:209-219, eps 1e-6) and for PointNet.final_mlp's nn.LayerNorm (pointnet.py:110, eps 1e-5).
#include "rowwise.cuh"

namespace pcrl {

__global__ void __launch_bounds__(256) ln_rows_kernel(const float* __restrict__ x, int64_t ldx,
                                                      const float* __restrict__ g, const float* __restrict__ b,
                                                      float* __restrict__ y, int64_t ldy, float* __restrict__ xhat,
                                                      float* __restrict__ rstd_out, int M, int D, float eps, int relu,
                                                      const int* rows_dev) {
  const int rows = rows_dev ? min(M, *rows_dev) : M;
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const float inv_d = 1.0f / (float)D;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < rows; i += warps) {
    const float* xr = x + (int64_t)i * ldx;
    float s = 0.f;
    for (int c = lane; c < D; c += 32) s += xr[c];
    const float mean = warp_sum(s) * inv_d;
    float v = 0.f;
    for (int c = lane; c < D; c += 32) {
      float d = xr[c] - mean;
      v = fmaf(d, d, v);
    }
    const float rstd = 1.0f / sqrtf(warp_sum(v) * inv_d + eps);
    if (rstd_out && lane == 0) rstd_out[i] = rstd;
    float* yr = y + (int64_t)i * ldy;
    for (int c = lane; c < D; c += 32) {
      float xh = (xr[c] - mean) * rstd;
      if (xhat) xhat[(int64_t)i * D + c] = xh;
      float o = fmaf(xh, g[c], b[c]);
      if (relu & 1) o = fmaxf(o, 0.f);
      if (relu & 2) o = round_tf32(o);  // the row feeds a TF32 GEMM: round instead of letting the tensor core truncate
      yr[c] = o;
    }
  }
}

template <int VPL>
__global__ void __launch_bounds__(256) ln_rows_bwd_kernel(const float* __restrict__ dy, int64_t lddy,
                                                          const float* __restrict__ xhat,
                                                          const float* __restrict__ rstd, const float* __restrict__ g,
                                                          float* __restrict__ dg, float* __restrict__ db,
                                                          float* __restrict__ dx, int64_t lddx, int M, int D,
                                                          const int* rows_dev) {
  const int rows = rows_dev ? min(M, *rows_dev) : M;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const float inv_d = 1.0f / (float)D;
  float pg[VPL], pb[VPL], gv[VPL];
#pragma unroll
  for (int q = 0; q < VPL; ++q) {
    pg[q] = pb[q] = 0.f;
    int c = lane + 32 * q;
    gv[q] = c < D ? g[c] : 0.f;
  }
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < rows; i += warps) {
    float d[VPL], xh[VPL];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int q = 0; q < VPL; ++q) {
      int c = lane + 32 * q;
      d[q] = c < D ? dy[(int64_t)i * lddy + c] : 0.f;
      xh[q] = c < D ? xhat[(int64_t)i * D + c] : 0.f;
      pg[q] = fmaf(d[q], xh[q], pg[q]);
      pb[q] += d[q];
      float t = d[q] * gv[q];
      s1 += t;
      s2 = fmaf(t, xh[q], s2);
    }
    const float m1 = warp_sum(s1) * inv_d, m2 = warp_sum(s2) * inv_d, r = rstd[i];
#pragma unroll
    for (int q = 0; q < VPL; ++q) {
      int c = lane + 32 * q;
      if (c < D) dx[(int64_t)i * lddx + c] = r * (d[q] * gv[q] - m1 - xh[q] * m2);
    }
  }
  // reduce the per-warp column partials through shared memory, one atomic per column per block
  __shared__ float sg[8][33], sb[8][33];
#pragma unroll
  for (int q = 0; q < VPL; ++q) {
    __syncthreads();
    sg[warp][lane] = pg[q];
    sb[warp][lane] = pb[q];
    __syncthreads();
    if (warp == 0) {
      float tg = 0.f, tb = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        tg += sg[w][lane];
        tb += sb[w][lane];
      }
      int c = lane + 32 * q;
      if (c < D) {
        if (dg && tg != 0.f) atomicAdd(dg + c, tg);
        if (db && tb != 0.f) atomicAdd(db + c, tb);
      }
    }
  }
}

// Vectorised variant for D == 128 * NV with 16-byte aligned rows: one float4 per lane and array covers 128 columns,
// and every warp keeps ROWS rows in flight (the backward over the ~100 k compacted PointNet rows is a pure HBM stream:
// bytes in flight per SM, not instruction count, set its speed).
template <int NV, int ROWS>
__global__ void __launch_bounds__(256) ln_rows_bwd_vec_kernel(const float* __restrict__ dy, int64_t lddy,
                                                              const float* __restrict__ xhat,
                                                              const float* __restrict__ rstd, const float* __restrict__ g,
                                                              float* __restrict__ dg, float* __restrict__ db,
                                                              float* __restrict__ dx, int64_t lddx, int M,
                                                              const int* rows_dev) {
  constexpr int D = 128 * NV;
  const int rows = rows_dev ? min(M, *rows_dev) : M;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const float inv_d = 1.0f / (float)D;
  float4 pg[NV], pb[NV], gv[NV];
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    pg[q] = pb[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    gv[q] = *reinterpret_cast<const float4*>(g + q * 128 + lane * 4);
  }
  for (int i0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * ROWS; i0 < rows; i0 += warps * ROWS) {
    float4 d[ROWS][NV], xh[ROWS][NV];
#pragma unroll
    for (int u = 0; u < ROWS; ++u) {
      const int i = min(i0 + u, rows - 1);
#pragma unroll
      for (int q = 0; q < NV; ++q) {
        d[u][q] = *reinterpret_cast<const float4*>(dy + (int64_t)i * lddy + q * 128 + lane * 4);
        xh[u][q] = *reinterpret_cast<const float4*>(xhat + (int64_t)i * D + q * 128 + lane * 4);
      }
    }
#pragma unroll
    for (int u = 0; u < ROWS; ++u) {
      if (i0 + u >= rows) break;
      const int i = i0 + u;
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int q = 0; q < NV; ++q) {
        const float dv[4] = {d[u][q].x, d[u][q].y, d[u][q].z, d[u][q].w};
        const float xv[4] = {xh[u][q].x, xh[u][q].y, xh[u][q].z, xh[u][q].w};
        const float gq[4] = {gv[q].x, gv[q].y, gv[q].z, gv[q].w};
        float* pgq = reinterpret_cast<float*>(&pg[q]);
        float* pbq = reinterpret_cast<float*>(&pb[q]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          pgq[e] = fmaf(dv[e], xv[e], pgq[e]);
          pbq[e] += dv[e];
          const float t = dv[e] * gq[e];
          s1 += t;
          s2 = fmaf(t, xv[e], s2);
        }
      }
      const float m1 = warp_sum(s1) * inv_d, m2 = warp_sum(s2) * inv_d, r = rstd[i];
#pragma unroll
      for (int q = 0; q < NV; ++q) {
        float4 o;
        o.x = r * (d[u][q].x * gv[q].x - m1 - xh[u][q].x * m2);
        o.y = r * (d[u][q].y * gv[q].y - m1 - xh[u][q].y * m2);
        o.z = r * (d[u][q].z * gv[q].z - m1 - xh[u][q].z * m2);
        o.w = r * (d[u][q].w * gv[q].w - m1 - xh[u][q].w * m2);
        *reinterpret_cast<float4*>(dx + (int64_t)i * lddx + q * 128 + lane * 4) = o;
      }
    }
  }
  // per-warp column partials -> shared memory -> one atomic per column per block
  __shared__ float sg[8][128 + 4], sb[8][128 + 4];
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    __syncthreads();
    *reinterpret_cast<float4*>(&sg[warp][lane * 4]) = pg[q];
    *reinterpret_cast<float4*>(&sb[warp][lane * 4]) = pb[q];
    __syncthreads();
    if (threadIdx.x < 128) {
      float tg = 0.f, tb = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        tg += sg[w][threadIdx.x];
        tb += sb[w][threadIdx.x];
      }
      const int c = q * 128 + threadIdx.x;
      if (dg && tg != 0.f) atomicAdd(dg + c, tg);
      if (db && tb != 0.f) atomicAdd(db + c, tb);
    }
  }
}

int launch_ln_rows(const float* x, int64_t ldx, const float* g, const float* b, float* y, int64_t ldy, float* xhat,
                   float* rstd, int M, int D, float eps, int relu, const int* rows_dev, cudaStream_t st) {
  if (M == 0) return PCRL_OK;
  int blocks = (int)std::min<int64_t>(cdiv(M, 8), (int64_t)sm_count() * 8);
  ln_rows_kernel<<<blocks, 256, 0, st>>>(x, ldx, g, b, y, ldy, xhat, rstd, M, D, eps, relu, rows_dev);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

int launch_ln_rows_bwd(const float* dy, int64_t lddy, const float* xhat, const float* rstd, const float* g, float* dg,
                       float* db, float* dx, int64_t lddx, int M, int D, const int* rows_dev, cudaStream_t st) {
  if (M == 0) return PCRL_OK;
  PCRL_CHECK_ARG(D <= 1024);
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if ((D == 128 || D == 256) && al(dy) && al(xhat) && al(dx) && al(g) && lddy % 4 == 0 && lddx % 4 == 0 && M >= 256) {
    const int blocks_v = (int)std::min<int64_t>(cdiv(M, 16), (int64_t)sm_count() * 4);
    if (D == 128)
      ln_rows_bwd_vec_kernel<1, 4><<<blocks_v, 256, 0, st>>>(dy, lddy, xhat, rstd, g, dg, db, dx, lddx, M, rows_dev);
    else
      ln_rows_bwd_vec_kernel<2, 2><<<blocks_v, 256, 0, st>>>(dy, lddy, xhat, rstd, g, dg, db, dx, lddx, M, rows_dev);
    PCRL_CHECK_LAUNCH();
    return PCRL_OK;
  }
  int blocks = (int)std::min<int64_t>(cdiv(M, 8), (int64_t)sm_count() * 2);
  if (D <= 128)
    ln_rows_bwd_kernel<4><<<blocks, 256, 0, st>>>(dy, lddy, xhat, rstd, g, dg, db, dx, lddx, M, D, rows_dev);
  else if (D <= 256)
    ln_rows_bwd_kernel<8><<<blocks, 256, 0, st>>>(dy, lddy, xhat, rstd, g, dg, db, dx, lddx, M, D, rows_dev);
  else
    ln_rows_bwd_kernel<32><<<blocks, 256, 0, st>>>(dy, lddy, xhat, rstd, g, dg, db, dx, lddx, M, D, rows_dev);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

}  // namespace pcrl
