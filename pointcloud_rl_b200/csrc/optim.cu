// Fused multi-tensor Adam + grad-norm + Polyak over flat fp32 buffers.  HBM-bound: per parameter it
// reads p,g,m,v (+target) and writes p,m,v (+target) = 28 B (+8 B) -- one pass, float4 vectorised.
#include "common.cuh"

namespace pcrl {

__global__ void bump_step_kernel(int32_t* step) { step[0] += 1; }

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, int64_t n, float lr,
                                                   float b1, float b2, float eps, float gscale,
                                                   const int32_t* __restrict__ step_dev, float* __restrict__ gradsq,
                                                   float* __restrict__ target, int64_t pb, int64_t pe, float tau) {
  // torch.optim.Adam: step_size = lr/(1-b1^t); denom = sqrt(v)/sqrt(1-b2^t) + eps
  const int t = step_dev[0];
  const float bc1 = 1.f - powf(b1, (float)t);
  const float bc2 = 1.f - powf(b2, (float)t);
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  float sq = 0.f;
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    float4 gv = reinterpret_cast<const float4*>(g)[i];
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float pa[4] = {pv.x, pv.y, pv.z, pv.w}, ga[4] = {gv.x, gv.y, gv.z, gv.w};
    float ma[4] = {mv.x, mv.y, mv.z, mv.w}, va[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gk = ga[k] * gscale;
      sq = fmaf(gk, gk, sq);
      ma[k] = b1 * ma[k] + (1.f - b1) * gk;
      va[k] = b2 * va[k] + (1.f - b2) * gk * gk;
      const float denom = sqrtf(va[k]) * inv_sqrt_bc2 + eps;
      pa[k] -= step_size * (ma[k] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = make_float4(pa[0], pa[1], pa[2], pa[3]);
    reinterpret_cast<float4*>(m)[i] = make_float4(ma[0], ma[1], ma[2], ma[3]);
    reinterpret_cast<float4*>(v)[i] = make_float4(va[0], va[1], va[2], va[3]);
    if (target) {
      const int64_t e = i << 2;
      if (e >= pb && e + 3 < pe) {
        float4 tv = reinterpret_cast<float4*>(target)[(e - pb) >> 2];
        tv.x = tv.x * (1.f - tau) + pa[0] * tau;  // ops.py:64
        tv.y = tv.y * (1.f - tau) + pa[1] * tau;
        tv.z = tv.z * (1.f - tau) + pa[2] * tau;
        tv.w = tv.w * (1.f - tau) + pa[3] * tau;
        reinterpret_cast<float4*>(target)[(e - pb) >> 2] = tv;
      }
    }
  }
  // tail (n not a multiple of 4)
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gk = g[i] * gscale;
    sq = fmaf(gk, gk, sq);
    const float mk = b1 * m[i] + (1.f - b1) * gk;
    const float vk = b2 * v[i] + (1.f - b2) * gk * gk;
    m[i] = mk;
    v[i] = vk;
    const float pn = p[i] - step_size * (mk / (sqrtf(vk) * inv_sqrt_bc2 + eps));
    p[i] = pn;
    if (target && i >= pb && i < pe) target[i - pb] = target[i - pb] * (1.f - tau) + pn * tau;
  }
  if (gradsq) {
    sq = warp_sum(sq);
    __shared__ float red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sq;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t2 = 0.f;
      for (int w = 0; w < 8; ++w) t2 += red[w];
      atomicAdd(gradsq, t2);
    }
  }
}

__global__ void polyak_kernel(float* __restrict__ t, const float* __restrict__ s, int64_t n, float tau) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    t[i] = t[i] * (1.f - tau) + s[i] * tau;
}

}  // namespace pcrl

using namespace pcrl;

extern "C" {

int pcrl_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                   float eps, float grad_scale, int32_t* step_dev, float* gradsq_out, float* target,
                   int64_t poly_begin, int64_t poly_end, float tau, void* stream) {
  PCRL_CHECK_ARG(p && g && m && v && step_dev && n >= 0);
  PCRL_CHECK_ARG(!target || (poly_begin % 4 == 0 && poly_begin >= 0 && poly_end <= n && (poly_end % 4 == 0 || poly_end == n)));
  if (n == 0) return PCRL_OK;
  cudaStream_t st = as_stream(stream);
  bump_step_kernel<<<1, 1, 0, st>>>(step_dev);
  PCRL_CHECK_LAUNCH();
  if (gradsq_out) PCRL_CHECK_CUDA(cudaMemsetAsync(gradsq_out, 0, sizeof(float), st));
  const int blocks = (int)std::min<int64_t>(cdiv(cdiv(n, 4), 256), (int64_t)sm_count() * 8);
  adam_kernel<<<std::max(blocks, 1), 256, 0, st>>>(p, g, m, v, n, lr, beta1, beta2, eps, grad_scale, step_dev,
                                                   gradsq_out, target, poly_begin, poly_end, tau);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

int pcrl_adam_step_part(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                        float eps, float grad_scale, int32_t* step_dev, int bump_step, float* gradsq_out, int zero_gradsq,
                        float* target, int64_t poly_begin, int64_t poly_end, float tau, void* stream) {
  PCRL_CHECK_ARG(p && g && m && v && step_dev && n >= 0);
  PCRL_CHECK_ARG(!target || (poly_begin % 4 == 0 && poly_begin >= 0 && poly_end <= n && (poly_end % 4 == 0 || poly_end == n)));
  cudaStream_t st = as_stream(stream);
  if (bump_step) {
    bump_step_kernel<<<1, 1, 0, st>>>(step_dev);
    PCRL_CHECK_LAUNCH();
  }
  if (gradsq_out && zero_gradsq) PCRL_CHECK_CUDA(cudaMemsetAsync(gradsq_out, 0, sizeof(float), st));
  if (n == 0) return PCRL_OK;
  const int blocks = (int)std::min<int64_t>(cdiv(cdiv(n, 4), 256), (int64_t)sm_count() * 8);
  adam_kernel<<<std::max(blocks, 1), 256, 0, st>>>(p, g, m, v, n, lr, beta1, beta2, eps, grad_scale, step_dev,
                                                   gradsq_out, target, poly_begin, poly_end, tau);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}
int pcrl_polyak(float* target, const float* source, int64_t n, float tau, void* stream) {
  PCRL_CHECK_ARG(target && source && n >= 0);
  if (n == 0) return PCRL_OK;
  const int blocks = (int)std::min<int64_t>(cdiv(n, 256), (int64_t)sm_count() * 8);
  polyak_kernel<<<blocks, 256, 0, as_stream(stream)>>>(target, source, n, tau);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

}  // extern "C"
