// Policy head, TD target and losses: the small elementwise/reduction pieces of the SAC/DrQ step.
// Everything the step logs is written into one device scalar array (single D2H per update).
#include "common.cuh"

namespace pcrl {

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x < 32) {
    t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
  }
  return t;  // valid in warp 0
}
__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = -INFINITY;
  if (threadIdx.x < 32) {
    t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : -INFINITY;
    t = warp_max(t);
  }
  return t;
}

// one warp per row, lanes over the action dimensions (A is small, <= 64): the per-element Philox + Box-Muller +
// exp/tanh/log chain is ~150 instructions, far too long to serialise A of them in one thread
__global__ void __launch_bounds__(256)
tanh_gaussian_fwd_kernel(const float* __restrict__ out, int M, int A, float ls_lo, float ls_hi, float scale,
                         float bias, const float* __restrict__ eps_in, uint64_t seed,
                         const uint64_t* __restrict__ counter_dev, uint32_t stream_id, float* __restrict__ action,
                         int ld_action, float* __restrict__ neglogp, float* __restrict__ eps_out) {
  const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (m >= M) return;
  const uint64_t cnt = counter_dev ? *counter_dev : 0ull;
  const float half_log_2pi = 0.91893853320467274178f;
  float acc = 0.f;
  for (int j = lane; j < A; j += 32) {
    const float mu = out[(int64_t)m * 2 * A + j];
    const float ls = out[(int64_t)m * 2 * A + A + j];
    const float log_std = fminf(fmaxf(ls, ls_lo), ls_hi);
    const float std = expf(log_std);
    float e;
    if (eps_in) {
      e = eps_in[(int64_t)m * A + j];
    } else {
      uint4 rnd = philox4x32_10(make_uint4((uint32_t)(j >> 1), (uint32_t)m, (uint32_t)cnt, (uint32_t)(cnt >> 32) ^ (stream_id << 24)),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
      float2 n2 = box_muller(rnd.x, rnd.y);
      e = (j & 1) ? n2.y : n2.x;
    }
    eps_out[(int64_t)m * A + j] = e;
    const float u = fmaf(std, e, mu);
    const float t = tanhf(u);
    // Normal.log_prob(u) = -(u-mu)^2/(2 std^2) - log std - log sqrt(2 pi); (u-mu)/std == e
    float logp = -0.5f * e * e - log_std - half_log_2pi;
    logp -= logf(scale * (1.f - t * t) + 1e-6f);  // distributions.py:89
    acc += logp;
    action[(int64_t)m * ld_action + j] = fmaf(t, scale, bias);
  }
  acc = warp_sum(acc);
  if (lane == 0) neglogp[m] = -acc;
}

// dout[m, j]   = d/dmu      = g_u
// dout[m, A+j] = d/dlogstd  = (g_u * std * eps + g_nlp) * 1[lo <= ls <= hi]
// with g_u = da * scale*(1-t^2) + g_nlp * d neglogp/du,  d neglogp/du = -2 t scale (1-t^2) / (scale(1-t^2)+1e-6)
__global__ void tanh_gaussian_bwd_kernel(const float* __restrict__ out, const float* __restrict__ eps,
                                         const float* __restrict__ da, int ld_da, float g_nlp_host,
                                         const float* __restrict__ alpha_dev, int M, int A, float ls_lo, float ls_hi,
                                         float scale, float* __restrict__ dout) {
  const int e_idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (e_idx >= M * A) return;
  const int m = e_idx / A, j = e_idx % A;
  const float g_nlp = alpha_dev ? -(*alpha_dev) / (float)M : g_nlp_host;
  const float mu = out[(int64_t)m * 2 * A + j];
  const float ls = out[(int64_t)m * 2 * A + A + j];
  const bool inside = (ls >= ls_lo) && (ls <= ls_hi);
  const float std = expf(fminf(fmaxf(ls, ls_lo), ls_hi));
  const float e = eps[e_idx];
  const float t = tanhf(fmaf(std, e, mu));
  const float s = scale * (1.f - t * t);
  const float dn_du = -2.f * t * s / (s + 1e-6f);
  const float g_u = da[(int64_t)m * ld_da + j] * s + g_nlp * dn_du;
  dout[(int64_t)m * 2 * A + j] = g_u;
  dout[(int64_t)m * 2 * A + A + j] = inside ? (g_u * std * e + g_nlp) : 0.f;
}

// one block; B*group rows.  y = r*rs + (1-d)*gamma*(min(q0,q1) + alpha*nlp), mean over group, broadcast
__global__ void td_target_kernel(const float* __restrict__ qt, const float* __restrict__ neglogp,
                                 const float* __restrict__ rewards, const uint8_t* __restrict__ dones, int B, int group,
                                 float gamma, float reward_scale, int ignore_dones, const float* __restrict__ alpha_dev,
                                 float* __restrict__ y) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float alpha = *alpha_dev;
  const float nd = (ignore_dones || !dones) ? 1.f : (1.f - (dones[b] ? 1.f : 0.f));
  float s = 0.f;
  for (int a = 0; a < group; ++a) {
    const int r = b * group + a;
    const float v = fminf(qt[2 * r], qt[2 * r + 1]) + alpha * neglogp[r];
    s += rewards[b] * reward_scale + nd * gamma * v;
  }
  s /= (float)group;
  for (int a = 0; a < group; ++a) y[b * group + a] = s;
}

__global__ void __launch_bounds__(256) critic_loss_kernel(const float* __restrict__ q, const float* __restrict__ y,
                                                          int R, float* __restrict__ dq, float* __restrict__ scalars) {
  __shared__ float red[8];
  float l = 0.f, mx = 0.f, qm = 0.f, ym = 0.f;
  const float inv = 1.f / (float)R;
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
    const float q0 = q[2 * r], q1 = q[2 * r + 1], t = y[r];
    const float e0 = q0 - t, e1 = q1 - t;
    l += e0 * e0 + e1 * e1;
    mx = fmaxf(mx, fmaxf(fabsf(e0), fabsf(e1)));
    qm += fminf(q0, q1);
    ym += t;
    dq[2 * r] = 2.f * e0 * inv;      // d/dq of (1/R) sum_rows sum_heads (q-y)^2 == mse*2 (sac.py:137)
    dq[2 * r + 1] = 2.f * e1 * inv;
  }
  l = block_sum(l, red);
  qm = block_sum(qm, red);
  ym = block_sum(ym, red);
  mx = block_max(mx, red);
  if (threadIdx.x == 0) {
    scalars[PCRL_S_CRITIC_LOSS] = l * inv;
    scalars[PCRL_S_MAX_ABS_ERR] = mx;
    scalars[PCRL_S_Q] = qm * inv;
    scalars[PCRL_S_Q_TARGET] = ym * inv;
  }
}

__global__ void __launch_bounds__(256) actor_loss_kernel(const float* __restrict__ q, const float* __restrict__ neglogp,
                                                         int M, const float* __restrict__ alpha_dev,
                                                         const float* __restrict__ log_alpha, float target_entropy,
                                                         float* __restrict__ dq, float* __restrict__ dlog_alpha,
                                                         float* __restrict__ scalars) {
  __shared__ float red[8];
  float qs = 0.f, es = 0.f;
  const float inv = 1.f / (float)M;
  for (int m = threadIdx.x; m < M; m += blockDim.x) {
    const float q0 = q[2 * m], q1 = q[2 * m + 1];
    const bool first = q0 <= q1;  // torch.min(dim) backward routes to the first minimal index
    qs += first ? q0 : q1;
    es += neglogp[m];
    dq[2 * m] = first ? -inv : 0.f;
    dq[2 * m + 1] = first ? 0.f : -inv;
  }
  qs = block_sum(qs, red);
  es = block_sum(es, red);
  if (threadIdx.x == 0) {
    const float alpha = *alpha_dev;
    const float entropy = es * inv;
    const float actor_loss = -(qs * inv + alpha * entropy);              // sac.py:183
    const float g = expf(log_alpha[0]) * (entropy - target_entropy);    // sac.py:190 (value == d/dlog_alpha)
    scalars[PCRL_S_ACTOR_LOSS] = actor_loss;
    scalars[PCRL_S_ENTROPY] = entropy;
    scalars[PCRL_S_ALPHA_LOSS] = g;
    scalars[PCRL_S_ALPHA_GRAD] = g;
    dlog_alpha[0] = g;
  }
}

__global__ void refresh_alpha_kernel(const float* __restrict__ log_alpha, float* __restrict__ alpha_dev,
                                     float* __restrict__ scalars) {
  const float a = expf(log_alpha[0]);
  alpha_dev[0] = a;
  if (scalars) scalars[PCRL_S_ALPHA] = a;
}

}  // namespace pcrl

using namespace pcrl;

extern "C" {

int pcrl_tanh_gaussian_fwd(const float* out, int M, int A, float ls_lo, float ls_hi, float scale, float bias,
                           const float* eps, uint64_t seed, const uint64_t* counter_dev, uint32_t stream_id,
                           float* action, int ld_action, float* neglogp, float* eps_out, void* stream) {
  PCRL_CHECK_ARG(out && action && neglogp && eps_out && M >= 0 && A > 0 && ld_action >= A);
  if (M == 0) return PCRL_OK;
  tanh_gaussian_fwd_kernel<<<(unsigned)cdiv((int64_t)M * 32, 256), 256, 0, as_stream(stream)>>>(
      out, M, A, ls_lo, ls_hi, scale, bias, eps, seed, counter_dev, stream_id, action, ld_action, neglogp, eps_out);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

int pcrl_tanh_gaussian_bwd(const float* out, const float* eps, const float* daction, int ld_daction, float g_nlp,
                           int M, int A, float ls_lo, float ls_hi, float scale, float* dout, void* stream) {
  PCRL_CHECK_ARG(out && eps && daction && dout && M >= 0 && A > 0);
  if (M == 0) return PCRL_OK;
  tanh_gaussian_bwd_kernel<<<(unsigned)cdiv((int64_t)M * A, 256), 256, 0, as_stream(stream)>>>(
      out, eps, daction, ld_daction, g_nlp, nullptr, M, A, ls_lo, ls_hi, scale, dout);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

int pcrl_tanh_gaussian_bwd_dev(const float* out, const float* eps, const float* daction, int ld_daction,
                               const float* alpha_dev, int M, int A, float ls_lo, float ls_hi, float scale,
                               float* dout, void* stream) {
  PCRL_CHECK_ARG(out && eps && daction && dout && alpha_dev && M >= 0 && A > 0);
  if (M == 0) return PCRL_OK;
  tanh_gaussian_bwd_kernel<<<(unsigned)cdiv((int64_t)M * A, 256), 256, 0, as_stream(stream)>>>(
      out, eps, daction, ld_daction, 0.f, alpha_dev, M, A, ls_lo, ls_hi, scale, dout);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

int pcrl_td_target(const float* qt, const float* neglogp, const float* rewards, const uint8_t* dones, int B,
                   int group, float gamma, float reward_scale, int ignore_dones, const float* alpha_dev, float* y,
                   void* stream) {
  PCRL_CHECK_ARG(qt && neglogp && rewards && alpha_dev && y && B >= 0 && group >= 1);
  if (B == 0) return PCRL_OK;
  td_target_kernel<<<(unsigned)cdiv(B, 128), 128, 0, as_stream(stream)>>>(qt, neglogp, rewards, dones, B, group, gamma,
                                                                           reward_scale, ignore_dones, alpha_dev, y);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

int pcrl_critic_loss(const float* q, const float* y, int R, float* dq, float* scalars, void* stream) {
  PCRL_CHECK_ARG(q && y && dq && scalars && R > 0);
  critic_loss_kernel<<<1, 256, 0, as_stream(stream)>>>(q, y, R, dq, scalars);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

int pcrl_actor_loss(const float* q, const float* neglogp, int M, const float* alpha_dev, const float* log_alpha,
                    float target_entropy, float* dq, float* dlog_alpha, float* scalars, void* stream) {
  PCRL_CHECK_ARG(q && neglogp && alpha_dev && log_alpha && dq && dlog_alpha && scalars && M > 0);
  actor_loss_kernel<<<1, 256, 0, as_stream(stream)>>>(q, neglogp, M, alpha_dev, log_alpha, target_entropy, dq,
                                                       dlog_alpha, scalars);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

int pcrl_refresh_alpha(const float* log_alpha, float* alpha_dev, float* scalars, void* stream) {
  PCRL_CHECK_ARG(log_alpha && alpha_dev);
  refresh_alpha_kernel<<<1, 1, 0, as_stream(stream)>>>(log_alpha, alpha_dev, scalars);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

}  // extern "C"
