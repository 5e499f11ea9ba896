// ColorJitterPoints (pyrl/utils/augmentations/pcd_aug.py:269-303): torchvision.transforms.ColorJitter applied to the
// uint8 point colours viewed as a [B', 3, 1, N] image batch.  ONE parameter set per call -- a random order of
// (brightness, contrast, saturation, hue) and one factor each -- is shared by every cloud of the batch, so the
// num_aug copies of a sample (repeat_interleave'd before the augmentation, drq.py:58) come out identical and the
// transform runs once per SOURCE cloud; the staging kernel repeats the result.
//
// Bit-exact restatement of torchvision 0.26's tensor path for uint8 input (_functional_tensor.py): every op rounds
// through uint8 (truncation after clamp), the products and sums are separate fp32 roundings (no FMA contraction:
// explicit __fmul_rn / __fadd_rn), contrast blends with the per-cloud mean of the truncated grey values.
#include "common.cuh"

namespace pcrl {

struct CjParams {
  int order[4];        // fn_idx: 0 brightness, 1 contrast, 2 saturation, 3 hue
  float b, c, s, h;    // factors
};

__device__ __forceinline__ float clamp255(float x) { return fminf(fmaxf(x, 0.f), 255.f); }
__device__ __forceinline__ float grey_of(float r, float g, float b) {
  // (0.2989 * r + 0.587 * g + 0.114 * b).to(uint8): left-to-right fp32, truncation
  const float v = __fadd_rn(__fadd_rn(__fmul_rn(0.2989f, r), __fmul_rn(0.587f, g)), __fmul_rn(0.114f, b));
  return (float)(uint8_t)v;
}
// _blend(img1, img2, ratio) for one channel: (ratio * a + (1 - ratio) * b).clamp(0, 255).to(uint8)
__device__ __forceinline__ float blend(float a, float b, float ratio, float one_minus) {
  return (float)(uint8_t)clamp255(__fadd_rn(__fmul_rn(ratio, a), __fmul_rn(one_minus, b)));
}

__device__ __forceinline__ void hue_adjust(float& r8, float& g8, float& b8, float hue) {
  // convert_image_dtype(uint8 -> float32): x / 255
  const float r = __fdiv_rn(r8, 255.f), g = __fdiv_rn(g8, 255.f), b = __fdiv_rn(b8, 255.f);
  // _rgb2hsv
  const float maxc = fmaxf(fmaxf(r, g), b), minc = fminf(fminf(r, g), b);
  const bool eqc = maxc == minc;
  const float cr = __fsub_rn(maxc, minc);
  const float s = __fdiv_rn(cr, eqc ? 1.f : maxc);
  const float div = eqc ? 1.f : cr;
  const float rc = __fdiv_rn(__fsub_rn(maxc, r), div), gc = __fdiv_rn(__fsub_rn(maxc, g), div),
              bc = __fdiv_rn(__fsub_rn(maxc, b), div);
  const float hr = (maxc == r) ? __fsub_rn(bc, gc) : 0.f;
  const float hg = ((maxc == g) && (maxc != r)) ? __fsub_rn(__fadd_rn(2.f, rc), bc) : 0.f;
  const float hb = ((maxc != g) && (maxc != r)) ? __fsub_rn(__fadd_rn(4.f, gc), rc) : 0.f;
  float h = __fadd_rn(__fadd_rn(hr, hg), hb);
  h = fmodf(__fadd_rn(__fdiv_rn(h, 6.f), 1.f), 1.f);
  // h = (h + hue_factor) % 1.0   (torch.remainder: result in [0, 1))
  h = __fadd_rn(h, hue);
  h = __fsub_rn(h, floorf(h));  // remainder(x, 1) = x - floor(x) exactly representable for |x| < 2
  if (h >= 1.f) h = 0.f;
  // _hsv2rgb
  const float v = maxc;
  const float h6 = __fmul_rn(h, 6.f);
  const float fi = floorf(h6);
  const float f = __fsub_rn(h6, fi);
  int i = (int)fi;
  const float p = fminf(fmaxf(__fmul_rn(v, __fsub_rn(1.f, s)), 0.f), 1.f);
  const float q = fminf(fmaxf(__fmul_rn(v, __fsub_rn(1.f, __fmul_rn(s, f))), 0.f), 1.f);
  const float t = fminf(fmaxf(__fmul_rn(v, __fsub_rn(1.f, __fmul_rn(s, __fsub_rn(1.f, f)))), 0.f), 1.f);
  i = ((i % 6) + 6) % 6;
  float ro, go, bo;
  switch (i) {
    case 0: ro = v; go = t; bo = p; break;
    case 1: ro = q; go = v; bo = p; break;
    case 2: ro = p; go = v; bo = t; break;
    case 3: ro = p; go = q; bo = v; break;
    case 4: ro = t; go = p; bo = v; break;
    default: ro = v; go = p; bo = q; break;
  }
  // convert_image_dtype(float32 -> uint8): x.mul(255 + 1 - 1e-3).to(uint8)
  const float k = 255.999f;
  r8 = (float)(uint8_t)__fmul_rn(ro, k);
  g8 = (float)(uint8_t)__fmul_rn(go, k);
  b8 = (float)(uint8_t)__fmul_rn(bo, k);
}

__device__ __forceinline__ void apply_op(int op, const CjParams& P, float& r, float& g, float& b) {
  if (op == 0) {
    const float om = (float)(1.0 - (double)P.b);
    r = blend(r, 0.f, P.b, om);
    g = blend(g, 0.f, P.b, om);
    b = blend(b, 0.f, P.b, om);
  } else if (op == 2) {
    const float om = (float)(1.0 - (double)P.s);
    const float gr = grey_of(r, g, b);
    r = blend(r, gr, P.s, om);
    g = blend(g, gr, P.s, om);
    b = blend(b, gr, P.s, om);
  } else if (op == 3) {
    hue_adjust(r, g, b, P.h);
  }
}

// One block per source cloud.  Pass 1 applies the ops that precede the contrast step and sums the grey values (an
// integer sum: exact), pass 2 blends with the cloud's mean and applies the rest.
__global__ void __launch_bounds__(256)
color_jitter_kernel(const uint8_t* __restrict__ rgb, int N, const float* __restrict__ params_dev, float b_lo, float b_hi,
                    float c_lo, float c_hi, float s_lo, float s_hi, float h_lo, float h_hi, uint64_t seed,
                    const uint64_t* __restrict__ counter_dev, uint32_t stream_id, uint8_t* __restrict__ out) {
  __shared__ CjParams P;
  __shared__ unsigned int s_sum;
  if (threadIdx.x == 0) {
    if (params_dev) {  // injected draws (parity mode): [order0..3, brightness, contrast, saturation, hue]
      for (int i = 0; i < 4; ++i) P.order[i] = (int)params_dev[i];
      P.b = params_dev[4]; P.c = params_dev[5]; P.s = params_dev[6]; P.h = params_dev[7];
    } else {  // one Philox draw per CALL (not per cloud): Fisher-Yates order + the four factors
      const uint64_t cnt = counter_dev ? *counter_dev : 0ull;
      const uint4 a = philox4x32_10(make_uint4(0xFFFFFFF0u, 0u, (uint32_t)cnt, (uint32_t)(cnt >> 32) ^ (stream_id << 24)),
                                    make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
      const uint4 d = philox4x32_10(make_uint4(0xFFFFFFF1u, 0u, (uint32_t)cnt, (uint32_t)(cnt >> 32) ^ (stream_id << 24)),
                                    make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
      int ord[4] = {0, 1, 2, 3};
      const uint32_t rr[3] = {a.x, a.y, a.z};
      for (int i = 3; i > 0; --i) {
        const int j = (int)(rr[3 - i] % (uint32_t)(i + 1));
        const int tmp = ord[i]; ord[i] = ord[j]; ord[j] = tmp;
      }
      for (int i = 0; i < 4; ++i) P.order[i] = ord[i];
      P.b = b_lo + (b_hi - b_lo) * u01(d.x);
      P.c = c_lo + (c_hi - c_lo) * u01(d.y);
      P.s = s_lo + (s_hi - s_lo) * u01(d.z);
      P.h = h_lo + (h_hi - h_lo) * u01(d.w);
    }
    s_sum = 0u;
  }
  __syncthreads();
  const uint8_t* src = rgb + (int64_t)blockIdx.x * 3 * N;
  uint8_t* dst = out + (int64_t)blockIdx.x * 3 * N;
  int cpos = 4;
  for (int i = 0; i < 4; ++i)
    if (P.order[i] == 1) cpos = i;
  unsigned int local = 0;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    float r = src[n], g = src[N + n], b = src[2 * N + n];
    for (int i = 0; i < cpos; ++i) apply_op(P.order[i], P, r, g, b);
    local += (unsigned int)grey_of(r, g, b);
    dst[n] = (uint8_t)r; dst[N + n] = (uint8_t)g; dst[2 * N + n] = (uint8_t)b;
  }
  atomicAdd(&s_sum, local);
  __syncthreads();
  if (cpos == 4) return;
  const float mean = __fdiv_rn((float)s_sum, (float)N);  // torch.mean: fp32 sum of integers (exact below 2^24) / N
  const float om = (float)(1.0 - (double)P.c);
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    float r = dst[n], g = dst[N + n], b = dst[2 * N + n];
    r = blend(r, mean, P.c, om);
    g = blend(g, mean, P.c, om);
    b = blend(b, mean, P.c, om);
    for (int i = cpos + 1; i < 4; ++i) apply_op(P.order[i], P, r, g, b);
    dst[n] = (uint8_t)r; dst[N + n] = (uint8_t)g; dst[2 * N + n] = (uint8_t)b;
  }
}

}  // namespace pcrl

using namespace pcrl;

extern "C" int pcrl_color_jitter_points(const uint8_t* rgb, int B, int N, const float* params_dev, float brightness,
                                        float contrast, float saturation, float hue, uint64_t seed,
                                        const uint64_t* counter_dev, uint32_t stream_id, uint8_t* out, void* stream) {
  PCRL_CHECK_ARG(rgb && out && B >= 0 && N > 0 && (int64_t)N * 255 < (1 << 24));
  PCRL_CHECK_ARG(brightness >= 0.f && contrast >= 0.f && saturation >= 0.f && hue >= 0.f && hue <= 0.5f);
  if (B == 0) return PCRL_OK;
  // torchvision ColorJitter._check_input: factor ranges [max(0, 1 - x), 1 + x]; hue [-h, h]
  color_jitter_kernel<<<B, 256, 0, as_stream(stream)>>>(rgb, N, params_dev, fmaxf(0.f, 1.f - brightness), 1.f + brightness,
                                                       fmaxf(0.f, 1.f - contrast), 1.f + contrast,
                                                       fmaxf(0.f, 1.f - saturation), 1.f + saturation, -hue, hue, seed,
                                                       counter_dev, stream_id, out);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}
