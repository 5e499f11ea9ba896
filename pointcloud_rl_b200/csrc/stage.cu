// Replay-minibatch staging with the DrQ augmentation fused into the load.
// One thread per (source cloud, point): the channel-major source is read once (coalesced along the
// point axis) and `repeat` augmented point-major rows are written.
#include <algorithm>
#include <cuda_bf16.h>

#include "common.cuh"

namespace pcrl {

// byte offset of (point p, channel ch) inside a 128x16 bf16 tile image: K-major core matrices of
// 8 rows x 16 bytes, K-adjacent cores 128 B apart, 8-row groups 256 B apart (the no-swizzle UMMA
// canonical layout the layer-0 MMA descriptor in pointnet_tc.cu describes: LBO=128, SBO=256).
__device__ __forceinline__ int xh_tile_offset(int p, int ch) {
  return (p >> 3) * 256 + (ch >> 3) * 128 + (p & 7) * 16 + (ch & 7) * 2;
}

template <int CP>
__global__ void __launch_bounds__(256) stage_points_kernel(
    const float* __restrict__ xyz, const void* __restrict__ rgb, int rgb_is_u8, const uint8_t* __restrict__ pos,
    int n_pos, const uint8_t* __restrict__ seg, int n_seg, int B, int N, int NP, int repeat, int aug_kind, int axis_mask, float lo,
    float hi, const float* __restrict__ noise, uint64_t seed, const uint64_t* __restrict__ counter_dev,
    uint32_t stream_id, float* __restrict__ xf, __nv_bfloat16* __restrict__ xh) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)B * NP) return;
  const int b = (int)(t / NP), n_out = (int)(t % NP);
  // padding rows (n_out >= N) replicate the cloud's point 0, augmentation draw included: a duplicate can only
  // tie with the real point and the smallest-index rule of the max-pool then drops it (no masking downstream)
  int n = n_out < N ? n_out : 0;
  // RandomDownSample: dropped points are replaced by a kept one (noise = int32 source map [N]); a duplicate can only tie
  // with its original in the max-pool, so the pooled features equal those of the sliced cloud
  if (aug_kind == PCRL_AUG_DOWNSAMPLE) n = reinterpret_cast<const int32_t*>(noise)[n];
  const int C = 3 + (rgb ? 3 : 0) + n_pos + n_seg;
  const uint64_t cnt = counter_dev ? *counter_dev : 0ull;

  float f[CP];
#pragma unroll
  for (int c = 0; c < CP; ++c) f[c] = 0.f;
  float rgb_raw[3] = {0.f, 0.f, 0.f};
  const bool real = true;
  {
    const float* px = xyz + (int64_t)b * 3 * N + n;
    f[0] = px[0];
    f[1] = px[N];
    f[2] = px[2 * (int64_t)N];
    int c = 3;
    if (rgb) {
      if (rgb_is_u8) {
        const uint8_t* pr = reinterpret_cast<const uint8_t*>(rgb) + (int64_t)b * 3 * N + n;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          rgb_raw[k] = (float)pr[(int64_t)k * N];
          f[3 + k] = rgb_raw[k] / 255.0f;  // pointnet.py:57 `rgb / 255.0`
        }
      } else {
        const float* pr = reinterpret_cast<const float*>(rgb) + (int64_t)b * 3 * N + n;
#pragma unroll
        for (int k = 0; k < 3; ++k) rgb_raw[k] = f[3 + k] = pr[(int64_t)k * N];
      }
      c = 6;
    }
    for (int k = 0; k < n_pos; ++k, ++c)
      if (c < CP) f[c] = (float)pos[((int64_t)b * n_pos + k) * N + n];
    for (int k = 0; k < n_seg; ++k, ++c)
      if (c < CP) f[c] = seg[((int64_t)b * n_seg + k) * N + n] ? 1.f : 0.f;
  }
  const float x0 = f[0], y0 = f[1], z0 = f[2];

  for (int a = 0; a < repeat; ++a) {
    const int r = b * repeat + a;  // repeat_interleave layout, drq.py:58 / array_ops.py:121
    float x = x0, y = y0, z = z0;
    if (real && aug_kind == PCRL_AUG_JITTER) {
      float j0, j1, j2;
      if (noise) {
        const float* pn = noise + (int64_t)r * 3 * N + n;
        j0 = pn[0];
        j1 = pn[N];
        j2 = pn[2 * (int64_t)N];
      } else {
        uint4 rnd = philox4x32_10(make_uint4((uint32_t)n, (uint32_t)r, (uint32_t)cnt, (uint32_t)(cnt >> 32) ^ (stream_id << 24)),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
        j0 = lo + (hi - lo) * u01(rnd.x);
        j1 = lo + (hi - lo) * u01(rnd.y);
        j2 = lo + (hi - lo) * u01(rnd.z);
      }
      x += j0;
      y += j1;
      z += j2;
    } else if (real && aug_kind == PCRL_AUG_ROTZ) {
      float ang;
      if (noise) {
        ang = noise[r];
      } else {
        uint4 rnd = philox4x32_10(make_uint4(0xFFFFFFFFu, (uint32_t)r, (uint32_t)cnt, (uint32_t)(cnt >> 32) ^ (stream_id << 24)),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
        ang = lo + (hi - lo) * u01(rnd.x);
      }
      float s, c;
      sincosf(ang, &s, &c);
      // x' = R x with R = [[c,-s,0],[s,c,0],[0,0,1]]  (ops.py:171-183, einsum 'bin,bji->bjn')
      float xr = c * x0 - s * y0, yr = s * x0 + c * y0;
      x = xr;
      y = yr;
    } else if (real && aug_kind == PCRL_AUG_SHIFT) {
      float t0, t1, t2;
      if (noise) {  // [R, 3] per-cloud translation
        t0 = noise[(int64_t)r * 3];
        t1 = noise[(int64_t)r * 3 + 1];
        t2 = noise[(int64_t)r * 3 + 2];
      } else {
        uint4 rnd = philox4x32_10(make_uint4(0xFFFFFFFEu, (uint32_t)r, (uint32_t)cnt, (uint32_t)(cnt >> 32) ^ (stream_id << 24)),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
        t0 = (axis_mask & 1) ? lo + (hi - lo) * u01(rnd.x) : 0.f;
        t1 = (axis_mask & 2) ? lo + (hi - lo) * u01(rnd.y) : 0.f;
        t2 = (axis_mask & 4) ? lo + (hi - lo) * u01(rnd.z) : 0.f;
      }
      x += t0;
      y += t1;
      z += t2;
    }
    f[0] = x;
    f[1] = y;
    f[2] = z;
    if (xf) {
      float4* dst = reinterpret_cast<float4*>(xf + ((int64_t)r * NP + n_out) * CP);
#pragma unroll
      for (int q = 0; q < CP / 4; ++q) dst[q] = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
    }

    if (xh) {
      // bf16 operand row for the layer-0 MMA (K=16): channels [0,C) hi parts (u8 rgb kept as the
      // exact integer, 1/255 folded into the packed weights), channel C = 1 (bias row), channels
      // C+1..C+3 = lo parts of xyz so coordinates keep ~16 mantissa bits through the bf16 MMA.
      __nv_bfloat16 h[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) h[c] = __float2bfloat16(0.f);
      if (real) {
#pragma unroll
        for (int c = 0; c < CP; ++c)
          if (c < C) h[c] = __float2bfloat16(f[c]);
        if (rgb && rgb_is_u8) {
#pragma unroll
          for (int k = 0; k < 3; ++k) h[3 + k] = __float2bfloat16(rgb_raw[k]);
        }
        h[C] = __float2bfloat16(1.f);
#pragma unroll
        for (int k = 0; k < 3; ++k) h[C + 1 + k] = __float2bfloat16(f[k] - __bfloat162float(h[k]));
      }
      const int64_t tile = ((int64_t)r * NP + n_out) >> 7;
      char* base = reinterpret_cast<char*>(xh) + tile * 4096;
      const int p = n_out & 127;
      *reinterpret_cast<uint4*>(base + xh_tile_offset(p, 0)) = *reinterpret_cast<uint4*>(&h[0]);
      *reinterpret_cast<uint4*>(base + xh_tile_offset(p, 8)) = *reinterpret_cast<uint4*>(&h[8]);
    }
  }
}

// RandomDownSample's draw on the device (pcd_aug.py:240-257; array_ops.py:659-673): one random subset of the points,
// shared by every cloud of the call.  n_drop = int(N*ratio) (fixed_ratio) or uniform in [0, int(N*ratio)); the kept
// points are the N - n_drop smallest of N random keys (the reference argsorts torch.rand).  Output: src[i] = i for a
// kept point, else the kept point of rank 0.  One block; N <= 4096 (keys in shared memory, O(N^2) ranking).
__global__ void __launch_bounds__(1024) downsample_map_kernel(int N, float ratio, int fixed_ratio, uint64_t seed,
                                                              const uint64_t* __restrict__ counter_dev, uint32_t stream_id,
                                                              int32_t* __restrict__ src) {
  __shared__ uint32_t key[4096];
  __shared__ int first;
  const uint64_t cnt = counter_dev ? *counter_dev : 0ull;
  const uint2 k2 = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  for (int i = threadIdx.x; i < N; i += blockDim.x)
    key[i] = philox4x32_10(make_uint4((uint32_t)i, 0xFFFFFFFDu, (uint32_t)cnt, (uint32_t)(cnt >> 32) ^ (stream_id << 24)), k2).x;
  const int hi = (int)((float)N * ratio);
  int n_drop = hi;
  if (!fixed_ratio) {
    const uint32_t r = philox4x32_10(make_uint4(0xFFFFFFFFu, 0xFFFFFFFDu, (uint32_t)cnt, (uint32_t)(cnt >> 32) ^ (stream_id << 24)), k2).y;
    n_drop = hi > 0 ? (int)(((uint64_t)r * (uint64_t)hi) >> 32) : 0;
  }
  const int n_keep = max(N - n_drop, 1);
  __syncthreads();
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const uint32_t ki = key[i];
    int rank = 0;
    for (int j = 0; j < N; ++j) rank += (key[j] < ki) || (key[j] == ki && j < i);
    if (rank == 0) first = i;
    src[i] = rank < n_keep ? i : -1;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < N; i += blockDim.x)
    if (src[i] < 0) src[i] = first;
}

// Replay sampling on the device: dst_leaf[b] = src_leaf[idx[b]] for every leaf of a transition, one launch.
// blockIdx.y = leaf, blockIdx.x = sampled row; rows are copied as 16-byte vectors when both sides allow it.
__global__ void __launch_bounds__(256) gather_transitions_kernel(const unsigned long long* __restrict__ src_ptrs,
                                                                 const unsigned long long* __restrict__ dst_ptrs,
                                                                 const long long* __restrict__ row_bytes,
                                                                 const long long* __restrict__ idx, int B) {
  const int leaf = blockIdx.y;
  const long long nb = row_bytes[leaf];
  const char* src = reinterpret_cast<const char*>(src_ptrs[leaf]);
  char* dst = reinterpret_cast<char*>(dst_ptrs[leaf]);
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    const char* s = src + idx[b] * nb;
    char* d = dst + (long long)b * nb;
    if (((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(d) | (uintptr_t)nb) & 15) == 0) {
      for (long long o = (long long)threadIdx.x * 16; o < nb; o += 256 * 16)
        *reinterpret_cast<uint4*>(d + o) = *reinterpret_cast<const uint4*>(s + o);
    } else {
      for (long long o = threadIdx.x; o < nb; o += 256) d[o] = s[o];
    }
  }
}

}  // namespace pcrl

using namespace pcrl;

extern "C" int pcrl_downsample_map(int N, float drop_ratio, int fixed_ratio, uint64_t seed, const uint64_t* counter_dev,
                                   uint32_t stream_id, int32_t* src_map, void* stream) {
  PCRL_CHECK_ARG(src_map && N >= 1 && drop_ratio >= 0.f && drop_ratio < 1.f);
  if (N > 4096) {
    set_error("pcrl_downsample_map: N = %d > 4096 points is not supported by the device-side draw", N);
    return PCRL_EUNSUPPORTED;
  }
  downsample_map_kernel<<<1, 1024, 0, as_stream(stream)>>>(N, drop_ratio, fixed_ratio, seed, counter_dev, stream_id, src_map);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

extern "C" int pcrl_gather_transitions(const uint64_t* src_ptrs, const uint64_t* dst_ptrs, const int64_t* row_bytes,
                                       int n_leaves, const int64_t* idx, int B, void* stream) {
  PCRL_CHECK_ARG(src_ptrs && dst_ptrs && row_bytes && idx && n_leaves >= 1 && B >= 0);
  if (B == 0) return PCRL_OK;
  dim3 grid((unsigned)std::min(B, 4096), (unsigned)n_leaves);
  gather_transitions_kernel<<<grid, 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const unsigned long long*>(src_ptrs), reinterpret_cast<const unsigned long long*>(dst_ptrs),
      reinterpret_cast<const long long*>(row_bytes), reinterpret_cast<const long long*>(idx), B);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

extern "C" int pcrl_stage_points(const float* xyz, const void* rgb, int rgb_is_u8, const uint8_t* pos, int n_pos,
                                 const uint8_t* seg, int n_seg, int B, int N, int repeat, int aug_kind, float aug_lo,
                                 float aug_hi, const float* noise, uint64_t seed, const uint64_t* counter_dev,
                                 uint32_t stream_id, float* xf, void* xh, int CP, void* stream) {
  PCRL_CHECK_ARG(xyz && (xf || xh) && B >= 0 && N > 0 && repeat >= 1);
  int axis_mask = (aug_kind >> 8) & 7;
  aug_kind &= 0xff;
  if (axis_mask == 0) axis_mask = 7;
  PCRL_CHECK_ARG(aug_kind == PCRL_AUG_NONE || aug_kind == PCRL_AUG_JITTER || aug_kind == PCRL_AUG_ROTZ ||
                 aug_kind == PCRL_AUG_SHIFT || aug_kind == PCRL_AUG_DOWNSAMPLE);
  PCRL_CHECK_ARG(aug_kind != PCRL_AUG_DOWNSAMPLE || noise != nullptr);
  const int C = 3 + (rgb ? 3 : 0) + (pos ? n_pos : 0) + (seg ? n_seg : 0);
  PCRL_CHECK_ARG((CP == 8 || CP == 16) && C <= CP);
  PCRL_CHECK_ARG(!xh || C + 4 <= 16);
  if (!pos) n_pos = 0;
  if (!seg) n_seg = 0;
  if (B == 0) return PCRL_OK;
  const int NP = (int)align_up(N, 128);
  const int64_t threads = (int64_t)B * NP;
  const unsigned blocks = (unsigned)cdiv(threads, 256);
  cudaStream_t st = as_stream(stream);
  if (CP == 8)
    stage_points_kernel<8><<<blocks, 256, 0, st>>>(xyz, rgb, rgb_is_u8, pos, n_pos, seg, n_seg, B, N, NP, repeat,
                                                   aug_kind, axis_mask, aug_lo, aug_hi, noise, seed, counter_dev, stream_id, xf,
                                                   reinterpret_cast<__nv_bfloat16*>(xh));
  else
    stage_points_kernel<16><<<blocks, 256, 0, st>>>(xyz, rgb, rgb_is_u8, pos, n_pos, seg, n_seg, B, N, NP, repeat,
                                                    aug_kind, axis_mask, aug_lo, aug_hi, noise, seed, counter_dev, stream_id, xf,
                                                    reinterpret_cast<__nv_bfloat16*>(xh));
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}
