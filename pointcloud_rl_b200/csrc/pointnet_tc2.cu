// Fused PointNet per-point MLP + max-pool on tcgen05 / TMEM, second generation (sm_100a).
//
// What bounded the first kernel (pointnet_tc.cu) was not the tensor pipe but the layer-2 epilogue: with points on the
// TMEM lanes, LayerNorm is thread-local but the max over points is the CROSS-lane direction, i.e. a transpose of every
// accumulator element through shared memory (256 KB of shared-memory traffic and ~3 instructions per element).
// Here layer 2 is issued TRANSPOSED,  D[channel, point] = W2 . h1^T  (channels on the TMEM lanes, points on the
// columns), so the max over points is a thread-local running maximum over the accumulator row -- no shuffles, no
// shared memory -- and LayerNorm's statistics, which would now be the cross-lane direction, never touch the
// accumulator at all:
//   * mean:      W1 and W2 are packed CENTRED over their output channels (W[c,:] - mean_c W[c,:]), so every
//                accumulator already holds y - mean(y);
//   * variance:  sum_c (y_c - mean)^2 = h1^T Gc h1 with the Gram matrix Gc = W2c^T W2c (c2 x c2, packed once per
//                weight update): one extra N = c2 MMA per tile, U = h1 . Gc, and a thread-local dot(h1, U) in the
//                warp that produced h1 (points on lanes there).
// The layer-2 epilogue is then  z = y * rstd[point]  (sign(gamma2) folded into the packed weights as before, so
// |gamma2|, beta2 and the ReLU commute with the max and run once per (cloud, channel) in the finalize kernel) and a
// 3-input max: ~1.25 instructions per accumulator element instead of ~4.5, and no shared-memory traffic.
//
// Per 128-point tile (c = (c1, c2, c3)):
//   acc0 = X W0'^T                 (M = points, K = 16)      -> relu            -> h0 (bf16, smem)
//   acc1 = h0 W1c^T                (M = points, N = c2)      -> LN + relu       -> h1 (bf16, smem; packed copy in regs)
//   U    = h1 Gc                   (M = points, N = c2)      -> rstd2 = rsqrt(dot(h1, U) / c3 + eps) -> smem
//   D_b  = W2c'[128 b .. ] h1^T    (M = channels, N = points) for b < c3 / 128  -> * rstd2[point] -> running max
//
// Persistent, one CTA per SM, 16 warps:
//   warp 0      bulk-TMA producer (weights once; 4 KB point tiles through a 3-stage ring)
//   warp 1, 2   MMA issuers of tile slots 0 / 1 (layers 0, 1 and the Gram MMA; odd and even tiles ping-pong)
//   warp 3      MMA issuer of the transposed layer 2
//   warps 4-7   front group of slot 0 (one warp per TMEM lane quadrant, a thread owns a whole point row)
//   warps 8-11  front group of slot 1
//   warps 12-15 pool group (a thread owns one channel of every 128-channel block)
// TMEM (512 columns): [0,128) slot 0 and [128,256) slot 1 (acc0 / acc1 / U alias: each is drained before the next MMA
// overwrites it), [256,384) and [384,512) the two transposed layer-2 accumulators (one per 128-channel block, ring).
#include <algorithm>
#include <cuda_bf16.h>

#include "common.cuh"
#include "pointnet_tc2.cuh"
#include "tc_ptx.cuh"

namespace pcrl {
namespace tc2 {
using namespace pcrl::tc;

constexpr int kStages = 3;
constexpr int kTileBytes = 128 * 16 * 2;  // one X tile: 128 points x 16 bf16
constexpr int kThreads = 512;

// profiling knobs (tools/probe_fwd2.py): knock out individual passes to attribute time.  0 in production.
//   1 pool math   2 Gram dot math   4 layer-1 normalise math   8 front-group TMEM loads   16 pool TMEM loads
__device__ int g_dbg2 = 0;
// clock64 trace of CTA 0 (flag 256): six roles x 512 (event, clock) pairs.  Compiled in only with -DPCRL_FWD_TRACE.
__device__ long long g_trace2[6 * 1024];
struct Tracer2 {
  int region, n;
  bool on;
  __device__ __forceinline__ void operator()(int ev) {
#ifdef PCRL_FWD_TRACE
    if (on && n < 512) {
      g_trace2[region * 1024 + 2 * n] = ev;
      g_trace2[region * 1024 + 2 * n + 1] = clock64();
      ++n;
    }
#endif
  }
};

// Packed weight buffer, second generation (appended to the first generation's buffer, which the backward's dump mode
// still reads):  [ W0' | W1c | Gc | W2c' | g1 be1 ]  <- one contiguous shared-memory image;  [ g2 be2 ] for the finalize
struct Wpack2 {
  uint32_t w0, w1, gc, w2, prm1, img_bytes, prm2, cm, total;
};
__host__ __device__ inline Wpack2 make_wpack2(int c1, int c2, int c3) {
  Wpack2 W;
  uint32_t o = 0;
  W.w0 = o;   o += c1 * 32;
  W.w1 = o;   o += c2 * c1 * 2;
  W.gc = o;   o += c2 * c2 * 2;
  W.w2 = o;   o += c3 * c2 * 2;
  W.prm1 = o; o += 2 * c2 * 4;
  W.img_bytes = o;
  W.prm2 = o; o += 2 * c3 * 4;
  W.cm = o;   o += (c1 + c2) * 4;  // column means of W1 and W2 (scratch of the packer)
  W.total = (o + 127) & ~127u;
  return W;
}

struct Smem2 {
  uint32_t img, xst, act, rbuf, bars, total;
};
__host__ __device__ inline Smem2 make_smem2(int c1, int c2, int c3) {
  Smem2 L;
  uint32_t o = 0;
  L.img = o;  o += make_wpack2(c1, c2, c3).img_bytes;
  o = (o + 127) & ~127u;
  L.xst = o;  o += kStages * kTileBytes;
  L.act = o;  o += 2 * 128 * (c1 > c2 ? c1 : c2) * 2;  // one h0/h1 buffer per tile slot
  L.rbuf = o; o += 4 * 128 * 4;                        // rstd2 of the last four tiles
  L.bars = o; o += 256;
  L.total = o;
  return L;
}

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

struct NoOp {
  __device__ __forceinline__ void operator()() const {}
};
// Walks NCH consecutive 32-column chunks of this warp's TMEM lanes with the loads software-pipelined over two register
// buffers: chunk c+1 is in flight while f(chunk c, first column) runs.  last_loaded() runs as soon as the final
// chunk's data is in registers (before its f).
template <int NCH, class F, class G = NoOp>
__device__ __forceinline__ void tmem_chunks(uint32_t taddr, bool skip_loads, F&& f, G&& last_loaded = NoOp()) {
  uint32_t va[32], vb[32];
  if (skip_loads) {  // profiling knock-out: no TMEM traffic, registers carry junk
#pragma unroll
    for (int i = 0; i < 32; ++i) va[i] = vb[i] = 0x3f800000u;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      if (c == NCH - 1) last_loaded();
      f(va, c * 32);
    }
    return;
  }
  tmem_ld32_async(taddr, va);
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    if (c & 1) {
      tmem_wait_ld_dep(vb);
      if (c + 1 < NCH) tmem_ld32_async(taddr + 32 * (c + 1), va);
      if (c == NCH - 1) last_loaded();
      f(vb, c * 32);
    } else {
      tmem_wait_ld_dep(va);
      if (c + 1 < NCH) tmem_ld32_async(taddr + 32 * (c + 1), vb);
      if (c == NCH - 1) last_loaded();
      f(va, c * 32);
    }
  }
}

// single register buffer (for the passes that keep a row of packed activations live beside the chunk)
template <int NCH, class F>
__device__ __forceinline__ void tmem_chunks1(uint32_t taddr, bool skip_loads, F&& f) {
  uint32_t v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = 0x3f800000u;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    if (!skip_loads) tmem_ld32(taddr + 32 * c, v);
    f(v, c * 32);
  }
}

template <int C1, int C2, int NBLK, bool ARGMAX>
__global__ void __launch_bounds__(kThreads, 1)
pointnet_fwd_tc2_kernel(const char* __restrict__ xh, const char* __restrict__ wpack, int n_tiles, int tiles_per_cloud,
                        int src_cloud_stride, float ln_eps, float key_bias, unsigned long long* __restrict__ pool_keys,
                        int c3_total, int ctas_per_group) {
  // Channel groups (wide last layers, c3_total > NBLK * 128): the layer-2 weights no longer fit in shared memory, so
  // the CTAs split into c3_total / (NBLK * 128) groups; a group keeps ITS NBLK blocks of W2c' resident, walks all the
  // tiles (recomputing layers 0 / 1 and the Gram product, whose variance covers all c3_total channels) and pools its
  // own channels.  Config 2 (c3 = 256) is one group.
  constexpr int C3G = NBLK * 128;                       // channels this CTA pools
  const int group = (int)blockIdx.x / ctas_per_group;   // which NBLK-block slice of W2c'
  const int cta_in_group = (int)blockIdx.x % ctas_per_group;
  extern __shared__ __align__(128) unsigned char smem[];
  const Smem2 L = make_smem2(C1, C2, C3G);
  const Wpack2 W = make_wpack2(C1, C2, C3G);            // shared-memory image (one group's W2c' slice)
  const Wpack2 WG = make_wpack2(C1, C2, c3_total);      // global image (all of W2c')
  const uint32_t sbase = smem_u32(smem);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;  // provably warp-uniform
  constexpr uint32_t kActBytes = 128 * (C1 > C2 ? C1 : C2) * 2;

  // mbarriers
  constexpr int WB = 0;             // weights landed
  constexpr int XF = 1;             // [kStages] x tile landed
  constexpr int XE = XF + kStages;  // [kStages] x stage free
  constexpr int F0 = XE + kStages;  // [slot] acc0 complete (tcgen05.commit)
  constexpr int F1 = F0 + 2;        // [slot] acc1 complete
  constexpr int FU = F1 + 2;        // [slot] U complete
  constexpr int E0 = FU + 2;        // [slot] h0 written, acc0 drained (128 arrivals)
  constexpr int E1 = E0 + 2;        // [slot] h1 written, acc1 drained
  constexpr int EU = E1 + 2;        // [slot] U drained, rstd2 written
  constexpr int HF = EU + 2;        // [slot] the transposed layer-2 MMAs have finished reading the slot's h1
  constexpr int F2 = HF + 2;        // [ring] transposed layer-2 accumulator complete
  constexpr int D2 = F2 + 2;        // [ring] ... drained by the pool group (128 arrivals)
  constexpr int UI = D2 + 2;        // [slot] the Gram MMA has been ISSUED (the transposed layer 2 queues behind it)
  constexpr int W1B = UI + 2;       // W1c + beta1' landed   } the weight image arrives in four pieces so that the first
  constexpr int WGB = W1B + 1;      // Gc landed             } tiles' layer 0 / 1 run while the 96 KB of Gc and W2c'
  constexpr int W2B = WGB + 1;      // W2c' slice landed     } are still on their way (WB: W0' only)
  constexpr int NBAR = W2B + 1;
  const uint32_t bar0 = sbase + L.bars;
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + L.bars + NBAR * 8);

  if (threadIdx.x == 0) {
    mbar_init(BAR(WB), 1);
    mbar_init(BAR(W1B), 1);
    mbar_init(BAR(WGB), 1);
    mbar_init(BAR(W2B), 1);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(BAR(XF + s), 1);
      mbar_init(BAR(XE + s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(BAR(F0 + s), 1);
      mbar_init(BAR(F1 + s), 1);
      mbar_init(BAR(FU + s), 1);
      mbar_init(BAR(E0 + s), 128);
      mbar_init(BAR(E1 + s), 128);
      mbar_init(BAR(EU + s), 128);
      mbar_init(BAR(HF + s), 1);
      mbar_init(BAR(F2 + s), 1);
      mbar_init(BAR(D2 + s), 128);
      mbar_init(BAR(UI + s), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
        smem_u32((const void*)tmem_ptr_smem)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
#if defined(PCRL_FWD_KNOCKOUT)
  const int dbg = g_dbg2;  // knock-outs put branches into the inner loops (no load hoisting across them): profiling builds only
#elif defined(PCRL_FWD_TRACE)
  const int dbg = g_dbg2 & 256;  // tracer only
#else
  constexpr int dbg = 0;
#endif

  // every CTA takes a contiguous range of tiles (consecutive tiles belong to the same cloud)
  const int t_q = n_tiles / ctas_per_group, t_r = n_tiles % ctas_per_group;
  const int n_local = t_q + (cta_in_group < t_r ? 1 : 0);
  const int64_t tile0 = (int64_t)cta_in_group * t_q + min(cta_in_group, t_r);
  const uint32_t s_w0 = sbase + L.img + W.w0, s_w1 = sbase + L.img + W.w1, s_gc = sbase + L.img + W.gc,
                 s_w2 = sbase + L.img + W.w2;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      // [W0' | W1c | Gc] is contiguous in both images; then this group's slice of W2c'; then beta1'
      auto load = [&](uint32_t dst_off, uint32_t src_off, uint32_t bytes, int bar) {
        mbar_expect_tx(BAR(bar), bytes);
        for (uint32_t off = 0; off < bytes; off += 32768u)
          bulk_g2s(sbase + L.img + dst_off + off, wpack + src_off + off, min(bytes - off, 32768u), BAR(bar));
      };
      const uint32_t w2_bytes = W.prm1 - W.w2;
      load(W.w0, WG.w0, W.w1 - W.w0, WB);
      mbar_expect_tx(BAR(W1B), (W.gc - W.w1) + (W.img_bytes - W.prm1));
      for (uint32_t off = 0; off < W.gc - W.w1; off += 32768u)
        bulk_g2s(sbase + L.img + W.w1 + off, wpack + WG.w1 + off, min(W.gc - W.w1 - off, 32768u), BAR(W1B));
      bulk_g2s(sbase + L.img + W.prm1, wpack + WG.prm1, W.img_bytes - W.prm1, BAR(W1B));
      // the first point tiles go out before the big weight pieces
      const int n_first = min(n_local, kStages);
      for (int i = 0; i < n_first; ++i) {
        mbar_expect_tx(BAR(XF + i), kTileBytes);
        const int64_t tile = tile0 + i;
        const int64_t src_tile = src_cloud_stride == 1 ? tile
                                                        : (tile / tiles_per_cloud) * src_cloud_stride * tiles_per_cloud + tile % tiles_per_cloud;
        bulk_g2s(sbase + L.xst + i * kTileBytes, xh + src_tile * kTileBytes, kTileBytes, BAR(XF + i));
      }
      load(W.gc, WG.gc, W.w2 - W.gc, WGB);
      load(W.w2, WG.w2 + (uint32_t)group * w2_bytes, w2_bytes, W2B);
      for (int i = n_first; i < n_local; ++i) {
        const int st = i % kStages;
        if (i >= kStages) mbar_wait_relaxed(BAR(XE + st), ((i / kStages) - 1) & 1, 64);
        mbar_expect_tx(BAR(XF + st), kTileBytes);
        const int64_t tile = tile0 + i;
        // cloud r reads the tiles of source cloud r * src_cloud_stride (the actor step encodes the first of the
        // num_aug staged copies of every sample in place)
        const int64_t src_tile = src_cloud_stride == 1 ? tile
                                                        : (tile / tiles_per_cloud) * src_cloud_stride * tiles_per_cloud + tile % tiles_per_cloud;
        bulk_g2s(sbase + L.xst + st * kTileBytes, xh + src_tile * kTileBytes, kTileBytes, BAR(XF + st));
      }
    }
  } else if (warp == 1 || warp == 2) {
    // ------------------------------------------------------------------ MMA issuer of one tile slot: layer 0, layer 1, Gram
    // (warp-uniform control flow, one elected lane issues: see elect_one())
    const int s = __shfl_sync(0xffffffffu, warp - 1, 0);
    const uint32_t tm = tmem_base + (uint32_t)(s * 128);
    const uint32_t act = sbase + L.act + (uint32_t)s * kActBytes;
    const uint32_t id0 = make_idesc(C1), id1 = make_idesc(C2);
    const uint64_t d_w0 = make_desc(s_w0, 256), d_w1 = make_desc(s_w1, C1 * 16), d_gc = make_desc(s_gc, C2 * 16);
    const uint64_t d_h0 = make_desc(act, C1 * 16), d_h1 = make_desc(act, C2 * 16);
    uint64_t d_x[kStages];
#pragma unroll
    for (int st = 0; st < kStages; ++st) d_x[st] = make_desc(sbase + L.xst + st * kTileBytes, 256);
    mbar_wait(BAR(WB), 0);
    Tracer2 tr{s, 0, (dbg & 256) && blockIdx.x == 0 && lane == 0};
    for (int j = s; j < n_local; j += 2) {
      const int n = j >> 1, st = j % kStages;
      mbar_wait(BAR(XF + st), (j / kStages) & 1);
      if (n > 0) mbar_wait(BAR(EU + s), (n - 1) & 1);  // the slot's previous tile has drained U
      tr(1000 * j + 100);
      tc_fence_after();
      if (elect_one()) {
        mma_bf16(tm, st == 0 ? d_x[0] : (st == 1 ? d_x[1] : d_x[2]), d_w0, id0, 0);
        mma_commit(BAR(XE + st));
        mma_commit(BAR(F0 + s));
      }
      __syncwarp();
      mbar_wait(BAR(E0 + s), n & 1);  // h0 in shared memory, acc0 drained
      if (n == 0) mbar_wait(BAR(W1B), 0);
      tr(1000 * j + 110);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < C1 / 16; ++ks) mma_bf16(tm, desc_kstep(d_h0, ks), desc_kstep(d_w1, ks), id1, ks > 0);
        mma_commit(BAR(F1 + s));
      }
      __syncwarp();
      mbar_wait(BAR(E1 + s), n & 1);  // h1 in shared memory, acc1 drained
      if (n == 0) mbar_wait(BAR(WGB), 0);
      tr(1000 * j + 120);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < C2 / 16; ++ks) mma_bf16(tm, desc_kstep(d_h1, ks), desc_kstep(d_gc, ks), id1, ks > 0);
        mma_commit(BAR(FU + s));
        mbar_arrive(BAR(UI + s));
      }
      __syncwarp();
      tr(1000 * j + 220);
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------------ MMA issuer of the transposed layer 2
    const uint32_t idT = make_idesc(128);  // N = the tile's 128 points
    const uint64_t d_w2 = make_desc(s_w2, C2 * 16);
    const uint64_t d_a0 = make_desc(sbase + L.act, C2 * 16), d_a1 = make_desc(sbase + L.act + kActBytes, C2 * 16);
    mbar_wait(BAR(W2B), 0);
    Tracer2 tr{2, 0, (dbg & 256) && blockIdx.x == 0 && lane == 0};
    for (int j = 0; j < n_local; ++j) {
      const int s = j & 1, n = j >> 1;
      const uint64_t d_h1 = s ? d_a1 : d_a0;
      // h1 of tile j is in shared memory AND its Gram MMA is already in the tensor pipe's queue: the front group's
      // chain (U -> rstd2 -> next tile of the slot) is the latency-critical one, the pool group only needs throughput
      mbar_wait(BAR(UI + s), n & 1);
#pragma unroll
      for (int blk = 0; blk < NBLK; ++blk) {
        const int g = j * NBLK + blk, r = g & 1, k = g >> 1;
        if (k > 0) mbar_wait(BAR(D2 + r), (k - 1) & 1);  // the pool group drained the ring slot's previous block
        tr(1000 * j + 130 + blk);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d = tmem_base + 256u + (uint32_t)(r * 128);
          const uint64_t wa = d_w2 + (uint64_t)(blk * 16 * C2);  // rows 128 blk .. of the W2c' image: 16 groups x C2*16 bytes, >> 4
#pragma unroll
          for (int ks = 0; ks < C2 / 16; ++ks) mma_bf16(d, desc_kstep(wa, ks), desc_kstep(d_h1, ks), idT, ks > 0);
          mma_commit(BAR(F2 + r));
          if (blk == NBLK - 1) mma_commit(BAR(HF + s));  // every MMA that reads the slot's h1 buffer has completed when this fires
        }
        __syncwarp();
      }
    }
  } else if (warp < 12) {
    // ------------------------------------------------------------------ front groups: layer-0 / layer-1 epilogues + variance of layer 2
    const int s = (warp - 4) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t taddr = tmem_base + (uint32_t)(s * 128) + ((uint32_t)(q * 32) << 16);
    const float* be1 = reinterpret_cast<const float*>(smem + L.img + W.prm1);  // beta1 / |gamma1|
    unsigned char* abuf = smem + L.act + (uint32_t)s * kActBytes;
    unsigned char* dst0 = abuf + (row >> 3) * (uint32_t)(C1 * 16) + (row & 7) * 16;
    unsigned char* dst1 = abuf + (row >> 3) * (uint32_t)(C2 * 16) + (row & 7) * 16;
    float* rbuf = reinterpret_cast<float*>(smem + L.rbuf);
    const float inv_c3 = 1.0f / (float)c3_total;
    const bool no_ld = dbg & 8;
    mbar_wait(BAR(W1B), 0);  // LN parameters landed
    Tracer2 tr{3 + s, 0, (dbg & 256) && blockIdx.x == 0 && lane == 0 && q == 0};
    for (int j = s; j < n_local; j += 2) {
      const int n = j >> 1;
      // ---- layer 0: ReLU -> bf16 operand of layer 1
      mbar_wait(BAR(F0 + s), n & 1);
      tr(1000 * j + 300);
      if (n > 0) mbar_wait(BAR(HF + s), (n - 1) & 1);  // the previous tile's transposed layer 2 no longer reads the buffer
      tr(1000 * j + 301);
      tc_fence_after();
      tmem_chunks<C1 / 32>(taddr, no_ld, [&](uint32_t(&v)[32], int ch) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          uint4 o;
          o.x = pack_relu_bf16x2(__uint_as_float(v[8 * jj + 0]), __uint_as_float(v[8 * jj + 1]));
          o.y = pack_relu_bf16x2(__uint_as_float(v[8 * jj + 2]), __uint_as_float(v[8 * jj + 3]));
          o.z = pack_relu_bf16x2(__uint_as_float(v[8 * jj + 4]), __uint_as_float(v[8 * jj + 5]));
          o.w = pack_relu_bf16x2(__uint_as_float(v[8 * jj + 6]), __uint_as_float(v[8 * jj + 7]));
          *reinterpret_cast<uint4*>(dst0 + ((ch >> 3) + jj) * 128) = o;
        }
      });
      fence_async_smem();
      tc_fence_before();
      mbar_arrive(BAR(E0 + s));
      tr(1000 * j + 310);

      // ---- layer 1: the accumulator holds y - mean(y) (centred weights): variance, normalise, affine, ReLU
      mbar_wait(BAR(F1 + s), n & 1);
      tr(1000 * j + 320);
      tc_fence_after();
      uint64_t q01 = pk2f(0.f, 0.f), q23 = q01;
      tmem_chunks<C2 / 32>(taddr, no_ld, [&](uint32_t(&v)[32], int) {
        if (dbg & 4) return;
#pragma unroll
        for (int jj = 0; jj < 32; jj += 4) {
          const uint64_t y01 = pk2(v[jj], v[jj + 1]), y23 = pk2(v[jj + 2], v[jj + 3]);
          q01 = fma2(y01, y01, q01);
          q23 = fma2(y23, y23, q23);
        }
      });
      float rstd1;
      {
        uint32_t a, b, c, d;
        unpk2(q01, a, b);
        unpk2(q23, c, d);
        const float sq = (__uint_as_float(a) + __uint_as_float(b)) + (__uint_as_float(c) + __uint_as_float(d));
        rstd1 = rsqrtf(sq * (1.0f / (float)C2) + ln_eps);
      }
      const uint64_t r2 = pk2f(rstd1, rstd1);
      uint32_t hp[C2 / 2];  // this row's h1 as packed bf16 pairs: the Gram dot below needs it again (registers, not a
                            // shared-memory read-back: the shared-memory pipe is what this kernel runs out of)
      tmem_chunks1<C2 / 32>(taddr, no_ld, [&](uint32_t(&v)[32], int ch) {
#pragma unroll
        for (int j4 = 0; j4 < 32; j4 += 4) {
          if (dbg & 4) {
            hp[(ch + j4) / 2] = v[j4];
            hp[(ch + j4) / 2 + 1] = v[j4 + 2];
            continue;
          }
          // gamma1 is folded away by the packer (sign into W1c's rows, |gamma1| into W2c's columns and the Gram matrix):
          // h1' = relu(xhat' + beta1 / |gamma1|) -- one FMA and one broadcast shared-memory load per 4 channels less
          const float4 bb = *reinterpret_cast<const float4*>(be1 + ch + j4);
          uint32_t x0, x1, x2, x3;
          unpk2(fma2(pk2(v[j4], v[j4 + 1]), r2, pk2f(bb.x, bb.y)), x0, x1);
          unpk2(fma2(pk2(v[j4 + 2], v[j4 + 3]), r2, pk2f(bb.z, bb.w)), x2, x3);
          hp[(ch + j4) / 2] = pack_relu_bf16x2(__uint_as_float(x0), __uint_as_float(x1));
          hp[(ch + j4) / 2 + 1] = pack_relu_bf16x2(__uint_as_float(x2), __uint_as_float(x3));
        }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
          *reinterpret_cast<uint4*>(dst1 + ((ch >> 3) + jj) * 128) =
              make_uint4(hp[ch / 2 + 4 * jj], hp[ch / 2 + 4 * jj + 1], hp[ch / 2 + 4 * jj + 2], hp[ch / 2 + 4 * jj + 3]);
      });
      fence_async_smem();
      tc_fence_before();
      mbar_arrive(BAR(E1 + s));
      tr(1000 * j + 330);

      // ---- variance of layer 2: sum_c (y2_c - mean)^2 = h1 . (Gc h1) = dot(h1, U)
      mbar_wait(BAR(FU + s), n & 1);
      tr(1000 * j + 340);
      tc_fence_after();
      uint64_t d01 = pk2f(0.f, 0.f), d23 = d01;
      tmem_chunks1<C2 / 32>(taddr, no_ld, [&](uint32_t(&v)[32], int ch) {
        if (dbg & 2) return;
#pragma unroll
        for (int jj = 0; jj < 32; jj += 4) {
          const uint32_t p0 = hp[(ch + jj) / 2], p1 = hp[(ch + jj) / 2 + 1];
          d01 = fma2(pk2(v[jj], v[jj + 1]), pk2(p0 << 16, p0 & 0xffff0000u), d01);
          d23 = fma2(pk2(v[jj + 2], v[jj + 3]), pk2(p1 << 16, p1 & 0xffff0000u), d23);
        }
      });
      {
        uint32_t a, b, c, d;
        unpk2(d01, a, b);
        unpk2(d23, c, d);
        const float ss = (__uint_as_float(a) + __uint_as_float(b)) + (__uint_as_float(c) + __uint_as_float(d));
        rbuf[(j & 3) * 128 + row] = rsqrtf(fmaxf(ss, 0.f) * inv_c3 + ln_eps);
      }
      tc_fence_before();
      mbar_arrive(BAR(EU + s));
      tr(1000 * j + 350);
    }
  } else {
    // ------------------------------------------------------------------ pool group: scale by rstd2[point], running max over points
    // thread (q, lane) owns channel 128 b + 32 q + lane of every block b; TMEM lane = channel, column = point
    const int q = warp & 3;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const float* rbuf = reinterpret_cast<const float*>(smem + L.rbuf);
    float m[NBLK];                   // values-only: running maximum over the cloud's points
    unsigned long long run[NBLK];    // argmax: running (key << 32 | ~index)
#pragma unroll
    for (int b = 0; b < NBLK; ++b) {
      m[b] = -3.0e38f;
      run[b] = 0ull;
    }
    int cur_cloud = n_local > 0 ? (int)(tile0 / tiles_per_cloud) : 0;
    auto flush = [&](int cloud) {
#pragma unroll
      for (int b = 0; b < NBLK; ++b) {
        unsigned long long key;
        if (ARGMAX) {
          key = run[b];
        } else {
          key = (unsigned long long)(__float_as_uint(m[b] + key_bias) & 0xffffff80u) << 32;  // same bits as the argmax variant
        }
        atomicMax(pool_keys + (int64_t)cloud * c3_total + group * C3G + b * 128 + q * 32 + lane, key);
        m[b] = -3.0e38f;
        run[b] = 0ull;
      }
    };
    const uint64_t bias2 = pk2f(key_bias, key_bias);
    Tracer2 tr{5, 0, (dbg & 256) && blockIdx.x == 0 && lane == 0 && q == 0};
    for (int j = 0; j < n_local; ++j) {
      const int s = j & 1, n = j >> 1;
      const int64_t tile = tile0 + j;
      const int cloud = (int)(tile / tiles_per_cloud);
      if (cloud != cur_cloud) {
        flush(cur_cloud);
        cur_cloud = cloud;
      }
      mbar_wait(BAR(EU + s), n & 1);  // rstd2 of the tile's points
      tr(1000 * j + 400);
      const float* rb = rbuf + (j & 3) * 128;
      const uint32_t idx_base = (uint32_t)((int)(tile % tiles_per_cloud) * 128);
      if (NBLK == 2) {
        // Both 128-channel blocks of the tile together, 32 points at a time: the per-point rstd2 (a broadcast
        // shared-memory load -- a full wavefront per point and warp, and the shared-memory pipe is what this kernel
        // runs out of) is loaded ONCE for the two blocks.
        mbar_wait(BAR(F2 + 0), j & 1);
        mbar_wait(BAR(F2 + 1), j & 1);
        tr(1000 * j + 410);
        tc_fence_after();
        const uint32_t ta = tmem_base + 256u + lane_off, tb = ta + 128u;
        uint32_t ka0 = 0, ka1 = 0, kb0 = 0, kb1 = 0;
        float ma0 = m[0], ma1 = -3.0e38f, mb0 = m[NBLK - 1], mb1 = -3.0e38f;
        uint32_t va[32], vb[32];
#pragma unroll
        for (int ch = 0; ch < 128; ch += 32) {
          if (!(dbg & 16)) {
            tmem_ld32_async(ta + ch, va);
            tmem_ld32_async(tb + ch, vb);
            tmem_wait_ld_dep(va);
            tmem_wait_ld_dep(vb);
          }
          if (ch == 96) {  // every column of both blocks has been read: hand the ring slots back before the last chunk's math
            tc_fence_before();
            mbar_arrive(BAR(D2 + 0));
            mbar_arrive(BAR(D2 + 1));
          }
          if (dbg & 1) continue;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 rr = *reinterpret_cast<const float4*>(rb + ch + i);  // broadcast
            const uint64_t r01 = pk2f(rr.x, rr.y), r23 = pk2f(rr.z, rr.w);
            if (ARGMAX) {
              uint32_t z0, z1, z2, z3;
              unpk2(fma2(pk2(va[i], va[i + 1]), r01, bias2), z0, z1);
              unpk2(fma2(pk2(va[i + 2], va[i + 3]), r23, bias2), z2, z3);
              z0 = (z0 & 0xffffff80u) | (uint32_t)(127 - (ch + i));
              z1 = (z1 & 0xffffff80u) | (uint32_t)(127 - (ch + i + 1));
              z2 = (z2 & 0xffffff80u) | (uint32_t)(127 - (ch + i + 2));
              z3 = (z3 & 0xffffff80u) | (uint32_t)(127 - (ch + i + 3));
              ka0 = __vimax3_u32(ka0, z0, z1);
              ka1 = __vimax3_u32(ka1, z2, z3);
              unpk2(fma2(pk2(vb[i], vb[i + 1]), r01, bias2), z0, z1);
              unpk2(fma2(pk2(vb[i + 2], vb[i + 3]), r23, bias2), z2, z3);
              z0 = (z0 & 0xffffff80u) | (uint32_t)(127 - (ch + i));
              z1 = (z1 & 0xffffff80u) | (uint32_t)(127 - (ch + i + 1));
              z2 = (z2 & 0xffffff80u) | (uint32_t)(127 - (ch + i + 2));
              z3 = (z3 & 0xffffff80u) | (uint32_t)(127 - (ch + i + 3));
              kb0 = __vimax3_u32(kb0, z0, z1);
              kb1 = __vimax3_u32(kb1, z2, z3);
            } else {
              const uint64_t zero2 = pk2f(0.f, 0.f);
              uint32_t z0, z1, z2, z3;
              unpk2(fma2(pk2(va[i], va[i + 1]), r01, zero2), z0, z1);
              unpk2(fma2(pk2(va[i + 2], va[i + 3]), r23, zero2), z2, z3);
              ma0 = fmax3(ma0, __uint_as_float(z0), __uint_as_float(z1));
              ma1 = fmax3(ma1, __uint_as_float(z2), __uint_as_float(z3));
              unpk2(fma2(pk2(vb[i], vb[i + 1]), r01, zero2), z0, z1);
              unpk2(fma2(pk2(vb[i + 2], vb[i + 3]), r23, zero2), z2, z3);
              mb0 = fmax3(mb0, __uint_as_float(z0), __uint_as_float(z1));
              mb1 = fmax3(mb1, __uint_as_float(z2), __uint_as_float(z3));
            }
          }
        }
        if (ARGMAX) {
          const uint32_t ka = max(ka0, ka1), kb = max(kb0, kb1);
          const uint32_t ia = idx_base + (127u - (ka & 127u)), ib = idx_base + (127u - (kb & 127u));
          run[0] = max(run[0], ((unsigned long long)(ka & 0xffffff80u) << 32) | (unsigned long long)(0xFFFFFFFFu - ia));
          run[NBLK - 1] = max(run[NBLK - 1], ((unsigned long long)(kb & 0xffffff80u) << 32) | (unsigned long long)(0xFFFFFFFFu - ib));
        } else {
          m[0] = fmaxf(ma0, ma1);
          m[NBLK - 1] = fmaxf(mb0, mb1);
        }
        tr(1000 * j + 421);
        continue;
      }
#pragma unroll
      for (int blk = 0; blk < NBLK; ++blk) {
        const int g = j * NBLK + blk, r = g & 1, k = g >> 1;
        mbar_wait(BAR(F2 + r), k & 1);
        tr(1000 * j + 410 + blk);
        tc_fence_after();
        const uint32_t t0 = tmem_base + 256u + (uint32_t)(r * 128) + lane_off;
        uint32_t mk0 = 0, mk1 = 0;        // two independent max chains
        float mm0 = m[blk], mm1 = -3.0e38f;
        tmem_chunks<4>(t0, (dbg & 16) != 0, [&](uint32_t(&v)[32], int ch) {
          if (dbg & 1) return;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 rr = *reinterpret_cast<const float4*>(rb + ch + i);  // broadcast
            if (ARGMAX) {
              // (bits(z + bias) & ~127) | (127 - point): z + bias > 0, so integer order is value order and the low 7
              // mantissa bits carry the point, ties resolving to the smallest index
              uint32_t z0, z1, z2, z3;
              unpk2(fma2(pk2(v[i], v[i + 1]), pk2f(rr.x, rr.y), bias2), z0, z1);
              unpk2(fma2(pk2(v[i + 2], v[i + 3]), pk2f(rr.z, rr.w), bias2), z2, z3);
              z0 = (z0 & 0xffffff80u) | (uint32_t)(127 - (ch + i));
              z1 = (z1 & 0xffffff80u) | (uint32_t)(127 - (ch + i + 1));
              z2 = (z2 & 0xffffff80u) | (uint32_t)(127 - (ch + i + 2));
              z3 = (z3 & 0xffffff80u) | (uint32_t)(127 - (ch + i + 3));
              mk0 = __vimax3_u32(mk0, z0, z1);
              mk1 = __vimax3_u32(mk1, z2, z3);
            } else {
              const uint64_t zero2 = pk2f(0.f, 0.f);
              uint32_t z0, z1, z2, z3;
              unpk2(fma2(pk2(v[i], v[i + 1]), pk2f(rr.x, rr.y), zero2), z0, z1);
              unpk2(fma2(pk2(v[i + 2], v[i + 3]), pk2f(rr.z, rr.w), zero2), z2, z3);
              mm0 = fmax3(mm0, __uint_as_float(z0), __uint_as_float(z1));
              mm1 = fmax3(mm1, __uint_as_float(z2), __uint_as_float(z3));
            }
          }
        }, [&]() {  // every column of the block has been read: hand the ring slot back before the last chunk's math
          tc_fence_before();
          mbar_arrive(BAR(D2 + r));
        });
        if (ARGMAX) {
          const uint32_t mk = max(mk0, mk1);
          const uint32_t idx = idx_base + (127u - (mk & 127u));
          const unsigned long long key = ((unsigned long long)(mk & 0xffffff80u) << 32) | (unsigned long long)(0xFFFFFFFFu - idx);
          run[blk] = max(run[blk], key);
        } else {
          m[blk] = fmaxf(mm0, mm1);
        }
        tr(1000 * j + 420 + blk);
      }
    }
    if (n_local > 0) flush(cur_cloud);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
  }
}

// pooled[r][c] = relu(|g2[c]| * max_p z + b2[c]) and the argmax, from the packed (key, ~index) maxima
__global__ void pool_finalize2_kernel(unsigned long long* __restrict__ keys, int R, int c3, const float* __restrict__ g2,
                                      const float* __restrict__ be2, float key_bias, float* __restrict__ pooled,
                                      int32_t* __restrict__ argmax) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)R * c3) return;
  const int c = (int)(i % c3);
  const unsigned long long k = keys[i];
  keys[i] = 0ull;  // leave the scratch zeroed for the next call
  const float z = __uint_as_float((uint32_t)(k >> 32)) - key_bias;
  pooled[i] = fmaxf(fmaf(fabsf(g2[c]), z, be2[c]), 0.f);
  if (argmax) argmax[i] = (int32_t)(0xFFFFFFFFu - (uint32_t)k);
}

__device__ __forceinline__ uint32_t img_off(int n, int k, int K) {
  return (uint32_t)((n >> 3) * (K * 16) + (k >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2);
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }
// relu(g x + b) = |g| relu(sign(g) x + b / |g|): |gamma1| moves into the next layer's weights.  A gain of exactly zero
// (relu(b), independent of x) is treated as 1e-12: x's contribution vanishes below fp32 resolution all the same.
__device__ __forceinline__ float gamma1_mag(float g) { return fmaxf(fabsf(g), 1e-12f); }

// Column means of W1 (over its c2 rows) and W2 (over its c3 rows): block 0 / block 1, 1024 threads = 8 row groups x
// (up to) 128 columns, reduced through shared memory.  ~1.5 us; the packer reads the result.
__global__ void __launch_bounds__(1024)
colmean_kernel(const float* __restrict__ w1, const float* __restrict__ w2, int c1, int c2, int c3, float* __restrict__ cm) {
  __shared__ float part[8][128];
  const float* w = blockIdx.x == 0 ? w1 : w2;
  const int rows = blockIdx.x == 0 ? c2 : c3, cols = blockIdx.x == 0 ? c1 : c2;
  const int k = threadIdx.x & 127, g = threadIdx.x >> 7;
  float s = 0.f;
  if (k < cols)
    for (int n = g; n < rows; n += 8) s += w[n * cols + k];
  part[g][k] = s;
  __syncthreads();
  if (g == 0 && k < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i][k];
    cm[(blockIdx.x == 0 ? 0 : c1) + k] = t / (float)rows;
  }
}

__global__ void __launch_bounds__(256)
pack_weights2_kernel(const float* __restrict__ w0, const float* __restrict__ b0, const float* __restrict__ w1,
                     const float* __restrict__ g1, const float* __restrict__ be1, const float* __restrict__ w2,
                     const float* __restrict__ g2, const float* __restrict__ be2, int C, int c1, int c2, int c3,
                     int rgb_u8, char* __restrict__ out) {
  const Wpack2 W = make_wpack2(c1, c2, c3);
  const float* m1 = reinterpret_cast<const float*>(out + W.cm);
  const float* m2 = m1 + c1;
  const int n0 = c1 * 16, n1 = c2 * c1, n2 = c3 * c2, n3 = c2 * c2, n4 = 2 * c2 + 2 * c3;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n0) {
    const int n = i / 16, k = i % 16;
    float v = 0.f;
    if (k < C) {
      v = w0[n * C + k];
      if (rgb_u8 && k >= 3 && k < 6) v *= (1.0f / 255.0f);  // staged rgb is the raw 0..255 integer
    } else if (k == C) {
      v = b0[n];  // multiplies the constant-1 channel
    } else if (k <= C + 3) {
      v = w0[n * C + (k - C - 1)];  // lo parts of xyz see the same weights
    }
    *reinterpret_cast<__nv_bfloat16*>(out + W.w0 + img_off(n, k, 16)) = __float2bfloat16(v);
  } else if (i < n0 + n1) {
    // layer 1: centred over the output channels, sign(gamma1) folded into the rows (the variance is unchanged)
    const int e = i - n0, n = e / c1, k = e % c1;
    const float v = w1[e] - m1[k];
    *reinterpret_cast<__nv_bfloat16*>(out + W.w1 + img_off(n, k, c1)) = __float2bfloat16(g1[n] >= 0.f ? v : -v);
  } else if (i < n0 + n1 + n2) {
    // layer 2: centred over the output channels, sign(gamma2) folded into the rows, |gamma1| into the columns
    const int e = i - n0 - n1, n = e / c2, k = e % c2;
    const float v = (w2[e] - m2[k]) * gamma1_mag(g1[k]);
    *reinterpret_cast<__nv_bfloat16*>(out + W.w2 + img_off(n, k, c2)) = __float2bfloat16(g2[n] >= 0.f ? v : -v);
  } else if (i >= n0 + n1 + n2 + n3 && i < n0 + n1 + n2 + n3 + n4) {
    const int e = i - n0 - n1 - n2 - n3;
    if (e < c2) reinterpret_cast<float*>(out + W.prm1)[e] = be1[e] / gamma1_mag(g1[e]);  // beta1 / |gamma1|
    else if (e < 2 * c2) reinterpret_cast<float*>(out + W.prm1)[e] = 0.f;
    else if (e < 2 * c2 + c3) reinterpret_cast<float*>(out + W.prm2)[e - 2 * c2] = g2[e - 2 * c2];
    else reinterpret_cast<float*>(out + W.prm2)[e - 2 * c2] = be2[e - 2 * c2 - c3];
  }
}

// Gram matrix Gc = W2c^T W2c of the centred, bf16-ROUNDED layer-2 weights (what the tensor core multiplies by), fp32
// accumulate.  One block = 32 consecutive outputs (k, kk .. kk+31) x 8 row groups: every thread sums c3/8 rows (the kk
// loads coalesce, the k column is a broadcast) and the partials meet in shared memory.  ~2 us, where one thread per
// output looping over all c3 rows is latency-bound at ~20 us.
__global__ void __launch_bounds__(256)
gram_kernel(const float* __restrict__ w2, const float* __restrict__ m2, const float* __restrict__ g1, int c2, int c3,
            char* __restrict__ gc_img) {
  __shared__ float part[8][33];
  const int o = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int e = blockIdx.x * 32 + o, k = e / c2, kk = e % c2;
  const float mk = m2[k], mkk = m2[kk], ak = gamma1_mag(g1[k]), akk = gamma1_mag(g1[kk]);
  float s0 = 0.f, s1 = 0.f;
  for (int n = g; n < c3; n += 16) {
    s0 = fmaf(bf16_round((w2[n * c2 + k] - mk) * ak), bf16_round((w2[n * c2 + kk] - mkk) * akk), s0);
    if (n + 8 < c3)
      s1 = fmaf(bf16_round((w2[(n + 8) * c2 + k] - mk) * ak), bf16_round((w2[(n + 8) * c2 + kk] - mkk) * akk), s1);
  }
  part[g][o] = s0 + s1;
  __syncthreads();
  if (g == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i][o];
    *reinterpret_cast<__nv_bfloat16*>(gc_img + img_off(k, kk, c2)) = __float2bfloat16(t);
  }
}

bool shapes_ok(int c1, int c2, int c3) {
  // c3 a multiple of 256 up to 2048: c3 / 256 channel groups of CTAs (the packer's column-mean kernel covers c3 <= 2048)
  return (c1 == 64 || c1 == 128) && c2 == 128 && (c3 == 128 || (c3 % 256 == 0 && c3 <= 2048)) &&
         make_smem2(c1, c2, std::min(c3, 256)).total <= 227 * 1024;
}

int64_t wpack_bytes(int c1, int c2, int c3) { return shapes_ok(c1, c2, c3) ? (int64_t)make_wpack2(c1, c2, c3).total : 0; }

int pack(const float* w0, const float* b0, const float* w1, const float* g1, const float* be1, const float* w2,
         const float* g2, const float* be2, int C, int c1, int c2, int c3, int rgb_u8, void* wpack2, cudaStream_t st) {
  const int n = c1 * 16 + c2 * c1 + c3 * c2 + c2 * c2 + 2 * c2 + 2 * c3;
  colmean_kernel<<<2, 1024, 0, st>>>(w1, w2, c1, c2, c3, reinterpret_cast<float*>((char*)wpack2 + make_wpack2(c1, c2, c3).cm));
  PCRL_CHECK_LAUNCH();
  pack_weights2_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(w0, b0, w1, g1, be1, w2, g2, be2, C, c1, c2, c3, rgb_u8,
                                                                (char*)wpack2);
  PCRL_CHECK_LAUNCH();
  const Wpack2 W = make_wpack2(c1, c2, c3);
  gram_kernel<<<(unsigned)(c2 * c2 / 32), 256, 0, st>>>(w2, reinterpret_cast<const float*>((char*)wpack2 + W.cm) + c1, g1, c2,
                                                       c3, (char*)wpack2 + W.gc);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

template <int C1, int C2, int NBLK>
static int launch(const void* xh, const void* wpack2, int n_tiles, int tiles_per_cloud, int src_cloud_stride, float ln_eps,
                  float key_bias, unsigned long long* keys, bool want_argmax, int c3_total, cudaStream_t st) {
  const Smem2 L = make_smem2(C1, C2, NBLK * 128);
  const int groups = c3_total / (NBLK * 128);
  const int ctas_per_group = std::max(1, std::min(sm_count() / groups, n_tiles));
  const int grid = groups * ctas_per_group;
  if (want_argmax) {
    auto* kern = pointnet_fwd_tc2_kernel<C1, C2, NBLK, true>;
    PCRL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    kern<<<grid, kThreads, L.total, st>>>((const char*)xh, (const char*)wpack2, n_tiles, tiles_per_cloud, src_cloud_stride,
                                          ln_eps, key_bias, keys, c3_total, ctas_per_group);
  } else {
    auto* kern = pointnet_fwd_tc2_kernel<C1, C2, NBLK, false>;
    PCRL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    kern<<<grid, kThreads, L.total, st>>>((const char*)xh, (const char*)wpack2, n_tiles, tiles_per_cloud, src_cloud_stride,
                                          ln_eps, key_bias, keys, c3_total, ctas_per_group);
  }
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

// The fused forward + finalize.  wpack2: the second-generation image (pack()).
int forward(const void* xh, int R, int src_cloud_stride, int NP, const void* wpack2, int c1, int c2, int c3, float ln_eps,
            uint64_t* pool_keys, float* pooled, int32_t* argmax, cudaStream_t st) {
  const int tiles_per_cloud = NP / 128;
  const int n_tiles = R * tiles_per_cloud;
  const float key_bias = 2.0f * sqrtf((float)c3);  // |xhat| < sqrt(c3): z + bias stays positive with margin
  auto* keys = reinterpret_cast<unsigned long long*>(pool_keys);
  int rc = PCRL_EUNSUPPORTED;
  const bool am = argmax != nullptr;
  // c3 = 128: one block; otherwise groups of two 128-channel blocks (c3 = 256: one group; 1024: four groups of CTAs)
  if (c1 == 128 && c3 % 256 == 0) rc = launch<128, 128, 2>(xh, wpack2, n_tiles, tiles_per_cloud, src_cloud_stride, ln_eps, key_bias, keys, am, c3, st);
  else if (c1 == 64 && c3 % 256 == 0) rc = launch<64, 128, 2>(xh, wpack2, n_tiles, tiles_per_cloud, src_cloud_stride, ln_eps, key_bias, keys, am, c3, st);
  else if (c1 == 128 && c3 == 128) rc = launch<128, 128, 1>(xh, wpack2, n_tiles, tiles_per_cloud, src_cloud_stride, ln_eps, key_bias, keys, am, c3, st);
  else if (c1 == 64 && c3 == 128) rc = launch<64, 128, 1>(xh, wpack2, n_tiles, tiles_per_cloud, src_cloud_stride, ln_eps, key_bias, keys, am, c3, st);
  if (rc != PCRL_OK) return rc;
  const int64_t n = (int64_t)R * c3;
  const Wpack2 W = make_wpack2(c1, c2, c3);
  const float* prm2 = reinterpret_cast<const float*>((const char*)wpack2 + W.prm2);
  pool_finalize2_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(keys, R, c3, prm2, prm2 + c3, key_bias, pooled, argmax);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

int set_debug_flags(int flags) {
  PCRL_CHECK_CUDA(cudaMemcpyToSymbol(g_dbg2, &flags, sizeof(int)));
  static long long zeros[6 * 1024] = {0};
  PCRL_CHECK_CUDA(cudaMemcpyToSymbol(g_trace2, zeros, sizeof(zeros)));
  return PCRL_OK;
}
int get_trace(long long* out_host) {
  PCRL_CHECK_CUDA(cudaMemcpyFromSymbol(out_host, g_trace2, sizeof(long long) * 6 * 1024));
  return 6 * 512;
}

}  // namespace tc2
}  // namespace pcrl
