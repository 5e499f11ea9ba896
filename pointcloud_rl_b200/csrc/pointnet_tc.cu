// Fused PointNet per-point MLP + max-pool on the 5th-gen tensor cores (tcgen05 / TMEM), sm_100a.
//
//   per 128-point tile:  acc0 = X W0'^T          (K=16: channels | 1 (bias) | xyz lo parts)   -> relu    -> h0 (bf16, smem)
//                        acc1 = h0 W1^T          -> LN over channels (thread-local) + relu   -> h1 (bf16, smem)
//                        acc2 = h1 W2^T          -> LN + relu -> max / argmax over the tile's points
//                        (key, ~index) packed u64 -> atomicMax into the cloud's pooled row
//
// Points sit on the MMA M axis (128 TMEM lanes = 128 points), channels along TMEM columns, so after a
// 32x32b tcgen05.ld every thread owns one point's channel row: LayerNorm statistics are thread-local
// and the max over points is the cross-lane direction (a warp-private shared-memory transpose).
//
// Persistent kernel, one CTA per SM, 18 warps:
//   warp 0       TMA producer: weights once (cp.async.bulk), then the 4 KB point tiles through a 4-stage ring
//   warp 1       MMA issuer (one elected thread), serves whichever of the two tile slots is ready
//   warps 2-9    epilogue of slot 0   } each slot owns 256 TMEM columns and one activation buffer; while one
//   warps 10-17  epilogue of slot 1   } slot's epilogue normalises layer k, the tensor core runs the other slot
//
// All operands use the no-swizzle K-major UMMA canonical layout (8x16B core matrices, LBO = 128 B
// between K-adjacent cores, SBO = K*16 B between 8-row groups); the staging kernel and the weight
// packer write global memory in exactly that byte image so plain 1-D bulk copies land MMA-ready tiles.
#include <cuda_bf16.h>

#include "common.cuh"
#include "pointnet_tc2.cuh"
#include "tc_ptx.cuh"

namespace pcrl {
namespace tc {

// profiling knobs (tools/probe_fwd.py): knock out individual epilogue passes to attribute time.  0 in production.
__device__ int g_dbg = 0;
__device__ long long g_trace[3 * 1024];  // CTA 0, g_dbg bit 7: three tracer threads x 512 (event, clock) pairs
__device__ int g_trace_n = 0;
struct Tracer {  // register-resident cursor, plain stores: no atomics or round trips on the traced path
  int region, n;
  bool on;
  __device__ __forceinline__ void operator()(int ev) {
#ifdef PCRL_FWD_TRACE  // compiled out of production builds (tools/probe_fwd.py documents how to enable)
    if (on && n < 512) {
      g_trace[region * 1024 + 2 * n] = ev;
      g_trace[region * 1024 + 2 * n + 1] = clock64();
      ++n;
    }
#endif
  }
};

constexpr int kStages = 3;
constexpr int kTileBytes = 128 * 16 * 2;  // one X tile: 128 points x 16 bf16
constexpr int kThreads = 128 + 256 + 256;  // warp 0 TMA, 1 MMA issuer (layers 0/1), 2 MMA issuer (layer 2), 3 idle; 8 warps layers 0/1; 8 warps layer 2

// Dump-mode store of one 32x32 fp32 chunk.  After tcgen05.ld every thread holds 32 consecutive floats of ITS row, so a
// direct store makes each instruction touch 32 different lines with 16 bytes each.  The chunk goes through a
// warp-private 2 KB scratch tile (two 16-column halves, XOR-swizzled 16-byte slots, conflict-free both ways) so that
// every global store instruction writes eight 64-byte row segments instead.
// x: this lane's row; g: address of (row 0 of the warp, first column of the chunk) in a row-major array of pitch ld.
__device__ __forceinline__ void dump_chunk32(float* scratch, float* g, int64_t ld, const uint32_t (&x)[32], int lane) {
  const int wsw = (lane >> 1) & 3;
  const int rr = lane >> 2, rs = lane & 3;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    __syncwarp();
#pragma unroll
    for (int sl = 0; sl < 4; ++sl)
      *reinterpret_cast<uint4*>(scratch + lane * 16 + ((sl ^ wsw) << 2)) =
          make_uint4(x[half * 16 + 4 * sl], x[half * 16 + 4 * sl + 1], x[half * 16 + 4 * sl + 2], x[half * 16 + 4 * sl + 3]);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = k * 8 + rr;
      const float4 val = *reinterpret_cast<const float4*>(scratch + r * 16 + ((rs ^ ((r >> 1) & 3)) << 2));
      *reinterpret_cast<float4*>(g + r * ld + half * 16 + rs * 4) = val;
    }
  }
}

// "Dump" mode (the sparse backward's recompute): instead of max-pooling, every row's intermediates are written out
// in fp32 for the LayerNorm / GEMM backward kernels.  rows_dev bounds the compacted rows actually present.
struct DumpOut {
  const int* rows_dev;
  float *h0, *xhat1, *rstd1, *h1, *xhat2, *rstd2;
  // If set, `xhat2` receives the dense part of the layer-2 LayerNorm backward instead of xhat2:
  //   dy2[a][c] = rstd2[a] * (-m1[a] - xhat2[a][c] * m2[a]),   m1 = mean_c(g2*dout), m2 = mean_c(g2*dout*xhat2)
  // (dout is nonzero only at the (row, channel) pairs that won the max-pool; their xhat is known from the pooled
  // value, so m1/m2 exist before the recompute and the separate LN-backward pass over [A, c3] disappears).
  const float *m1, *m2;
};

// Packed weight buffer (pcrl_pointnet_pack_weights):
//   [ W0' | W1 | W2s | g1 be1 g2 be2 (fp32) ]   <- loaded into shared memory as one image by the max-pool forward
//   [ pos[c3] (int32) | n_pos (int32, padded to 16 B) ]
//   [ W2 ]                                       <- plain layer-2 image, used by the dump (backward recompute) mode
// W2s is W2 with its rows permuted and sign-flipped: row pos[c] = sign(g2[c]) * W2[c, :], channels with g2 >= 0 first
// (n_pos of them).  max_p(g*x + b) = |g| * max_p(sign(g)*x) + b, so the layer-2 epilogue max-pools z = sign(g2)*xhat
// directly -- no gamma/beta loads and one FFMA less per element in the hot loop -- and the affine part runs once per
// (cloud, channel) in pool_finalize_kernel.
struct WpackLayout {
  uint32_t w0, w1, w2s, prm, pos, npos, w2, total, img_bytes;
};
__host__ __device__ inline WpackLayout make_wpack(int c1, int c2, int c3) {
  WpackLayout W;
  uint32_t o = 0;
  W.w0 = o;   o += c1 * 32;
  W.w1 = o;   o += c2 * c1 * 2;
  W.w2s = o;  o += c3 * c2 * 2;
  W.prm = o;  o += (2 * c2 + 2 * c3) * 4;
  W.img_bytes = o;
  W.pos = o;  o += c3 * 4;
  W.npos = o; o += 16;
  W.w2 = o;   o += c3 * c2 * 2;
  W.total = o;
  return W;
}
constexpr float kKeyBias = 16.0f;  // |xhat| <= sqrt(C) <= 16 for C <= 256: z + 16 > 0, so float order == integer order

struct SmemLayout {
  uint32_t w0, w1, w2, prm, xst, act0, act1, tr, wkey, stat, bars, total;
};
__host__ __device__ inline SmemLayout make_layout(int c1, int c2, int c3) {
  SmemLayout L;
  uint32_t o = 0;
  L.w0 = o;   o += c1 * 32;
  L.w1 = o;   o += c2 * c1 * 2;
  L.w2 = o;   o += c3 * c2 * 2;
  L.prm = o;  o += (2 * c2 + 2 * c3) * 4;
  o = (o + 127) & ~127u;
  L.xst = o;  o += kStages * kTileBytes;
  L.act0 = o; o += 128 * c1 * 2;  // h0 (bf16 UMMA image): layer-0 epilogue -> layer-1 MMA
  L.act1 = o; o += 128 * c2 * 2;  // h1: layer-1 epilogue -> layer-2 MMA
  L.tr = o;   o += 8 * 4096;      // max-pool transpose scratch: 4 KB per layer-2 warp
  L.wkey = o; o += 4 * c3 * 8;
  L.stat = o; o += 2 * 2 * 128 * 8;  // per group, per channel half: (sum, sumsq) of every row
  L.bars = o; o += 256;
  L.total = o;
  return L;
}

// The layers run as a pipeline over consecutive 128-point tiles.
//   * warps 4..11 ("front" group) do the layer-0 and layer-1 epilogues of tile i+1 while
//   * warps 12..19 ("pool" group) normalise and max-pool layer 2 of tile i, and the tensor pipe runs under both.
// TMEM (512 columns): [0, max(c1,c2)) is the accumulator of layers 0 and 1 (they alias: the front group is a serial
// chain anyway), followed by a ring of three half-accumulators of c3/2 columns for layer 2.  Layer 2 is issued as
// two N = c3/2 MMAs per tile; tile i uses ring slots (2i)%3 and (2i+1)%3, so the first half of tile i+1 can be
// computed while the pool group still reads tile i, and the second half as soon as the pool group has drained the
// first half of tile i: the pool group, which is the bottleneck stage, never waits for the tensor pipe.
__global__ void __launch_bounds__(kThreads, 1)
pointnet_fwd_tc_kernel(const char* __restrict__ xh, const char* __restrict__ wpack, int n_tiles, int tiles_per_cloud,
                       int src_cloud_stride, int c1, int c2, int c3, float ln_eps, int want_argmax,
                       unsigned long long* __restrict__ pool_keys, DumpOut dump) {
  extern __shared__ __align__(128) unsigned char smem[];
  const SmemLayout L = make_layout(c1, c2, c3);
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // mbarriers
  constexpr int WB = 0;              // weights landed
  constexpr int XF = 1;              // [kStages] x tile landed
  constexpr int XE = XF + kStages;   // [kStages] x stage free
  constexpr int F0 = XE + kStages;   // layer-0 accumulator complete (tcgen05.commit)
  constexpr int F1 = F0 + 1;         // layer-1 accumulator complete
  constexpr int F2 = F1 + 1;         // both layer-2 halves complete (2 commits)
  constexpr int E0 = F2 + 1;         // front group: h0 written, accumulator drained (256 arrivals)
  constexpr int E1 = E0 + 1;         // front group: h1 written, accumulator drained
  constexpr int DA = E1 + 1;         // [2, by tile parity] pool group drained the first layer-2 half
  constexpr int DB = DA + 2;         // [2, by tile parity] ... the second half
  constexpr int NBAR = DB + 2;
  const uint32_t bar0 = sbase + L.bars;
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + L.bars + NBAR * 8);

  if (threadIdx.x == 0) {
    mbar_init(BAR(WB), 1);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(BAR(XF + s), 1);
      mbar_init(BAR(XE + s), 1);
    }
    mbar_init(BAR(F0), 1);
    mbar_init(BAR(F1), 1);
    mbar_init(BAR(F2), 2);
    mbar_init(BAR(E0), 256);
    mbar_init(BAR(E1), 256);
    for (int k = 0; k < 2; ++k) {
      mbar_init(BAR(DA + k), 256);
      mbar_init(BAR(DB + k), 256);
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

  if (dump.rows_dev) n_tiles = min(n_tiles, (*dump.rows_dev + 127) >> 7);  // compacted set: size known on the device only
  // every CTA takes a contiguous range of tiles: consecutive tiles belong to the same cloud, so the pool group keeps a
  // running per-cloud maximum in shared memory and touches global memory only when the cloud changes
  const int t_q = n_tiles / (int)gridDim.x, t_r = n_tiles % (int)gridDim.x;
  const int n_local = t_q + ((int)blockIdx.x < t_r ? 1 : 0);
  const int64_t tile0 = (int64_t)blockIdx.x * t_q + min((int)blockIdx.x, t_r);
  const uint32_t wbytes = (uint32_t)(c1 * 32 + c2 * c1 * 2 + c3 * c2 * 2 + (2 * c2 + 2 * c3) * 4);
  const int dbg = g_dbg;
  const int ch2 = c3 >> 1;                                   // channels per layer-2 half
  const uint32_t ring0 = tmem_base + (uint32_t)(c1 > c2 ? c1 : c2);  // first ring slot
  // ring slot (TMEM column base) of layer-2 half `half` of local tile i
  auto ring = [&](int i, int half) { return ring0 + (uint32_t)(((2 * i + half) % 3) * ch2); };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_expect_tx(BAR(WB), wbytes);
      // weights + LN parameters are one contiguous image laid out exactly like the smem carve-up; the dump mode
      // swaps in the plain (unpermuted, unsigned) layer-2 weights
      const WpackLayout W = make_wpack(c1, c2, c3);
      uint32_t off = 0;
      while (off < wbytes) {
        uint32_t n = min(wbytes - off, 32768u);
        uint32_t src = off;
        if (dump.xhat2 && off >= W.w2s && off < W.prm) {
          n = min(n, W.prm - off);
          src = W.w2 + (off - W.w2s);
        } else if (dump.xhat2 && off < W.w2s) {
          n = min(n, W.w2s - off);
        }
        bulk_g2s(sbase + L.w0 + off, wpack + src, n, BAR(WB));
        off += n;
      }
      for (int i = 0; i < n_local; ++i) {
        const int st = i % kStages;
        if (i >= kStages) mbar_wait_relaxed(BAR(XE + st), ((i / kStages) - 1) & 1, 100);
        mbar_expect_tx(BAR(XF + st), kTileBytes);
        const int64_t tile = tile0 + i;
        // cloud r reads the tiles of source cloud r * src_cloud_stride (the actor step encodes the first of the
        // num_aug staged copies of every sample in place)
        const int64_t src_tile = src_cloud_stride == 1 ? tile
                                                        : (tile / tiles_per_cloud) * src_cloud_stride * tiles_per_cloud + tile % tiles_per_cloud;
        bulk_g2s(sbase + L.xst + st * kTileBytes, xh + src_tile * kTileBytes, kTileBytes, BAR(XF + st));
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer, layers 0 and 1 (a serial chain per tile)
    if (lane == 0) {
      mbar_wait(BAR(WB), 0);
      const uint32_t id0 = make_idesc(c1), id1 = make_idesc(c2);
      Tracer trace{0, 0, (dbg & 128) && blockIdx.x == 0};
      for (int j = 0; j < n_local; ++j) {
        const int st = j % kStages;
        mbar_wait(BAR(XF + st), (j / kStages) & 1);
        if (j > 0) mbar_wait(BAR(E1), (j - 1) & 1);  // layer-1 epilogue of the previous tile drained the shared accumulator
        trace(100);
        tc_fence_after();
        mma_bf16(tmem_base, make_desc(sbase + L.xst + st * kTileBytes, 256), make_desc(sbase + L.w0, 256), id0, 0);
        mma_commit(BAR(XE + st));
        mma_commit(BAR(F0));
        trace(200);
        mbar_wait(BAR(E0), j & 1);  // h0 in shared memory, accumulator drained
        trace(101);
        tc_fence_after();
        for (int ks = 0; ks < c1 / 16; ++ks)
          mma_bf16(tmem_base, make_desc(sbase + L.act0 + ks * 256, c1 * 16), make_desc(sbase + L.w1 + ks * 256, c1 * 16), id1, ks > 0);
        mma_commit(BAR(F1));
        trace(201);
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ MMA issuer, layer 2 (two N = c3/2 halves into the ring)
    if (lane == 0) {
      mbar_wait(BAR(WB), 0);
      const uint32_t id2 = make_idesc(ch2);
      const uint32_t w2b = sbase + L.w2 + (uint32_t)(ch2 >> 3) * (uint32_t)(c2 * 16);  // rows ch2.. of the W2 image
      Tracer trace{1, 0, (dbg & 128) && blockIdx.x == 0};
      for (int j = 0; j < n_local; ++j) {
        mbar_wait(BAR(E1), j & 1);  // h1 of tile j in shared memory
        // first half -> ring slot last used by the second half of tile j-2
        if (j >= 2) mbar_wait(BAR(DB + (j & 1)), ((j - 2) >> 1) & 1);
        trace(102);
        tc_fence_after();
        for (int ks = 0; ks < c2 / 16; ++ks)
          mma_bf16(ring(j, 0), make_desc(sbase + L.act1 + ks * 256, c2 * 16), make_desc(sbase + L.w2 + ks * 256, c2 * 16), id2, ks > 0);
        mma_commit(BAR(F2));
        trace(202);
        // second half -> ring slot last used by the first half of tile j-1
        if (j >= 1) mbar_wait(BAR(DA + ((j - 1) & 1)), ((j - 1) >> 1) & 1);
        trace(103);
        tc_fence_after();
        for (int ks = 0; ks < c2 / 16; ++ks)
          mma_bf16(ring(j, 1), make_desc(sbase + L.act1 + ks * 256, c2 * 16), make_desc(w2b + ks * 256, c2 * 16), id2, ks > 0);
        mma_commit(BAR(F2));
        trace(203);
      }
    }
  } else if (warp == 3) {
    // idle: the register file is allocated in units of four warps anyway
  } else if (warp < 12) {
    // ------------------------------------------------------------------ front group: layer-0 and layer-1 epilogues
    // 8 warps: two per TMEM lane quadrant (hardware rule: a warp reaches lanes 32*(warp%4)..+31), each taking half
    // of the channels.  LayerNorm needs whole-row statistics, so the pair swaps its partial (sum, sum of squares)
    // through shared memory around a 64-thread named barrier.
    const int h = (warp - 4) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    const float* prm = reinterpret_cast<const float*>(smem + L.prm);
    const float *g1 = prm, *be1 = prm + c2;
    unsigned char* dst0 = smem + L.act0 + (row >> 3) * (uint32_t)(c1 * 16) + (row & 7) * 16;
    unsigned char* dst1 = smem + L.act1 + (row >> 3) * (uint32_t)(c2 * 16) + (row & 7) * 16;
    float2* stat_mine = reinterpret_cast<float2*>(smem + L.stat) + (2 + h) * 128 + row;
    const float2* stat_peer = reinterpret_cast<const float2*>(smem + L.stat) + (2 + (h ^ 1)) * 128 + row;
    const int pair_bar = 6 + q;
    float* dscr = reinterpret_cast<float*>(smem + L.tr) + (warp - 4) * 512;  // dump mode: 2 KB of the (idle) max-pool scratch
    mbar_wait(BAR(WB), 0);  // LN parameters landed
    for (int i = 0; i < n_local; ++i) {
      const int64_t tile = tile0 + i;
      uint32_t v[32];
      // ---- layer 0: ReLU -> bf16 operand of layer 1
      mbar_wait(BAR(F0), i & 1);
      tc_fence_after();
      if (!(dbg & 64)) {
        const int col0 = h * (c1 >> 1);
        float* d_h0 = dump.h0 ? dump.h0 + (tile * 128 + q * 32) * (int64_t)c1 : nullptr;
        for (int ch = col0; ch < col0 + (c1 >> 1); ch += 32) {
          tmem_ld32(tlane + ch, v);
          if (d_h0) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(fmaxf(__uint_as_float(v[j]), 0.f));
            dump_chunk32(dscr, d_h0 + ch, c1, v, lane);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 o;
            o.x = pack_relu_bf16x2(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1]));
            o.y = pack_relu_bf16x2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3]));
            o.z = pack_relu_bf16x2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5]));
            o.w = pack_relu_bf16x2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7]));
            *reinterpret_cast<uint4*>(dst0 + ((ch >> 3) + j) * 128) = o;
          }
        }
      }
      fence_async_smem();
      tc_fence_before();
      mbar_arrive(BAR(E0));

      // ---- layer 1: LN over channels + ReLU -> bf16 operand of layer 2
      mbar_wait(BAR(F1), i & 1);
      tc_fence_after();
      float rstd = 1.f, nmr = 0.f;
      const int col0 = h * (c2 >> 1);
      if (!(dbg & 32)) {
        float s4[4] = {0.f, 0.f, 0.f, 0.f}, q4[4] = {0.f, 0.f, 0.f, 0.f};
        for (int ch = 0; ch < (c2 >> 1); ch += 32) {
          tmem_ld32(tlane + col0 + ch, v);
          stats32(v, s4, q4);
        }
        float sum = (s4[0] + s4[1]) + (s4[2] + s4[3]), sq = (q4[0] + q4[1]) + (q4[2] + q4[3]);
        *stat_mine = make_float2(sum, sq);
        named_bar(pair_bar, 64);
        const float2 o = *stat_peer;
        sum += o.x;
        sq += o.y;
        const float mean = sum / (float)c2;
        rstd = rsqrtf(fmaxf(sq / (float)c2 - mean * mean, 0.f) + ln_eps);
        nmr = -mean * rstd;
      }
      if (i > 0) mbar_wait(BAR(F2), (i - 1) & 1);  // both layer-2 MMAs of the previous tile have finished reading h1
      if (!(dbg & 32)) {
        float* d_x1 = dump.xhat1 ? dump.xhat1 + (tile * 128 + q * 32) * (int64_t)c2 : nullptr;
        float* d_h1 = dump.h1 ? dump.h1 + (tile * 128 + q * 32) * (int64_t)c2 : nullptr;
        if (d_x1 && h == 0) dump.rstd1[tile * 128 + row] = rstd;
        for (int ch = col0; ch < col0 + (c2 >> 1); ch += 32) {
          tmem_ld32(tlane + ch, v);
          {
            const uint64_t r2 = pk2f(rstd, rstd), n2 = pk2f(nmr, nmr);
#pragma unroll
            for (int j = 0; j < 32; j += 2) unpk2(fma2(pk2(v[j], v[j + 1]), r2, n2), v[j], v[j + 1]);  // xhat
          }
          if (d_x1) dump_chunk32(dscr, d_x1 + ch, c2, v, lane);
#pragma unroll
          for (int j4 = 0; j4 < 32; j4 += 4) {
            const float4 gg = *reinterpret_cast<const float4*>(g1 + ch + j4);
            const float4 bb = *reinterpret_cast<const float4*>(be1 + ch + j4);
            unpk2(fma2(pk2(v[j4], v[j4 + 1]), pk2f(gg.x, gg.y), pk2f(bb.x, bb.y)), v[j4], v[j4 + 1]);
            unpk2(fma2(pk2(v[j4 + 2], v[j4 + 3]), pk2f(gg.z, gg.w), pk2f(bb.z, bb.w)), v[j4 + 2], v[j4 + 3]);
          }
          if (d_h1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(fmaxf(__uint_as_float(v[j]), 0.f));
            dump_chunk32(dscr, d_h1 + ch, c2, v, lane);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 pk;
            pk.x = pack_relu_bf16x2(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1]));
            pk.y = pack_relu_bf16x2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3]));
            pk.z = pack_relu_bf16x2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5]));
            pk.w = pack_relu_bf16x2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7]));
            *reinterpret_cast<uint4*>(dst1 + ((ch >> 3) + j) * 128) = pk;
          }
        }
      }
      fence_async_smem();
      tc_fence_before();
      mbar_arrive(BAR(E1));
    }
  } else {
    // ------------------------------------------------------------------ pool group: layer-2 LN, then max (and argmax) over the tile's points
    // 8 warps, two per TMEM lane quadrant; warp (q, h) owns channel quarter h of BOTH layer-2 halves, so the first
    // half's ring slot is released after half of the work.
    const int h = (warp - 12) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;  // point within the tile
    const int cq = ch2 >> 1;        // channels per (half, warp)
    unsigned long long* wkey = reinterpret_cast<unsigned long long*>(smem + L.wkey) + q * c3;
    const unsigned long long* wkey_all = reinterpret_cast<unsigned long long*>(smem + L.wkey);
    float2* stat_mine = reinterpret_cast<float2*>(smem + L.stat) + h * 128 + row;
    const float2* stat_peer = reinterpret_cast<const float2*>(smem + L.stat) + (h ^ 1) * 128 + row;
    const int pair_bar = 2 + q;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const int tid_grp = threadIdx.x - (128 + 256);
    // dump mode runs on the plain layer-2 weights: every channel counts as positive
    const int n_pos = dump.xhat2 ? c3 : *reinterpret_cast<const int*>(wpack + make_wpack(c1, c2, c3).npos);
    mbar_wait(BAR(WB), 0);  // LN parameters landed
    Tracer trace_e{2, 0, (dbg & 128) && blockIdx.x == 0 && warp == 12 && lane == 0};

    uint32_t* tr = reinterpret_cast<uint32_t*>(smem + L.tr) + (warp - 12) * 1024;  // 4 KB per warp
    const uint32_t lane_tag = 31u - (uint32_t)lane;
    uint32_t st_off[8], ld_off[8];  // XOR-swizzled word offsets of the transpose tile
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      st_off[k] = (uint32_t)(lane * 32 + ((k ^ (lane & 7)) << 2));
      ld_off[k] = (uint32_t)((((lane >> 2) ^ k) << 2) + (lane & 3));
    }

    // flush: combine the four quadrants' running maxima of `cloud` into global memory and clear them
    auto flush = [&](int cloud) {
      named_bar(1, 256);
      for (int c = tid_grp; c < c3; c += 256) {
        unsigned long long k = wkey_all[c];
        k = max(k, wkey_all[c3 + c]);
        k = max(k, wkey_all[2 * c3 + c]);
        k = max(k, wkey_all[3 * c3 + c]);
        atomicMax(pool_keys + (int64_t)cloud * c3 + c, k);
      }
      for (int c = tid_grp; c < 4 * c3; c += 256) reinterpret_cast<unsigned long long*>(smem + L.wkey)[c] = 0ull;
      named_bar(1, 256);
    };
    if (!dump.xhat2) {
      for (int c = tid_grp; c < 4 * c3; c += 256) reinterpret_cast<unsigned long long*>(smem + L.wkey)[c] = 0ull;
      named_bar(1, 256);
    }
    int cur_cloud = n_local > 0 ? (int)(tile0 / tiles_per_cloud) : 0;

    for (int i = 0; i < n_local; ++i) {
      const int64_t tile = tile0 + i;
      const int cloud = (int)(tile / tiles_per_cloud);
      const uint32_t idx_base = (uint32_t)((int)(tile % tiles_per_cloud) * 128 + q * 32);
      mbar_wait(BAR(F2), i & 1);
      trace_e(300);
      tc_fence_after();
      uint32_t v[32];
      float rstd = 1.f, nmr = 0.f;
      if (!(dbg & 4)) {
        // the accumulator holds y' = sign(g2) * y (channels with g2 >= 0 first): sum of squares is unchanged, the
        // mean needs the signs back
        float sum = 0.f, q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          const uint32_t t0 = ring(i, half) + lane_off + (uint32_t)(h * cq);
          const int cbase = half * ch2 + h * cq;
          for (int ch = 0; ch < cq; ch += 32) {
            tmem_ld32(t0 + ch, v);
            const int nb = n_pos - (cbase + ch);  // columns [0, nb) of this chunk are positive-gamma channels
            float s4[4] = {0.f, 0.f, 0.f, 0.f};
            if (nb >= 32 || nb <= 0) {
              stats32(v, s4, q4);
              const float cs = (s4[0] + s4[1]) + (s4[2] + s4[3]);
              sum += nb > 0 ? cs : -cs;
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float y = __uint_as_float(v[j]);
                s4[j & 3] += (j < nb) ? y : -y;
                q4[j & 3] = fmaf(y, y, q4[j & 3]);
              }
              sum += (s4[0] + s4[1]) + (s4[2] + s4[3]);
            }
          }
        }
        float sq = (q4[0] + q4[1]) + (q4[2] + q4[3]);
        *stat_mine = make_float2(sum, sq);
        named_bar(pair_bar, 64);
        const float2 o = *stat_peer;
        sum += o.x;
        sq += o.y;
        const float mean = sum / (float)c3;
        rstd = rsqrtf(fmaxf(sq / (float)c3 - mean * mean, 0.f) + ln_eps);
        nmr = -mean * rstd;
      }
      trace_e(310);
      if (dump.xhat2) {
        float* d_x2 = dump.xhat2 + (tile * 128 + q * 32) * (int64_t)c3;
        float* dscr = reinterpret_cast<float*>(smem + L.tr) + (warp - 4) * 512;
        if (h == 0) dump.rstd2[tile * 128 + row] = rstd;
        const float dm1 = dump.m1 ? dump.m1[tile * 128 + row] : 0.f, dm2 = dump.m1 ? dump.m2[tile * 128 + row] : 0.f;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          const uint32_t t0 = ring(i, half) + lane_off + (uint32_t)(h * cq);
          const int cbase = half * ch2 + h * cq;
          for (int ch = 0; ch < cq; ch += 32) {
            tmem_ld32(t0 + ch, v);
            if (dump.m1) {
              // rstd * (-m1 - (y*rstd + nmr) * m2) = y * (-rstd^2 m2) + rstd * (-m1 - nmr * m2)
              const float ka = -rstd * rstd * dm2, kb = rstd * (-dm1 - nmr * dm2);
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(fmaf(__uint_as_float(v[j]), ka, kb));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(fmaf(__uint_as_float(v[j]), rstd, nmr));
            }
            dump_chunk32(dscr, d_x2 + cbase + ch, c3, v, lane);
          }
          tc_fence_before();
          mbar_arrive(BAR((half ? DB : DA) + (i & 1)));
        }
        continue;
      }
      if (cloud != cur_cloud) {
        flush(cur_cloud);
        cur_cloud = cloud;
      }
      // Max over the warp's 32 points per channel without cross-lane reductions: every lane packs
      // (bits(z + 16) & ~31) | (31 - lane), z = sign(g2) * xhat -- z + 16 is a positive float, so integer order is
      // value order, and the low 5 mantissa bits carry the lane so ties resolve to the smallest point index --
      // writes its 32 keys as one row of a warp-private 32x32 tile (XOR-swizzled float4 chunks, conflict-free), then
      // reads back one COLUMN.  Padding rows (n >= N) are staged as copies of the cloud's point 0, so they can only
      // tie with a real point and the smallest-index rule drops them: no masking.  |g2|, b2 and the ReLU are applied
      // once per (cloud, channel) in the finalize kernel (they commute with the max).
      const float add_pos = nmr + kKeyBias, add_neg = kKeyBias - nmr;
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        const uint32_t t0 = ring(i, half) + lane_off + (uint32_t)(h * cq);
        const int cbase = half * ch2 + h * cq;
        for (int ch = 0; ch < cq; ch += 32) {
          if (dbg & 1) break;
          tmem_ld32(t0 + ch, v);
          const int nb = n_pos - (cbase + ch);
          if (nb >= 32 || nb <= 0) {
            const float add = nb > 0 ? add_pos : add_neg;
            const uint64_t r2 = pk2f(rstd, rstd), a2 = pk2f(add, add);
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              uint32_t z0, z1;
              unpk2(fma2(pk2(v[j], v[j + 1]), r2, a2), z0, z1);
              asm("lop3.b32 %0, %1, 0xffffffe0, %2, 0xea;" : "=r"(v[j]) : "r"(z0), "r"(lane_tag));  // (z & ~31) | tag
              asm("lop3.b32 %0, %1, 0xffffffe0, %2, 0xea;" : "=r"(v[j + 1]) : "r"(z1), "r"(lane_tag));
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              asm("lop3.b32 %0, %1, 0xffffffe0, %2, 0xea;"
                  : "=r"(v[j])
                  : "r"(__float_as_uint(fmaf(__uint_as_float(v[j]), rstd, (j < nb) ? add_pos : add_neg))), "r"(lane_tag));
          }
          __syncwarp();  // previous chunk's column reads are done
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4)
            *reinterpret_cast<uint4*>(tr + st_off[c4]) = make_uint4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]);
          __syncwarp();
          uint32_t m0 = 0, m1 = 0;  // two independent max chains
#pragma unroll
          for (int p = 0; p < 32; p += 4) {
            m0 = __vimax3_u32(m0, tr[p * 32 + ld_off[p & 7]], tr[(p + 1) * 32 + ld_off[(p + 1) & 7]]);
            m1 = __vimax3_u32(m1, tr[(p + 2) * 32 + ld_off[(p + 2) & 7]], tr[(p + 3) * 32 + ld_off[(p + 3) & 7]]);
          }
          const uint32_t m = max(m0, m1);
          // lane now owns (permuted) channel cbase + ch + lane: m = max key over this warp's 32 points
          const uint32_t idx = idx_base + (31u - (m & 31u));
          const unsigned long long key = ((unsigned long long)(m & ~31u) << 32) | (unsigned long long)(0xFFFFFFFFu - idx);
          unsigned long long* slot = wkey + cbase + ch + lane;  // owned by this lane: running maximum over the cloud's tiles
          *slot = max(*slot, key);
        }
        tc_fence_before();
        mbar_arrive(BAR((half ? DB : DA) + (i & 1)));  // ring slot drained: the next tile's MMA may overwrite it
      }
      trace_e(400);
    }
    if (!dump.xhat2 && n_local > 0) flush(cur_cloud);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
  }
}

// pooled[r][c] = relu(|g2[c]| * max_p z + b2[c]) and the argmax, from the packed (key, ~index) maxima stored at the
// permuted channel position pos[c]
__global__ void pool_finalize_kernel(unsigned long long* __restrict__ keys, int R, int c3,
                                     const float* __restrict__ g2, const float* __restrict__ be2,
                                     const int* __restrict__ pos, float* __restrict__ pooled,
                                     int32_t* __restrict__ argmax) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)R * c3) return;
  const int r = (int)(i / c3), c = (int)(i % c3);
  unsigned long long* kp = keys + (int64_t)r * c3 + pos[c];
  const unsigned long long k = *kp;
  *kp = 0ull;  // leave the scratch zeroed for the next call (pos is a permutation: every key is read exactly once)
  const float z = __uint_as_float((uint32_t)(k >> 32)) - kKeyBias;
  pooled[i] = fmaxf(fmaf(fabsf(g2[c]), z, be2[c]), 0.f);
  if (argmax) argmax[i] = (int32_t)(0xFFFFFFFFu - (uint32_t)k);
}

// xha[a] = xh[src[a]] for the compacted rows a < *count: 32 bytes (two 16-byte core-matrix rows) per point
__global__ void gather_xh_kernel(const char* __restrict__ xh, const int32_t* __restrict__ src, const int* __restrict__ count,
                                 char* __restrict__ xha) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= *count) return;
  const int pidx = src[a];
  const char* sp_ = xh + (int64_t)(pidx >> 7) * kTileBytes + ((pidx & 127) >> 3) * 256 + (pidx & 7) * 16;
  char* dp = xha + (int64_t)(a >> 7) * kTileBytes + ((a & 127) >> 3) * 256 + (a & 7) * 16;
  *reinterpret_cast<uint4*>(dp) = *reinterpret_cast<const uint4*>(sp_);
  *reinterpret_cast<uint4*>(dp + 128) = *reinterpret_cast<const uint4*>(sp_ + 128);
}

// byte offset of element (row n, k) in a K-major no-swizzle image with K columns
__device__ __forceinline__ uint32_t img_off(int n, int k, int K) {
  return (uint32_t)((n >> 3) * (K * 16) + (k >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2);
}

__global__ void __launch_bounds__(256)
pack_weights_kernel(const float* __restrict__ w0, const float* __restrict__ b0, const float* __restrict__ w1,
                    const float* __restrict__ g1, const float* __restrict__ be1, const float* __restrict__ w2,
                    const float* __restrict__ g2, const float* __restrict__ be2, int C, int c1, int c2, int c3,
                    int rgb_u8, char* __restrict__ out) {
  // every block derives the channel permutation of layer 2 (g2 >= 0 first, stable): c3 <= 256 = blockDim
  __shared__ int s_pos[256];
  __shared__ int s_npos;
  {
    const int c = threadIdx.x;
    if (c < c3) {
      int before_same = 0, n_pos = 0;
      const bool mine = g2[c] >= 0.f;
      for (int o = 0; o < c3; ++o) {
        const bool p = g2[o] >= 0.f;
        n_pos += p;
        if (o < c && p == mine) ++before_same;
      }
      s_pos[c] = mine ? before_same : n_pos + before_same;
      if (c == 0) s_npos = n_pos;
    }
  }
  __syncthreads();
  const WpackLayout W = make_wpack(c1, c2, c3);
  const int n0 = c1 * 16, n1 = c2 * c1, n2 = c3 * c2, n3 = 2 * c2 + 2 * c3;
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
    const int e = i - n0, n = e / c1, k = e % c1;
    *reinterpret_cast<__nv_bfloat16*>(out + W.w1 + img_off(n, k, c1)) = __float2bfloat16(w1[e]);
  } else if (i < n0 + n1 + n2) {
    const int e = i - n0 - n1, n = e / c2, k = e % c2;
    const float v = w2[e];
    *reinterpret_cast<__nv_bfloat16*>(out + W.w2 + img_off(n, k, c2)) = __float2bfloat16(v);
    *reinterpret_cast<__nv_bfloat16*>(out + W.w2s + img_off(s_pos[n], k, c2)) = __float2bfloat16(g2[n] >= 0.f ? v : -v);
  } else if (i < n0 + n1 + n2 + n3) {
    const int e = i - n0 - n1 - n2;
    float v;
    if (e < c2) v = g1[e];
    else if (e < 2 * c2) v = be1[e - c2];
    else if (e < 2 * c2 + c3) v = g2[e - 2 * c2];
    else v = be2[e - 2 * c2 - c3];
    reinterpret_cast<float*>(out + W.prm)[e] = v;
  } else if (i < n0 + n1 + n2 + n3 + c3) {
    const int c = i - (n0 + n1 + n2 + n3);
    reinterpret_cast<int*>(out + W.pos)[c] = s_pos[c];
    if (c == 0) *reinterpret_cast<int*>(out + W.npos) = s_npos;
  }
}

static bool shapes_ok(int c1, int c2, int c3) {
  auto ok = [](int c, int m) { return c >= m && c <= 256 && c % m == 0; };
  // TMEM: max(c1,c2) columns for layers 0/1 plus a ring of three c3/2-column half-accumulators for layer 2;
  // channels are split between warp pairs in whole 32-column chunks (layers 0/1: halves, layer 2: quarters)
  return ok(c1, 64) && ok(c2, 64) && ok(c3, 128) && std::max(c1, c2) + 3 * (c3 / 2) <= 512 &&
         make_layout(c1, c2, c3).total <= 227 * 1024;
}

bool recompute_supported(int c1, int c2, int c3) { return shapes_ok(c1, c2, c3); }

// Recompute of the compacted active points on the fused tensor-core kernel (called by pcrl_pointnet_bwd in fast mode):
// gathers their bf16 tile rows, runs the three layers and writes h0 / xhat1 / rstd1 / h1 / xhat2 / rstd2 (fp32).
int recompute_active_tc(const void* xh, const void* wpack, const int32_t* src, const int* count_dev, int capacity,
                        int c1, int c2, int c3, float ln_eps, void* xha_scratch, float* h0, float* xhat1, float* rstd1,
                        float* h1, float* xhat2, float* rstd2, const float* m1, const float* m2, cudaStream_t st) {
  if (!shapes_ok(c1, c2, c3)) {
    set_error("recompute_active_tc: widths unsupported by the tcgen05 path");
    return PCRL_EUNSUPPORTED;
  }
  if (src) {  // src == nullptr: the caller has already gathered the tile rows into xha_scratch
    gather_xh_kernel<<<(unsigned)cdiv(capacity, 256), 256, 0, st>>>((const char*)xh, src, count_dev, (char*)xha_scratch);
    PCRL_CHECK_LAUNCH();
  }
  const SmemLayout L = make_layout(c1, c2, c3);
  PCRL_CHECK_CUDA(cudaFuncSetAttribute(pointnet_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  const int n_tiles = (int)cdiv(capacity, 128);
  DumpOut d{count_dev, h0, xhat1, rstd1, h1, xhat2, rstd2, m1, m2};
  pointnet_fwd_tc_kernel<<<std::min(sm_count(), n_tiles), kThreads, L.total, st>>>(
      (const char*)xha_scratch, (const char*)wpack, n_tiles, /*tiles_per_cloud=*/1 << 30, /*src_cloud_stride=*/1, c1, c2, c3, ln_eps, 0,
      nullptr, d);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

}  // namespace tc
}  // namespace pcrl

using namespace pcrl;

extern "C" {

// not part of the public header: profiling switches for tools/probe_fwd.py
int pcrl_debug_set_fwd_flags(int flags) {
  int zero = 0;
  PCRL_CHECK_CUDA(cudaMemcpyToSymbol(tc::g_dbg, &flags, sizeof(int)));
  PCRL_CHECK_CUDA(cudaMemcpyToSymbol(tc::g_trace_n, &zero, sizeof(int)));
  static long long zeros[3 * 1024] = {0};
  PCRL_CHECK_CUDA(cudaMemcpyToSymbol(tc::g_trace, zeros, sizeof(zeros)));
  return PCRL_OK;
}
int pcrl_debug_get_trace(long long* out_host, int max_events) {
  PCRL_CHECK_CUDA(cudaMemcpyFromSymbol(out_host, tc::g_trace, sizeof(long long) * 3 * 1024));
  return 3 * 512;
}

// which fused forward pcrl_pointnet_fwd_bf16 runs: 0 = newest the shape supports, 1 = first generation (pointnet_tc.cu)
static int g_fwd_version = 0;
int pcrl_debug_set_fwd_version(int v) {
  g_fwd_version = v;
  return PCRL_OK;
}
int pcrl_debug_set_fwd2_flags(int flags) { return tc2::set_debug_flags(flags); }
int pcrl_debug_get_trace2(long long* out_host) { return tc2::get_trace(out_host); }
// first-generation image (also read by the backward's recompute), then the second-generation image, 128-byte aligned
static int64_t wpack1_bytes(int c1, int c2, int c3) {
  return tc::shapes_ok(c1, c2, c3) ? align_up((int64_t)tc::make_wpack(c1, c2, c3).total, 128) : 0;
}

int64_t pcrl_pointnet_wpack_bytes(int c1, int c2, int c3) { return wpack1_bytes(c1, c2, c3) + tc2::wpack_bytes(c1, c2, c3); }

int pcrl_pointnet_pack_weights_part(const float* w0, const float* b0, const float* w1, const float* g1, const float* be1,
                                    const float* w2, const float* g2, const float* be2, int C, int c1, int c2, int c3,
                                    int rgb_u8, int which, void* wpack, void* stream) {
  PCRL_CHECK_ARG(w0 && b0 && w1 && g1 && be1 && w2 && g2 && be2 && wpack);
  const bool v1 = tc::shapes_ok(c1, c2, c3), v2 = tc2::shapes_ok(c1, c2, c3);
  if (C + 4 > 16 || !(v1 || v2)) {
    set_error("pcrl_pointnet_pack_weights: shape unsupported by the tcgen05 path (C=%d <= 12; widths (%d,%d,%d): c1 in {64,128}, "
              "c2 = 128, c3 = 128 or a multiple of 256 up to 2048 -- or c1, c2 multiples of 64 and c3 of 128 up to 256 with "
              "max(c1,c2) + 1.5*c3 <= 512 TMEM columns); use the tf32 / fp32 path", C, c1, c2, c3);
    return PCRL_EUNSUPPORTED;
  }
  // the first-generation image feeds the backward's recompute (shapes it covers; wider PointNets recompute on the TF32
  // GEMMs instead), and the forward too where the second generation does not cover the shape
  if (v1 && ((which & PCRL_WPACK_BWD) || ((which & PCRL_WPACK_FWD) && !v2))) {
    const int n = c1 * 16 + c2 * c1 + c3 * c2 + 2 * c2 + 2 * c3 + c3;
    tc::pack_weights_kernel<<<(unsigned)cdiv(n, 256), 256, 0, as_stream(stream)>>>(w0, b0, w1, g1, be1, w2, g2, be2, C,
                                                                                     c1, c2, c3, rgb_u8, (char*)wpack);
    PCRL_CHECK_LAUNCH();
  }
  if ((which & PCRL_WPACK_FWD) && v2)
    return tc2::pack(w0, b0, w1, g1, be1, w2, g2, be2, C, c1, c2, c3, rgb_u8, (char*)wpack + wpack1_bytes(c1, c2, c3),
                     as_stream(stream));
  return PCRL_OK;
}

int pcrl_pointnet_pack_weights(const float* w0, const float* b0, const float* w1, const float* g1, const float* be1,
                               const float* w2, const float* g2, const float* be2, int C, int c1, int c2, int c3,
                               int rgb_u8, void* wpack, void* stream) {
  return pcrl_pointnet_pack_weights_part(w0, b0, w1, g1, be1, w2, g2, be2, C, c1, c2, c3, rgb_u8,
                                         PCRL_WPACK_FWD | PCRL_WPACK_BWD, wpack, stream);
}

int pcrl_pointnet_fwd_bf16(const void* xh, int R, int N, int NP, const void* wpack, int c1, int c2, int c3,
                           float ln_eps, uint64_t* pool_keys, float* pooled, int32_t* argmax, void* stream) {
  return pcrl_pointnet_fwd_bf16_strided(xh, R, 1, N, NP, wpack, c1, c2, c3, ln_eps, pool_keys, pooled, argmax, stream);
}

int pcrl_pointnet_fwd_bf16_strided(const void* xh, int R, int src_cloud_stride, int N, int NP, const void* wpack, int c1,
                                   int c2, int c3, float ln_eps, uint64_t* pool_keys, float* pooled, int32_t* argmax,
                                   void* stream) {
  PCRL_CHECK_ARG(xh && wpack && pool_keys && pooled && R >= 0 && N > 0 && NP >= N && NP % 128 == 0 && src_cloud_stride >= 1);
  if (!tc::shapes_ok(c1, c2, c3) && !tc2::shapes_ok(c1, c2, c3)) {
    set_error("pcrl_pointnet_fwd_bf16: widths (%d,%d,%d) unsupported (multiples of 64/64/128 up to 256, max(c1,c2) + 1.5*c3 <= 512, smem budget)", c1,
              c2, c3);
    return PCRL_EUNSUPPORTED;
  }
  if (R == 0) return PCRL_OK;
  cudaStream_t st = as_stream(stream);
  if (g_fwd_version != 1 && tc2::shapes_ok(c1, c2, c3))
    return tc2::forward(xh, R, src_cloud_stride, NP, (const char*)wpack + wpack1_bytes(c1, c2, c3), c1, c2, c3, ln_eps,
                        pool_keys, pooled, argmax, st);
  const tc::SmemLayout L = tc::make_layout(c1, c2, c3);
  PCRL_CHECK_CUDA(cudaFuncSetAttribute(tc::pointnet_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       227 * 1024));
  const int tiles_per_cloud = NP / 128;
  const int n_tiles = R * tiles_per_cloud;
  const int grid = std::min(sm_count(), n_tiles);
  tc::pointnet_fwd_tc_kernel<<<grid, tc::kThreads, L.total, st>>>((const char*)xh, (const char*)wpack, n_tiles,
                                                                  tiles_per_cloud, src_cloud_stride, c1, c2, c3, ln_eps,
                                                                  argmax != nullptr,
                                                                  (unsigned long long*)pool_keys, tc::DumpOut{});
  PCRL_CHECK_LAUNCH();
  const int64_t n = (int64_t)R * c3;
  const tc::WpackLayout W = tc::make_wpack(c1, c2, c3);
  const float* prm = reinterpret_cast<const float*>((const char*)wpack + W.prm);
  tc::pool_finalize_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(
      (unsigned long long*)pool_keys, R, c3, prm + 2 * c2, prm + 2 * c2 + c3,
      reinterpret_cast<const int*>((const char*)wpack + W.pos), pooled, argmax);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

}  // extern "C"
