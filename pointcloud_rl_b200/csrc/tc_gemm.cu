// Generic TF32 GEMM on tcgen05 tensor cores, operands streamed by TMA (tensor maps over the plain
// row-major fp32 matrices, 128B swizzle), accumulator in TMEM:
//
//     C[i][j] (op)= sum_l A(i,l) * B(l,j) (+ bias[j]) (ReLU)
//
// A is given either K-major (A[i*lda + l], the forward `x W^T`) or MN-major (A[l*lda + i], the `dY^T X`
// weight gradient); likewise B (B[j*ldb + l] K-major / B[l*ldb + j] MN-major).  Both majors are native
// UMMA operand layouts, so the backward GEMMs need no transposed copies.  fp32 inputs are consumed as
// TF32 directly (no conversion pass).  Used for the actor/critic MLP layers (fwd + bwd) and for the
// compacted PointNet backward; the exact-fp32 parity path stays on sgemm.cu.
//
// One 128 x BN output tile per CTA (optionally a K-split slice of it): warp 0 = TMA producer,
// warp 1 = MMA issuer, warps 2-5 = epilogue (TMEM -> registers -> bias/ReLU -> global).
//
// Small-M problems (the MLP heads: M = 256..512 rows, K up to 1024) have too few output tiles to fill 148 SMs and a
// K loop that is a serial chain of shared-memory-operand MMAs on each of them.  Those run in "cluster split-K" mode:
// a thread-block cluster of S CTAs owns one (wide) output tile, each CTA contracts 1/S of K into its own TMEM
// accumulator, writes the partial tile to its shared memory, and after a cluster barrier every CTA sums a 128/S-row
// band of the tile over all S peers through distributed shared memory and applies the epilogue.  No atomics, no
// global workspace, fixed summation order.
#include <cuda.h>

#include "common.cuh"
#include "tc_gemm.cuh"

namespace pcrl {
namespace tcg {

constexpr int BM = 128;
constexpr int BK = 32;  // fp32 elements: one 128-byte swizzle row
constexpr int kThreads = 192;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptors (version 1 = sm_100)
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;  // 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
  return d;
}
// K-major fp32/TF32 tile: 128 B rows, 16 B-granular 128B swizzle, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t saddr) { return desc_sw128(saddr, 16, 1024, 2); }
// MN-major TF32 tile: the only legal layout for 32-bit MN-major operands is the 128B swizzle with a 32 B
// base (Swizzle<2,5,2>, atoms of 4 contraction rows x 128 B): SBO = 512 B between 4-row groups along K,
// LBO = bytes between successive 32-element blocks along M/N (one TMA box each).
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t saddr) { return desc_sw128(saddr, 4096, 512, 1); }
__device__ __forceinline__ uint32_t idesc_tf32(int n, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
      "%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  __syncwarp();  // the single-lane producer / issuer roles rejoin their warps first (.aligned barrier)
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float4 ld_peer_v4(uint32_t saddr, uint32_t rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(saddr), "r"(rank));
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(ra));
  return v;
}

struct Params {
  float* C;
  int64_t ldc;
  const float* bias;
  int M, N, K;
  int bn;        // output tile width (16..128, multiple of 16)
  int a_mn, b_mn;
  int relu;
  int mode;      // 0 store, 1 C += (owned tile), 2 atomic add (split-K)
  int split_k;
  int cluster_k;  // > 1: cluster split-K mode (one tile per cluster of cluster_k CTAs, DSMEM reduction)
  int stages;
  const int* k_dev;  // optional device-side bound on the contraction length
  const int* m_dev;  // optional device-side bound on M
  const float* mask;  // optional [M, N] post-activation tensor: C = (mask > 0) ? value : 0  (fused ReLU backward)
  int64_t ldmask;
};

// Persistent: each CTA walks output tiles t = blockIdx.x, += gridDim.x.  The smem ring and its phases run
// continuously across tiles, and the accumulator is double-buffered in TMEM (2 x bn columns) so the epilogue
// of tile n overlaps the loads and MMAs of tile n+1.
template <int TMEM_COLS>
__global__ void __launch_bounds__(kThreads, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, Params P) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // 1024-byte alignment for the 128B swizzle atoms
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t a_bytes = BM * BK * 4, b_bytes = (uint32_t)P.bn * BK * 4;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const uint32_t bar0 = sbase + (uint32_t)P.stages * stage_bytes;  // full[s], empty[s], acc_full[2], acc_empty[2]
  auto FULL = [&](int s) { return bar0 + 8u * s; };
  auto EMPTY = [&](int s) { return bar0 + 8u * (P.stages + s); };
  auto ACC_FULL = [&](int b) { return bar0 + 8u * (2 * P.stages + b); };
  auto ACC_EMPTY = [&](int b) { return bar0 + 8u * (2 * P.stages + 2 + b); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + (size_t)P.stages * stage_bytes + 8 * (2 * P.stages + 4));
  const uint32_t epi_off = (uint32_t)(P.stages * stage_bytes + 8 * (2 * P.stages + 5) + 16 + 15) & ~15u;  // 4 x 2 KB epilogue scratch

  const int K = P.k_dev ? min(P.K, *P.k_dev) : P.K;
  const int Mlim = P.m_dev ? min(P.M, *P.m_dev) : P.M;
  const int kb_total = (K + BK - 1) / BK;
  const int kb_per = (kb_total + P.split_k - 1) / P.split_k;
  const int nt = (P.N + P.bn - 1) / P.bn, mt = (P.M + BM - 1) / BM;
  const int S = P.cluster_k;
  const int total_tiles = S > 1 ? (int)blockIdx.x + 1 : nt * mt * P.split_k;  // cluster mode: exactly one tile per CTA
  const int kb_per_c = (kb_total + S - 1) / S;
  // tile -> (k-split z, M tile, N tile); N fastest so concurrently running CTAs share the same A rows in L2
  auto tile_coords = [&](int t, int& z, int& i0, int& j0, int& kb0, int& nkb) {
    if (S > 1) {  // the S CTAs of a cluster share tile t / S and take consecutive K slices
      z = t % S;
      const int rem = t / S;
      i0 = (rem / nt) * BM;
      j0 = (rem % nt) * P.bn;
      kb0 = z * kb_per_c;
      nkb = max(min(kb_total, kb0 + kb_per_c) - kb0, 0);
      return nkb > 0;
    }
    z = t / (nt * mt);
    const int rem = t - z * nt * mt;
    i0 = (rem / nt) * BM;
    j0 = (rem % nt) * P.bn;
    kb0 = z * kb_per;
    nkb = max(min(kb_total, kb0 + kb_per) - kb0, 0);
    return i0 < Mlim && nkb > 0;  // tiles past the device-side row count / empty K slices are skipped by every role
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < P.stages; ++s) {
      mbar_init(FULL(s), 1);
      mbar_init(EMPTY(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(ACC_FULL(b), 1);
      mbar_init(ACC_EMPTY(b), 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32((const void*)tmem_ptr_smem)),
                 "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t it = 0;  // running k-block count: ring slot = it % stages, phase = (it / stages) & 1
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int z, i0, j0, kb0, nkb;
        if (!tile_coords(t, z, i0, j0, kb0, nkb)) continue;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % P.stages;
          mbar_wait(EMPTY(s), ((it / P.stages) & 1) ^ 1);  // passes immediately on the first lap
          mbar_expect_tx(FULL(s), stage_bytes);
          const uint32_t sa = sbase + s * stage_bytes, sb = sa + a_bytes;
          const int l0 = (kb0 + kb) * BK;
          if (!P.a_mn) {
            tma_load_2d(sa, &map_a, l0, i0, FULL(s));  // box {32 (K), 128 rows}
          } else {
            for (int b = 0; b < BM / 32; ++b) tma_load_2d(sa + b * 4096, &map_a, i0 + b * 32, l0, FULL(s));  // {32 (M), 32 K-rows}
          }
          if (!P.b_mn) {
            tma_load_2d(sb, &map_b, l0, j0, FULL(s));  // box {32 (K), bn rows}
          } else {
            for (int b = 0; b < P.bn / 32; ++b) tma_load_2d(sb + b * 4096, &map_b, j0 + b * 32, l0, FULL(s));
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = idesc_tf32(P.bn, P.a_mn, P.b_mn);
      uint32_t it = 0, n = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int z, i0, j0, kb0, nkb;
        if (!tile_coords(t, z, i0, j0, kb0, nkb)) continue;
        const uint32_t buf = n & 1;
        mbar_wait(ACC_EMPTY(buf), ((n >> 1) & 1) ^ 1);  // epilogue drained this accumulator (free on first use)
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * (uint32_t)P.bn;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % P.stages;
          mbar_wait(FULL(s), (it / P.stages) & 1);
          tc_fence_after();
          const uint32_t sa = sbase + s * stage_bytes, sb = sa + a_bytes;
#pragma unroll
          for (int ks = 0; ks < BK / 8; ++ks) {
            // K-major: +32 B inside the 128 B swizzle row per 8-element K step; MN-major: next 8-row group
            const uint64_t ad = P.a_mn ? desc_mnmajor(sa + ks * 1024) : desc_kmajor(sa + ks * 32);
            const uint64_t bd = P.b_mn ? desc_mnmajor(sb + ks * 1024) : desc_kmajor(sb + ks * 32);
            mma_tf32(d_tmem, ad, bd, idesc, (kb | ks) != 0);
          }
          mma_commit(EMPTY(s));
        }
        mma_commit(ACC_FULL(buf));
        ++n;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: warp (2..5) -> TMEM lane quadrant warp%4
    const int q = warp & 3;
    uint32_t n = 0;
    uint32_t v[32];
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int z, i0, j0, kb0, nkb;
      if (!tile_coords(t, z, i0, j0, kb0, nkb)) continue;
      const uint32_t buf = n & 1;
      // ReLU-backward mask of this tile as bits, fetched BEFORE waiting for the accumulator: the loads do not depend on
      // the MMAs, and issued inside the store loop they are eight dependent HBM round trips per tile (+35 us on the
      // 100 k-row dgrad GEMMs).  Bit (k*4+e) of mb[c/32][half]: mask > 0 at row k*8 + lane/4, column
      // c + half*16 + (lane%4)*4 + e -- the post-transpose ownership used below.  Loads are unconditional (clamped
      // addresses) so that all eight of a chunk are in flight together.
      uint32_t mb00 = 0, mb01 = 0, mb10 = 0, mb11 = 0, mb20 = 0, mb21 = 0, mb30 = 0, mb31 = 0;
      const bool mask_fast = S == 1 && P.mask && P.bn <= 128 && (P.ldmask & 3) == 0 && (P.N & 3) == 0 && P.N >= 4 &&
                             (reinterpret_cast<uintptr_t>(P.mask) & 15) == 0 && (j0 & 3) == 0;
      if (mask_fast) {
        const int rr_ = lane >> 2, rs_ = lane & 3;
        auto fetch = [&](int cc, uint32_t& b0, uint32_t& b1) {
          if (cc * 32 >= P.bn) return;
          float4 m[2][4];
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int col = min(j0 + cc * 32 + half * 16 + rs_ * 4, P.N - 4);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int grow = min(i0 + q * 32 + k * 8 + rr_, P.M - 1);
              m[half][k] = __ldg(reinterpret_cast<const float4*>(P.mask + (int64_t)grow * P.ldmask + col));
            }
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            b0 |= ((m[0][k].x > 0.f ? 1u : 0u) | (m[0][k].y > 0.f ? 2u : 0u) | (m[0][k].z > 0.f ? 4u : 0u) | (m[0][k].w > 0.f ? 8u : 0u)) << (4 * k);
            b1 |= ((m[1][k].x > 0.f ? 1u : 0u) | (m[1][k].y > 0.f ? 2u : 0u) | (m[1][k].z > 0.f ? 4u : 0u) | (m[1][k].w > 0.f ? 8u : 0u)) << (4 * k);
          }
        };
        fetch(0, mb00, mb01);
        fetch(1, mb10, mb11);
        fetch(2, mb20, mb21);
        fetch(3, mb30, mb31);
      }
      mbar_wait(ACC_FULL(buf), (n >> 1) & 1);
      tc_fence_after();
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + buf * (uint32_t)P.bn;
      if (S > 1) {
        // partial tile -> shared memory (the pipeline ring is idle now), row pitch bn + 4 floats: conflict-free
        float* red = reinterpret_cast<float*>(smem) + (size_t)(q * 32 + lane) * (P.bn + 4);
        for (int c = 0; c < P.bn; c += 32) {
          tmem_ld32(tbase + c, v);
#pragma unroll
          for (int j4 = 0; j4 < 32; j4 += 4)
            *reinterpret_cast<uint4*>(red + c + j4) = make_uint4(v[j4], v[j4 + 1], v[j4 + 2], v[j4 + 3]);
        }
        ++n;
        continue;
      }
      // After tcgen05.ld a thread holds 32 consecutive columns of ITS row: stored directly, every instruction would
      // touch 32 different rows with 16 bytes each.  Each 32x32 chunk goes through a warp-private, XOR-swizzled 2 KB
      // scratch tile (two 16-column halves) so that a lane ends up with a float4 of row k*8 + lane/4, columns
      // (lane%4)*4.. : bias / ReLU / mask / accumulate then run on 64-byte row segments, loads and stores coalesced.
      float* scr = reinterpret_cast<float*>(smem + epi_off) + q * 512;
      const int wsw = (lane >> 1) & 3, rr = lane >> 2, rs = lane & 3;
      for (int c = 0; c < P.bn; c += 32) {
        if (c + 32 <= P.bn) {
          tmem_ld32(tbase + c, v);
        } else {
          // bn == 16 / 48 / ...: a 16-column tail
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
              : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
              : "r"(tbase + c));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
        const int ncol = min(32, P.bn - c);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          if (half * 16 >= ncol) break;
          __syncwarp();
#pragma unroll
          for (int sl = 0; sl < 4; ++sl)
            *reinterpret_cast<uint4*>(scr + lane * 16 + ((sl ^ wsw) << 2)) =
                make_uint4(v[half * 16 + 4 * sl], v[half * 16 + 4 * sl + 1], v[half * 16 + 4 * sl + 2], v[half * 16 + 4 * sl + 3]);
          __syncwarp();
          const int col = j0 + c + half * 16 + rs * 4;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int r = k * 8 + rr;
            const int grow = i0 + q * 32 + r;
            const float4 a4 = *reinterpret_cast<const float4*>(scr + r * 16 + ((rs ^ ((r >> 1) & 3)) << 2));
            if (grow >= P.M || col >= P.N) continue;
            float o[4] = {a4.x, a4.y, a4.z, a4.w};
            float* crow = P.C + (int64_t)grow * P.ldc + col;
            const bool full = col + 4 <= P.N;
            float mk[4] = {1.f, 1.f, 1.f, 1.f};
            if (mask_fast) {
              const int cc = c >> 5;
              const uint32_t word = half == 0 ? (cc == 0 ? mb00 : cc == 1 ? mb10 : cc == 2 ? mb20 : mb30)
                                              : (cc == 0 ? mb01 : cc == 1 ? mb11 : cc == 2 ? mb21 : mb31);
              const uint32_t b4 = word >> (4 * k);
#pragma unroll
              for (int e = 0; e < 4; ++e) mk[e] = (b4 >> e) & 1u ? 1.f : 0.f;
            } else if (P.mask) {
              const float* mrow = P.mask + (int64_t)grow * P.ldmask + col;
              if (full && (reinterpret_cast<uintptr_t>(mrow) & 15) == 0) {
                const float4 m4 = __ldg(reinterpret_cast<const float4*>(mrow));
                mk[0] = m4.x; mk[1] = m4.y; mk[2] = m4.z; mk[3] = m4.w;
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  if (col + e < P.N) mk[e] = __ldg(mrow + e);
              }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float x = o[e];
              if (P.bias && z == 0 && col + e < P.N) x += __ldg(P.bias + col + e);
              if (P.relu) x = fmaxf(x, 0.f);
              if (!(mk[e] > 0.f)) x = 0.f;
              o[e] = x;
            }
            if (full && (reinterpret_cast<uintptr_t>(crow) & 15) == 0 && P.mode != 2) {
              float4 w4 = make_float4(o[0], o[1], o[2], o[3]);
              if (P.mode == 1) {
                const float4 old = *reinterpret_cast<const float4*>(crow);
                w4.x += old.x; w4.y += old.y; w4.z += old.z; w4.w += old.w;
              }
              *reinterpret_cast<float4*>(crow) = w4;
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                if (col + e < P.N) {
                  if (P.mode == 0) crow[e] = o[e];
                  else if (P.mode == 1) crow[e] += o[e];
                  else atomicAdd(crow + e, o[e]);
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(ACC_EMPTY(buf));  // all 128 epilogue threads: this accumulator may be overwritten
      ++n;
    }
  }

  if (S > 1) {
    int z, i0, j0, kb0, nkb;
    const bool have = tile_coords((int)blockIdx.x, z, i0, j0, kb0, nkb);
    const int pitch = P.bn + 4;
    if (!have) {  // empty K slice: contribute zeros
      for (int e = threadIdx.x; e < BM * pitch / 4; e += kThreads) reinterpret_cast<float4*>(smem)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    cluster_sync_all();  // every CTA's partial tile is in its shared memory
    const int rows_per = BM / S, c4n = P.bn / 4;
    const int items = rows_per * c4n;
    // two output float4s per thread and step, all peers' loads in flight before the first add (remote shared-memory
    // latency, not bandwidth, bounds this loop)
    for (int e0 = threadIdx.x; e0 < items; e0 += 2 * kThreads) {
      float4 part[2][8];
      uint32_t a[2];
      int rr[2], cc[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int e = min(e0 + u * kThreads, items - 1);
        rr[u] = z * rows_per + e / c4n;
        cc[u] = (e % c4n) * 4;
        a[u] = sbase + (uint32_t)(rr[u] * pitch + cc[u]) * 4;
#pragma unroll
        for (int p = 0; p < 8; ++p)
          if (p < S) part[u][p] = ld_peer_v4(a[u], (uint32_t)p);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (e0 + u * kThreads >= items) break;
        float4 acc = part[u][0];
#pragma unroll
        for (int p = 1; p < 8; ++p)
          if (p < S) { acc.x += part[u][p].x; acc.y += part[u][p].y; acc.z += part[u][p].z; acc.w += part[u][p].w; }
        const int row = i0 + rr[u], col = j0 + cc[u];
        if (row >= P.M || col >= P.N) continue;
        float o[4] = {acc.x, acc.y, acc.z, acc.w};
        float* crow = P.C + (int64_t)row * P.ldc + col;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          if (col + q4 >= P.N) break;
          float x = o[q4];
          if (P.bias) x += __ldg(P.bias + col + q4);
          if (P.relu) x = fmaxf(x, 0.f);
          if (P.mask && !(__ldg(P.mask + (int64_t)row * P.ldmask + col + q4) > 0.f)) x = 0.f;
          if (P.mode != 0) x += crow[q4];
          o[q4] = x;
        }
        if (col + 4 <= P.N && (reinterpret_cast<uintptr_t>(crow) & 15) == 0) {
          *reinterpret_cast<float4*>(crow) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
          for (int q4 = 0; q4 < 4 && col + q4 < P.N; ++q4) crow[q4] = o[q4];
        }
      }
    }
    cluster_sync_all();  // nobody exits while a peer may still read its shared memory
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}

// ---------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D fp32 tensor map: inner dimension `inner` elements (contiguous), `outer` rows of pitch ld elements.
static bool make_map(CUtensorMap* m, const float* base, int64_t inner, int64_t outer, int64_t ld, int box_inner,
                     int box_outer, bool mn_major) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

bool tc_gemm_supported(const TcGemmArgs& g) {
  auto ok = [](const float* p, int64_t ld) { return p && (reinterpret_cast<uintptr_t>(p) & 15) == 0 && ld % 4 == 0 && ld > 0; };
  if (!ok(g.A, g.lda) || !ok(g.B, g.ldb)) return false;
  if (g.M < 1 || g.N < 1 || g.K < 1) return false;
  return get_encode() != nullptr;
}

int launch_tc_gemm(const TcGemmArgs& g, cudaStream_t st) {
  if (!tc_gemm_supported(g)) {
    set_error("tc_gemm: operands must be 16-byte aligned with row pitch %% 4 == 0");
    return PCRL_EINVAL;
  }
  // output tile width: wide tiles when there is enough work, narrower ones to keep >= ~1 wave of CTAs
  int bn = 128;
  const int64_t mt = cdiv(g.M, BM);
  const int sms = sm_count();
  while (bn > 32 && mt * cdiv(g.N, bn) * g.split_k < sms) bn >>= 1;
  if (g.N <= 16) bn = 16;
  else if (g.N <= 32) bn = 32;
  else if (g.N <= 64 && bn > 64) bn = 64;
  if (g.b_mn && bn < 32) bn = 32;  // MN-major B is loaded in 32-column boxes
  if (g.bn_hint) bn = g.bn_hint;
  // cluster split-K for short, wide problems (see the header comment): widest tile the accumulator allows, then as
  // many K slices as keep every slice >= 2 k-blocks and the grid around one wave
  int cluster_k = 1;
  const int64_t kb_total = cdiv(g.K, BK);
  if (!g.k_dev && !g.m_dev && !g.bn_hint && kb_total >= 16 && kb_total <= 64 && g.N >= 32) {
    const int bn_w = g.N >= 128 ? 128 : (int)align_up(g.N, 32);
    const int64_t tiles_w = mt * cdiv(g.N, bn_w);
    int s_k = 1;
    // clusters of 4 place freely on the GPCs; clusters of 8 with ~200 KB of shared memory each do not all fit at once
    while (s_k < 4 && tiles_w * (s_k * 2) <= sms && kb_total / (s_k * 2) >= 4) s_k *= 2;
    if (s_k >= 2) {
      cluster_k = s_k;
      bn = bn_w;
    }
  }
  Params P{};
  P.C = g.C; P.ldc = g.ldc; P.bias = g.bias; P.M = g.M; P.N = g.N; P.K = g.K; P.bn = bn; P.a_mn = g.a_mn; P.b_mn = g.b_mn;
  P.relu = g.relu; P.mode = g.mode; P.split_k = g.split_k; P.k_dev = g.k_dev; P.m_dev = g.m_dev;
  P.mask = g.mask; P.ldmask = g.ldmask; P.cluster_k = cluster_k;
  PCRL_CHECK_ARG(g.split_k >= 1 && (g.split_k == 1 || (g.mode == 2 && !g.relu)));
  if (cluster_k > 1) {
    P.split_k = 1;  // the caller's atomic split-K request is served by the cluster instead: C (+)= reduced tile
    const uint32_t stage_b = BM * BK * 4 + bn * BK * 4;
    const int64_t kb_slice = cdiv(kb_total, cluster_k);
    const size_t red_bytes = (size_t)BM * (bn + 4) * 4;
    int stages = (int)std::min<int64_t>(3, std::max<int64_t>(2, kb_slice));  // <= ~105 KB: two CTAs per SM, the heads' GEMMs run 2-3 at a time
    while ((size_t)stages * stage_b < red_bytes) ++stages;
    while ((size_t)stages * stage_b + 12288 > 224 * 1024 && stages > 2) --stages;
    if ((size_t)stages * stage_b < red_bytes) {
      set_error("tc_gemm: cluster split-K tile does not fit shared memory");
      return PCRL_EINVAL;
    }
    P.stages = stages;
    const size_t smem_c = (size_t)stages * stage_b + 8 * (2 * stages + 5) + 32 + 8192 + 1024;
    CUtensorMap ma, mb;
    bool okm;
    if (!g.a_mn) okm = make_map(&ma, g.A, g.K, g.M, g.lda, BK, BM, false);
    else         okm = make_map(&ma, g.A, g.M, g.K, g.lda, 32, BK, true);
    if (!g.b_mn) okm = okm && make_map(&mb, g.B, g.K, g.N, g.ldb, BK, bn, false);
    else         okm = okm && make_map(&mb, g.B, g.N, g.K, g.ldb, 32, BK, true);
    if (!okm) {
      set_error("tc_gemm: cuTensorMapEncodeTiled failed");
      return PCRL_ECUDA;
    }
    // function attributes are per device: set on every launch (host-side, ~1 us) instead of behind a process-wide flag
    PCRL_CHECK_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(mt * cdiv(g.N, bn) * cluster_k), 1, 1);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = smem_c;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)cluster_k;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    PCRL_CHECK_CUDA(cudaLaunchKernelEx(&cfg, tc_gemm_kernel<256>, ma, mb, P));
    return PCRL_OK;
  }
  const uint32_t stage_bytes = BM * BK * 4 + bn * BK * 4;
  const int64_t n_tiles = mt * cdiv(g.N, bn) * g.split_k;
  const int64_t kb_max = cdiv(cdiv(g.K, BK), g.split_k);
  if (kb_max < 1) return PCRL_OK;
  // ring depth: enough to cover the K loop of a tile (plus prefetch into the next tile), capped by smem; short-K
  // tall-skinny problems keep the ring small so two CTAs fit on an SM
  // ring budget: short K loops keep the ring small so that two CTAs fit on an SM -- tall-skinny streaming problems get
  // twice the CTAs per SM, and the small MLP-head GEMMs, which run two or three at a time on forked streams, do not
  // lock each other out of the SMs
  const int64_t budget = (kb_max <= 8) ? 100 * 1024 : 200 * 1024;
  P.stages = (int)std::min<int64_t>(std::min<int64_t>(8, std::max<int64_t>(3, 2 * kb_max)), std::max<int64_t>(2, budget / stage_bytes));
  const size_t smem = (size_t)P.stages * stage_bytes + 8 * (2 * P.stages + 5) + 32 + 8192 + 1024;
  const int ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(2, (227 * 1024) / smem));

  CUtensorMap ma, mb;
  bool okm;
  if (!g.a_mn) okm = make_map(&ma, g.A, g.K, g.M, g.lda, BK, BM, false);        // A[i*lda + l]
  else         okm = make_map(&ma, g.A, g.M, g.K, g.lda, 32, BK, true);        // A[l*lda + i]
  if (!g.b_mn) okm = okm && make_map(&mb, g.B, g.K, g.N, g.ldb, BK, bn, false); // B[j*ldb + l]
  else         okm = okm && make_map(&mb, g.B, g.N, g.K, g.ldb, 32, BK, true); // B[l*ldb + j]
  if (!okm) {
    set_error("tc_gemm: cuTensorMapEncodeTiled failed");
    return PCRL_ECUDA;
  }
  PCRL_CHECK_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  const unsigned grid = (unsigned)std::min<int64_t>(n_tiles, (int64_t)sms * ctas_per_sm);
  tc_gemm_kernel<256><<<grid, kThreads, smem, st>>>(ma, mb, P);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

}  // namespace tcg
}  // namespace pcrl

using namespace pcrl;

// Test / benchmark hook for the raw GEMM: C = act(op(A) op(B) + bias), see include/pcrl.h
extern "C" int pcrl_gemm_tf32(const float* A, int lda, int a_mn, const float* B, int ldb, int b_mn, const float* bias,
                              float* C, int ldc, int M, int N, int K, int relu, int mode, int split_k, void* stream) {
  PCRL_CHECK_ARG(A && B && C && M > 0 && N > 0 && K > 0);
  tcg::TcGemmArgs g{};
  g.A = A; g.lda = lda; g.a_mn = a_mn; g.B = B; g.ldb = ldb; g.b_mn = b_mn; g.bias = bias; g.C = C; g.ldc = ldc;
  g.M = M; g.N = N; g.K = K; g.relu = relu; g.mode = mode; g.split_k = split_k;
  return tcg::launch_tc_gemm(g, as_stream(stream));
}
