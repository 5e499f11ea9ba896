// Gradient all-reduce over NVLink peer memory: ONE kernel per reduction, no NCCL call on the path.
//
// The reference wraps the agent in DistributedDataParallel (module_utils.py:322-349): bucketed NCCL ring all-reduces.
// Here every rank's flat gradient buffer is "symmetric": allocated with the same layout on every GPU of the node and
// mapped into every peer's address space (torch.distributed._symmetric_memory hands out the peer pointers), so a kernel
// can load a peer's gradients and store into a peer's buffer directly through NVLink / NVSwitch.
//
// Two-shot, in place, deterministic:
//   1. rank r tells every peer "my gradients of this call are final" (a flag store into the peer's flag block) and
//      waits until every peer has said so;
//   2. r owns the r-th slice of the range: it loads that slice from all ranks (all loads in flight, summed in rank
//      order, so every rank would compute the same bits) and stores the sum into the slice of EVERY rank's buffer --
//      in place: the only reader of a peer's copy of slice r is r itself, and it has finished reading;
//   3. r fences, tells every peer "my slice is written everywhere" and waits for the same from everybody; the kernel
//      then ends, so whatever follows on the stream (the fused Adam step) sees the complete sum.
// For the 0.3 MB PointNet gradient this is two flag round trips (~10 us at 8 GPUs where NCCL's LL ring needs 47 us);
// for the Q heads' 10.5 MB every rank moves 7/8 of its slice over NVLink twice instead of 14 ring steps.
//
// Flags are per CHANNEL (a call site: the PointNet range, the Q-head range, each actor bucket ...), so reductions that
// are in flight on different streams at the same time cannot see each other's flags; within a channel the value is the
// call count ("epoch", kept in local device memory and advanced by the kernel itself, so CUDA-graph replays work).
// A rank cannot run ahead: it needs every peer's "ready" of epoch e+1, which a peer only sends after leaving epoch e.
#include <algorithm>

#include "common.cuh"

namespace pcrl {
namespace p2p {

constexpr int kMaxWorld = 16;
constexpr int kChannels = PCRL_P2P_CHANNELS;
constexpr int kThreads = 512;

// flag block of one rank (int32): [channel][0 = ready, 1 = done][source rank, padded to kMaxWorld]
__host__ __device__ inline int flag_index(int channel, int which, int src) { return (channel * 2 + which) * kMaxWorld + src; }

__device__ __forceinline__ void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer_v4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_peer_f(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

struct Args {
  const unsigned long long* bufs;   // [world] base address of every rank's symmetric gradient buffer
  const unsigned long long* flags;  // [world] base address of every rank's flag block
  int rank, world, channel;
  long long off, n;                 // element range [off, off + n) of the buffer
  int* state;                       // local: [channel][0 = epoch, 1 = CTAs finished, 2 = error]
  long long timeout_ns;
};

// spin until flag[src] >= e for every source rank (one thread per source), then the whole CTA proceeds
__device__ __forceinline__ bool wait_all(const int* my_flags, int channel, int which, int world, int e, long long timeout_ns,
                                         int* err) {
  __shared__ int bad;
  if (threadIdx.x == 0) bad = 0;
  __syncthreads();
  if ((int)threadIdx.x < world) {
    const int* f = my_flags + flag_index(channel, which, (int)threadIdx.x);
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(f) < e) {
      if ((long long)(global_ns() - t0) > timeout_ns) {
        bad = 1;
        break;
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
  if (bad && threadIdx.x == 0) atomicExch(err, 1 + which);
  return bad == 0;
}

__global__ void __launch_bounds__(kThreads) allreduce_kernel(Args a) {
  int* st = a.state + a.channel * 4;
  const int e = *reinterpret_cast<volatile int*>(st) + 1;  // this call's epoch (st[0] is advanced by the last CTA at the very end)
  const int* my_flags = reinterpret_cast<const int*>(a.flags[a.rank]);

  // 1. my gradients are final (they were written by earlier kernels of this stream): tell every peer, CTA 0 only
  if (blockIdx.x == 0 && (int)threadIdx.x < a.world) {
    __threadfence_system();
    st_release_sys(reinterpret_cast<int*>(a.flags[threadIdx.x]) + flag_index(a.channel, 0, a.rank), e);
  }
  if (!wait_all(my_flags, a.channel, 0, a.world, e, a.timeout_ns, st + 2)) return;

  // 2. reduce my slice, store it everywhere.  Slices are whole float4s of the 16-byte aligned middle of the range; the
  //    unaligned head / tail elements (if any) belong to rank 0.
  const long long lo = a.off, hi = a.off + a.n;
  const long long v_lo = min((lo + 3) & ~3ll, hi), v_hi = max(hi & ~3ll, v_lo);
  const long long nv = (v_hi - v_lo) >> 2;
  const long long per = (nv + a.world - 1) / a.world;
  const long long s_lo = min(per * a.rank, nv), s_hi = min(per * (a.rank + 1), nv);
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
  for (long long i = s_lo + tid; i < s_hi; i += nthr) {
    const long long el = v_lo + 4 * i;
    float4 v[kMaxWorld];
#pragma unroll
    for (int p = 0; p < kMaxWorld; ++p)
      if (p < a.world) v[p] = ld_peer_v4(reinterpret_cast<const float*>(a.bufs[p]) + el);
    float4 s = v[0];
#pragma unroll
    for (int p = 1; p < kMaxWorld; ++p)
      if (p < a.world) {
        s.x += v[p].x;
        s.y += v[p].y;
        s.z += v[p].z;
        s.w += v[p].w;
      }
#pragma unroll
    for (int p = 0; p < kMaxWorld; ++p)
      if (p < a.world) *reinterpret_cast<float4*>(reinterpret_cast<float*>(a.bufs[p]) + el) = s;
  }
  if (a.rank == 0) {
    const long long n_edge = (v_lo - lo) + (hi - v_hi);
    for (long long i = tid; i < n_edge; i += nthr) {
      const long long el = i < v_lo - lo ? lo + i : v_hi + (i - (v_lo - lo));
      float s = 0.f;
      for (int p = 0; p < a.world; ++p) s += ld_peer_f(reinterpret_cast<const float*>(a.bufs[p]) + el);
      for (int p = 0; p < a.world; ++p) reinterpret_cast<float*>(a.bufs[p])[el] = s;
    }
  }

  // 3. my slice is written everywhere once every CTA of this grid has fenced its stores: the last one tells the peers,
  //    waits for theirs and closes the epoch
  __threadfence_system();
  __syncthreads();
  __shared__ int last;
  if (threadIdx.x == 0) last = (atomicAdd(st + 1, 1) == (int)gridDim.x - 1);
  __syncthreads();
  if (!last) return;
  __threadfence_system();
  if ((int)threadIdx.x < a.world)
    st_release_sys(reinterpret_cast<int*>(a.flags[threadIdx.x]) + flag_index(a.channel, 1, a.rank), e);
  wait_all(my_flags, a.channel, 1, a.world, e, a.timeout_ns, st + 2);
  if (threadIdx.x == 0) {
    st[1] = 0;
    __threadfence();
    *reinterpret_cast<volatile int*>(st) = e;
  }
}

}  // namespace p2p
}  // namespace pcrl

using namespace pcrl;

extern "C" {
int64_t pcrl_p2p_flag_bytes(void) { return (int64_t)p2p::kChannels * 2 * p2p::kMaxWorld * 4; }
int64_t pcrl_p2p_state_bytes(void) { return (int64_t)p2p::kChannels * 4 * 4; }

int pcrl_p2p_allreduce(const uint64_t* bufs_dev, const uint64_t* flags_dev, int rank, int world, int64_t off, int64_t n,
                       int channel, int32_t* state_dev, int max_ctas, void* stream) {
  PCRL_CHECK_ARG(bufs_dev && flags_dev && state_dev);
  PCRL_CHECK_ARG(world >= 1 && world <= p2p::kMaxWorld && rank >= 0 && rank < world);
  PCRL_CHECK_ARG(channel >= 0 && channel < p2p::kChannels && off >= 0 && n >= 0);
  if (n == 0 || world == 1) return PCRL_OK;
  p2p::Args a{};
  a.bufs = reinterpret_cast<const unsigned long long*>(bufs_dev);
  a.flags = reinterpret_cast<const unsigned long long*>(flags_dev);
  a.rank = rank;
  a.world = world;
  a.channel = channel;
  a.off = off;
  a.n = n;
  a.state = state_dev;
  a.timeout_ns = 120ll * 1000 * 1000 * 1000;  // a peer that never arrives: give up after two minutes (state[2] != 0)
  // every thread keeps `world` 16-byte loads in flight; a few CTAs saturate the NVLink ports without taking the SMs
  // away from the kernels this reduction overlaps
  const int64_t per_rank_v4 = cdiv(cdiv(n, 4), world);
  int ctas = (int)std::min<int64_t>(std::max<int64_t>(1, cdiv(per_rank_v4, p2p::kThreads)), max_ctas > 0 ? max_ctas : 32);
  p2p::allreduce_kernel<<<ctas, p2p::kThreads, 0, as_stream(stream)>>>(a);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}
}
