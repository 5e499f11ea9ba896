// Generic strided fp32 GEMM on CUDA cores (FFMA).  This is the exact-fp32 building block: the parity
// path of the PointNet forward, the compacted sparse backward and the small MLP heads all run on it.
// Tensor-core (tcgen05) kernels live in pointnet_tc.cu.
#include <stdarg.h>

#include <atomic>
#include <mutex>

#include "common.cuh"

namespace pcrl {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

// ---- per-device contexts ------------------------------------------------------------------------
static std::mutex g_ctx_mutex;
static DeviceCtx* g_ctx[64] = {nullptr};
static std::atomic<long long> g_tf32_fallbacks{0};
static std::atomic<int> g_strict_tf32{0};

static DeviceCtx* make_ctx(int dev) {
  auto* c = new DeviceCtx();
  c->device = dev;
  if (cudaDeviceGetAttribute(&c->sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) c->sms = 148;
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  if (cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, lo) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming) != cudaSuccess) {
    set_error("pcrl: could not create the device context of device %d: %s", dev, cudaGetErrorString(cudaGetLastError()));
    delete c;
    return nullptr;
  }
  return c;
}

DeviceCtx* device_ctx() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
    set_error("pcrl: no current CUDA device");
    return nullptr;
  }
  std::lock_guard<std::mutex> lock(g_ctx_mutex);
  if (!g_ctx[dev]) g_ctx[dev] = make_ctx(dev);
  return g_ctx[dev];
}

int sm_count() {
  DeviceCtx* c = device_ctx();
  return c ? c->sms : 148;
}

void note_tf32_fallback() { g_tf32_fallbacks.fetch_add(1); }
bool strict_tf32() { return g_strict_tf32.load() != 0; }

constexpr int BM = 64, BN = 64, BK = 16, PAD = 4;

__global__ void __launch_bounds__(256) sgemm_kernel(GemmArgs g) {
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int i0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
  const int M = g.m_dev ? min(g.M, *g.m_dev) : g.M;
  const int K = g.k_dev ? min(g.K, *g.k_dev) : g.K;
  if (i0 >= M) return;

  int kchunk = (K + g.split_k - 1) / g.split_k;
  kchunk = (kchunk + BK - 1) / BK * BK;
  const int k0 = blockIdx.z * kchunk;
  const int k1 = min(K, k0 + kchunk);
  if (k0 >= k1 && !(blockIdx.z == 0)) return;

  const bool a_l_contig = (g.a_sl == 1);
  const bool b_j_contig = (g.b_sj == 1);

  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

  float ra[4], rb[4];
  auto load_tiles = [&](int kt) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int e = tid + 256 * q;
      int i, l;
      if (a_l_contig) { l = e & 15; i = e >> 4; } else { i = e & 63; l = e >> 6; }
      int gi = i0 + i, gl = kt + l;
      ra[q] = (gi < M && gl < k1) ? __ldg(g.A + (int64_t)gi * g.a_si + (int64_t)gl * g.a_sl) : 0.f;
      int j;
      if (b_j_contig) { j = e & 63; l = e >> 6; } else { l = e & 15; j = e >> 4; }
      int gj = j0 + j;
      gl = kt + l;
      rb[q] = (gj < g.N && gl < k1) ? __ldg(g.B + (int64_t)gl * g.b_sl + (int64_t)gj * g.b_sj) : 0.f;
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int e = tid + 256 * q;
      int i, l;
      if (a_l_contig) { l = e & 15; i = e >> 4; } else { i = e & 63; l = e >> 6; }
      As[l][i] = ra[q];
      int j;
      if (b_j_contig) { j = e & 63; l = e >> 6; } else { l = e & 15; j = e >> 4; }
      Bs[l][j] = rb[q];
    }
  };

  if (k0 < k1) {
    load_tiles(k0);
    store_tiles();
    __syncthreads();
    for (int kt = k0; kt < k1; kt += BK) {
      const bool more = (kt + BK < k1);
      if (more) load_tiles(kt + BK);
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[p][q] = fmaf(av[p], bv[q], acc[p][q]);
      }
      __syncthreads();
      if (more) {
        store_tiles();
        __syncthreads();
      }
    }
  }

#pragma unroll
  for (int p = 0; p < 4; ++p) {
    int gi = i0 + ty * 4 + p;
    if (gi >= M) continue;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int gj = j0 + tx * 4 + q;
      if (gj >= g.N) continue;
      float v = acc[p][q];
      if (g.bias && blockIdx.z == 0) v += g.bias[gj];
      float* c = g.C + (int64_t)gi * g.ldc + gj;
      if (g.accumulate) {
        atomicAdd(c, v);
      } else {
        if (g.relu & 1) v = fmaxf(v, 0.f);
        if (g.relu & 2) v = round_tf32(v);
        if (g.mask && !(g.mask[(int64_t)gi * g.ldmask + gj] > 0.f)) v = 0.f;
        *c = v;
      }
    }
  }
}

int launch_sgemm(const GemmArgs& g, cudaStream_t st) {
  PCRL_CHECK_ARG(g.M >= 0 && g.N >= 0 && g.K >= 0 && g.split_k >= 1);
  PCRL_CHECK_ARG(g.split_k == 1 || (g.accumulate && !g.relu));
  PCRL_CHECK_ARG(!g.mask || !g.accumulate);
  if (g.M == 0 || g.N == 0) return PCRL_OK;
  dim3 grid((unsigned)cdiv(g.N, BN), (unsigned)cdiv(g.M, BM), (unsigned)g.split_k);
  sgemm_kernel<<<grid, 256, 0, st>>>(g);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

}  // namespace pcrl

extern "C" {
int pcrl_abi_version(void) { return PCRL_ABI_VERSION; }
const char* pcrl_last_error(void) { return pcrl::last_error(); }
int pcrl_sm_count(void) { return pcrl::sm_count(); }

int64_t pcrl_create(int device) {
  int prev = 0;
  if (cudaGetDevice(&prev) != cudaSuccess || cudaSetDevice(device) != cudaSuccess) {
    pcrl::set_error("pcrl_create: cannot select device %d", device);
    return 0;
  }
  pcrl::DeviceCtx* c = pcrl::device_ctx();
  cudaSetDevice(prev);
  return (int64_t)reinterpret_cast<intptr_t>(c);
}

int pcrl_destroy(int64_t handle) {
  auto* c = reinterpret_cast<pcrl::DeviceCtx*>((intptr_t)handle);
  if (!c) return PCRL_OK;
  std::lock_guard<std::mutex> lock(pcrl::g_ctx_mutex);
  if (c->device < 0 || c->device >= 64 || pcrl::g_ctx[c->device] != c) {
    pcrl::set_error("pcrl_destroy: not a live handle");
    return PCRL_EINVAL;
  }
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->side);
  cudaStreamDestroy(c->side);
  cudaEventDestroy(c->ev_fork);
  cudaEventDestroy(c->ev_join);
  cudaSetDevice(prev);
  pcrl::g_ctx[c->device] = nullptr;
  delete c;
  return PCRL_OK;
}

int64_t pcrl_tf32_fallbacks(void) { return (int64_t)pcrl::g_tf32_fallbacks.load(); }
int pcrl_set_strict_tf32(int on) {
  pcrl::g_strict_tf32.store(on ? 1 : 0);
  return PCRL_OK;
}
}
