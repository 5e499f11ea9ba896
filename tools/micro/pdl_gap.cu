// Microbenchmark: per-kernel cost of a chain of dependent tiny kernels replayed from a CUDA graph, with and without
// programmatic dependent launch (PDL: griddepcontrol.launch_dependents / griddepcontrol.wait).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pdl_gap pdl_gap.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void step_plain(float* x, int work) {
  float v = x[threadIdx.x];
  for (int i = 0; i < work; ++i) v = v * 1.0001f + 0.5f;
  x[threadIdx.x] = v;
}
__global__ void step_pdl(float* x, int work) {
  asm volatile("griddepcontrol.launch_dependents;");  // let the next kernel's CTAs get scheduled right away
  asm volatile("griddepcontrol.wait;" ::: "memory");  // ... but do not touch memory before the previous grid is done
  float v = x[threadIdx.x];
  for (int i = 0; i < work; ++i) v = v * 1.0001f + 0.5f;
  x[threadIdx.x] = v;
}

static float run(bool pdl, int n, int work, int blocks) {
  float* x;
  cudaMalloc(&x, 1024 * 4);
  cudaMemset(x, 0, 1024 * 4);
  cudaStream_t st;
  cudaStreamCreate(&st);
  cudaGraph_t g;
  cudaGraphExec_t ge;
  cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
  for (int i = 0; i < n; ++i) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(blocks);
    cfg.blockDim = dim3(256);
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    if (pdl) cudaLaunchKernelEx(&cfg, step_pdl, x, work);
    else cudaLaunchKernelEx(&cfg, step_plain, x, work);
  }
  cudaStreamEndCapture(st, &g);
  cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
  if (e != cudaSuccess) { printf("instantiate: %s\n", cudaGetErrorString(e)); return -1; }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int w = 0; w < 3; ++w) cudaGraphLaunch(ge, st);
  cudaStreamSynchronize(st);
  cudaEventRecord(e0, st);
  for (int w = 0; w < 20; ++w) cudaGraphLaunch(ge, st);
  cudaEventRecord(e1, st);
  cudaStreamSynchronize(st);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms * 1e3f / (20 * n);
}

int main() {
  for (int blocks : {1, 148}) {
    for (int work : {0, 2000}) {
      printf("blocks=%3d work=%4d: plain %.2f us/kernel   PDL %.2f us/kernel\n", blocks, work, run(false, 100, work, blocks),
             run(true, 100, work, blocks));
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
