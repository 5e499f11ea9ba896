// tcgen05 TF32 GEMM (tc_gemm.cu): C[i][j] (op)= sum_l A(i,l) B(l,j) (+bias) (ReLU)
#pragma once
#include "common.cuh"

namespace pcrl {
namespace tcg {

struct TcGemmArgs {
  const float* A;  // a_mn == 0: A[i*lda + l] (K-major);  a_mn == 1: A[l*lda + i] (MN-major)
  int64_t lda;
  int a_mn;
  const float* B;  // b_mn == 0: B[j*ldb + l];            b_mn == 1: B[l*ldb + j]
  int64_t ldb;
  int b_mn;
  const float* bias;
  float* C;
  int64_t ldc;
  int M, N, K;
  int relu;
  int mode;     // 0 store, 1 C += (tile owned by one CTA), 2 atomicAdd (needed when split_k > 1)
  int split_k;
  const int* k_dev;  // optional device bound on the contraction length (rows beyond it must be zero up to a multiple of 32)
  const int* m_dev;  // optional device bound on M: tiles past it exit immediately
  const float* mask; // optional [M,N] post-activation tensor: output zeroed where mask <= 0 (fused ReLU backward)
  int64_t ldmask;
  int bn_hint;       // 0 = choose the output tile width automatically
};

bool tc_gemm_supported(const TcGemmArgs& g);
int launch_tc_gemm(const TcGemmArgs& g, cudaStream_t st);

}  // namespace tcg
}  // namespace pcrl
