// Row-wise kernels shared by dense.cu and pointnet_f32.cu.
#pragma once
#include <algorithm>

#include "common.cuh"

namespace pcrl {

// y[i,:] = act(LN(x[i,:]) * g + b) for i < rows (rows_dev optional bound).  xhat/rstd optional saves.
int launch_ln_rows(const float* x, int64_t ldx, const float* g, const float* b, float* y, int64_t ldy, float* xhat,
                   float* rstd, int M, int D, float eps, int relu, const int* rows_dev, cudaStream_t st);
// dx = rstd * (dy*g - mean(dy*g) - xhat * mean(dy*g*xhat));  dg += sum dy*xhat;  db += sum dy
int launch_ln_rows_bwd(const float* dy, int64_t lddy, const float* xhat, const float* rstd, const float* g, float* dg,
                       float* db, float* dx, int64_t lddx, int M, int D, const int* rows_dev, cudaStream_t st);
// GEMM dispatch: tf32 != 0 and TMA-compatible operands -> tcgen05 TF32 kernel (tc_gemm.cu), else FFMA sgemm.
int launch_gemm(const GemmArgs& g, int tf32, cudaStream_t st);
// zero rows [*count, round_up(*count, 32)) of x[rows, width] (contraction tails of the compacted backward)
int launch_zero_tail(float* x, int width, const int* count_dev, int capacity_rows, cudaStream_t st);
int launch_colsum(const float* x, int64_t ld, int M, int N, const int* rows_dev, float* out, cudaStream_t st);

}  // namespace pcrl
