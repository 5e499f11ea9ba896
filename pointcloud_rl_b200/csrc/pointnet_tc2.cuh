// Second-generation fused tcgen05 PointNet forward (pointnet_tc2.cu): transposed layer 2, Gram-matrix variance.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pcrl {
namespace tc2 {
bool shapes_ok(int c1, int c2, int c3);
int64_t wpack_bytes(int c1, int c2, int c3);
int pack(const float* w0, const float* b0, const float* w1, const float* g1, const float* be1, const float* w2,
         const float* g2, const float* be2, int C, int c1, int c2, int c3, int rgb_u8, void* wpack2, cudaStream_t st);
int forward(const void* xh, int R, int src_cloud_stride, int NP, const void* wpack2, int c1, int c2, int c3, float ln_eps,
            uint64_t* pool_keys, float* pooled, int32_t* argmax, cudaStream_t st);
int set_debug_flags(int flags);
int get_trace(long long* out_host);
}  // namespace tc2
}  // namespace pcrl
