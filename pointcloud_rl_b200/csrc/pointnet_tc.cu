// placeholder until the tcgen05 kernel lands (next commit)
#include "common.cuh"
extern "C" {
int64_t pcrl_pointnet_wpack_bytes(int c1, int c2, int c3) { return 0; }
int pcrl_pointnet_pack_weights(const float*, const float*, const float*, const float*, const float*, const float*,
                               const float*, const float*, int, int, int, int, int, void*, void*) {
  pcrl::set_error("bf16 path not built");
  return PCRL_EUNSUPPORTED;
}
int pcrl_pointnet_fwd_bf16(const void*, int, int, int, const void*, int, int, int, float, uint64_t*, float*, int32_t*,
                           void*) {
  pcrl::set_error("bf16 path not built");
  return PCRL_EUNSUPPORTED;
}
}
