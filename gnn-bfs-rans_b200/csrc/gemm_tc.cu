// gemm_tc.cu — K6 (tensor-core path): tcgen05 / TMEM / TMA GEMM for the per-node Linear.
// Placeholder until the tcgen05 kernel lands: reports "not supported" so linear.cu uses the SIMT path.
#include "common.cuh"

namespace b2g {
bool tc_linear_supported(int64_t, int, int, int, int) { return false; }
int64_t tc_linear_ws_bytes(int64_t, int, int, int, int) { return 0; }
int tc_linear_fwd(const void*, int64_t, const void*, int64_t, const float*, const float*, void*, int64_t,
                  float*, int64_t, int64_t, int, int, int, int, int, void*, cudaStream_t) {
  return B2G_E_UNSUPPORTED;
}
}  // namespace b2g
