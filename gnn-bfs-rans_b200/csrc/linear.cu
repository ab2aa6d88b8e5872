// linear.cu — C ABI of K6 (dense per-node Linear inside the conv layers) and the
// choice between the tcgen05 tensor-core kernel (gemm_tc.cu) and the exact-fp32 SIMT kernel
// (gemm_simt.cu).  impl: 0 = auto (tensor core when the shape/dtype is covered), 1 = SIMT, 2 = TC.
#include "common.cuh"

namespace b2g {
int simt_linear_fwd(const void*, int64_t, const void*, int64_t, const float*, const float*, void*, int64_t, float*, int64_t, int64_t, int, int, int, int, int, cudaStream_t);
int simt_linear_dgrad(const void*, int64_t, const void*, int64_t, void*, int64_t, int64_t, int, int, int, cudaStream_t);
int64_t simt_wgrad_ws_bytes(int64_t, int, int);
int simt_linear_wgrad(const void*, int64_t, const void*, int64_t, float*, int64_t, float*, int64_t, int, int, int, void*, cudaStream_t);
// gemm_tc.cu
bool tc_linear_supported(int64_t n, int m, int k, int dt, int which);
int64_t tc_linear_ws_bytes(int64_t n, int m, int k, int dt, int which);
int tc_linear_fwd(const void*, int64_t, const void*, int64_t, const float*, const float*, void*, int64_t, float*, int64_t, int64_t, int, int, int, int, int, int, const void*, int64_t, void*, cudaStream_t);
int64_t tc_wgrad_ws_bytes(int64_t n, int m, int k);
int tc_linear_wgrad(const void*, int64_t, const void*, int64_t, float*, int64_t, int64_t, int, int, void*, cudaStream_t);
}  // namespace b2g

using namespace b2g;

static inline bool dt_ok(int dt) { return dt == B2G_F32 || dt == B2G_BF16; }

extern "C" {

int64_t b2g_linear_workspace_bytes(int64_t n, int m, int k, int dt, int which) {
  if (n < 0 || m <= 0 || k <= 0 || !dt_ok(dt)) return B2G_E_ARG;
  int64_t b = 256;
  if (which == 2) b += simt_wgrad_ws_bytes(n, m, k);
  if (tc_linear_supported(n, m, k, dt, which)) {
    const int64_t t = which == 2 ? tc_wgrad_ws_bytes(n, m, k) : tc_linear_ws_bytes(n, m, k, dt, which);
    if (t > b) b = t;
  }
  return b;
}

int b2g_linear_impl(int64_t n, int m, int k, int dt, int which) {
  if (n < 0 || m <= 0 || k <= 0 || !dt_ok(dt)) return B2G_E_ARG;
  return tc_linear_supported(n, m, k, dt, which) ? 2 : 1;
}

int b2g_linear_fwd(const void* X, int64_t ldx, const void* W, int64_t ldw, const float* bias,
                   const float* row_scale, void* Y, int64_t ldy, float* aux, int64_t ldaux, int64_t n,
                   int m, int m_main, int k, int dt, int act, int impl, void* ws, void* stream) {
  // bits 8..15 of impl: SMs the persistent tensor-core kernel leaves idle (a collective running on another stream needs SMs:
  // the GEMM's CTAs hold all of an SM's shared memory, so nothing else is scheduled next to them)
  const int reserve = (impl >> 8) & 0xff;
  impl &= 0xff;
  if (n < 0 || m <= 0 || k <= 0 || !dt_ok(dt) || act < 0 || act > 1 || impl < 0 || impl > 2) return B2G_E_ARG;
  if (m_main < 0 || m_main > m || (m_main < m && (!aux || (m_main % 8) != 0))) return B2G_E_ARG;
  if (n == 0) return B2G_OK;
  if (!X || !W || (m_main > 0 && !Y)) return B2G_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const bool tc = tc_linear_supported(n, m, k, dt, 0);
  if (impl == 2 && !tc) return B2G_E_UNSUPPORTED;
  if (tc && impl != 1) {
    const int rc = tc_linear_fwd(X, ldx, W, ldw, bias, row_scale, Y, ldy, aux, ldaux, n, m, m_main, k, dt, act, reserve, nullptr, 0, ws, st);
    // shapes the tensor-core kernels decline at launch time (fp32 aux split with k > 256, tensors TMA cannot map):
    // auto mode falls through to the exact-fp32 SIMT kernel instead of failing the request
    if (rc != B2G_E_UNSUPPORTED || impl == 2) return rc;
  }
  return simt_linear_fwd(X, ldx, W, ldw, bias, row_scale, Y, ldy, aux, ldaux, n, m, m_main, k, dt, act, st);
}

/* Y = (X W^T) where mask > 0, else 0 (bf16, tensor-core path only): the ReLU backward fused into the dgrad GEMM of the layer
 * behind it — X = dY of that layer, W = its weight transposed ([k_out, m_in] row-major), mask = the ReLU's output [n, m]. */
int b2g_linear_fwd_masked(const void* X, int64_t ldx, const void* W, int64_t ldw, const void* mask, int64_t ldmask, void* Y,
                          int64_t ldy, int64_t n, int m, int k, int dt, void* ws, void* stream) {
  if (n < 0 || m <= 0 || k <= 0 || !dt_ok(dt)) return B2G_E_ARG;
  if (n == 0) return B2G_OK;
  if (!X || !W || !Y || !mask) return B2G_E_ARG;
  if (dt != B2G_BF16 || !tc_linear_supported(n, m, k, dt, 0)) return B2G_E_UNSUPPORTED;
  return tc_linear_fwd(X, ldx, W, ldw, nullptr, nullptr, Y, ldy, nullptr, 0, n, m, m, k, dt, 0, 0, mask, ldmask, ws, (cudaStream_t)stream);
}

int b2g_linear_dgrad(const void* dY, int64_t lddy, const void* W, int64_t ldw, void* dX,
                     int64_t lddx, int64_t n, int m, int k, int dt, int impl, void* ws,
                     void* stream) {
  if (n < 0 || m <= 0 || k <= 0 || !dt_ok(dt) || impl < 0 || impl > 2) return B2G_E_ARG;
  if (n == 0) return B2G_OK;
  if (!dY || !W || !dX) return B2G_E_ARG;
  if (impl == 2) return B2G_E_UNSUPPORTED;  // dgrad runs on the tensor cores through fwd with W^T (host side)
  (void)ws;
  return simt_linear_dgrad(dY, lddy, W, ldw, dX, lddx, n, m, k, dt, (cudaStream_t)stream);
}

int b2g_linear_wgrad(const void* dY, int64_t lddy, const void* X, int64_t ldx, float* dW,
                     int64_t lddw, float* db, int64_t n, int m, int k, int dt, int impl, void* ws,
                     void* stream) {
  if (n < 0 || m <= 0 || k <= 0 || !dt_ok(dt) || impl < 0 || impl > 2) return B2G_E_ARG;
  if (!dW || !ws) return B2G_E_ARG;
  if (n && (!dY || !X)) return B2G_E_ARG;
  const bool tc = n > 0 && tc_linear_supported(n, m, k, dt, 2);
  if (impl == 2 && !tc) return B2G_E_UNSUPPORTED;
  if (tc && impl != 1) {
    if (db) return B2G_E_ARG;   // the tensor-core wgrad leaves the bias gradient to b2g_colsum (host side)
    return tc_linear_wgrad(dY, lddy, X, ldx, dW, lddw, n, m, k, ws, (cudaStream_t)stream);
  }
  return simt_linear_wgrad(dY, lddy, X, ldx, dW, lddw, db, n, m, k, dt, ws, (cudaStream_t)stream);
}

}  // extern "C"
