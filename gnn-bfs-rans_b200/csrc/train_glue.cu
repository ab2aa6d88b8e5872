// train_glue.cu — the pieces of the reference's training step that sit around the hot path:
//   * device side of `Batch.from_data_list` (train.py:155): node offsets added to the concatenated edge_index and the
//     per-node graph id vector, after the samples' tensors were copied into place (one kernel instead of B host-side adds);
//   * the reference's field-wise weighted MSE loss with the pressure-mean anchor (normalization.py:177-236) as one
//     reduction + one elementwise gradient kernel (the reference: ~25 torch kernels forward, as many backward);
//   * global gradient norm + clip + Adam (train.py:188-189: clip_grad_norm_(max_norm = 1), Adam(lr, weight_decay)) over ONE
//     flat parameter / gradient / moment buffer: three launches instead of the foreach kernels' dozens.
// All reductions are two-stage with a fixed block count and a fixed in-block order: deterministic.
#include "common.cuh"

namespace b2g {

// ------------------------------------------------------------------------------------------ Batch.from_data_list
__global__ void __launch_bounds__(256) batch_finalize_kernel(int64_t* __restrict__ ei, int64_t e_tot,
                                                             const int64_t* __restrict__ edge_ptr,
                                                             const int64_t* __restrict__ node_ptr, int n_graphs,
                                                             int64_t* __restrict__ batch, int64_t n_tot) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < e_tot; t += stride) {
    int lo = 0, hi = n_graphs;                       // graph s with edge_ptr[s] <= t < edge_ptr[s + 1]
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (edge_ptr[mid] <= t) lo = mid; else hi = mid;
    }
    const int64_t off = node_ptr[lo];
    if (off) {
      ei[t] += off;
      ei[e_tot + t] += off;
    }
  }
  if (batch)
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_tot; t += stride) {
      int lo = 0, hi = n_graphs;
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (node_ptr[mid] <= t) lo = mid; else hi = mid;
      }
      batch[t] = lo;
    }
}

// ------------------------------------------------------------------------------------------ weighted MSE (7 fields)
constexpr int WM_BLOCKS = 296;   // 2 x 148
constexpr int WM_COLS = 7;       // U(3) p k epsilon nut — normalization.py:196-226
template <typename T> __device__ __forceinline__ float ld_as_float(const T* p) { return (float)*p; }
template <> __device__ __forceinline__ float ld_as_float<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// partial[block][0..6] = sum (pred - target)^2 per column, [7] = sum pred_p, [8] = sum target_p
template <typename T>
__global__ void __launch_bounds__(256) wmse_partial_kernel(const T* __restrict__ pred, int64_t ldp, const T* __restrict__ tgt,
                                                           int64_t ldt, int64_t n, float* __restrict__ partial) {
  float acc[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) acc[k] = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int c = 0; c < WM_COLS; ++c) {
      const float p = ld_as_float(pred + i * ldp + c), t = ld_as_float(tgt + i * ldt + c);
      const float d = p - t;
      acc[c] += d * d;
      if (c == 3) { acc[7] += p; acc[8] += t; }
    }
  }
  __shared__ float sh[8][9];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const float s = warp_sum(acc[k]);
    if (lane == 0) sh[w][k] = s;
  }
  __syncthreads();
  if (threadIdx.x < 9) {
    float s = 0.f;
    for (int ww = 0; ww < 8; ++ww) s += sh[ww][threadIdx.x];
    partial[blockIdx.x * 9 + threadIdx.x] = s;
  }
}
// loss[0] = sum_f w_f mean_f((pred - target)^2) + w_p * prw * (mean pred_p - mean target_p)^2; coef[0..6] = d loss / d d_ic per
// unit (pred - target) of column c, coef[7] = constant added to d loss / d pred_p (the anchor term)
__global__ void wmse_final_kernel(const float* __restrict__ partial, int blocks, int64_t n, const float* __restrict__ fw /*[5]*/,
                                  float prw, float* __restrict__ loss, float* __restrict__ coef) {
  __shared__ float tot[9];
  if (threadIdx.x < 9) {
    float s = 0.f;
    for (int b = 0; b < blocks; ++b) s += partial[b * 9 + threadIdx.x];
    tot[threadIdx.x] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const float fn = (float)n;
    const float wU = fw[0], wp = fw[1], wk = fw[2], we = fw[3], wn = fw[4];
    const float u_loss = (tot[0] + tot[1] + tot[2]) / (3.f * fn);
    float p_loss = tot[3] / fn;
    const float dm = tot[7] / fn - tot[8] / fn;
    if (prw > 0.f) p_loss += prw * dm * dm;
    loss[0] = wU * u_loss + wp * p_loss + wk * tot[4] / fn + we * tot[5] / fn + wn * tot[6] / fn;
    coef[0] = coef[1] = coef[2] = wU * 2.f / (3.f * fn);
    coef[3] = wp * 2.f / fn;
    coef[4] = wk * 2.f / fn;
    coef[5] = we * 2.f / fn;
    coef[6] = wn * 2.f / fn;
    coef[7] = prw > 0.f ? wp * prw * 2.f * dm / fn : 0.f;
  }
}
template <typename T>
__global__ void __launch_bounds__(256) wmse_bwd_kernel(const T* __restrict__ pred, int64_t ldp, const T* __restrict__ tgt,
                                                       int64_t ldt, int64_t n, const float* __restrict__ coef,
                                                       const float* __restrict__ gout, T* __restrict__ dpred, int64_t ldd) {
  const float g = gout ? gout[0] : 1.f;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n * WM_COLS; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / WM_COLS;
    const int c = (int)(t - i * WM_COLS);
    float d = (ld_as_float(pred + i * ldp + c) - ld_as_float(tgt + i * ldt + c)) * coef[c];
    if (c == 3) d += coef[7];
    dpred[i * ldd + c] = (T)(d * g);
  }
}

// ------------------------------------------------------------------------------------------ clip + Adam over a flat buffer
constexpr int AD_BLOCKS = 296;
__global__ void __launch_bounds__(256) sqsum_partial_kernel(const float* __restrict__ g, int64_t n, float* __restrict__ partial) {
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) acc += g[i] * g[i];
  __shared__ float sh[8];
  const float s = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += sh[w];
    partial[blockIdx.x] = t;
  }
}
// state[0] = step count (as float), state[1] = total gradient norm of this step (output)
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, int64_t n, const float* __restrict__ partial,
                                                   int blocks, float max_norm, float lr, float b1, float b2, float eps,
                                                   float wd, float* __restrict__ state) {
  __shared__ float s_scale, s_bc1, s_bc2;
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int b = 0; b < blocks; ++b) tot += partial[b];
    const float norm = sqrtf(tot);
    float scale = 1.f;
    if (max_norm > 0.f) scale = fminf(1.f, max_norm / (norm + 1e-6f));        // torch.nn.utils.clip_grad_norm_
    const float step = state[0] + 1.f;                                          // every block reads the pre-update count
    s_scale = scale;
    s_bc1 = 1.f - powf(b1, step);
    s_bc2 = 1.f - powf(b2, step);
  }
  __syncthreads();
  const float scale = s_scale, bc1 = s_bc1, bc2s = sqrtf(s_bc2);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i] * scale;
    const float pi = p[i];
    if (wd != 0.f) gi += wd * pi;                                               // torch.optim.Adam: L2 term added to the gradient
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2s + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
  }
}
__global__ void adam_state_kernel(const float* __restrict__ partial, int blocks, float* __restrict__ state) {
  float tot = 0.f;
  for (int b = 0; b < blocks; ++b) tot += partial[b];
  state[0] += 1.f;
  state[1] = sqrtf(tot);
}

}  // namespace b2g

using namespace b2g;

extern "C" {

int b2g_batch_finalize(int64_t* edge_index, int64_t e_tot, const int64_t* edge_ptr, const int64_t* node_ptr, int n_graphs,
                       int64_t* batch, int64_t n_tot, void* stream) {
  if (e_tot < 0 || n_tot < 0 || n_graphs < 1 || !edge_ptr || !node_ptr || (e_tot > 0 && !edge_index)) return B2G_E_ARG;
  const int64_t work = e_tot > n_tot ? e_tot : n_tot;
  if (work == 0) return B2G_OK;
  const int64_t want = ceil_div(work, 256), cap = (int64_t)B2G_NUM_SMS * 8;
  batch_finalize_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(edge_index, e_tot, edge_ptr, node_ptr,
                                                                                               n_graphs, batch, n_tot);
  count_launch();
  return cuda_status();
}

int64_t b2g_wmse_workspace_bytes(void) { return (int64_t)WM_BLOCKS * 9 * 4 + 256; }

int b2g_wmse_fwd(const void* pred, int64_t ldp, const void* target, int64_t ldt, int64_t n, int dt, const float* field_weights,
                 float pressure_ref_weight, float* loss, float* coef, void* ws, void* stream) {
  if (n <= 0 || !pred || !target || !field_weights || !loss || !coef || !ws || ldp < WM_COLS || ldt < WM_COLS) return B2G_E_ARG;
  if (dt != B2G_F32 && dt != B2G_BF16) return B2G_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  float* partial = static_cast<float*>(ws);
  if (dt == B2G_F32)
    wmse_partial_kernel<float><<<WM_BLOCKS, 256, 0, st>>>((const float*)pred, ldp, (const float*)target, ldt, n, partial);
  else
    wmse_partial_kernel<__nv_bfloat16><<<WM_BLOCKS, 256, 0, st>>>((const __nv_bfloat16*)pred, ldp, (const __nv_bfloat16*)target, ldt, n, partial);
  wmse_final_kernel<<<1, 32, 0, st>>>(partial, WM_BLOCKS, n, field_weights, pressure_ref_weight, loss, coef);
  count_launch(2);
  return cuda_status();
}

int b2g_wmse_bwd(const void* pred, int64_t ldp, const void* target, int64_t ldt, int64_t n, int dt, const float* coef,
                 const float* grad_out, void* dpred, int64_t ldd, void* stream) {
  if (n <= 0 || !pred || !target || !coef || !dpred || ldp < WM_COLS || ldt < WM_COLS || ldd < WM_COLS) return B2G_E_ARG;
  if (dt != B2G_F32 && dt != B2G_BF16) return B2G_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t want = ceil_div(n * WM_COLS, 256), cap = (int64_t)B2G_NUM_SMS * 8;
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  if (dt == B2G_F32)
    wmse_bwd_kernel<float><<<grid, 256, 0, st>>>((const float*)pred, ldp, (const float*)target, ldt, n, coef, grad_out, (float*)dpred, ldd);
  else
    wmse_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)pred, ldp, (const __nv_bfloat16*)target, ldt, n, coef,
                                                         grad_out, (__nv_bfloat16*)dpred, ldd);
  count_launch();
  return cuda_status();
}

int64_t b2g_adam_workspace_bytes(void) { return (int64_t)AD_BLOCKS * 4 + 256; }

/* One optimisation step over flat fp32 buffers: total gradient norm, clip to max_norm (<= 0: no clipping), Adam with L2
 * weight decay (torch.optim.Adam semantics), step counter and the norm in `state` (fp32 [2], device). */
int b2g_clip_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float max_norm, float lr,
                       float beta1, float beta2, float eps, float weight_decay, float* state, void* ws, void* stream) {
  if (n <= 0 || !params || !grads || !exp_avg || !exp_avg_sq || !state || !ws) return B2G_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  float* partial = static_cast<float*>(ws);
  sqsum_partial_kernel<<<AD_BLOCKS, 256, 0, st>>>(grads, n, partial);
  const int64_t want = ceil_div(n, 256);
  adam_kernel<<<(unsigned)(want < AD_BLOCKS ? want : AD_BLOCKS), 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, n, partial, AD_BLOCKS,
                                                                               max_norm, lr, beta1, beta2, eps, weight_decay, state);
  adam_state_kernel<<<1, 1, 0, st>>>(partial, AD_BLOCKS, state);
  count_launch(3);
  return cuda_status();
}

}  // extern "C"
