// aggregate.cu — K2/K3: deterministic CSR segment-sum with fused normalisation, self term, bias and
// ReLU.  Replaces PyG MessagePassing.propagate = x.index_select(0, edge_index[0]) (an [E,F]
// temporary) + zeros(N,F).scatter_add_(0, edge_index[1], msg) for
//   GCNConv  (gnn_model.py:63,166): out_i = sum_p dinv_i*dinv_j * xw_j + bias          
//   GINConv  (gnn_model.py:75,166): h_i   = sum_p x_j + (1+eps) x_i                    
// and, on the transposed CSR, their backward.  The [E,F] message tensor is never materialised.
//
// HBM-bound.  Algorithmic bytes per launch: 2*N*F*s + 4*nnz + 4*(N+1) (+4*N dinv).
// Mapping: one group of LANES lanes per target row, each lane owns VPL 16-byte vectors of the row;
// neighbour rows are read with 16-byte L1-bypassing loads, U rows in flight per lane, and re-used
// across neighbouring targets through the 126 MB L2.  fp32 accumulation in CSR (= edge) order.
#include "common.cuh"

namespace b2g {

constexpr int SEG_ITERS = 16;  // rows per row-group per CTA chunk (128 consecutive rows per CTA step at 8 groups)

template <int LANES>
__device__ __forceinline__ unsigned group_mask() {
  if (LANES >= 32) return 0xffffffffu;
  const int g = (threadIdx.x & 31) / LANES;
  return ((1u << (LANES & 31)) - 1u) << (g * LANES);
}

// kScale: 0 = plain sum, 1 = per-row scale only (applied once at the end), 2 = per-edge weight rs_i * cs_j
// kFull : the row is exactly LANES*VPL vectors wide (no per-vector bounds checks)
template <typename T, int LANES, int VPL, int kScale, bool kFull>
__global__ void __launch_bounds__(256, (VPL <= 2 ? 3 : (VPL == 4 ? 2 : 1)))
seg_sum_kernel(const T* __restrict__ x, int64_t ldx, const T* __restrict__ x_self, int64_t ldxs,
               T* __restrict__ out, int64_t ldo, int64_t n_rows, int nvec,
               const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
               const float* __restrict__ row_scale, const float* __restrict__ col_scale,
               float self_coef, const float* __restrict__ bias, int relu) {
  constexpr int VN = Vec<T>::N;
  constexpr int U = (VPL == 1) ? 8 : (VPL == 2 ? 4 : 2);   // neighbour rows in flight per lane (mesh rows: 7)
  const int gl = threadIdx.x % LANES;                       // lane within the row group
  constexpr int GPB = 256 / LANES;                          // row groups per CTA
  constexpr int64_t CHUNK = (int64_t)GPB * SEG_ITERS;       // consecutive rows a CTA owns per grid stride
  const int gi = threadIdx.x / LANES;
  const uint64_t pol_stream = l2_policy_evict_first();      // outputs are written once: do not let them push x rows out of L2

  // A CTA walks CHUNK consecutive rows (adjacent rows share neighbours -> L1 hits), then strides by the
  // whole (co-resident) grid, so at any time the chip works on one compact window of rows whose
  // neighbour rows are still in L2.
  for (int64_t c0 = (int64_t)blockIdx.x * CHUNK; c0 < n_rows; c0 += (int64_t)gridDim.x * CHUNK) {
    int64_t i = c0 + gi;
    int b = 0, e = 0;
    if (i < n_rows) {
      b = __ldg(rowptr + i);
      e = __ldg(rowptr + i + 1);
    }
    for (int it = 0; it < SEG_ITERS; ++it, i += GPB) {
      if (i >= n_rows) break;
      const int64_t i2 = i + GPB;
      int b2 = 0, e2 = 0;
      if (it + 1 < SEG_ITERS && i2 < n_rows) {                // next row's extent: off the critical path
        b2 = __ldg(rowptr + i2);
        e2 = __ldg(rowptr + i2 + 1);
      }
      const float rs = (kScale && row_scale) ? __ldg(row_scale + i) : 1.0f;
      float acc[VPL][VN];
#pragma unroll
      for (int v = 0; v < VPL; ++v)
#pragma unroll
        for (int k = 0; k < VN; ++k) acc[v][k] = 0.f;

      // Branch-free gather, U neighbour rows in flight.  Every lane of the group reads the same col[] entry
      // (one broadcast transaction, no shuffles); slots past the end of the row re-read row i itself with
      // weight 0, so the unrolled body carries no predicates, and w = 1 reproduces a plain fp32 add bit for bit.
      // The self term (GIN: (1+eps) x_i) rides along as a virtual first neighbour of the same batch.
      const int ns = (self_coef != 0.f && !x_self && kScale != 1) ? 1 : 0;
      for (int j = b - ns; j < e; j += U) {
        Vec<T> buf[U][VPL];
        int c[U];
        // phase 1: all U column indices; phase 2: all U row loads; phase 3: the FMAs.
#pragma unroll
        for (int u = 0; u < U; ++u) c[u] = (j + u < b || j + u >= e) ? (int)i : ldg_i32_ordered(col + j + u);
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const T* __restrict__ row = x + (int64_t)c[u] * ldx;
#pragma unroll
          for (int v = 0; v < VPL; ++v) {
            const int vi = gl + v * LANES;
            if (kFull || vi < nvec) buf[u][v] = ldg_vec_l1<T>(row + vi * VN);
          }
        }
        if (j == b && e2 > b2 && gl == 0) prefetch_l1(col + b2);   // next row's indices land in L1 meanwhile
        // ptxas otherwise sinks each FMA group next to its load (load -> use -> load -> use: ONE row in flight
        // per warp; ncu showed a long-scoreboard stall on every buffer).  Make every add depend on all U loads
        // through a value ptxas cannot fold: XOR one word of each buffer, XOR it with its own identity shuffle
        // (always 0, but opaque), and OR that zero into the weight / predicate of every slot.
        uint32_t dep = 0;
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int v = 0; v < VPL; ++v)
            if (kFull || gl + v * LANES < nvec) dep ^= first_word(buf[u][v]);
        dep ^= __shfl_sync(group_mask<LANES>(), dep, threadIdx.x & 31);
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (kScale == 2 || ns) {
            float w = (j + u < e) ? 1.0f : 0.f;
            if (kScale == 2 && j + u >= b && j + u < e) w = (col_scale ? __ldg(col_scale + c[u]) : 1.0f) * rs;
            if (j + u < b) w = self_coef;
            w = __uint_as_float(__float_as_uint(w) | dep);
#pragma unroll
            for (int v = 0; v < VPL; ++v)
              if (kFull || gl + v * LANES < nvec) fma_vec(acc[v], w, buf[u][v]);
          } else {
            const uint32_t pred = (uint32_t)(j + u < e) | dep;      // plain sum: predicated packed adds, no weights
#pragma unroll
            for (int v = 0; v < VPL; ++v)
              if (kFull || gl + v * LANES < nvec) add_vec_if(acc[v], buf[u][v], pred);
          }
        }
      }
      // epilogue: row scale, self term, bias, relu, store
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int vi = gl + v * LANES;
        if (kFull || vi < nvec) {
          if (kScale == 1) {
#pragma unroll
            for (int k = 0; k < VN; ++k) acc[v][k] *= rs;
          }
          if (self_coef != 0.f && !ns) {
            const Vec<T> sv = ldg_vec_l1<T>((x_self ? x_self + i * ldxs : x + i * ldx) + vi * VN);
            fma_vec(acc[v], self_coef, sv);
          }
          if (bias) {   // per-row reload from L1 (registers are reserved for rows in flight)
#pragma unroll
            for (int k = 0; k < VN; k += 4) {
              const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + vi * VN + k));
              acc[v][k] += bb.x; acc[v][k + 1] += bb.y; acc[v][k + 2] += bb.z; acc[v][k + 3] += bb.w;
            }
          }
          if (relu) {
#pragma unroll
            for (int k = 0; k < VN; ++k) acc[v][k] = fmaxf(acc[v][k], 0.f);
          }
          Vec<T> o;
          o.from_float(acc[v]);
          stg_vec_hint<T>(out + i * ldo + vi * VN, o, pol_stream);
        }
      }
      b = b2; e = e2;
    }
  }
}

template <typename T, int LANES, int VPL>
static int launch_seg_sum(const void* x, int64_t ldx, const void* x_self, int64_t ldxs, void* out,
                          int64_t ldo, int64_t n_rows, int nvec, const int32_t* rowptr,
                          const int32_t* col, const float* row_scale, const float* col_scale,
                          float self_coef, const float* bias, int relu, cudaStream_t st) {
  const int64_t chunk = (int64_t)(256 / LANES) * SEG_ITERS;
  int64_t blocks = ceil_div(n_rows, chunk);
  const int mode = col_scale ? 2 : (row_scale ? 1 : 0);
  const bool full = nvec == LANES * VPL;
#define B2G_LAUNCH(MODE, FULL)                                                                            \
  {                                                                                                       \
    const int64_t cap = resident_ctas(seg_sum_kernel<T, LANES, VPL, MODE, FULL>, 256);                    \
    if (blocks > cap) blocks = cap;                                                                       \
    seg_sum_kernel<T, LANES, VPL, MODE, FULL><<<(unsigned)blocks, 256, 0, st>>>(                          \
        (const T*)x, ldx, (const T*)x_self, ldxs, (T*)out, ldo, n_rows, nvec, rowptr, col, row_scale,     \
        col_scale, self_coef, bias, relu);                                                                \
  }
  if (full) {
    if (mode == 2) B2G_LAUNCH(2, true) else if (mode == 1) B2G_LAUNCH(1, true) else B2G_LAUNCH(0, true)
  } else {
    if (mode == 2) B2G_LAUNCH(2, false) else if (mode == 1) B2G_LAUNCH(1, false) else B2G_LAUNCH(0, false)
  }
#undef B2G_LAUNCH
  count_launch();
  return cuda_status();
}

template <typename T>
static int dispatch_seg_sum(int nvec, const void* x, int64_t ldx, const void* x_self, int64_t ldxs,
                            void* out, int64_t ldo, int64_t n_rows, const int32_t* rowptr,
                            const int32_t* col, const float* rs, const float* cs, float self_coef,
                            const float* bias, int relu, cudaStream_t st) {
#define B2G_SS(L, V) \
  return launch_seg_sum<T, L, V>(x, ldx, x_self, ldxs, out, ldo, n_rows, nvec, rowptr, col, rs, cs, self_coef, bias, relu, st)
  if (nvec <= 8) B2G_SS(8, 1);
  if (nvec <= 16) B2G_SS(16, 1);
  if (nvec <= 32) B2G_SS(32, 1);
  if (nvec <= 64) B2G_SS(32, 2);
  if (nvec <= 128) B2G_SS(32, 4);
  if (nvec <= 256) B2G_SS(32, 8);
#undef B2G_SS
  return B2G_E_SHAPE;
}

// ---------------------------------------------------------------- halo row gather / scatter-add
template <typename T>
__global__ void __launch_bounds__(256) rows_gather_kernel(const T* __restrict__ x, int64_t ldx,
                                                          const int32_t* __restrict__ idx, int64_t n_idx,
                                                          T* __restrict__ out, int64_t ldo, int nvec) {
  constexpr int VN = Vec<T>::N;
  const int64_t total = n_idx * nvec;
  if (total < (1ll << 32)) {                                    // 32-bit index arithmetic (the 64-bit division below ran the
    const uint32_t tot = (uint32_t)total, step = gridDim.x * blockDim.x, nv = (uint32_t)nvec;   // identity copy at 3.5 TB/s)
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < tot; t += step) {
      const uint32_t r = t / nv, v = t - r * nv;
      const int64_t src = idx ? (int64_t)__ldg(idx + r) : (int64_t)r;   // idx == NULL: identity (a strided row copy)
      stg_vec<T>(out + (int64_t)r * ldo + v * VN, ldg_vec<T>(x + src * ldx + v * VN));
      if (t + step < t) break;                                  // uint32 wrap
    }
    return;
  }
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / nvec;
    const int v = (int)(t - r * nvec);
    const int64_t src = idx ? (int64_t)__ldg(idx + r) : r;
    stg_vec<T>(out + r * ldo + v * VN, ldg_vec<T>(x + src * ldx + v * VN));
  }
}
template <typename T>
__global__ void __launch_bounds__(256) rows_scatter_add_kernel(T* __restrict__ x, int64_t ldx,
                                                               const int32_t* __restrict__ idx, int64_t n_idx,
                                                               const T* __restrict__ in, int64_t ldi, int nvec) {
  constexpr int VN = Vec<T>::N;
  const int64_t total = n_idx * nvec;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / nvec;
    const int v = (int)(t - r * nvec);
    T* dst = x + (int64_t)__ldg(idx + r) * ldx + v * VN;
    float a[VN], b[VN];
    Vec<T> va = *reinterpret_cast<const Vec<T>*>(dst);
    va.to_float(a);
    ldg_vec<T>(in + r * ldi + v * VN).to_float(b);
#pragma unroll
    for (int k = 0; k < VN; ++k) a[k] += b[k];
    va.from_float(a);
    *reinterpret_cast<Vec<T>*>(dst) = va;
  }
}

// ---------------------------------------------------------------- column sums (bias gradients)
constexpr int COLSUM_BLOCKS = B2G_NUM_SMS * 4;
template <typename T>
__global__ void __launch_bounds__(256) colsum_partial_kernel(const T* __restrict__ x, int64_t ldx, int64_t n_rows,
                                                             int nvec, float* __restrict__ partial) {
  constexpr int VN = Vec<T>::N;
  __shared__ float red[256 * VN];
  const int rows_per_iter = 256 / nvec;  // nvec <= 256
  const int r_in = threadIdx.x / nvec, v = threadIdx.x % nvec;
  float acc[VN];
#pragma unroll
  for (int k = 0; k < VN; ++k) acc[k] = 0.f;
  if (r_in < rows_per_iter) {
    for (int64_t row = (int64_t)blockIdx.x * rows_per_iter + r_in; row < n_rows;
         row += (int64_t)gridDim.x * rows_per_iter) {
      float f[VN];
      ldg_vec<T>(x + row * ldx + v * VN).to_float(f);
#pragma unroll
      for (int k = 0; k < VN; ++k) acc[k] += f[k];
    }
  }
#pragma unroll
  for (int k = 0; k < VN; ++k) red[threadIdx.x * VN + k] = acc[k];
  __syncthreads();
  if (threadIdx.x < nvec) {
#pragma unroll
    for (int k = 0; k < VN; ++k) {
      float s = 0.f;
      for (int r = 0; r < rows_per_iter; ++r) s += red[(r * nvec + threadIdx.x) * VN + k];  // fixed order
      partial[(int64_t)blockIdx.x * nvec * VN + threadIdx.x * VN + k] = s;
    }
  }
}
__global__ void __launch_bounds__(256) colsum_final_kernel(const float* __restrict__ partial, int nb, int F,
                                                           float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= F) return;
  float s = 0.f;
  for (int b = 0; b < nb; ++b) s += partial[(int64_t)b * F + c];
  out[c] = s;
}

static inline unsigned grid_for(int64_t n, int threads, int per_sm = 8) {
  int64_t b = ceil_div(n > 0 ? n : 1, threads);
  const int64_t cap = (int64_t)B2G_NUM_SMS * per_sm;
  return (unsigned)(b < cap ? b : cap);
}

}  // namespace b2g

namespace b2g {
int rows_seg_sum(const void*, int64_t, void*, int64_t, int64_t, int, int, const int32_t*, const int32_t*, const float*,
                 const float*, float, const float*, int, int64_t, int64_t, int, int, cudaStream_t);
}  // namespace b2g

using namespace b2g;

static inline int elem_size(int dt) { return dt == B2G_F32 ? 4 : 2; }
static inline bool row_ok(const void* p, int64_t ld, int dt) {
  return aligned16(p) && ((ld * elem_size(dt)) % 16 == 0);
}

extern "C" {

static int seg_sum_impl(const void* x, int64_t ldx, const void* x_self, int64_t ldxs, void* out, int64_t ldo, int64_t n_rows,
                        int F, int dt, const int32_t* rowptr, const int32_t* col, const float* row_scale,
                        const float* col_scale, float self_coef, const float* bias, int relu, int64_t band,
                        int64_t max_row_len, void* stream, int impl = 0, int chunk_rows = 0, int panel_rows = 0) {
  if (impl < 0 || impl > 1) return B2G_E_ARG;
  if (n_rows < 0 || F <= 0 || (dt != B2G_F32 && dt != B2G_BF16)) return B2G_E_ARG;
  if (n_rows == 0) return B2G_OK;
  if (!x || !out || !rowptr) return B2G_E_ARG;
  if ((F * elem_size(dt)) % 16 != 0) return B2G_E_SHAPE;
  if (!row_ok(x, ldx, dt) || !row_ok(out, ldo, dt) || (x_self && !row_ok(x_self, ldxs, dt))) return B2G_E_ALIGN;
  const int nvec = F * elem_size(dt) / 16;
  cudaStream_t st = (cudaStream_t)stream;
  // rows of whole 512-byte multiples (F = 256 bf16, F = 128/256 fp32, ...): warp-per-row fast path (aggregate_rows.cu)
  if (impl != 1 && !x_self) {
    const int rc = rows_seg_sum(x, ldx, out, ldo, n_rows, nvec, dt, rowptr, col, row_scale, col_scale, self_coef, bias, relu,
                                band > 0 ? band : 0, max_row_len > 0 ? max_row_len : 0, chunk_rows, panel_rows, st);
    if (rc != B2G_E_UNSUPPORTED) return rc;
  }
  if (dt == B2G_F32)
    return dispatch_seg_sum<float>(nvec, x, ldx, x_self, ldxs, out, ldo, n_rows, rowptr, col, row_scale, col_scale, self_coef, bias, relu, st);
  return dispatch_seg_sum<__nv_bfloat16>(nvec, x, ldx, x_self, ldxs, out, ldo, n_rows, rowptr, col, row_scale, col_scale, self_coef, bias, relu, st);
}

int b2g_seg_sum(const void* x, int64_t ldx, const void* x_self, int64_t ldxs, void* out,
                int64_t ldo, int64_t n_rows, int F, int dt, const int32_t* rowptr,
                const int32_t* col, const float* row_scale, const float* col_scale,
                float self_coef, const float* bias, int relu, void* stream) {
  return seg_sum_impl(x, ldx, x_self, ldxs, out, ldo, n_rows, F, dt, rowptr, col, row_scale, col_scale, self_coef, bias, relu,
                      0, 0, stream);
}

int b2g_seg_sum_banded(const void* x, int64_t ldx, const void* x_self, int64_t ldxs, void* out,
                       int64_t ldo, int64_t n_rows, int F, int dt, const int32_t* rowptr,
                       const int32_t* col, const float* row_scale, const float* col_scale,
                       float self_coef, const float* bias, int relu, int64_t band, void* stream) {
  return seg_sum_impl(x, ldx, x_self, ldxs, out, ldo, n_rows, F, dt, rowptr, col, row_scale, col_scale, self_coef, bias, relu,
                      band, 0, stream);
}

int b2g_seg_sum_hinted(const void* x, int64_t ldx, const void* x_self, int64_t ldxs, void* out,
                       int64_t ldo, int64_t n_rows, int F, int dt, const int32_t* rowptr,
                       const int32_t* col, const float* row_scale, const float* col_scale,
                       float self_coef, const float* bias, int relu, int64_t band, int64_t max_row_len, void* stream) {
  return seg_sum_impl(x, ldx, x_self, ldxs, out, ldo, n_rows, F, dt, rowptr, col, row_scale, col_scale, self_coef, bias, relu,
                      band, max_row_len, stream);
}

int b2g_seg_sum_tuned(const void* x, int64_t ldx, const void* x_self, int64_t ldxs, void* out,
                      int64_t ldo, int64_t n_rows, int F, int dt, const int32_t* rowptr,
                      const int32_t* col, const float* row_scale, const float* col_scale,
                      float self_coef, const float* bias, int relu, int64_t band, int64_t max_row_len, int impl,
                      int chunk_rows, int panel_rows, void* stream) {
  return seg_sum_impl(x, ldx, x_self, ldxs, out, ldo, n_rows, F, dt, rowptr, col, row_scale, col_scale, self_coef, bias, relu,
                      band, max_row_len, stream, impl, chunk_rows, panel_rows);
}

int64_t b2g_colsum_workspace_bytes(int F) { return F > 0 ? (int64_t)COLSUM_BLOCKS * F * 4 : B2G_E_ARG; }

int b2g_colsum(const void* x, int64_t ldx, int64_t n_rows, int F, int dt, float* out, void* ws,
               void* stream) {
  if (n_rows < 0 || F <= 0 || (dt != B2G_F32 && dt != B2G_BF16) || !out || !ws) return B2G_E_ARG;
  if ((F * elem_size(dt)) % 16 != 0) return B2G_E_SHAPE;
  if (n_rows && !row_ok(x, ldx, dt)) return B2G_E_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  const int vn = 16 / elem_size(dt);
  const int nvec_total = F / vn;
  // wide rows (fused q|k|v|skip gradients: F = 3328) are reduced 256 vectors at a time; launches are stream-ordered
  // so the partial buffer is reused.
  for (int v0 = 0; v0 < nvec_total; v0 += 256) {
    const int nvec = nvec_total - v0 < 256 ? nvec_total - v0 : 256;
    const int c0 = v0 * vn, Fc = nvec * vn;
    if (dt == B2G_F32)
      colsum_partial_kernel<float><<<COLSUM_BLOCKS, 256, 0, st>>>((const float*)x + c0, ldx, n_rows, nvec, (float*)ws);
    else
      colsum_partial_kernel<__nv_bfloat16><<<COLSUM_BLOCKS, 256, 0, st>>>((const __nv_bfloat16*)x + c0, ldx, n_rows, nvec, (float*)ws);
    colsum_final_kernel<<<(unsigned)ceil_div(Fc, 256), 256, 0, st>>>((const float*)ws, COLSUM_BLOCKS, Fc, out + c0);
    count_launch(2);
  }
  return cuda_status();
}

int b2g_rows_gather(const void* x, int64_t ldx, const int32_t* idx, int64_t n_idx, void* out,
                    int64_t ldo, int F, int dt, void* stream) {
  if (n_idx < 0 || F <= 0 || (dt != B2G_F32 && dt != B2G_BF16)) return B2G_E_ARG;
  if (n_idx == 0) return B2G_OK;
  if (!x || !out) return B2G_E_ARG;                            // idx == NULL: rows 0 .. n_idx-1 (copy into a column block)
  if ((F * elem_size(dt)) % 16 != 0) return B2G_E_SHAPE;
  if (!row_ok(x, ldx, dt) || !row_ok(out, ldo, dt)) return B2G_E_ALIGN;
  const int nvec = F * elem_size(dt) / 16;
  cudaStream_t st = (cudaStream_t)stream;
  if (dt == B2G_F32)
    rows_gather_kernel<float><<<grid_for(n_idx * nvec, 256), 256, 0, st>>>((const float*)x, ldx, idx, n_idx, (float*)out, ldo, nvec);
  else
    rows_gather_kernel<__nv_bfloat16><<<grid_for(n_idx * nvec, 256), 256, 0, st>>>((const __nv_bfloat16*)x, ldx, idx, n_idx, (__nv_bfloat16*)out, ldo, nvec);
  count_launch();
  return cuda_status();
}

int b2g_rows_scatter_add(void* x, int64_t ldx, const int32_t* idx, int64_t n_idx, const void* in,
                         int64_t ldi, int F, int dt, void* stream) {
  if (n_idx < 0 || F <= 0 || (dt != B2G_F32 && dt != B2G_BF16)) return B2G_E_ARG;
  if (n_idx == 0) return B2G_OK;
  if (!x || !idx || !in) return B2G_E_ARG;
  if ((F * elem_size(dt)) % 16 != 0) return B2G_E_SHAPE;
  if (!row_ok(x, ldx, dt) || !row_ok(in, ldi, dt)) return B2G_E_ALIGN;
  const int nvec = F * elem_size(dt) / 16;
  cudaStream_t st = (cudaStream_t)stream;
  if (dt == B2G_F32)
    rows_scatter_add_kernel<float><<<grid_for(n_idx * nvec, 256), 256, 0, st>>>((float*)x, ldx, idx, n_idx, (const float*)in, ldi, nvec);
  else
    rows_scatter_add_kernel<__nv_bfloat16><<<grid_for(n_idx * nvec, 256), 256, 0, st>>>((__nv_bfloat16*)x, ldx, idx, n_idx, (const __nv_bfloat16*)in, ldi, nvec);
  count_launch();
  return cuda_status();
}

}  // extern "C"
