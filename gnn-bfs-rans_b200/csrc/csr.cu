// csr.cu — K1: device CSR builder for the edge list a layer aggregates over.
// No reference counterpart: PyG re-derives gcn_norm / remove+add self loops on every forward and
// scatters by edge (SURVEY §2 "implicit kernel inventory").  Here the effective edge list
//   self_loops=0: edge_index as given                       (GINConv, TransformerConv)
//   self_loops=1: non-loop edges, then one (v,v) per node   (GCNConv gcn_norm, GATConv)
// is grouped by target (or source) ONCE per distinct edge_index and cached by the host.
// Deterministic: count -> scan -> atomic fill -> per-row sort by edge id == a stable sort by key,
// so fp32 summation order equals edge order (what torch's CPU scatter_add_ does).
// Edge ids: original position e for kept edges, E+v for the appended loop of node v.
#include "scan.cuh"

namespace b2g {

constexpr int ROWSORT_MAX = 32;  // rows up to this degree are sorted by one thread in registers

__global__ void __launch_bounds__(256) csr_count_kernel(const int64_t* __restrict__ ei, int64_t E, int64_t N,
                                                        int self_loops, int by_source,
                                                        int32_t* __restrict__ cnt,
                                                        unsigned long long* __restrict__ n_bad) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = ei[e], d = ei[E + e];
    if (s < 0 || s >= N || d < 0 || d >= N) { atomicAdd(n_bad, 1ull); continue; }
    if (self_loops && s == d) continue;
    atomicAdd(&cnt[by_source ? s : d], 1);
  }
}

struct DegFlag {
  const int32_t* cnt;
  int extra;
  __device__ __forceinline__ int operator()(int64_t i) const { return cnt[i] + extra; }
};
struct RowptrWrite {
  int32_t* rowptr;
  __device__ __forceinline__ void operator()(int64_t i, int64_t pos) const { rowptr[i] = (int32_t)pos; }
};
__global__ void csr_finish_count(const int64_t* total, const unsigned long long* n_bad, int64_t N,
                                 int32_t* rowptr, int64_t* nnz_out) {
  rowptr[N] = (int32_t)*total;
  nnz_out[0] = *total;
  nnz_out[1] = (int64_t)*n_bad;
}

__global__ void __launch_bounds__(256) csr_fill_kernel(const int64_t* __restrict__ ei, int64_t E, int64_t N,
                                                       int self_loops, int by_source,
                                                       const int32_t* __restrict__ rowptr,
                                                       int32_t* __restrict__ cursor,
                                                       int32_t* __restrict__ eid) {
  const int64_t total = self_loops ? E + N : E;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    int64_t key;
    if (e < E) {
      const int64_t s = ei[e], d = ei[E + e];
      if (s < 0 || s >= N || d < 0 || d >= N) continue;
      if (self_loops && s == d) continue;
      key = by_source ? s : d;
    } else {
      key = e - E;  // appended loop of node key
    }
    const int slot = atomicAdd(&cursor[key], 1);
    eid[rowptr[key] + slot] = (int32_t)e;
  }
}

// Sort one row's edge ids ascending (== stable order) with a branch-free rank sort held in
// registers, and emit col.  Ids are unique, so ranks are a permutation.
template <int MAXD>
__device__ __forceinline__ void sort_row(int b, int deg, int32_t* __restrict__ eid,
                                         int32_t* __restrict__ col, const int64_t* __restrict__ other,
                                         int64_t E) {
  int32_t v[MAXD];
#pragma unroll
  for (int k = 0; k < MAXD; ++k) v[k] = k < deg ? eid[b + k] : 0x7fffffff;
#pragma unroll
  for (int k = 0; k < MAXD; ++k) {
    if (k < deg) {
      const int32_t x = v[k];
      int r = 0;
#pragma unroll
      for (int u = 0; u < MAXD; ++u) r += (v[u] < x);
      eid[b + r] = x;
      col[b + r] = (int32_t)(x < E ? other[x] : (int64_t)(x - E));
    }
  }
}

// One thread per row (mesh rows have <= ~8 entries); rows above ROWSORT_MAX go to the heavy list.
__global__ void __launch_bounds__(256) csr_rowsort_kernel(const int64_t* __restrict__ ei, int64_t E, int64_t N,
                                                          int by_source, const int32_t* __restrict__ rowptr,
                                                          int32_t* __restrict__ eid, int32_t* __restrict__ col,
                                                          float* __restrict__ dinv,
                                                          int32_t* __restrict__ heavy, int32_t* __restrict__ n_heavy) {
  const int64_t* __restrict__ other = by_source ? ei + E : ei;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int b = rowptr[i], deg = rowptr[i + 1] - b;
    if (dinv) dinv[i] = deg > 0 ? __fdiv_rn(1.0f, __fsqrt_rn((float)deg)) : 0.0f;
    if (deg <= 8) sort_row<8>(b, deg, eid, col, other, E);
    else if (deg <= 16) sort_row<16>(b, deg, eid, col, other, E);
    else if (deg <= ROWSORT_MAX) sort_row<ROWSORT_MAX>(b, deg, eid, col, other, E);
    else heavy[atomicAdd(n_heavy, 1)] = (int32_t)i;
  }
}

// Heavy rows (deg > ROWSORT_MAX): one CTA per row from the heavy list, rank sort through a copy.
__global__ void __launch_bounds__(256) csr_heavysort_kernel(const int64_t* __restrict__ ei, int64_t E,
                                                            int by_source, const int32_t* __restrict__ rowptr,
                                                            int32_t* __restrict__ eid, int32_t* __restrict__ col,
                                                            int32_t* __restrict__ tmp,
                                                            const int32_t* __restrict__ heavy,
                                                            const int32_t* __restrict__ n_heavy) {
  const int64_t* __restrict__ other = by_source ? ei + E : ei;
  const int nh = *n_heavy;
  for (int h = blockIdx.x; h < nh; h += gridDim.x) {
    const int i = heavy[h];
    const int b = rowptr[i], deg = rowptr[i + 1] - b;
    for (int t = threadIdx.x; t < deg; t += blockDim.x) tmp[b + t] = eid[b + t];
    __syncthreads();
    for (int t = threadIdx.x; t < deg; t += blockDim.x) {
      const int32_t x = tmp[b + t];
      int r = 0;
      for (int u = 0; u < deg; ++u) r += (tmp[b + u] < x);
      eid[b + r] = x;
      col[b + r] = (int32_t)(x < E ? other[x] : (int64_t)(x - E));
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) csr_inverse_kernel(const int32_t* __restrict__ eid_a, int64_t nnz,
                                                          int32_t* __restrict__ pos_of_id) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < nnz;
       p += (int64_t)gridDim.x * blockDim.x)
    pos_of_id[eid_a[p]] = (int32_t)p;
}
__global__ void __launch_bounds__(256) csr_perm_kernel(const int32_t* __restrict__ eid_b, int64_t nnz,
                                                       const int32_t* __restrict__ pos_of_id,
                                                       int32_t* __restrict__ perm) {
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nnz;
       q += (int64_t)gridDim.x * blockDim.x)
    perm[q] = pos_of_id[eid_b[q]];
}

static inline unsigned grid_for(int64_t n, int threads, int per_sm = 8) {
  int64_t b = ceil_div(n > 0 ? n : 1, threads);
  const int64_t cap = (int64_t)B2G_NUM_SMS * per_sm;
  return (unsigned)(b < cap ? b : cap);
}

// workspace: [cnt/cursor int32 N] [heavy int32 N] [n_heavy + n_bad: 256 B] [tmp int32 E+N] [tiles]
struct CsrWs {
  int32_t* cnt;
  int32_t* heavy;
  int32_t* n_heavy;
  unsigned long long* n_bad;
  int32_t* tmp;
  int64_t* tiles;
  int64_t bytes;
};
static CsrWs carve(void* ws, int64_t E, int64_t N) {
  CsrWs w;
  uint8_t* p = static_cast<uint8_t*>(ws);
  int64_t off = 0;
  auto pad = [](int64_t b) { return ceil_div(b > 0 ? b : 1, 256) * 256; };
  w.cnt = reinterpret_cast<int32_t*>(p + off); off += pad(4 * N);
  w.heavy = reinterpret_cast<int32_t*>(p + off); off += pad(4 * N);
  w.n_heavy = reinterpret_cast<int32_t*>(p + off);
  w.n_bad = reinterpret_cast<unsigned long long*>(p + off + 8); off += 256;
  w.tmp = reinterpret_cast<int32_t*>(p + off); off += pad(4 * (E + N));
  w.tiles = reinterpret_cast<int64_t*>(p + off); off += pad(scan_ws_bytes(N));
  w.bytes = off;
  return w;
}

}  // namespace b2g

using namespace b2g;

extern "C" {

int64_t b2g_csr_workspace_bytes(int64_t E, int64_t N) {
  if (E < 0 || N < 0) return B2G_E_ARG;
  return carve(nullptr, E, N).bytes;
}

int b2g_csr_count(const int64_t* edge_index, int64_t E, int64_t N, int self_loops, int by_source,
                  int32_t* rowptr, int64_t* nnz_out, void* ws, void* stream) {
  if (E < 0 || N < 0 || !rowptr || !nnz_out || !ws || (E && !edge_index)) return B2G_E_ARG;
  if (E + N >= (int64_t)0x7fffffff) return B2G_E_RANGE;
  cudaStream_t st = (cudaStream_t)stream;
  CsrWs w = carve(ws, E, N);
  cudaError_t e = cudaMemsetAsync(w.cnt, 0, (uint8_t*)w.tmp - (uint8_t*)w.cnt, st);  // cnt, heavy, counters
  if (e != cudaSuccess) return (int)e;
  if (E) {
    csr_count_kernel<<<grid_for(E, 256), 256, 0, st>>>(edge_index, E, N, self_loops, by_source, w.cnt, w.n_bad);
    count_launch();
  }
  DegFlag f{w.cnt, self_loops ? 1 : 0};
  scan_count(N, w.tiles, nullptr, f, st);
  if (N) scan_consume<true>(N, w.tiles, f, RowptrWrite{rowptr}, st);
  csr_finish_count<<<1, 1, 0, st>>>(w.tiles + scan_num_tiles(N), w.n_bad, N, rowptr, nnz_out);
  count_launch();
  return cuda_status();
}

int b2g_csr_fill(const int64_t* edge_index, int64_t E, int64_t N, int self_loops, int by_source,
                 const int32_t* rowptr, int64_t nnz, int32_t* col, int32_t* eid, float* dinv,
                 void* ws, void* stream) {
  if (E < 0 || N < 0 || nnz < 0 || !rowptr || !ws || (E && !edge_index)) return B2G_E_ARG;
  if (nnz && (!col || !eid)) return B2G_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  CsrWs w = carve(ws, E, N);
  cudaError_t e = cudaMemsetAsync(w.cnt, 0, (uint8_t*)w.tmp - (uint8_t*)w.cnt, st);  // cursor, heavy, counters
  if (e != cudaSuccess) return (int)e;
  const int64_t total = self_loops ? E + N : E;
  if (total) {
    csr_fill_kernel<<<grid_for(total, 256), 256, 0, st>>>(edge_index, E, N, self_loops, by_source, rowptr, w.cnt, eid);
    count_launch();
  }
  if (N) {
    csr_rowsort_kernel<<<grid_for(N, 256), 256, 0, st>>>(edge_index, E, N, by_source, rowptr, eid, col, dinv, w.heavy, w.n_heavy);
    csr_heavysort_kernel<<<B2G_NUM_SMS * 2, 256, 0, st>>>(edge_index, E, by_source, rowptr, eid, col, w.tmp, w.heavy, w.n_heavy);
    count_launch(2);
  }
  return cuda_status();
}

int b2g_csr_perm(const int32_t* eid_a, const int32_t* eid_b, int64_t nnz, int32_t* scratch,
                 int32_t* perm, void* stream) {
  if (nnz < 0) return B2G_E_ARG;
  if (nnz == 0) return B2G_OK;
  if (!eid_a || !eid_b || !scratch || !perm) return B2G_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  csr_inverse_kernel<<<grid_for(nnz, 256), 256, 0, st>>>(eid_a, nnz, scratch);
  csr_perm_kernel<<<grid_for(nnz, 256), 256, 0, st>>>(eid_b, nnz, scratch, perm);
  count_launch(2);
  return cuda_status();
}

}  // extern "C"
