// builder.cu — K0: OpenFOAM owner/neighbour faces -> PyG edge_index (int64 [2,E]), bit-exact with
// /root/reference/graph_constructor.py:28-56 (build_edge_index) and :109-187,220-227 (build_graph
// edge part: filtering, remap, range validation, isolated-node self loops), plus K0c edge
// attributes (:58-90, :190-219).  Pure HBM-bound integer work: ordered compaction by recomputed
// flags (scan.cuh) — the face arrays are read twice, the output is written once.
#include "scan.cuh"

namespace b2g {

// ---------------------------------------------------------------- build_edge_index (no filtering)
template <bool kVec>
__global__ void __launch_bounds__(256) build_edge_index_kernel(const int32_t* __restrict__ owner,
                                                               const int32_t* __restrict__ nei,
                                                               int64_t n_owner, int64_t n_nei,
                                                               int64_t* __restrict__ out) {
  const int64_t E = 2 * n_nei + (n_owner - n_nei);
  int64_t* __restrict__ src = out;
  int64_t* __restrict__ dst = out + E;
  for (int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; f < n_owner;
       f += (int64_t)gridDim.x * blockDim.x) {
    const int64_t o = owner[f];
    if (f < n_nei) {  // internal face: (o,n) then (n,o), interleaved            :39-45
      const int64_t n = nei[f];
      if (kVec) {  // 16-byte stores: positions 2f, 2f+1 are adjacent; row 1 is aligned iff E is even
        *reinterpret_cast<longlong2*>(src + 2 * f) = make_longlong2(o, n);
        *reinterpret_cast<longlong2*>(dst + 2 * f) = make_longlong2(n, o);
      } else {
        src[2 * f] = o; src[2 * f + 1] = n;
        dst[2 * f] = n; dst[2 * f + 1] = o;
      }
    } else {          // boundary face: one (o,o) self loop                        :49-53
      const int64_t p = 2 * n_nei + (f - n_nei);
      src[p] = o;
      dst[p] = o;
    }
  }
}

// ---------------------------------------------------------------- build_graph candidates
// Candidate c in [0, L): c < 2*n_nei -> internal face c>>1, direction c&1; else boundary face.
struct Cand {
  const int32_t* owner;
  const int32_t* nei;
  const int32_t* map;   // old_to_new or nullptr
  int64_t n_nei, n_cells, n_nodes;
  int mode;             // 0 = C, 1 = A/B
  uint8_t* touched;     // [n_nodes] (phase 1 marks; may be nullptr in phase 3)
  unsigned long long* n_bad;

  __device__ __forceinline__ bool eval(int64_t c, int64_t& s, int64_t& d) const {
    if (c < 2 * n_nei) {
      const int64_t f = c >> 1;
      int64_t o = owner[f], n = nei[f];
      if (mode == 1) {
        if (o < 0 || o >= n_cells || n < 0 || n >= n_cells) {  // reference: IndexError (:144)
          if (n_bad && (c & 1) == 0) atomicAdd(n_bad, 1ull);
          return false;
        }
        if (map) {
          o = map[o];
          n = map[n];
          if (o < 0 || n < 0) return false;                   // :144,149
        } else if (o >= n_nodes || n >= n_nodes) {
          return false;                                        // identity map on [0,n)  :113-115
        }
      }
      s = (c & 1) ? n : o;
      d = (c & 1) ? o : n;
    } else {
      s = d = owner[n_nei + (c - 2 * n_nei)];
    }
    return s < n_nodes && d < n_nodes;                         // :168-173
  }
};

struct CandCount {
  Cand k;
  __device__ __forceinline__ int operator()(int64_t c) const {
    int64_t s, d;
    if (!k.eval(c, s, d)) return 0;
    if (s >= 0) k.touched[s] = 1;   // benign same-value race                       :178
    if (d >= 0) k.touched[d] = 1;
    return 1;
  }
};
struct CandFlag {
  Cand k;
  __device__ __forceinline__ int operator()(int64_t c) const {
    int64_t s, d;
    return k.eval(c, s, d) ? 1 : 0;
  }
};
struct CandWrite {
  Cand k;
  int64_t* out;
  int64_t E_total;
  __device__ __forceinline__ void operator()(int64_t c, int64_t pos) const {
    int64_t s, d;
    k.eval(c, s, d);
    out[pos] = s;
    out[E_total + pos] = d;
  }
};
struct IsoFlag {
  const uint8_t* touched;
  __device__ __forceinline__ int operator()(int64_t v) const { return touched[v] ? 0 : 1; }
};
struct IsoWrite {
  int64_t* out;
  int64_t E_total;
  const int64_t* e_kept;  // device scalar
  __device__ __forceinline__ void operator()(int64_t v, int64_t pos) const {
    const int64_t p = *e_kept + pos;                          // :186-187 appended, ascending
    out[p] = v;
    out[E_total + p] = v;
  }
};

// workspace layout: [touched: n_nodes bytes, padded to 256] [tilesA: scan_ws(L)] [tilesB: scan_ws(n_nodes)] [n_bad: 8]
struct BuildWs {
  uint8_t* touched;
  int64_t* tilesA;
  int64_t* tilesB;
  unsigned long long* n_bad;
  int64_t bytes;
};
static BuildWs carve(void* ws, int64_t L, int64_t n_nodes) {
  BuildWs w;
  uint8_t* p = static_cast<uint8_t*>(ws);
  int64_t off = 0;
  w.touched = p + off; off += ceil_div(n_nodes > 0 ? n_nodes : 1, 256) * 256;
  w.tilesA = reinterpret_cast<int64_t*>(p + off); off += ceil_div(scan_ws_bytes(L), 256) * 256;
  w.tilesB = reinterpret_cast<int64_t*>(p + off); off += ceil_div(scan_ws_bytes(n_nodes), 256) * 256;
  w.n_bad = reinterpret_cast<unsigned long long*>(p + off); off += 256;
  w.bytes = off;
  return w;
}

__global__ void builder_finish_counts(const int64_t* tilesA_total, const int64_t* tilesB_total,
                                      const unsigned long long* n_bad, int64_t* counts) {
  counts[0] = *tilesA_total;
  counts[1] = *tilesB_total;
  counts[2] = (int64_t)*n_bad;
}

// ---------------------------------------------------------------- mask -> old_to_new
struct MaskFlag {
  const uint8_t* m;
  __device__ __forceinline__ int operator()(int64_t i) const { return m[i] ? 1 : 0; }
};
struct MapWrite {
  int32_t* map;
  __device__ __forceinline__ void operator()(int64_t i, int64_t pos) const { map[i] = (int32_t)pos; }
};

// ---------------------------------------------------------------- edge attributes
// float64 arithmetic with explicit round-to-nearest mul/add (no FMA contraction) so the fp32
// result matches numpy's sqrt(dx*dx + dy*dy + dz*dz) and d/dist bit for bit.
__global__ void __launch_bounds__(256) edge_attr_kernel(const double* __restrict__ cc, int64_t n_cc,
                                                        const int64_t* __restrict__ ei, int64_t E,
                                                        float4* __restrict__ out) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = ei[e], d = ei[E + e];
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (s != d && s >= 0 && s < n_cc && d >= 0 && d < n_cc) {   // :203-209
      const double dx = __dsub_rn(cc[3 * d + 0], cc[3 * s + 0]);
      const double dy = __dsub_rn(cc[3 * d + 1], cc[3 * s + 1]);
      const double dz = __dsub_rn(cc[3 * d + 2], cc[3 * s + 2]);
      const double q = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
      const double dist = __dsqrt_rn(q);                          // :214
      if (dist > 0.0)                                             // :215-216
        r = make_float4((float)__ddiv_rn(dx, dist), (float)__ddiv_rn(dy, dist),
                        (float)__ddiv_rn(dz, dist), (float)dist);
      else
        r = make_float4((float)dx, (float)dy, (float)dz, (float)dist);
    }
    out[e] = r;
  }
}

static inline unsigned grid_for(int64_t n, int threads, int per_sm = 8) {
  int64_t b = ceil_div(n > 0 ? n : 1, threads);
  const int64_t cap = (int64_t)B2G_NUM_SMS * per_sm;  // grid-stride: whole waves of resident CTAs
  return (unsigned)(b < cap ? b : cap);
}

}  // namespace b2g

using namespace b2g;

extern "C" {

int b2g_build_edge_index(const int32_t* owner, const int32_t* neighbour, int64_t n_owner,
                         int64_t n_nei, int64_t* out_ei, void* stream) {
  if (n_owner < 0 || n_nei < 0 || n_nei > n_owner) return B2G_E_ARG;
  if (n_owner == 0) return B2G_OK;
  if (!owner || !out_ei || (n_nei && !neighbour)) return B2G_E_ARG;
  const int64_t E = n_owner + n_nei;
  if (!(E & 1) && aligned16(out_ei))
    build_edge_index_kernel<true><<<grid_for(n_owner, 256), 256, 0, (cudaStream_t)stream>>>(owner, neighbour, n_owner, n_nei, out_ei);
  else
    build_edge_index_kernel<false><<<grid_for(n_owner, 256), 256, 0, (cudaStream_t)stream>>>(owner, neighbour, n_owner, n_nei, out_ei);
  count_launch();
  return cuda_status();
}

int64_t b2g_mask_to_map_workspace_bytes(int64_t n_cells) { return scan_ws_bytes(n_cells) + 256; }

int b2g_mask_to_map(const uint8_t* mask, int64_t n_cells, int32_t* old_to_new, int64_t* n_set_out,
                    void* ws, void* stream) {
  if (n_cells < 0 || !n_set_out || !ws) return B2G_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (n_cells == 0) return (int)cudaMemsetAsync(n_set_out, 0, 8, st);
  if (!mask || !old_to_new) return B2G_E_ARG;
  cudaError_t e = cudaMemsetAsync(old_to_new, 0xff, n_cells * sizeof(int32_t), st);  // -1
  if (e != cudaSuccess) return (int)e;
  int64_t* tiles = static_cast<int64_t*>(ws);
  scan_count(n_cells, tiles, n_set_out, MaskFlag{mask}, st);
  scan_consume(n_cells, tiles, MaskFlag{mask}, MapWrite{old_to_new}, st);
  return cuda_status();
}

int64_t b2g_build_graph_workspace_bytes(int64_t n_owner, int64_t n_nei, int64_t n_nodes) {
  if (n_owner < 0 || n_nei < 0 || n_nodes < 0) return B2G_E_ARG;
  return carve(nullptr, n_owner + n_nei, n_nodes).bytes;
}

int b2g_build_graph_count(const int32_t* owner, const int32_t* neighbour, int64_t n_owner,
                          int64_t n_nei, int mode, const int32_t* old_to_new, int64_t n_cells,
                          int64_t n_nodes, void* ws, int64_t* counts_out, void* stream) {
  if (n_owner < 0 || n_nei < 0 || n_nei > n_owner || n_nodes < 0 || n_cells < 0) return B2G_E_ARG;
  if (mode != 0 && mode != 1) return B2G_E_ARG;
  if (mode == 0 && old_to_new) return B2G_E_ARG;
  if (!ws || !counts_out || (n_owner && !owner) || (n_nei && !neighbour)) return B2G_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t L = mode == 0 ? n_owner + n_nei : 2 * n_nei;
  BuildWs w = carve(ws, n_owner + n_nei, n_nodes);
  cudaError_t e = cudaMemsetAsync(ws, 0, w.bytes, st);
  if (e != cudaSuccess) return (int)e;
  Cand k{owner, neighbour, old_to_new, n_nei, n_cells, n_nodes, mode, w.touched, w.n_bad};
  scan_count(L, w.tilesA, nullptr, CandCount{k}, st);
  scan_count(n_nodes, w.tilesB, nullptr, IsoFlag{w.touched}, st);
  builder_finish_counts<<<1, 1, 0, st>>>(w.tilesA + scan_num_tiles(L), w.tilesB + scan_num_tiles(n_nodes), w.n_bad, counts_out);
  count_launch();
  return cuda_status();
}

int b2g_build_graph_fill(const int32_t* owner, const int32_t* neighbour, int64_t n_owner,
                         int64_t n_nei, int mode, const int32_t* old_to_new, int64_t n_cells,
                         int64_t n_nodes, const void* ws, int64_t E_total, int64_t* out_ei,
                         void* stream) {
  if (n_owner < 0 || n_nei < 0 || n_nei > n_owner || n_nodes < 0 || E_total < 0) return B2G_E_ARG;
  if (mode != 0 && mode != 1) return B2G_E_ARG;
  if (E_total == 0) return B2G_OK;
  if (!ws || !out_ei) return B2G_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t L = mode == 0 ? n_owner + n_nei : 2 * n_nei;
  BuildWs w = carve(const_cast<void*>(ws), n_owner + n_nei, n_nodes);
  Cand k{owner, neighbour, old_to_new, n_nei, n_cells, n_nodes, mode, nullptr, nullptr};
  if (L > 0) scan_consume(L, w.tilesA, CandFlag{k}, CandWrite{k, out_ei, E_total}, st);
  if (n_nodes > 0)
    scan_consume(n_nodes, w.tilesB, IsoFlag{w.touched},
                 IsoWrite{out_ei, E_total, w.tilesA + scan_num_tiles(L)}, st);
  return cuda_status();
}

int b2g_edge_attr(const double* cell_centers, int64_t n_centers, const int64_t* edge_index,
                  int64_t E, float* out, void* stream) {
  if (E < 0 || n_centers < 0) return B2G_E_ARG;
  if (E == 0) return B2G_OK;
  if (!edge_index || !out || (n_centers && !cell_centers)) return B2G_E_ARG;
  if (!aligned16(out)) return B2G_E_ALIGN;
  edge_attr_kernel<<<grid_for(E, 256), 256, 0, (cudaStream_t)stream>>>(cell_centers, n_centers, edge_index, E, reinterpret_cast<float4*>(out));
  count_launch();
  return cuda_status();
}

}  // extern "C"
