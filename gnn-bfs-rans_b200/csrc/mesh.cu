// mesh.cu — device-side mesh ingest: the two arrays the reference loader derives from the polyMesh
// connectivity with Python loops over every face (/root/reference/openfoam_loader.py):
//   get_cell_centers   (:191-227)  cell centre = mean of the UNIQUE vertices of the faces a cell owns or neighbours
//   get_internal_cells (:229-248)  cells named by `neighbour` + owners of the first len(neighbour) faces
// plus n_cells = max(max(owner), max(neighbour)) + 1 (:197).
//
// Same shape as the CSR builder (csr.cu): count vertices per cell -> exclusive scan -> atomic fill of the per-cell vertex
// lists -> one thread per cell walks its list in ASCENDING vertex id, skipping duplicates (repeated "smallest id greater
// than the last one": no sort, no scratch, independent of the order the atomics filled the list), so the fp64 sum is
// deterministic.  The reference adds in CPython set order; fp64 addition is not associative, so parity is 1e-13
// absolute, not bit equality (oracle/mesh_oracle.py states the same).  All HBM-bound integer work: 4 B per vertex slot
// read twice, 4 B written and read once, 24 B per unique vertex gathered.
#include <climits>

#include "scan.cuh"

namespace b2g {

struct MeshWs {
  int32_t* cnt;       // [n_cells] vertex slots per cell (with duplicates)
  int32_t* cursor;    // [n_cells]
  int64_t* start;     // [n_cells + 1]
  int32_t* verts;     // [n_slots]
  int64_t* tiles;
  unsigned long long* n_bad;
  int64_t bytes;
};

static MeshWs mesh_ws(void* base, int64_t n_cells, int64_t n_slots) {
  auto pad = [](int64_t b) { return (b + 255) / 256 * 256; };
  char* p = static_cast<char*>(base);
  int64_t off = 0;
  MeshWs w;
  w.n_bad = reinterpret_cast<unsigned long long*>(p + off); off += 256;
  w.cnt = reinterpret_cast<int32_t*>(p + off); off += pad(n_cells * 4);
  w.cursor = reinterpret_cast<int32_t*>(p + off); off += pad(n_cells * 4);
  w.start = reinterpret_cast<int64_t*>(p + off); off += pad((n_cells + 1) * 8);
  w.verts = reinterpret_cast<int32_t*>(p + off); off += pad(n_slots * 4);
  w.tiles = reinterpret_cast<int64_t*>(p + off); off += pad(scan_ws_bytes(n_cells));
  w.bytes = off;
  return w;
}

struct MeshIn {
  const int32_t* owner;
  const int32_t* neighbour;
  const int64_t* face_off;
  const int32_t* face_pts;
  int64_t n_owner, n_nb, n_faces, n_cells, n_points, n_slots;
};

// item t < n_owner: face t on its owner's side; item n_owner + t: face t on its neighbour's side (:203-214)
__device__ __forceinline__ bool mesh_item(const MeshIn& m, int64_t t, int64_t& face, int64_t& cell) {
  if (t < m.n_owner) { face = t; cell = m.owner[t]; }
  else { face = t - m.n_owner; cell = m.neighbour[face]; }
  return face < m.n_faces && cell >= 0 && cell < m.n_cells;
}

__global__ void __launch_bounds__(256) mesh_count_kernel(const MeshIn m, int32_t* __restrict__ cnt,
                                                         unsigned long long* __restrict__ n_bad) {
  const int64_t items = m.n_owner + m.n_nb;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < items; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t f, c;
    if (!mesh_item(m, t, f, c)) { atomicAdd(n_bad, 1ull); continue; }
    atomicAdd(&cnt[c], (int32_t)(m.face_off[f + 1] - m.face_off[f]));
  }
}

struct CntRead {
  const int32_t* cnt;
  __device__ __forceinline__ int operator()(int64_t i) const { return cnt[i]; }
};
struct StartWrite {
  int64_t* start;
  __device__ __forceinline__ void operator()(int64_t i, int64_t pos) const { start[i] = pos; }
};
__global__ void mesh_finish_scan(const int64_t* total, int64_t n_cells, int64_t* start) { start[n_cells] = *total; }

__global__ void __launch_bounds__(256) mesh_fill_kernel(const MeshIn m, const int64_t* __restrict__ start,
                                                        int32_t* __restrict__ cursor, int32_t* __restrict__ verts,
                                                        unsigned long long* __restrict__ n_bad) {
  const int64_t items = m.n_owner + m.n_nb;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < items; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t f, c;
    if (!mesh_item(m, t, f, c)) continue;                       // counted as bad by mesh_count_kernel
    const int64_t b = m.face_off[f], n = m.face_off[f + 1] - b;
    const int64_t pos = start[c] + atomicAdd(&cursor[c], (int32_t)n);
    if (pos + n > m.n_slots) { atomicAdd(n_bad, 1ull); continue; }
    for (int64_t k = 0; k < n; ++k) {
      const int32_t p = m.face_pts[b + k];
      if (p < 0 || p >= m.n_points) atomicAdd(n_bad, 1ull);
      verts[pos + k] = p;
    }
  }
}

__global__ void __launch_bounds__(128) mesh_center_kernel(const double* __restrict__ points, int64_t n_points,
                                                          const int64_t* __restrict__ start,
                                                          const int32_t* __restrict__ verts, int64_t n_slots,
                                                          int64_t n_cells, double* __restrict__ out) {
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_cells; c += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = start[c], e = min(start[c + 1], n_slots);
    double sx = 0.0, sy = 0.0, sz = 0.0;
    int64_t uniq = 0;
    int32_t last = -1;
    while (true) {                                              // next unique vertex: the smallest id above `last`
      int32_t nxt = 0x7fffffff;
      for (int64_t t = b; t < e; ++t) {
        const int32_t p = verts[t];
        if (p > last && p < nxt) nxt = p;
      }
      if (nxt == 0x7fffffff || nxt >= n_points) break;          // ids past n_points were reported through n_bad
      sx += points[3 * (int64_t)nxt + 0];
      sy += points[3 * (int64_t)nxt + 1];
      sz += points[3 * (int64_t)nxt + 2];
      ++uniq;
      last = nxt;
    }
    const double inv = (double)uniq;
    out[3 * c + 0] = uniq ? sx / inv : 0.0;                     // :218 np.mean = sum / count; :221-223 zeros otherwise
    out[3 * c + 1] = uniq ? sy / inv : 0.0;
    out[3 * c + 2] = uniq ? sz / inv : 0.0;
  }
}

__global__ void __launch_bounds__(256) mesh_max_kernel(const int32_t* __restrict__ a, int64_t na,
                                                       const int32_t* __restrict__ b, int64_t nb,
                                                       int* __restrict__ out) {
  int m = INT_MIN;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < na + nb; t += (int64_t)gridDim.x * blockDim.x)
    m = max(m, t < na ? a[t] : b[t - na]);
#pragma unroll
  for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m != INT_MIN) atomicMax(out, m);
}
__global__ void mesh_max_finish(const int* m, int64_t* out) { out[0] = (int64_t)*m + 1; }

__global__ void __launch_bounds__(256) mesh_internal_kernel(const int32_t* __restrict__ owner,
                                                            const int32_t* __restrict__ neighbour, int64_t n_nb,
                                                            int64_t n_cells, uint8_t* __restrict__ mask,
                                                            unsigned long long* __restrict__ n_bad) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_nb; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t a = neighbour[t], b = owner[t];
    if (a < 0 || a >= n_cells || b < 0 || b >= n_cells) { atomicAdd(n_bad, 1ull); continue; }
    mask[a] = 1;                                                // :239-240
    mask[b] = 1;                                                // :243-244
  }
}
__global__ void mesh_copy_bad(const unsigned long long* n_bad, int64_t* out) { out[0] = (int64_t)*n_bad; }

static unsigned mesh_grid(int64_t n, int threads) {
  const int64_t want = ceil_div(n > 0 ? n : 1, threads);
  const int64_t cap = (int64_t)B2G_NUM_SMS * 16;
  return (unsigned)(want < cap ? want : cap);
}

}  // namespace b2g

using namespace b2g;

extern "C" {

int b2g_mesh_num_cells(const int32_t* owner, int64_t n_owner, const int32_t* neighbour, int64_t n_nb, int64_t* n_cells_out,
                       void* ws, void* stream) {
  if (n_owner < 0 || n_nb < 0 || !n_cells_out || !ws || (n_owner && !owner) || (n_nb && !neighbour)) return B2G_E_ARG;
  if (n_owner + n_nb == 0) return B2G_E_ARG;                    // np.max of an empty array raises in the reference
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int* m = static_cast<int*>(ws);
  const int init = INT_MIN;
  cudaMemcpyAsync(m, &init, sizeof(int), cudaMemcpyHostToDevice, st);
  mesh_max_kernel<<<mesh_grid(n_owner + n_nb, 256), 256, 0, st>>>(owner, n_owner, neighbour, n_nb, m);
  mesh_max_finish<<<1, 1, 0, st>>>(m, n_cells_out);
  count_launch(2);
  return cuda_status();
}

int64_t b2g_mesh_workspace_bytes(int64_t n_cells, int64_t n_slots) {
  if (n_cells < 0 || n_slots < 0) return B2G_E_ARG;
  return mesh_ws(nullptr, n_cells, n_slots).bytes;
}

int b2g_mesh_cell_centers(const double* points, int64_t n_points, const int32_t* owner, int64_t n_owner,
                          const int32_t* neighbour, int64_t n_nb, const int64_t* face_off, const int32_t* face_pts,
                          int64_t n_faces, int64_t n_slots, int64_t n_cells, double* centers, int64_t* n_bad_out,
                          void* ws, int64_t ws_bytes, void* stream) {
  if (n_points < 0 || n_owner < 0 || n_nb < 0 || n_faces < 0 || n_cells < 0 || !face_off || !n_bad_out || !ws) return B2G_E_ARG;
  if ((n_points && !points) || (n_owner && !owner) || (n_nb && !neighbour) || (n_cells && !centers)) return B2G_E_ARG;
  if (n_owner > n_faces || n_nb > n_faces) return B2G_E_ARG;   // faces[i] IndexError in the reference (:205, :212)
  if (n_points >= 0x7fffffff) return B2G_E_RANGE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // n_slots = face_off[n_owner] + face_off[n_nb]: the vertex slots of both sides (the host knows it from flattening the
  // faces); a device-side total that disagrees is reported through n_bad instead of overrunning the workspace
  if (n_slots < 0 || (n_slots && !face_pts)) return B2G_E_ARG;
  MeshWs w = mesh_ws(ws, n_cells, n_slots);
  if (ws_bytes < w.bytes) return B2G_E_ARG;
  MeshIn m{owner, neighbour, face_off, face_pts, n_owner, n_nb, n_faces, n_cells, n_points, n_slots};
  cudaMemsetAsync(w.n_bad, 0, 256, st);
  cudaMemsetAsync(w.cnt, 0, (size_t)n_cells * 4, st);
  cudaMemsetAsync(w.cursor, 0, (size_t)n_cells * 4, st);
  const int64_t items = n_owner + n_nb;
  if (items) mesh_count_kernel<<<mesh_grid(items, 256), 256, 0, st>>>(m, w.cnt, w.n_bad);
  CntRead f{w.cnt};
  scan_count(n_cells, w.tiles, nullptr, f, st);
  if (n_cells) scan_consume<true>(n_cells, w.tiles, f, StartWrite{w.start}, st);
  mesh_finish_scan<<<1, 1, 0, st>>>(w.tiles + scan_num_tiles(n_cells), n_cells, w.start);
  if (items) mesh_fill_kernel<<<mesh_grid(items, 256), 256, 0, st>>>(m, w.start, w.cursor, w.verts, w.n_bad);
  if (n_cells) mesh_center_kernel<<<mesh_grid(n_cells, 128), 128, 0, st>>>(points, n_points, w.start, w.verts, n_slots, n_cells, centers);
  mesh_copy_bad<<<1, 1, 0, st>>>(w.n_bad, n_bad_out);
  count_launch(7);
  return cuda_status();
}

int b2g_mesh_internal_cells(const int32_t* owner, int64_t n_owner, const int32_t* neighbour, int64_t n_nb, int64_t n_cells,
                            uint8_t* mask, int64_t* n_bad_out, void* ws, void* stream) {
  if (n_owner < 0 || n_nb < 0 || n_cells < 0 || !n_bad_out || !ws || (n_cells && !mask)) return B2G_E_ARG;
  if (n_nb > n_owner) return B2G_E_ARG;                         // owner[i] IndexError in the reference (:244)
  if (n_nb && (!owner || !neighbour)) return B2G_E_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned long long* n_bad = static_cast<unsigned long long*>(ws);
  cudaMemsetAsync(n_bad, 0, 8, st);
  cudaMemsetAsync(mask, 0, (size_t)n_cells, st);
  if (n_nb) mesh_internal_kernel<<<mesh_grid(n_nb, 256), 256, 0, st>>>(owner, neighbour, n_nb, n_cells, mask, n_bad);
  mesh_copy_bad<<<1, 1, 0, st>>>(n_bad, n_bad_out);
  count_launch(2);
  return cuda_status();
}

}  // extern "C"
