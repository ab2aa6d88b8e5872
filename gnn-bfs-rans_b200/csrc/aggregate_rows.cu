// aggregate_rows.cu — K2/K3 fast path: CSR segment-sum for rows that are whole multiples of 512 bytes
// (F = 256 bf16, F = 128 / 256 fp32): one warp per target row, VPL 16-byte vectors per lane.
// Same contract as b2g_seg_sum (aggregate.cu); replaces PyG's index_select + scatter_add_ for GCNConv / GINConv
// (gnn_model.py:63,75,166) and their backward on the transposed CSR.
//
// What shaped this kernel (ncu evidence under profiles/, numbers for cfg4 = 10 M rows x 7 entries, bf16 F = 256):
//  * the first version executed 284 warp instructions per row (64-bit address arithmetic, per-slot predicates,
//    64-bit divisions in the chunk mapping, runtime-disabled epilogue code if-converted into predicated-off
//    instructions, spills) at 58 % issue-slot utilisation: with 8 warps per scheduler the ISSUE stream, not DRAM,
//    set the pace (8 x 284 issue cycles per 8 rows >= one DRAM latency).  Here every row takes ONE warp-uniform
//    branch on its length into straight-line code for exactly K entries: K x (SHFL, IMAD.WIDE.U32, LDG.128), then
//    8K FHADD.BF16 (bf16: `add.rn.f32.bf16`, exact) or 4K packed add.f32x2 (fp32); compile-time epilogue variants;
//    long rows (> BU entries) leave through a cold, non-inlined function;
//  * the issue stream is in order: an index load placed after the adds queues behind this row's data (two DRAM
//    latencies back to back per row), and one placed right before them stalls on ITS address chain.  The loop keeps
//    three rows in the pipe (gather row i / column indices of row i+1 / rowptr pair of row i+2) so that every
//    prefetch is issued ahead of the gathers with its address operands long arrived;
//  * a linear sweep re-read every feature row 2.3x from DRAM (fp32: 3x): the co-resident CTAs must work on one
//    narrow front and, on band-structured meshes, panel by panel (RowSched below): 14.3 -> 5.9 GB of reads.
#include "rows.cuh"

namespace b2g {

struct RowsArgs {
  const void* x;
  void* out;
  const int32_t* rowptr;
  const int32_t* col;
  const float* row_scale;
  const float* col_scale;
  const float* bias;
  uint32_t xrow_bytes, orow_bytes, n_rows;
  float self_coef;
  int relu;
  RowSched ord;
};

// Kernel variants (compile time; anything else goes to the generic seg_sum_kernel):
//   kW    per-entry weights row_scale[i] * col_scale[j] (FMA path) instead of plain adds
//   kSelf the row itself is entry 0 of its list (weight self_coef when kW, else exactly 1)
//   kRS   multiply the sum by row_scale[i] at the end (unweighted GCN forward; with kW the scale is in the weights)
//   kEpi  bias / ReLU epilogue (runtime flags inside)
template <int VPL>
struct RowsCfg {
  static constexpr int BU = 8;               // longest row handled by straight-line code = rows in flight per warp
};
#ifndef B2G_ROWS_MINB
#define B2G_ROWS_MINB 4
#endif
#ifndef B2G_ROWS_WARPS
#define B2G_ROWS_WARPS 8                     // warps per CTA (measured: 4 warps x 8-9 CTAs/SM is slower, 2.60-2.86 vs 2.56 ms)
#endif
constexpr int ROWS_THREADS = 32 * B2G_ROWS_WARPS;
template <int VPL, bool kW>
constexpr int rows_minb() { return (VPL == 1) ? (kW ? 3 : B2G_ROWS_MINB) : 2; }   // kW at 4 CTAs/SM measured slower (3.69 vs 3.35 ms)   // CTAs per SM the register budget is planned for

// The bias slice of a lane is the same for every row: it is loaded ONCE per warp (bias_regs) instead of two 16-byte
// loads per row in the epilogue.
template <typename T, int VPL>
__device__ __forceinline__ void rows_load_bias(const RowsArgs& a, int lane, float (&bia)[VPL][Vec<T>::N]) {
  constexpr int VN = Vec<T>::N;
#pragma unroll
  for (int v = 0; v < VPL; ++v)
#pragma unroll
    for (int k = 0; k < VN; ++k) bia[v][k] = a.bias ? __ldg(a.bias + (lane + 32 * v) * VN + k) : 0.f;
}

template <typename T, int VPL, bool kEpi>
__device__ __forceinline__ void rows_store(const RowsArgs& a, float (&acc)[VPL][Vec<T>::N], uint32_t i, int lane,
                                           const float (&bia)[VPL][Vec<T>::N]) {
  constexpr int VN = Vec<T>::N;
  char* ob = reinterpret_cast<char*>(a.out) + (uint64_t)i * a.orow_bytes + lane * 16;
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    if (kEpi) {
#pragma unroll
      for (int k = 0; k < VN; ++k) acc[v][k] = __fadd_rn(acc[v][k], bia[v][k]);   // never contracted with the row-scale multiply:
                                                                                   // same bits as the generic kernel
      if (a.relu) {
#pragma unroll
        for (int k = 0; k < VN; ++k) acc[v][k] = fmaxf(acc[v][k], 0.f);
      }
    }
    Vec<T> o;
    o.from_float(acc[v]);
    __stcs(reinterpret_cast<uint4*>(ob + 512 * v), *reinterpret_cast<uint4*>(&o.v));
  }
}

// Rows longer than BU entries: a plain loop, one entry at a time per warp step of 32 prefetched indices.  Cold on
// meshes (hex: 7, tet: 5, polyhedral: ~15 -> this path); kept out of line so that its registers and code do not
// weigh on the straight-line path.
template <typename T, int VPL, bool kW, bool kSelf, bool kRS, bool kEpi>
__device__ __noinline__ void rows_long(const RowsArgs a, uint32_t i, int b, int e) {
  constexpr int VN = Vec<T>::N;
  const int lane = threadIdx.x & 31;
  const char* xb = reinterpret_cast<const char*>(a.x) + lane * 16;
  float acc[VPL][VN];
#pragma unroll
  for (int v = 0; v < VPL; ++v)
#pragma unroll
    for (int k = 0; k < VN; ++k) acc[v][k] = 0.f;
  const float rs = ((kW || kRS) && a.row_scale) ? __ldg(a.row_scale + i) : 1.0f;
  if (kSelf) {
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const uint4 u = ldg_row16(xb + (uint64_t)i * a.xrow_bytes + 512 * v);
      if (kW) fma_row16<T>(acc[v], a.self_coef, u);
      else add_row16(acc[v], u, T());
    }
  }
  for (int j = b; j < e; j += 32) {
    const int n = min(32, e - j);
    int cl = 0;
    float wl = 0.f;
    if (lane < n) {
      cl = __ldg(a.col + j + lane);
      if (kW) wl = (a.col_scale ? __ldg(a.col_scale + cl) : 1.0f) * rs;
    }
    int u = 0;
    for (; u + 4 <= n; u += 4) {                               // 4 rows in flight
      uint4 buf[4][VPL];
      float w[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const uint32_t c = (uint32_t)__shfl_sync(0xffffffffu, cl, u + t);
        if (kW) w[t] = __shfl_sync(0xffffffffu, wl, u + t);
#pragma unroll
        for (int v = 0; v < VPL; ++v) buf[t][v] = ldg_row16(xb + (uint64_t)c * a.xrow_bytes + 512 * v);
      }
#pragma unroll
      for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          if (kW) fma_row16<T>(acc[v], w[t], buf[t][v]);
          else add_row16(acc[v], buf[t][v], T());
        }
    }
    for (; u < n; ++u) {
      const uint32_t c = (uint32_t)__shfl_sync(0xffffffffu, cl, u);
      const float w = kW ? __shfl_sync(0xffffffffu, wl, u) : 1.0f;
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const uint4 t = ldg_row16(xb + (uint64_t)c * a.xrow_bytes + 512 * v);
        if (kW) fma_row16<T>(acc[v], w, t);
        else add_row16(acc[v], t, T());
      }
    }
  }
  if (kRS && !kW) {
#pragma unroll
    for (int v = 0; v < VPL; ++v)
#pragma unroll
      for (int k = 0; k < VN; ++k) acc[v][k] = __fmul_rn(acc[v][k], rs);
  }
  float bia[VPL][VN];
  if (kEpi) rows_load_bias<T, VPL>(a, lane, bia);
  rows_store<T, VPL, kEpi>(a, acc, i, lane, bia);
}

// K entries held one per lane in (cl, wl): K loads, then the K adds / FMAs.
// `wfn` turns the lane's loaded scale into its entry weight; it runs AFTER the K row loads have been issued: the issue
// stream is in order, and arithmetic on a just-requested col_scale value ahead of the gathers would hold them back for a
// full memory latency (measured on the weighted GCN-backward variant: 3.43 ms against 2.64 ms unweighted).
template <typename T, int VPL, bool kW, int K, typename WFn>
__device__ __forceinline__ void rows_batch(float (&acc)[VPL][Vec<T>::N], const char* xb, uint32_t xrow_bytes, int cl,
                                           WFn&& wfn) {
  uint4 buf[K][VPL];
#pragma unroll
  for (int u = 0; u < K; ++u) {
    const uint32_t c = (uint32_t)__shfl_sync(0xffffffffu, cl, u);
    const char* p = xb + (uint64_t)c * xrow_bytes;             // one IMAD.WIDE.U32
#pragma unroll
    for (int v = 0; v < VPL; ++v) buf[u][v] = ldg_row16(p + 512 * v);
  }
  const float wl = kW ? wfn() : 1.0f;
#pragma unroll
  for (int u = 0; u < K; ++u) {
    const float w = kW ? __shfl_sync(0xffffffffu, wl, u) : 1.0f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      if (kW) fma_row16<T>(acc[v], w, buf[u][v]);
      else add_row16(acc[v], buf[u][v], T());
    }
  }
}

template <typename T, int VPL, bool kW, bool kSelf, bool kRS, bool kEpi>
__global__ void __launch_bounds__(ROWS_THREADS, (rows_minb<VPL, kW>())) seg_rows_kernel(const RowsArgs a) {
  constexpr int VN = Vec<T>::N;
  constexpr int BU = RowsCfg<VPL>::BU;
  constexpr uint32_t END = 0xffffffffu;
  constexpr int NS = kSelf ? 1 : 0;
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  const char* xb = reinterpret_cast<const char*>(a.x) + lane * 16;
  // opaque to the optimiser: otherwise nvcc re-derives base + lane * 16 per neighbour (IMAD.WIDE + IADD3 + IADD3.X instead
  // of ONE IMAD.WIDE.U32 with the per-lane base as the 64-bit addend)
  asm volatile("" : "+l"(xb));
  const int32_t* __restrict__ rowptr = a.rowptr;
  const int32_t* __restrict__ col = a.col;

  float bia[VPL][VN];
  if (kEpi) rows_load_bias<T, VPL>(a, lane, bia);

  // this warp's rows: wi, wi + 8, ... of chunk q, then of chunk q + grid, ...
  uint32_t q = blockIdx.x, iend = 0;
  auto first_row_of_next_chunk = [&]() -> uint32_t {
    while (q < a.ord.n_chunks) {
      uint32_t rows;
      const uint32_t c0 = a.ord.chunk(q, a.n_rows, rows);
      q += gridDim.x;
      if ((uint32_t)wi < rows) {
        iend = c0 + rows;
        return c0 + wi;
      }
    }
    return END;
  };
  // entry `lane` of row i_: entry 0 is the row itself when kSelf, the others are col[b_ ...].  Branch-free (the index
  // is clamped into the row instead of predicated; lanes past the row's end hold a valid but unused index): a
  // conditional load becomes a BSSY/BSYNC region that ptxas schedules AFTER the adds waiting for this row's data.
  auto entry = [&](uint32_t i_, int b_, int e_) -> int {
    const int t = max(min(b_ + lane - NS, e_ - 1), 0);
    const int c = ldg_i32_ordered(col + t);
    return (kSelf && lane == 0) ? (int)i_ : c;
  };

  // Three rows in the pipe: row i is gathered and summed; the column indices of row i2 are requested with the
  // rowptr pair fetched one iteration earlier; the rowptr pair of row i3 is requested.  Every prefetch has its
  // address operands ready when it is issued, ahead of row i's gathers: nothing in the in-order issue stream waits
  // for an index chain.  END rows are clamped to a valid row so that the prefetch loads stay unconditional.
  uint32_t i = first_row_of_next_chunk();
  if (i == END) return;
  auto advance = [&](uint32_t i_) -> uint32_t {
    if (i_ == END) return END;
    const uint32_t n = i_ + (uint32_t)B2G_ROWS_WARPS;
    return n < iend ? n : first_row_of_next_chunk();
  };
  uint32_t i2 = advance(i);
  uint32_t i2c = min(i2, a.n_rows - 1u);
  int b = __ldg(rowptr + i), e = __ldg(rowptr + i + 1);
  int b2 = __ldg(rowptr + i2c), e2 = __ldg(rowptr + i2c + 1);
  int cl = entry(i, b, e);

  while (true) {
    const int cl2 = entry(i2c, b2, e2);
    const uint32_t i3 = advance(i2);
    const uint32_t i3c = min(i3, a.n_rows - 1u);
    const int b3 = __ldg(rowptr + i3c), e3 = __ldg(rowptr + i3c + 1);
    const int len = e - b + NS;
    if (len > BU) {
      rows_long<T, VPL, kW, kSelf, kRS, kEpi>(a, i, b, e);
    } else {
      float rs = 1.0f, csv = 1.0f;
      if ((kW || kRS) && a.row_scale) rs = __ldg(a.row_scale + i);
      if (kW && a.col_scale) csv = __ldg(a.col_scale + cl);        // cl is a valid (clamped) index on every lane
      auto wfn = [&]() -> float {
        const int t = lane - NS;
        return (t < 0) ? a.self_coef : ((t < e - b) ? csv * rs : 0.f);
      };
      float acc[VPL][VN];
#pragma unroll
      for (int v = 0; v < VPL; ++v)
#pragma unroll
        for (int k = 0; k < VN; ++k) acc[v][k] = 0.f;
      switch (len) {
#define B2G_CASE(KK) case KK: rows_batch<T, VPL, kW, KK>(acc, xb, a.xrow_bytes, cl, wfn); break;
        B2G_CASE(1) B2G_CASE(2) B2G_CASE(3) B2G_CASE(4)
        B2G_CASE(5) B2G_CASE(6) B2G_CASE(7) B2G_CASE(8)
#undef B2G_CASE
        default: break;
      }
      if (kRS && !kW) {
#pragma unroll
        for (int v = 0; v < VPL; ++v)
#pragma unroll
          for (int k = 0; k < VN; ++k) acc[v][k] = __fmul_rn(acc[v][k], rs);
      }
      rows_store<T, VPL, kEpi>(a, acc, i, lane, bia);
    }
    if (i2 == END) break;
    i = i2; b = b2; e = e2; cl = cl2;
    i2 = i3; i2c = i3c; b2 = b3; e2 = e3;
  }
}

template <typename T, int VPL, bool kW, bool kSelf, bool kRS, bool kEpi>
static int launch_rows_variant(const RowsArgs& a, cudaStream_t st) {
  int64_t blocks = a.ord.n_chunks;
  const int64_t cap = resident_ctas(seg_rows_kernel<T, VPL, kW, kSelf, kRS, kEpi>, ROWS_THREADS);
  if (blocks > cap) blocks = cap;
  seg_rows_kernel<T, VPL, kW, kSelf, kRS, kEpi><<<(unsigned)blocks, ROWS_THREADS, 0, st>>>(a);
  count_launch();
  return cuda_status();
}

template <typename T, int VPL>
static int launch_rows(const RowsArgs& a, cudaStream_t st) {
  const bool self = a.self_coef != 0.f;
  const bool weighted = a.col_scale != nullptr || (self && (a.self_coef != 1.f || a.row_scale != nullptr));
  const bool epi = a.bias != nullptr || a.relu != 0;
  const bool rs = a.row_scale != nullptr;
  if (!weighted) {
    if (!self && rs && epi) return launch_rows_variant<T, VPL, false, false, true, true>(a, st);     // GCN forward
    if (!self && rs && !epi) return launch_rows_variant<T, VPL, false, false, true, false>(a, st);
    if (!self && !rs && !epi) return launch_rows_variant<T, VPL, false, false, false, false>(a, st);  // plain sum
    if (self && !rs && !epi) return launch_rows_variant<T, VPL, false, true, false, false>(a, st);    // GIN (eps = 0) fwd / bwd
    return B2G_E_UNSUPPORTED;
  }
  if (!epi) {
    // (a one-"head" mma.sync variant of this sum was measured at 5.44 ms against 3.17 ms here: 32 m16n8k8 HMMAs per row with
    // 1/16 of the tile used make the legacy tensor path the bound — DESIGN §4 point 11)
    if (!self) return launch_rows_variant<T, VPL, true, false, false, false>(a, st);                  // GCN backward
    return launch_rows_variant<T, VPL, true, true, false, false>(a, st);                              // GIN with eps != 0
  }
  return B2G_E_UNSUPPORTED;
}

// B2G_E_UNSUPPORTED = not a case of this fast path (the caller falls back to the generic kernel).
int rows_seg_sum(const void* x, int64_t ldx, void* out, int64_t ldo, int64_t n_rows, int nvec, int dt,
                 const int32_t* rowptr, const int32_t* col, const float* row_scale, const float* col_scale,
                 float self_coef, const float* bias, int relu, int64_t band, int64_t /*max_row_len*/, int chunk_rows,
                 int panel_rows, cudaStream_t st) {
  const int es = dt == B2G_F32 ? 4 : 2;
  if (nvec != 32 && nvec != 64) return B2G_E_UNSUPPORTED;
  if (n_rows < 1024 || n_rows >= (1ll << 32) - (1ll << 25) || ldx * es >= (1ll << 32) || ldo * es >= (1ll << 32))
    return B2G_E_UNSUPPORTED;
  RowsArgs a{};
  if (!make_row_sched(n_rows, band, a.ord, chunk_rows, panel_rows)) return B2G_E_ARG;
  a.x = x; a.out = out; a.rowptr = rowptr; a.col = col; a.row_scale = row_scale; a.col_scale = col_scale; a.bias = bias;
  a.xrow_bytes = (uint32_t)(ldx * es);
  a.orow_bytes = (uint32_t)(ldo * es);
  a.n_rows = (uint32_t)n_rows;
  a.self_coef = self_coef;
  a.relu = relu;
  if (dt == B2G_F32) return nvec == 32 ? launch_rows<float, 1>(a, st) : launch_rows<float, 2>(a, st);
  return nvec == 32 ? launch_rows<__nv_bfloat16, 1>(a, st) : launch_rows<__nv_bfloat16, 2>(a, st);
}

}  // namespace b2g
