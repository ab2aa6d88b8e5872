// gat_rows.cu — K4 "aggregate first" path of GATConv(heads = 4, concat = False) (gnn_model.py:65-68,168;
// SURVEY §8a rows 5, 8) for feature rows of 512 / 1024 bytes (F = 256 bf16, F = 128 / 256 fp32).
//
// PyG computes  out_i = 1/H sum_h sum_j alpha_ijh (W_h x_j) + b  by projecting first (an [N, H*C] matrix) and
// gathering H*C-wide rows per edge (2 KB per edge at H = 4, C = 256, bf16: 143 GB through L2 at cfg4, measured
// 24 ms, 17 % of the HBM roofline).  The sum over j is linear, so
//      out_i = 1/H sum_h W_h (sum_j alpha_ijh x_j) + b = z_i Wc^T + b,   z_i = [sum_j alpha_ij1 x_j | ... | sum_j alpha_ijH x_j]
// and the scores only need  a_src[j,h] = x_j . (W_h^T att_src_h),  a_dst[i,h] = x_i . (W_h^T att_dst_h):
//   b2g_rowdot        a[N, 2H] = x V^T                          (one pass over x, 8 dot products per row)
//   b2g_gatz_fwd      z[N, H*F]: per target row the exact max-subtracted segment softmax (one entry per lane) and
//                     ONE gather of the F-wide neighbour rows shared by all heads (4x fewer gathered bytes)
//   K6 GEMM           out = z Wc^T + b                           (gemm_tc.cu, k = H*F)
// Backward, by the same linearity (g = d out):  dz = g Wc (GEMM);
//   b2g_gatz_bwd_dst  d alpha_ijh = dz_ih . x_j (x rows gathered once per edge), softmax / LeakyReLU backward,
//                     per-edge alpha and d e in target-major order, d a_dst
//   b2g_gatz_bwd_src  y_j = [sum_i alpha_ij1 g_i | ...] over the transposed CSR (F-wide rows of g), d a_src
//   K6 GEMM           dx = [y | d a] [W/H ; V]                   (one GEMM, k = H*C + 2H)
// Deterministic: fp32 accumulation in CSR (= edge) order; attention dropout is the same counter-based Philox stream as
// attention.cu (keyed by the target-major edge position), regenerated in backward.
#include <cstdlib>
#include "rows.cuh"

namespace b2g {

constexpr uint32_t ROW_END = 0xffffffffu;
constexpr int GH = 4;                         // heads (the reference's GATConv: heads = 4)

// This warp's rows (wi, wi + 8, ... of chunk q, then of chunk q + grid, ...) with the rowptr pair fetched two rows ahead
// (see aggregate_rows.cu: every prefetch is issued with its address operands long arrived).
struct WarpRows {
  uint32_t q, iend, n_rows;
  uint32_t i, i2, i2c, i3, i3c;
  int b, e, b2, e2, b3, e3;
  __device__ __forceinline__ uint32_t next_chunk_first(const RowSched& ord, int wi) {
    while (q < ord.n_chunks) {
      uint32_t rows;
      const uint32_t c0 = ord.chunk(q, n_rows, rows);
      q += gridDim.x;
      if ((uint32_t)wi < rows) {
        iend = c0 + rows;
        return c0 + wi;
      }
    }
    return ROW_END;
  }
  __device__ __forceinline__ uint32_t advance(uint32_t i_, const RowSched& ord, int wi) {
    if (i_ == ROW_END) return ROW_END;
    const uint32_t n = i_ + 8u;
    return n < iend ? n : next_chunk_first(ord, wi);
  }
  __device__ __forceinline__ bool begin(const RowSched& ord, uint32_t n, int wi, const int32_t* __restrict__ rowptr) {
    q = blockIdx.x; iend = 0; n_rows = n;
    i = next_chunk_first(ord, wi);
    if (i == ROW_END) return false;
    i2 = advance(i, ord, wi);
    i2c = min(i2, n_rows - 1u);
    b = __ldg(rowptr + i); e = __ldg(rowptr + i + 1);
    b2 = __ldg(rowptr + i2c); e2 = __ldg(rowptr + i2c + 1);
    return true;
  }
  __device__ __forceinline__ void look_ahead(const RowSched& ord, int wi, const int32_t* __restrict__ rowptr) {
    i3 = advance(i2, ord, wi);
    i3c = min(i3, n_rows - 1u);
    b3 = __ldg(rowptr + i3c); e3 = __ldg(rowptr + i3c + 1);
  }
  __device__ __forceinline__ bool shift() {
    if (i2 == ROW_END) return false;
    i = i2; b = b2; e = e2;
    i2 = i3; i2c = i3c; b2 = b3; e2 = e3;
    return true;
  }
};

// entry `lane` of the index window starting at position p0 of a row ending at e_: clamped into the row (lanes past the
// end repeat the last entry: a valid address whose weight is 0), branch-free so that the load stays where it is written
__device__ __forceinline__ int window_entry(const int32_t* __restrict__ idx, int p0, int e_, int lane) {
  return ldg_i32_ordered(idx + max(min(p0 + lane, e_ - 1), 0));
}
__device__ __forceinline__ float4 ldg_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float lrelu(float s, float slope) { return s > 0.f ? s : s * slope; }
// max over the warp through the integer reduction unit (REDUX): monotone float <-> int map, one instruction per head
__device__ __forceinline__ float warp_max_redux(float v) {
  int k = __float_as_int(v);
  k ^= (k >> 31) & 0x7fffffff;
  k = __reduce_max_sync(0xffffffffu, k);
  k ^= (k >> 31) & 0x7fffffff;
  return __int_as_float(k);
}
__device__ __forceinline__ float pick4(const float (&v)[4], int h) {   // v[h] for a per-lane h without local memory
  return h == 0 ? v[0] : (h == 1 ? v[1] : (h == 2 ? v[2] : v[3]));
}

template <typename T>
__device__ __forceinline__ void ld16_rows(const char* p, float (&f)[Vec<T>::N]) {
  uint4 u;
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "l"(p));
  unpack_row16(u, f, T());
}

struct GatzArgs {
  const void* x; uint32_t xrow_bytes;        // gathered rows [*, F] (fwd / bwd_dst: x; bwd_src: g = d out)
  const void* x_self;                        // tz_fwd: the target rows' own features (row i of THIS call); == x unless the call
                                             // covers a row range of a larger problem (streaming.py)
  const float* a; uint32_t lda;              // fp32 [N, >= 2H]: columns 0..H-1 a_src, H..2H-1 a_dst (stride in floats)
  void* z; uint32_t zrow_bytes;              // fwd: z [N, H*F];  bwd_src: y [N, H*C]
  const void* dz; uint32_t dzrow_bytes;      // bwd_dst: dz [N, H*F]
  const int32_t* rowptr; const int32_t* col; const int32_t* perm;
  float* smax; float* ssum;                  // [N, H] softmax statistics (ssum includes PyG's + 1e-16)
  float* alpha_e; float* de_e;               // [nnz, H] target-major
  const float* alpha_in;                     // TransformerConv backward: the forward pass's (pre-dropout) alpha [nnz, H]
  const float* ebias;                        // TransformerConv with edge features (edge_dim): per-(entry, head) term added to
                                             // the logits (forward) / to d alpha (backward), fp32 [nnz, H] target-major; or NULL
  float* d_a; uint32_t ldda;                 // [N, >= 2H] (row stride in ELEMENTS of its type)
  int alpha_only;                            // tz_fwd: attention weights only (fused TransformerConv forward): alpha_e = pre-dropout
                                             // weights, de_e = post-dropout weights (or NULL when p_drop == 0), smax = weight sums
  int da_bf16;                               // d_a holds bf16 (a column block of the bf16 dgrad operand) instead of fp32
  int only_long;                             // bwd_dst: rows of <= 8 entries were done by gatz_bwd_dst_mma_kernel: skip them
  uint32_t n_rows;
  float slope, p_drop;
  uint64_t seed;
  const uint64_t* epoch;                     // dropout epoch word (common.cuh), mixed into seed
  RowSched ord;
};

// d a[i, k] = v, in the buffer's type (fp32, or bf16 when the caller hands the column block of a bf16 GEMM operand)
__device__ __forceinline__ void store_da(const GatzArgs& a, uint64_t i, int k, float v) {
  if (a.da_bf16) reinterpret_cast<__nv_bfloat16*>(a.d_a)[i * a.ldda + k] = __float2bfloat16_rn(v);
  else a.d_a[i * a.ldda + k] = v;
}

#ifndef B2G_GATZ_BU
#define B2G_GATZ_BU 8
#endif
#ifndef B2G_GATZ_MINB
#define B2G_GATZ_MINB 2
#endif
template <int VPL> struct GatzCfg { static constexpr int BU = (VPL == 1) ? B2G_GATZ_BU : 4; };   // gathered rows in flight per warp

// acc[h] += w[h](entry off+u) * row(entry off+u) for u < BU (default: a full batch; entries past the row's end carry
// weight 0 and re-read the row's last entry)
template <typename T, int VPL, int BU = GatzCfg<VPL>::BU, typename Mid>
__device__ __forceinline__ void gatz_gather_fma(float (&acc)[GH][VPL][Vec<T>::N], const char* xb, uint32_t xrow_bytes,
                                                int cl, const float (&w)[GH], int off, Mid&& mid) {
  constexpr int VN = Vec<T>::N;
  uint4 buf[BU][VPL];
#pragma unroll
  for (int u = 0; u < BU; ++u) {
    const uint32_t c = (uint32_t)__shfl_sync(0xffffffffu, cl, off + u);
    const char* p = xb + (uint64_t)c * xrow_bytes;
#pragma unroll
    for (int v = 0; v < VPL; ++v) buf[u][v] = ldg_row16(p + 512 * v);
  }
  mid();                                     // the softmax of this row: needs a_src only, runs while the rows are in flight
#pragma unroll
  for (int u = 0; u < BU; ++u) {
    float wu[GH];
#pragma unroll
    for (int h = 0; h < GH; ++h) wu[h] = __shfl_sync(0xffffffffu, w[h], off + u);
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      float f[VN];
      unpack_row16(buf[u][v], f, T());
#pragma unroll
      for (int h = 0; h < GH; ++h)
#pragma unroll
        for (int k = 0; k < VN; k += 2) ffma2_acc(acc[h][v][k], acc[h][v][k + 1], wu[h], f[k], f[k + 1]);
    }
  }
}

// ---- rows of <= 8 entries: ONE (entry, head) pair per lane (lane = 4 * entry + head).  The H segment softmaxes of the
// row are then 3 + 3 xor-shuffles and one exp per lane instead of 4 x (reduce, exp, reduce) on entry-per-lane data
// (~35 instead of ~110 instructions per row), per-edge [nnz, H] arrays are read / written fully coalesced, and the
// transposed reduction of the logit / d-alpha dots (warp_transpose_sum32) already delivers this layout.
__device__ __forceinline__ float head_max8(float v) {          // max over the 8 lanes that share lane & 3
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 4));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 8));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 16));
}
__device__ __forceinline__ float head_sum8(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  return v + __shfl_xor_sync(0xffffffffu, v, 16);
}
__device__ __forceinline__ float packed_keep_scale(uint64_t seed, uint64_t pos, float p_drop, int h) {
  float sc[4];
  dropout_scale4(seed, pos, p_drop, sc);                      // the same Philox draw as the entry-per-lane kernels
  return pick4(sc, h);
}
// acc[h] += wp(entry u, head h) * buf[u] for the K gathered rows of a <= 8-entry row; wp is the packed weight
template <typename T, int VPL, int K>
__device__ __forceinline__ void packed_fma(float (&acc)[GH][VPL][Vec<T>::N], const uint4 (&buf)[8][VPL], float wp) {
  constexpr int VN = Vec<T>::N;
#pragma unroll
  for (int u = 0; u < K; ++u) {
    float wu[GH];
#pragma unroll
    for (int h = 0; h < GH; ++h) wu[h] = __shfl_sync(0xffffffffu, wp, 4 * u + h);
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      float f[VN];
      unpack_row16(buf[u][v], f, T());
#pragma unroll
      for (int h = 0; h < GH; ++h)
#pragma unroll
        for (int k = 0; k < VN; k += 2) ffma2_acc(acc[h][v][k], acc[h][v][k + 1], wu[h], f[k], f[k + 1]);
    }
  }
}

template <typename T, int VPL>
__device__ __forceinline__ void gatz_store(char* zrow, const float (&acc)[GH][VPL][Vec<T>::N], int lane) {
#pragma unroll
  for (int h = 0; h < GH; ++h)
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      Vec<T> o;
      o.from_float(acc[h][v]);
      __stcs(reinterpret_cast<uint4*>(zrow + (h * VPL + v) * 512 + lane * 16), *reinterpret_cast<uint4*>(&o.v));
    }
}

// ------------------------------------------------------------------------------------------ forward
// Rows longer than 32 entries (cold on meshes): three sweeps over the scores (max, sum, weights + gather).
template <typename T, int VPL>
__device__ __noinline__ void gatz_fwd_long(const GatzArgs a, uint32_t i, int b, int e) {
  constexpr int VN = Vec<T>::N;
  const int lane = threadIdx.x & 31;
  const char* xb = reinterpret_cast<const char*>(a.x) + lane * 16;
  asm volatile("" : "+l"(xb));     // keep the per-lane base in a register pair: one IMAD.WIDE.U32 per gathered row
  const float4 ad4 = ldg_f4(a.a + (uint64_t)i * a.lda + GH);
  const float ad[GH] = {ad4.x, ad4.y, ad4.z, ad4.w};
  float m[GH], zs[GH];
#pragma unroll
  for (int h = 0; h < GH; ++h) { m[h] = -INFINITY; zs[h] = 0.f; }
  for (int p0 = b; p0 < e; p0 += 32) {
    const int c = window_entry(a.col, p0, e, lane);
    const float4 as4 = ldg_f4(a.a + (uint64_t)(uint32_t)c * a.lda);
    const float as[GH] = {as4.x, as4.y, as4.z, as4.w};
#pragma unroll
    for (int h = 0; h < GH; ++h) m[h] = fmaxf(m[h], warp_max_redux(p0 + lane < e ? lrelu(as[h] + ad[h], a.slope) : -INFINITY));
  }
  for (int p0 = b; p0 < e; p0 += 32) {
    const int c = window_entry(a.col, p0, e, lane);
    const float4 as4 = ldg_f4(a.a + (uint64_t)(uint32_t)c * a.lda);
    const float as[GH] = {as4.x, as4.y, as4.z, as4.w};
    float p[GH];
#pragma unroll
    for (int h = 0; h < GH; ++h) p[h] = p0 + lane < e ? __expf(lrelu(as[h] + ad[h], a.slope) - m[h]) : 0.f;
    warp_sum4(p[0], p[1], p[2], p[3]);
#pragma unroll
    for (int h = 0; h < GH; ++h) zs[h] += p[h];
  }
  float inv[GH];
#pragma unroll
  for (int h = 0; h < GH; ++h) { zs[h] += 1e-16f; inv[h] = 1.0f / zs[h]; }
  float acc[GH][VPL][VN];
#pragma unroll
  for (int h = 0; h < GH; ++h)
#pragma unroll
    for (int v = 0; v < VPL; ++v)
#pragma unroll
      for (int k = 0; k < VN; ++k) acc[h][v][k] = 0.f;
  for (int p0 = b; p0 < e; p0 += 32) {
    const int c = window_entry(a.col, p0, e, lane);
    const float4 as4 = ldg_f4(a.a + (uint64_t)(uint32_t)c * a.lda);
    const float as[GH] = {as4.x, as4.y, as4.z, as4.w};
    float w[GH];
#pragma unroll
    for (int h = 0; h < GH; ++h) w[h] = p0 + lane < e ? __expf(lrelu(as[h] + ad[h], a.slope) - m[h]) * inv[h] : 0.f;
    if (a.p_drop > 0.f) {
      float sc[4];
      dropout_scale4(mix_epoch(a.seed, a.epoch), (uint64_t)(p0 + lane), a.p_drop, sc);
#pragma unroll
      for (int h = 0; h < GH; ++h) w[h] *= sc[h];
    }
    const int n = min(32, e - p0);
    for (int j = 0; j < n; j += GatzCfg<VPL>::BU) gatz_gather_fma<T, VPL>(acc, xb, a.xrow_bytes, c, w, j, []() {});
  }
  if (a.smax && lane < GH) {
    a.smax[(uint64_t)i * GH + lane] = pick4(m, lane);
    a.ssum[(uint64_t)i * GH + lane] = pick4(zs, lane);
  }
  gatz_store<T, VPL>(reinterpret_cast<char*>(a.z) + (uint64_t)i * a.zrow_bytes, acc, lane);
}

template <typename T, int VPL>
__global__ void __launch_bounds__(256, B2G_GATZ_MINB) gatz_fwd_kernel(const GatzArgs a) {
  constexpr int VN = Vec<T>::N;
  constexpr int BU = GatzCfg<VPL>::BU;
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  const char* xb = reinterpret_cast<const char*>(a.x) + lane * 16;
  asm volatile("" : "+l"(xb));     // keep the per-lane base in a register pair: one IMAD.WIDE.U32 per gathered row
  WarpRows r;
  if (!r.begin(a.ord, a.n_rows, wi, a.rowptr)) return;
  int cl = window_entry(a.col, r.b, r.e, lane);
  while (true) {
    const int cl2 = window_entry(a.col, r.b2, r.e2, lane);
    r.look_ahead(a.ord, wi, a.rowptr);
    const int len = r.e - r.b;
    if (len > 32) {
      gatz_fwd_long<T, VPL>(a, r.i, r.b, r.e);
    } else if (VPL == 1 && len > 0 && len <= 8) {
      // every mesh row: packed (entry, head) softmax, one gather, exact-length FMA
      const int u = lane >> 2, h = lane & 3;
      const uint32_t cu = (uint32_t)__shfl_sync(0xffffffffu, cl, u);
      float as, ad;
      asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(as) : "l"(a.a + (uint64_t)cu * a.lda + h));
      asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(ad) : "l"(a.a + (uint64_t)r.i * a.lda + GH + h));
      uint4 buf[8][VPL];
#pragma unroll
      for (int e8 = 0; e8 < 8; ++e8) {                         // entries past the row's end re-read its last row (L1 hit)
        const uint32_t c = (uint32_t)__shfl_sync(0xffffffffu, cl, e8);
        const char* p = xb + (uint64_t)c * a.xrow_bytes;
#pragma unroll
        for (int v = 0; v < VPL; ++v) buf[e8][v] = ldg_row16(p + 512 * v);
      }
      const float sc = u < len ? lrelu(as + ad, a.slope) : -INFINITY;
      const float m = head_max8(sc);
      float wp = u < len ? __expf(sc - m) : 0.f;
      const float zs = head_sum8(wp) + 1e-16f;
      wp *= 1.0f / zs;
      if (a.p_drop > 0.f) wp *= packed_keep_scale(mix_epoch(a.seed, a.epoch), (uint64_t)(r.b + u), a.p_drop, h);
      if (a.smax && lane < GH) {                               // lanes 0..3 = entry 0, head = lane
        a.smax[(uint64_t)r.i * GH + lane] = m;
        a.ssum[(uint64_t)r.i * GH + lane] = zs;
      }
      float acc[GH][VPL][VN];
#pragma unroll
      for (int hh = 0; hh < GH; ++hh)
#pragma unroll
        for (int v = 0; v < VPL; ++v)
#pragma unroll
          for (int k = 0; k < VN; ++k) acc[hh][v][k] = 0.f;
      switch (len) {
#define B2G_CASE(KK) case KK: packed_fma<T, VPL, KK>(acc, buf, wp); break;
        B2G_CASE(1) B2G_CASE(2) B2G_CASE(3) B2G_CASE(4) B2G_CASE(5) B2G_CASE(6) B2G_CASE(7) B2G_CASE(8)
#undef B2G_CASE
        default: break;
      }
      gatz_store<T, VPL>(reinterpret_cast<char*>(a.z) + (uint64_t)r.i * a.zrow_bytes, acc, lane);
    } else {
      const float4 as4 = ldg_f4(a.a + (uint64_t)(uint32_t)cl * a.lda);          // a_src of this lane's entry
      const float4 ad4 = ldg_f4(a.a + (uint64_t)r.i * a.lda + GH);               // a_dst of the row (uniform)
      float acc[GH][VPL][VN];
#pragma unroll
      for (int h = 0; h < GH; ++h)
#pragma unroll
        for (int v = 0; v < VPL; ++v)
#pragma unroll
          for (int k = 0; k < VN; ++k) acc[h][v][k] = 0.f;
      float w[GH], m[GH], zs[GH];
      auto softmax = [&]() {                  // exact two-pass softmax: all scores of the row sit one per lane
        const float as[GH] = {as4.x, as4.y, as4.z, as4.w};
        const float ad[GH] = {ad4.x, ad4.y, ad4.z, ad4.w};
#pragma unroll
        for (int h = 0; h < GH; ++h) {
          const float s = lane < len ? lrelu(as[h] + ad[h], a.slope) : -INFINITY;
          m[h] = warp_max_redux(s);
          w[h] = lane < len ? __expf(s - m[h]) : 0.f;
          zs[h] = w[h];
        }
        warp_sum4(zs[0], zs[1], zs[2], zs[3]);
#pragma unroll
        for (int h = 0; h < GH; ++h) {
          zs[h] += 1e-16f;
          w[h] *= 1.0f / zs[h];
        }
        if (a.p_drop > 0.f) {
          float sc[4];
          dropout_scale4(mix_epoch(a.seed, a.epoch), (uint64_t)(r.b + lane), a.p_drop, sc);
#pragma unroll
          for (int h = 0; h < GH; ++h) w[h] *= sc[h];
        }
      };
      if (len > BU) {
        gatz_gather_fma<T, VPL>(acc, xb, a.xrow_bytes, cl, w, 0, softmax);
        for (int j = BU; j < len; j += BU) gatz_gather_fma<T, VPL>(acc, xb, a.xrow_bytes, cl, w, j, []() {});
      } else if (len > 0) {                   // every mesh row at VPL = 1: straight-line code for exactly `len` entries
        switch (len) {
#define B2G_CASE(KK) case KK: if (KK <= BU) gatz_gather_fma<T, VPL, (KK <= BU ? KK : 1)>(acc, xb, a.xrow_bytes, cl, w, 0, softmax); break;
          B2G_CASE(1) B2G_CASE(2) B2G_CASE(3) B2G_CASE(4) B2G_CASE(5) B2G_CASE(6) B2G_CASE(7) B2G_CASE(8)
#undef B2G_CASE
          default: break;
        }
      } else {
#pragma unroll
        for (int h = 0; h < GH; ++h) { m[h] = 0.f; zs[h] = 1e-16f; }
      }
      if (a.smax && lane < GH) {
        a.smax[(uint64_t)r.i * GH + lane] = len > 0 ? pick4(m, lane) : 0.f;
        a.ssum[(uint64_t)r.i * GH + lane] = pick4(zs, lane);
      }
      gatz_store<T, VPL>(reinterpret_cast<char*>(a.z) + (uint64_t)r.i * a.zrow_bytes, acc, lane);
    }
    if (!r.shift()) break;
    cl = cl2;
  }
}

// ------------------------------------------------------------------------------------------ backward, target side
// d alpha for the entries [off, off + 8) of the window held one per lane: gather 8 rows (2 x 4 when BU = 4), dot each
// with the row's dz (4 heads), reduce the 32 partials across the warp with ONE transposed reduction (31 shuffles),
// hand entry (off + u)'s four values to lane off + u.
// `mid` runs once, after the first rows have been requested: whatever has to be done to the row's own just-loaded data
// (unpacking dz / u) must not sit in the in-order issue stream ahead of the gathers.
template <typename T, int VPL, typename Mid>
__device__ __forceinline__ void gatz_dalpha8(float (&dal)[GH], const float (&dzf)[GH][VPL][Vec<T>::N], const char* xb,
                                             uint32_t xrow_bytes, int cl, int off, int lane, Mid&& mid) {
  constexpr int VN = Vec<T>::N;
  constexpr int BU = GatzCfg<VPL>::BU;
  float part[32];
#pragma unroll
  for (int s0 = 0; s0 < 8; s0 += BU) {
    uint4 buf[BU][VPL];
#pragma unroll
    for (int u = 0; u < BU; ++u) {
      const uint32_t c = (uint32_t)__shfl_sync(0xffffffffu, cl, off + s0 + u);
      const char* p = xb + (uint64_t)c * xrow_bytes;
#pragma unroll
      for (int v = 0; v < VPL; ++v) buf[u][v] = ldg_row16(p + 512 * v);
    }
    if (s0 == 0) mid();
#pragma unroll
    for (int u = 0; u < BU; ++u) {
      float p0[GH], p1[GH];
#pragma unroll
      for (int h = 0; h < GH; ++h) { p0[h] = 0.f; p1[h] = 0.f; }
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        float f[VN];
        unpack_row16(buf[u][v], f, T());
#pragma unroll
        for (int h = 0; h < GH; ++h)
#pragma unroll
          for (int k = 0; k < VN; k += 2) ffma2_mul(p0[h], p1[h], dzf[h][v][k], dzf[h][v][k + 1], f[k], f[k + 1]);
      }
#pragma unroll
      for (int h = 0; h < GH; ++h) part[(s0 + u) * GH + h] = p0[h] + p1[h];
    }
  }
  const float red = warp_transpose_sum32(part);          // lane 4u + h: d alpha of entry off + u, head h
  const int src = ((lane - off) & 7) * GH;
#pragma unroll
  for (int h = 0; h < GH; ++h) {
    const float v = __shfl_sync(0xffffffffu, red, src + h);
    if (lane >= off && lane < off + 8) dal[h] = v;
  }
}

// One window of <= 32 entries [p0, p0 + n): alpha from the saved statistics, d alpha, and - given t = sum_k alpha_k
// d alpha_k over the WHOLE row - d e; writes alpha_e / de_e, returns this window's contribution to d a_dst.
// kT = false: GATConv (alpha recomputed from a_src / a_dst and the saved statistics; LeakyReLU in front of the softmax).
// kT = true : TransformerConv (alpha read back from the forward pass; `ad` carries d s_alpha, the gradient of the
//             per-head weight sums that multiply the value bias; the logits enter the softmax directly: sraw = 1).
template <typename T, int VPL, bool kT, typename First>
__device__ __forceinline__ void gatz_bwd_window(const GatzArgs& a, const float (&dzf)[GH][VPL][Vec<T>::N], const char* xb,
                                                int cl, int p0, int n, int lane, const float (&ad)[GH],
                                                const float (&sm)[GH], const float (&rinv)[GH], float (&alpha)[GH],
                                                float (&dal)[GH], float (&sraw)[GH], float (&mask)[GH], First&& first) {
  // requests first (nothing below may consume them before the gathers of gatz_dalpha8 are in flight)
  const float4 in4 = kT ? ldg_f4(a.alpha_in + (uint64_t)max(min(p0 + lane, p0 + n - 1), 0) * GH)
                        : ldg_f4(a.a + (uint64_t)(uint32_t)cl * a.lda);
#pragma unroll
  for (int h = 0; h < GH; ++h) dal[h] = 0.f;
  gatz_dalpha8<T, VPL>(dal, dzf, xb, a.xrow_bytes, cl, 0, lane, first);
  for (int j = 8; j < n; j += 8) gatz_dalpha8<T, VPL>(dal, dzf, xb, a.xrow_bytes, cl, j, lane, []() {});
  const float in[GH] = {in4.x, in4.y, in4.z, in4.w};
  float eb[GH] = {0.f, 0.f, 0.f, 0.f};          // GATConv(edge_dim): edge term of the logit, in front of the LeakyReLU
  if (!kT && a.ebias && lane < n) {
    const float4 b4 = ldg_f4(a.ebias + (uint64_t)(p0 + lane) * GH);
    eb[0] = b4.x; eb[1] = b4.y; eb[2] = b4.z; eb[3] = b4.w;
  }
#pragma unroll
  for (int h = 0; h < GH; ++h) {
    if (kT) {
      sraw[h] = 1.0f;
      alpha[h] = lane < n ? in[h] : 0.f;
    } else {
      sraw[h] = in[h] + ad[h] + eb[h];
      alpha[h] = lane < n ? __expf(lrelu(sraw[h], a.slope) - sm[h]) * rinv[h] : 0.f;
    }
    mask[h] = 1.0f;
  }
  if (a.p_drop > 0.f) dropout_scale4(mix_epoch(a.seed, a.epoch), (uint64_t)(p0 + lane), a.p_drop, mask);
#pragma unroll
  for (int h = 0; h < GH; ++h) dal[h] = (dal[h] + (kT ? ad[h] : 0.f)) * mask[h];   // d(alpha) of the pre-dropout probability
  if (kT && a.ebias && lane < n) {            // edge-feature term of d alpha' (dm_ih . a_ij), same dropout mask
    const float4 b4 = ldg_f4(a.ebias + (uint64_t)(p0 + lane) * GH);
    dal[0] += b4.x * mask[0]; dal[1] += b4.y * mask[1]; dal[2] += b4.z * mask[2]; dal[3] += b4.w * mask[3];
  }
}

// the row's per-head side inputs: GAT (a_dst, softmax max, 1 / softmax sum) or TransformerConv (d s_alpha from dz_aug)
template <typename T, bool kT>
__device__ __forceinline__ void gatz_bwd_row_inputs(const GatzArgs& a, uint32_t i, int hf_bytes, float (&ad)[GH],
                                                    float (&sm)[GH], float (&rinv)[GH]) {
  if (kT) {
    float f[Vec<T>::N];
    ld16_rows<T>(reinterpret_cast<const char*>(a.dz) + (uint64_t)i * a.dzrow_bytes + hf_bytes, f);
#pragma unroll
    for (int h = 0; h < GH; ++h) { ad[h] = f[h]; sm[h] = 0.f; rinv[h] = 1.f; }
  } else {
    const float4 ad4 = ldg_f4(a.a + (uint64_t)i * a.lda + GH);
    const float4 sm4 = ldg_f4(a.smax + (uint64_t)i * GH), ss4 = ldg_f4(a.ssum + (uint64_t)i * GH);
    ad[0] = ad4.x; ad[1] = ad4.y; ad[2] = ad4.z; ad[3] = ad4.w;
    sm[0] = sm4.x; sm[1] = sm4.y; sm[2] = sm4.z; sm[3] = sm4.w;
    rinv[0] = 1.0f / ss4.x; rinv[1] = 1.0f / ss4.y; rinv[2] = 1.0f / ss4.z; rinv[3] = 1.0f / ss4.w;
  }
}

// the row's own H*F-wide vector (dz_i, or u_i in the TransformerConv forward): requested raw, unpacked later
template <int VPL>
__device__ __forceinline__ void gatz_load_dz_raw(const GatzArgs& a, uint32_t i, int lane, uint4 (&raw)[GH][VPL]) {
  const char* dzr = reinterpret_cast<const char*>(a.dz) + (uint64_t)i * a.dzrow_bytes + lane * 16;
#pragma unroll
  for (int h = 0; h < GH; ++h)
#pragma unroll
    for (int v = 0; v < VPL; ++v)
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(raw[h][v].x), "=r"(raw[h][v].y), "=r"(raw[h][v].z), "=r"(raw[h][v].w)
                   : "l"(dzr + (h * VPL + v) * 512));
}
template <typename T, int VPL>
__device__ __forceinline__ void gatz_unpack_dz(const uint4 (&raw)[GH][VPL], float (&dzf)[GH][VPL][Vec<T>::N]) {
#pragma unroll
  for (int h = 0; h < GH; ++h)
#pragma unroll
    for (int v = 0; v < VPL; ++v) unpack_row16(raw[h][v], dzf[h][v], T());
}
template <typename T, int VPL>
__device__ __forceinline__ void gatz_load_dz(const GatzArgs& a, uint32_t i, int lane, float (&dzf)[GH][VPL][Vec<T>::N]) {
  uint4 raw[GH][VPL];
  gatz_load_dz_raw<VPL>(a, i, lane, raw);
  gatz_unpack_dz<T, VPL>(raw, dzf);
}

template <typename T, int VPL>
__device__ __forceinline__ void gatz_bwd_finish(const GatzArgs& a, int p0, int n, int lane, const float (&alpha)[GH],
                                                const float (&dal)[GH], const float (&sraw)[GH], const float (&mask)[GH],
                                                const float (&t)[GH], float (&dad)[GH], float (&de)[GH]) {
#pragma unroll
  for (int h = 0; h < GH; ++h) {
    const float ds = alpha[h] * (dal[h] - t[h]);
    de[h] = ds * (sraw[h] > 0.f ? 1.0f : a.slope);
    dad[h] += de[h];
  }
  if (lane < n) {
    *reinterpret_cast<float4*>(a.alpha_e + (uint64_t)(p0 + lane) * GH) =
        make_float4(alpha[0] * mask[0], alpha[1] * mask[1], alpha[2] * mask[2], alpha[3] * mask[3]);
    *reinterpret_cast<float4*>(a.de_e + (uint64_t)(p0 + lane) * GH) = make_float4(de[0], de[1], de[2], de[3]);
  }
}

template <typename T, int VPL, bool kT>
__device__ __noinline__ void gatz_bwd_dst_long(const GatzArgs a, uint32_t i, int b, int e) {
  constexpr int VN = Vec<T>::N;
  const int lane = threadIdx.x & 31;
  const char* xb = reinterpret_cast<const char*>(a.x) + lane * 16;
  asm volatile("" : "+l"(xb));     // keep the per-lane base in a register pair: one IMAD.WIDE.U32 per gathered row
  float dzf[GH][VPL][VN];
  gatz_load_dz<T, VPL>(a, i, lane, dzf);
  float ad[GH], sm[GH], rinv[GH];
  gatz_bwd_row_inputs<T, kT>(a, i, GH * VPL * 512, ad, sm, rinv);
  float t[GH] = {0.f, 0.f, 0.f, 0.f};
  // sweep 1: d alpha of every entry (parked in de_e), t = sum alpha * d alpha
  for (int p0 = b; p0 < e; p0 += 32) {
    const int n = min(32, e - p0);
    const int cl = window_entry(a.col, p0, e, lane);
    float alpha[GH], dal[GH], sraw[GH], mask[GH];
    gatz_bwd_window<T, VPL, kT>(a, dzf, xb, cl, p0, n, lane, ad, sm, rinv, alpha, dal, sraw, mask, []() {});
    if (lane < n) *reinterpret_cast<float4*>(a.de_e + (uint64_t)(p0 + lane) * GH) = make_float4(dal[0], dal[1], dal[2], dal[3]);
    float pr[GH];
#pragma unroll
    for (int h = 0; h < GH; ++h) pr[h] = alpha[h] * dal[h];
    warp_sum4(pr[0], pr[1], pr[2], pr[3]);
#pragma unroll
    for (int h = 0; h < GH; ++h) t[h] += pr[h];
  }
  // sweep 2: d e
  float dad[GH] = {0.f, 0.f, 0.f, 0.f};
  float du[GH][VPL][VN];
#pragma unroll
  for (int h = 0; h < GH; ++h)
#pragma unroll
    for (int v = 0; v < VPL; ++v)
#pragma unroll
      for (int k = 0; k < VN; ++k) du[h][v][k] = 0.f;
  for (int p0 = b; p0 < e; p0 += 32) {
    const int n = min(32, e - p0);
    const int cl = window_entry(a.col, p0, e, lane);
    float alpha[GH], dal[GH] = {0.f, 0.f, 0.f, 0.f}, sraw[GH], mask[GH] = {1.f, 1.f, 1.f, 1.f};
    if (lane < n) {
      const float4 d4 = *reinterpret_cast<const float4*>(a.de_e + (uint64_t)(p0 + lane) * GH);
      dal[0] = d4.x; dal[1] = d4.y; dal[2] = d4.z; dal[3] = d4.w;
    }
    if (a.p_drop > 0.f) dropout_scale4(mix_epoch(a.seed, a.epoch), (uint64_t)(p0 + lane), a.p_drop, mask);
    if (kT) {
      const float4 al4 = ldg_f4(a.alpha_in + (uint64_t)max(min(p0 + lane, p0 + n - 1), 0) * GH);
      const float al[GH] = {al4.x, al4.y, al4.z, al4.w};
#pragma unroll
      for (int h = 0; h < GH; ++h) { sraw[h] = 1.0f; alpha[h] = lane < n ? al[h] : 0.f; }
    } else {
      const float4 as4 = ldg_f4(a.a + (uint64_t)(uint32_t)cl * a.lda);
      float as[GH] = {as4.x, as4.y, as4.z, as4.w};
      if (a.ebias && lane < n) {
        const float4 b4 = ldg_f4(a.ebias + (uint64_t)(p0 + lane) * GH);
        as[0] += b4.x; as[1] += b4.y; as[2] += b4.z; as[3] += b4.w;
      }
#pragma unroll
      for (int h = 0; h < GH; ++h) {
        sraw[h] = as[h] + ad[h];
        alpha[h] = lane < n ? __expf(lrelu(sraw[h], a.slope) - sm[h]) * rinv[h] : 0.f;
      }
    }
    float de[GH];
    gatz_bwd_finish<T, VPL>(a, p0, n, lane, alpha, dal, sraw, mask, t, dad, de);
    if (kT && a.z)                              // du_i += sum_j de_ij x_j (TransformerConv: gradient of u_i)
      for (int j = 0; j < n; j += GatzCfg<VPL>::BU) gatz_gather_fma<T, VPL>(du, xb, a.xrow_bytes, cl, de, j, []() {});
  }
  if (!kT) {
    warp_sum4(dad[0], dad[1], dad[2], dad[3]);
    if (lane < GH) store_da(a, i, GH + lane, pick4(dad, lane));
  } else if (a.z) {
    gatz_store<T, VPL>(reinterpret_cast<char*>(a.z) + (uint64_t)i * a.zrow_bytes, du, lane);
  }
}

template <typename T, int VPL, bool kT>
__global__ void __launch_bounds__(256, 2) gatz_bwd_dst_kernel(const GatzArgs a) {
  constexpr int VN = Vec<T>::N;
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  const char* xb = reinterpret_cast<const char*>(a.x) + lane * 16;
  asm volatile("" : "+l"(xb));     // keep the per-lane base in a register pair: one IMAD.WIDE.U32 per gathered row
  WarpRows r;
  if (!r.begin(a.ord, a.n_rows, wi, a.rowptr)) return;
  int cl = window_entry(a.col, r.b, r.e, lane);
  while (true) {
    const int cl2 = window_entry(a.col, r.b2, r.e2, lane);
    r.look_ahead(a.ord, wi, a.rowptr);
    const int len = r.e - r.b;
    if (a.only_long && len <= 8) {
      // done by gatz_bwd_dst_mma_kernel
    } else if (len > 32) {
      gatz_bwd_dst_long<T, VPL, kT>(a, r.i, r.b, r.e);
    } else if (VPL == 1 && len > 0 && len <= 8 && (!kT || a.z)) {
      // every mesh row, packed (entry, head)-per-lane layout (see head_max8): ONE gather serves the d alpha dots and -
      // TransformerConv - du_i = sum_j de_ij x_j (the rows stay in registers; dz_i is dead after the dots).
      const int u = lane >> 2, h = lane & 3;
      uint4 buf[8][VPL];
#pragma unroll
      for (int e8 = 0; e8 < 8; ++e8) {
        const uint32_t c = (uint32_t)__shfl_sync(0xffffffffu, cl, e8);
        const char* p = xb + (uint64_t)c * a.xrow_bytes;
#pragma unroll
        for (int v = 0; v < VPL; ++v) buf[e8][v] = ldg_row16(p + 512 * v);
      }
      // per-(entry, head) side inputs, requested before anything consumes the gathers
      float in0 = 0.f, in1 = 0.f, in2 = 0.f, in3 = 0.f;
      const int pos = max(min(r.b + u, r.e - 1), 0);
      if (kT) {
        asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(in0) : "l"(a.alpha_in + (uint64_t)pos * GH + h));      // alpha (forward)
      } else {
        const uint32_t cu = (uint32_t)__shfl_sync(0xffffffffu, cl, u);
        asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(in0) : "l"(a.a + (uint64_t)cu * a.lda + h));           // a_src
        asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(in1) : "l"(a.a + (uint64_t)r.i * a.lda + GH + h));     // a_dst
        asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(in2) : "l"(a.smax + (uint64_t)r.i * GH + h));
        asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(in3) : "l"(a.ssum + (uint64_t)r.i * GH + h));
      }
      float dal;
      {
        uint4 dzraw[GH][VPL];
        gatz_load_dz_raw<VPL>(a, r.i, lane, dzraw);
        float dzf[GH][VPL][VN];
        gatz_unpack_dz<T, VPL>(dzraw, dzf);
        float part[32];
#pragma unroll
        for (int e8 = 0; e8 < 8; ++e8) {
          float p0[GH], p1[GH];
#pragma unroll
          for (int hh = 0; hh < GH; ++hh) { p0[hh] = 0.f; p1[hh] = 0.f; }
#pragma unroll
          for (int v = 0; v < VPL; ++v) {
            float f[VN];
            unpack_row16(buf[e8][v], f, T());
#pragma unroll
            for (int hh = 0; hh < GH; ++hh)
#pragma unroll
              for (int k = 0; k < VN; k += 2) ffma2_mul(p0[hh], p1[hh], dzf[hh][v][k], dzf[hh][v][k + 1], f[k], f[k + 1]);
          }
#pragma unroll
          for (int hh = 0; hh < GH; ++hh) part[e8 * GH + hh] = p0[hh] + p1[hh];
        }
        dal = warp_transpose_sum32(part);                        // lane 4u + h: dz_ih . x_u
      }
      float alpha, sraw = 1.0f;
      if (kT) {
        float f[Vec<T>::N];                                       // d s_alpha of the row: columns H*F .. H*F+3 of dz_aug
        ld16_rows<T>(reinterpret_cast<const char*>(a.dz) + (uint64_t)r.i * a.dzrow_bytes + GH * VPL * 512, f);
        float dsa[GH];
#pragma unroll
        for (int hh = 0; hh < GH; ++hh) dsa[hh] = f[hh];
        dal += pick4(dsa, h);
        if (a.ebias && u < len) dal += __ldg(a.ebias + (uint64_t)(r.b + u) * GH + h);
        alpha = u < len ? in0 : 0.f;
      } else {
        sraw = in0 + in1;
        if (a.ebias && u < len) sraw += __ldg(a.ebias + (uint64_t)(r.b + u) * GH + h);
        alpha = u < len ? __expf(lrelu(sraw, a.slope) - in2) * (1.0f / in3) : 0.f;
      }
      const float mask = a.p_drop > 0.f ? packed_keep_scale(mix_epoch(a.seed, a.epoch), (uint64_t)(r.b + u), a.p_drop, h) : 1.0f;
      dal *= mask;                                                // d(alpha) of the pre-dropout probability
      const float t = head_sum8(alpha * dal);
      const float de = alpha * (dal - t) * (sraw > 0.f ? 1.0f : a.slope);
      if (u < len) {                                              // coalesced 4-byte stores, target-major
        a.alpha_e[(uint64_t)(r.b + u) * GH + h] = alpha * mask;
        a.de_e[(uint64_t)(r.b + u) * GH + h] = de;
      }
      if (!kT) {
        const float dad = head_sum8(de);
        if (lane < GH) store_da(a, r.i, GH + lane, dad);
      } else {
        float du[GH][VPL][VN];
#pragma unroll
        for (int hh = 0; hh < GH; ++hh)
#pragma unroll
          for (int v = 0; v < VPL; ++v)
#pragma unroll
            for (int k = 0; k < VN; ++k) du[hh][v][k] = 0.f;
        switch (len) {
#define B2G_CASE(KK) case KK: packed_fma<T, VPL, KK>(du, buf, de); break;
          B2G_CASE(1) B2G_CASE(2) B2G_CASE(3) B2G_CASE(4) B2G_CASE(5) B2G_CASE(6) B2G_CASE(7) B2G_CASE(8)
#undef B2G_CASE
          default: break;
        }
        gatz_store<T, VPL>(reinterpret_cast<char*>(a.z) + (uint64_t)r.i * a.zrow_bytes, du, lane);
      }
    } else {
      float dzf[GH][VPL][VN];
      uint4 dzraw[GH][VPL];
      gatz_load_dz_raw<VPL>(a, r.i, lane, dzraw);
      float ad[GH], sm[GH], rinv[GH];
      float alpha[GH], dal[GH], sraw[GH], mask[GH];
      if (len > 0) {
        gatz_bwd_window<T, VPL, kT>(a, dzf, xb, cl, r.b, len, lane, ad, sm, rinv, alpha, dal, sraw, mask, [&]() {
          gatz_unpack_dz<T, VPL>(dzraw, dzf);                     // after the first gathers are in flight
          gatz_bwd_row_inputs<T, kT>(a, r.i, GH * VPL * 512, ad, sm, rinv);
        });
      } else {
#pragma unroll
        for (int h = 0; h < GH; ++h) { alpha[h] = 0.f; dal[h] = 0.f; sraw[h] = 1.f; mask[h] = 1.f; }
      }
      float t[GH];
#pragma unroll
      for (int h = 0; h < GH; ++h) t[h] = alpha[h] * dal[h];
      warp_sum4(t[0], t[1], t[2], t[3]);
      float dad[GH] = {0.f, 0.f, 0.f, 0.f}, de[GH];
      gatz_bwd_finish<T, VPL>(a, r.b, len, lane, alpha, dal, sraw, mask, t, dad, de);
      if (!kT) {
        warp_sum4(dad[0], dad[1], dad[2], dad[3]);
        if (lane < GH) store_da(a, r.i, GH + lane, pick4(dad, lane));
      } else if (a.z) {                          // du_i = sum_j de_ij x_j (rows of 9..32 entries, or VPL = 2: second gather)
        float du[GH][VPL][VN];
#pragma unroll
        for (int h = 0; h < GH; ++h)
#pragma unroll
          for (int v = 0; v < VPL; ++v)
#pragma unroll
            for (int k = 0; k < VN; ++k) du[h][v][k] = 0.f;
        for (int j = 0; j < len; j += GatzCfg<VPL>::BU) gatz_gather_fma<T, VPL>(du, xb, a.xrow_bytes, cl, de, j, []() {});
        gatz_store<T, VPL>(reinterpret_cast<char*>(a.z) + (uint64_t)r.i * a.zrow_bytes, du, lane);
      }
    }
    if (!r.shift()) break;
    cl = cl2;
  }
}

// ------------------------------------------------------------------------------------------ TransformerConv forward
// TransformerConv(heads = 4, concat = False) (gnn_model.py:77-80,170) by the same linearity:
//   q_ih . k_jh = x_j . (Wk_h^T q_ih) + q_ih . bk_h; the second term is constant along the softmax axis and drops out, so
//   with u_i = x_i Mq + cq (ONE GEMM; Mq_h = Wq_h^T Wk_h / sqrt(C), cq_h = bq_h^T Wk_h / sqrt(C)) the logits are
//   e_ijh = u_ih . x_j, and  out_i = 1/H sum_h (Wv_h z_ih + bv_h s_ih) + Ws x_i + bs  with z_ih = sum_j alpha_ijh x_j and
//   s_ih = sum_j alpha_ijh (1 without dropout).  The kernel writes  z_aug_i = [z_i1 .. z_iH | s_i1 .. s_iH 0 0 0 0 | x_i]
//   so that the value projection, its bias and the skip connection are ONE GEMM (k = H*F + 8 + F).
// Only F-wide rows of x are gathered (twice: logits, then weighted sum; the second pass hits L1) instead of the H*C-wide
// rows of k and v (4 KB per edge at H = 4, C = 256, bf16).
struct TzScratch { float v[GH]; };

// edge-feature term of window [p0, p0 + n): s[h] += ebias[p0 + lane, h] (one entry per lane)
__device__ __forceinline__ void tz_add_ebias(const GatzArgs& a, int p0, int n, int lane, float (&s)[GH]) {
  if (a.ebias && lane < n) {
    const float4 b4 = ldg_f4(a.ebias + (uint64_t)(p0 + lane) * GH);
    s[0] += b4.x; s[1] += b4.y; s[2] += b4.z; s[3] += b4.w;
  }
}

template <typename T, int VPL>
__device__ __forceinline__ void tz_store_tail(const GatzArgs& a, uint32_t i, int lane, const float (&ssum)[GH]) {
  constexpr int VN = Vec<T>::N;
  char* zrow = reinterpret_cast<char*>(a.z) + (uint64_t)i * a.zrow_bytes + GH * VPL * 512;
  if (lane < 8 / VN) {                        // s_alpha (H values) + zero padding: 8 elements of T
    float f[VN];
#pragma unroll
    for (int k = 0; k < VN; ++k) {
      const int idx = lane * VN + k;
      f[k] = idx < GH ? pick4(ssum, idx) : 0.f;
    }
    Vec<T> o;
    o.from_float(f);
    *reinterpret_cast<uint4*>(zrow + lane * 16) = *reinterpret_cast<uint4*>(&o.v);
  }
  const char* xi = reinterpret_cast<const char*>(a.x_self) + (uint64_t)i * a.xrow_bytes + lane * 16;
#pragma unroll
  for (int v = 0; v < VPL; ++v)
    __stcs(reinterpret_cast<uint4*>(zrow + 8 * sizeof(T) + v * 512 + lane * 16), ldg_row16(xi + 512 * v));
}

template <typename T, int VPL>
__device__ __noinline__ void tz_fwd_long(const GatzArgs a, uint32_t i, int b, int e) {
  constexpr int VN = Vec<T>::N;
  const int lane = threadIdx.x & 31;
  const char* xb = reinterpret_cast<const char*>(a.x) + lane * 16;
  asm volatile("" : "+l"(xb));     // keep the per-lane base in a register pair: one IMAD.WIDE.U32 per gathered row
  float uf[GH][VPL][VN];
  gatz_load_dz<T, VPL>(a, i, lane, uf);
  float m[GH], zs[GH];
#pragma unroll
  for (int h = 0; h < GH; ++h) { m[h] = -INFINITY; zs[h] = 0.f; }
  // sweep 1: logits (parked in alpha_e or, without a backward pass, recomputed), running max
  for (int p0 = b; p0 < e; p0 += 32) {
    const int n = min(32, e - p0);
    const int cl = window_entry(a.col, p0, e, lane);
    float s[GH] = {0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < n; j += 8) gatz_dalpha8<T, VPL>(s, uf, xb, a.xrow_bytes, cl, j, lane, []() {});
    tz_add_ebias(a, p0, n, lane, s);
#pragma unroll
    for (int h = 0; h < GH; ++h) m[h] = fmaxf(m[h], warp_max_redux(lane < n ? s[h] : -INFINITY));
  }
  for (int p0 = b; p0 < e; p0 += 32) {
    const int n = min(32, e - p0);
    const int cl = window_entry(a.col, p0, e, lane);
    float s[GH] = {0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < n; j += 8) gatz_dalpha8<T, VPL>(s, uf, xb, a.xrow_bytes, cl, j, lane, []() {});
    tz_add_ebias(a, p0, n, lane, s);
    float p[GH];
#pragma unroll
    for (int h = 0; h < GH; ++h) p[h] = lane < n ? __expf(s[h] - m[h]) : 0.f;
    warp_sum4(p[0], p[1], p[2], p[3]);
#pragma unroll
    for (int h = 0; h < GH; ++h) zs[h] += p[h];
  }
  float inv[GH], ssum[GH] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int h = 0; h < GH; ++h) inv[h] = 1.0f / (zs[h] + 1e-16f);
  float acc[GH][VPL][VN];
#pragma unroll
  for (int h = 0; h < GH; ++h)
#pragma unroll
    for (int v = 0; v < VPL; ++v)
#pragma unroll
      for (int k = 0; k < VN; ++k) acc[h][v][k] = 0.f;
  for (int p0 = b; p0 < e; p0 += 32) {
    const int n = min(32, e - p0);
    const int cl = window_entry(a.col, p0, e, lane);
    float s[GH] = {0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < n; j += 8) gatz_dalpha8<T, VPL>(s, uf, xb, a.xrow_bytes, cl, j, lane, []() {});
    tz_add_ebias(a, p0, n, lane, s);
    float w[GH];
#pragma unroll
    for (int h = 0; h < GH; ++h) w[h] = lane < n ? __expf(s[h] - m[h]) * inv[h] : 0.f;
    if (a.alpha_e && lane < n) *reinterpret_cast<float4*>(a.alpha_e + (uint64_t)(p0 + lane) * GH) = make_float4(w[0], w[1], w[2], w[3]);
    if (a.p_drop > 0.f) {
      float sc[4];
      dropout_scale4(mix_epoch(a.seed, a.epoch), (uint64_t)(p0 + lane), a.p_drop, sc);
#pragma unroll
      for (int h = 0; h < GH; ++h) w[h] *= sc[h];
    }
    float ws[GH] = {w[0], w[1], w[2], w[3]};
    warp_sum4(ws[0], ws[1], ws[2], ws[3]);
#pragma unroll
    for (int h = 0; h < GH; ++h) ssum[h] += ws[h];
    if (a.alpha_only) {
      if (a.de_e && lane < n) *reinterpret_cast<float4*>(a.de_e + (uint64_t)(p0 + lane) * GH) = make_float4(w[0], w[1], w[2], w[3]);
      continue;
    }
    for (int j = 0; j < n; j += GatzCfg<VPL>::BU) gatz_gather_fma<T, VPL>(acc, xb, a.xrow_bytes, cl, w, j, []() {});
  }
  if (a.alpha_only) {
    if (a.smax && lane < GH) a.smax[(uint64_t)i * GH + lane] = pick4(ssum, lane);
    return;
  }
  gatz_store<T, VPL>(reinterpret_cast<char*>(a.z) + (uint64_t)i * a.zrow_bytes, acc, lane);
  tz_store_tail<T, VPL>(a, i, lane, ssum);
}

template <typename T, int VPL>
__global__ void __launch_bounds__(256, 2) tz_fwd_kernel(const GatzArgs a) {
  constexpr int VN = Vec<T>::N;
  constexpr int BU = GatzCfg<VPL>::BU;
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  const char* xb = reinterpret_cast<const char*>(a.x) + lane * 16;
  asm volatile("" : "+l"(xb));     // keep the per-lane base in a register pair: one IMAD.WIDE.U32 per gathered row
  WarpRows r;
  if (!r.begin(a.ord, a.n_rows, wi, a.rowptr)) return;
  int cl = window_entry(a.col, r.b, r.e, lane);
  while (true) {
    const int cl2 = window_entry(a.col, r.b2, r.e2, lane);
    r.look_ahead(a.ord, wi, a.rowptr);
    const int len = r.e - r.b;
    if (len > 32) {
      tz_fwd_long<T, VPL>(a, r.i, r.b, r.e);
    } else if (VPL == 1 && len > 0 && len <= 8) {
      // every mesh row: ONE gather of the (<= 8) neighbour rows serves both the logits and the weighted sums (the rows
      // stay in registers across the softmax; u_i is dead by then).  ncu on the two-gather version: 32 GB of reads for
      // 26 GB compulsory, and the second gather's latency sat in the middle of every row's dependency chain.
      uint4 buf[8][VPL];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const uint32_t c = (uint32_t)__shfl_sync(0xffffffffu, cl, u);
        const char* p = xb + (uint64_t)c * a.xrow_bytes;
#pragma unroll
        for (int v = 0; v < VPL; ++v) buf[u][v] = ldg_row16(p + 512 * v);
      }
      float wp;
      {
        uint4 uraw[GH][VPL];
        gatz_load_dz_raw<VPL>(a, r.i, lane, uraw);
        float uf[GH][VPL][VN];
        gatz_unpack_dz<T, VPL>(uraw, uf);
        float part[32];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          float p0[GH], p1[GH];
#pragma unroll
          for (int h = 0; h < GH; ++h) { p0[h] = 0.f; p1[h] = 0.f; }
#pragma unroll
          for (int v = 0; v < VPL; ++v) {
            float f[VN];
            unpack_row16(buf[u][v], f, T());
#pragma unroll
            for (int h = 0; h < GH; ++h)
#pragma unroll
              for (int k = 0; k < VN; k += 2) ffma2_mul(p0[h], p1[h], uf[h][v][k], uf[h][v][k + 1], f[k], f[k + 1]);
          }
#pragma unroll
          for (int h = 0; h < GH; ++h) part[u * GH + h] = p0[h] + p1[h];
        }
        wp = warp_transpose_sum32(part);                     // lane 4u + h: logit of entry u, head h
      }
      const int pu = lane >> 2, ph = lane & 3;                // packed (entry, head) softmax, see head_max8
      if (a.ebias && pu < len) wp += __ldg(a.ebias + (uint64_t)(r.b + pu) * GH + ph);
      {
        const float sc = pu < len ? wp : -INFINITY;
        const float m = head_max8(sc);
        wp = pu < len ? __expf(sc - m) : 0.f;
        wp *= 1.0f / (head_sum8(wp) + 1e-16f);
      }
      if (a.alpha_e && pu < len) a.alpha_e[(uint64_t)(r.b + pu) * GH + ph] = wp;      // coalesced 4-byte stores
      if (a.p_drop > 0.f) wp *= packed_keep_scale(mix_epoch(a.seed, a.epoch), (uint64_t)(r.b + pu), a.p_drop, ph);
      const float ss = head_sum8(wp);                          // per-head weight sums (every lane of the head has it)
      if (a.alpha_only) {
        if (a.de_e && pu < len) a.de_e[(uint64_t)(r.b + pu) * GH + ph] = wp;
        if (a.smax && lane < GH) a.smax[(uint64_t)r.i * GH + lane] = ss;       // lanes 0..3: entry 0, head = lane
        if (!r.shift()) break;
        cl = cl2;
        continue;
      }
      float ssum[GH];
#pragma unroll
      for (int h = 0; h < GH; ++h) ssum[h] = __shfl_sync(0xffffffffu, ss, h);
      float acc[GH][VPL][VN];
#pragma unroll
      for (int h = 0; h < GH; ++h)
#pragma unroll
        for (int v = 0; v < VPL; ++v)
#pragma unroll
          for (int k = 0; k < VN; ++k) acc[h][v][k] = 0.f;
      switch (len) {                                            // exact-length FMA phase
#define B2G_CASE(KK) case KK: packed_fma<T, VPL, KK>(acc, buf, wp); break;
        B2G_CASE(1) B2G_CASE(2) B2G_CASE(3) B2G_CASE(4) B2G_CASE(5) B2G_CASE(6) B2G_CASE(7) B2G_CASE(8)
#undef B2G_CASE
        default: break;
      }
      gatz_store<T, VPL>(reinterpret_cast<char*>(a.z) + (uint64_t)r.i * a.zrow_bytes, acc, lane);
      tz_store_tail<T, VPL>(a, r.i, lane, ssum);
    } else {
      float w[GH] = {0.f, 0.f, 0.f, 0.f}, ssum[GH] = {0.f, 0.f, 0.f, 0.f};
      if (len > 0) {                          // logits: u_i (registers) . x_j, one entry per lane
        float uf[GH][VPL][VN];
        uint4 uraw[GH][VPL];
        gatz_load_dz_raw<VPL>(a, r.i, lane, uraw);
        gatz_dalpha8<T, VPL>(w, uf, xb, a.xrow_bytes, cl, 0, lane, [&]() { gatz_unpack_dz<T, VPL>(uraw, uf); });
        for (int j = 8; j < len; j += 8) gatz_dalpha8<T, VPL>(w, uf, xb, a.xrow_bytes, cl, j, lane, []() {});
        tz_add_ebias(a, r.b, len, lane, w);
      }
      float zs[GH];
#pragma unroll
      for (int h = 0; h < GH; ++h) {
        const float s = lane < len ? w[h] : -INFINITY;
        const float m = warp_max_redux(s);
        w[h] = lane < len ? __expf(s - m) : 0.f;
        zs[h] = w[h];
      }
      warp_sum4(zs[0], zs[1], zs[2], zs[3]);
#pragma unroll
      for (int h = 0; h < GH; ++h) w[h] *= 1.0f / (zs[h] + 1e-16f);
      if (a.alpha_e && lane < len) *reinterpret_cast<float4*>(a.alpha_e + (uint64_t)(r.b + lane) * GH) = make_float4(w[0], w[1], w[2], w[3]);
      if (a.p_drop > 0.f) {
        float sc[4];
        dropout_scale4(mix_epoch(a.seed, a.epoch), (uint64_t)(r.b + lane), a.p_drop, sc);
#pragma unroll
        for (int h = 0; h < GH; ++h) w[h] *= sc[h];
      }
#pragma unroll
      for (int h = 0; h < GH; ++h) ssum[h] = w[h];
      warp_sum4(ssum[0], ssum[1], ssum[2], ssum[3]);
      if (a.alpha_only) {
        if (a.de_e && lane < len) *reinterpret_cast<float4*>(a.de_e + (uint64_t)(r.b + lane) * GH) = make_float4(w[0], w[1], w[2], w[3]);
        if (a.smax && lane < GH) a.smax[(uint64_t)r.i * GH + lane] = pick4(ssum, lane);
        if (!r.shift()) break;
        cl = cl2;
        continue;
      }
      float acc[GH][VPL][VN];
#pragma unroll
      for (int h = 0; h < GH; ++h)
#pragma unroll
        for (int v = 0; v < VPL; ++v)
#pragma unroll
          for (int k = 0; k < VN; ++k) acc[h][v][k] = 0.f;
      for (int j = 0; j < len; j += BU) gatz_gather_fma<T, VPL>(acc, xb, a.xrow_bytes, cl, w, j, []() {});
      gatz_store<T, VPL>(reinterpret_cast<char*>(a.z) + (uint64_t)r.i * a.zrow_bytes, acc, lane);
      tz_store_tail<T, VPL>(a, r.i, lane, ssum);
    }
    if (!r.shift()) break;
    cl = cl2;
  }
}

// ------------------------------------------------------------------------------------------ attention weights on the tensor cores
// tz_alpha for bf16 rows of 512 bytes (F = 256, H = 4): the logits of a target row are the 8 x 4 matrix
//     L[e, h] = sum_f x_{j(e)}[f] * u_ih[f]         (u already carries 1/sqrt(C), see tz_fwd_kernel)
// = one m16n8k16 tile product chain (rows 8..15 and columns 4..7 are padding) with K = 256 in 16 steps.  The SIMT version
// spent 592 warp instructions per row on it (unpack + packed FMAs + a 31-shuffle transposed reduction; ncu: issue 43 % at
// 16 warps per SM, 12.3 ms at cfg4 for 30 GB = 2.4 TB/s); here it is 16 HMMAs on operands that are ALREADY in fragment layout:
// the order of k is free in a dot product, so lane (g, q) = (entry, quarter) simply loads the 16-byte pieces q, q + 4, ...
// of row x_{j(g)} (A fragment: a0 / a2 of step s = registers 2s, 2s + 1 of those 32) and of u_i's head g & 3 (B fragment),
// 64 contiguous bytes per row and instruction.  The accumulator fragment leaves entry g / heads 2q, 2q + 1 in lanes q < 2;
// two shuffles bring them to the packed (entry, head) = lane layout of the softmax (head_max8).
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a2, uint32_t b0, uint32_t b1) {
  const uint32_t zero = 0u;                       // rows 8..15 of the A tile
  asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(zero), "r"(a2), "r"(zero), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(256, 2) tz_alpha_mma_kernel(const GatzArgs a) {
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  const int g = lane >> 2, q = lane & 3;
  const char* xb = reinterpret_cast<const char*>(a.x) + q * 16;
  const char* ub = reinterpret_cast<const char*>(a.dz) + (g & 3) * 512 + q * 16;
  asm volatile("" : "+l"(xb));
  asm volatile("" : "+l"(ub));
  WarpRows r;
  if (!r.begin(a.ord, a.n_rows, wi, a.rowptr)) return;
  // entry g of the 8-entry window at p0 (lanes past the row's end repeat its last entry: a valid address, logit masked)
  auto entry = [&](int p0, int e_) -> int { return ldg_i32_ordered(a.col + max(min(p0 + g, e_ - 1), 0)); };
  int cl = entry(r.b, r.e);
  while (true) {
    const int cl2 = entry(r.b2, r.e2);
    r.look_ahead(a.ord, wi, a.rowptr);
    const int len = r.e - r.b;
    uint4 U[8];
    {
      const char* pu = ub + (uint64_t)r.i * a.dzrow_bytes;
#pragma unroll
      for (int m = 0; m < 8; ++m)
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(U[m].x), "=r"(U[m].y), "=r"(U[m].z), "=r"(U[m].w) : "l"(pu + 64 * m));
    }
    // logit (+ edge term) of entry p0 + g, head q, for the window whose entry-g column index is c; -inf past the row's end
    auto window_logit = [&](int c, int p0) -> float {
      uint4 X[8];
      const char* px = xb + (uint64_t)(uint32_t)c * a.xrow_bytes;
#pragma unroll
      for (int m = 0; m < 8; ++m) X[m] = ldg_row16(px + 64 * m);
      float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};     // two chains: the HMMA latencies overlap
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        mma_bf16_16816(c0, X[m].x, X[m].y, U[m].x, U[m].y);
        mma_bf16_16816(c1, X[m].z, X[m].w, U[m].z, U[m].w);
      }
      const float l0 = c0[0] + c1[0], l1 = c0[1] + c1[1];          // lane (g, q < 2): entry g, heads 2q and 2q + 1
      const int src = (lane & ~3) | (q >> 1);
      const float v0 = __shfl_sync(0xffffffffu, l0, src), v1 = __shfl_sync(0xffffffffu, l1, src);
      float s = (q & 1) ? v1 : v0;                                  // lane 4g + h: logit of entry g, head h
      const bool in = p0 + g < r.e;
      if (a.ebias && in) s += __ldg(a.ebias + (uint64_t)(p0 + g) * GH + q);
      return in ? s : -INFINITY;
    };
    // weights w of the window at p0 (0 past the end): alpha_e, dropout, de_e; returns the per-head sum of what was kept
    auto emit = [&](float w, int p0) -> float {
      const bool in = p0 + g < r.e;
      if (a.alpha_e && in) a.alpha_e[(uint64_t)(p0 + g) * GH + q] = w;            // coalesced 4-byte stores
      if (a.p_drop > 0.f) w *= packed_keep_scale(mix_epoch(a.seed, a.epoch), (uint64_t)(p0 + g), a.p_drop, q);
      if (a.de_e && in) a.de_e[(uint64_t)(p0 + g) * GH + q] = w;
      return head_sum8(w);
    };
    float ss = 0.f;
    if (len > 0 && len <= 8) {                                      // every mesh row
      const float sc = window_logit(cl, r.b);
      const float m = head_max8(sc);
      float w = __expf(sc - m);                                     // exp(-inf) = 0 past the end
      w *= 1.0f / (head_sum8(w) + 1e-16f);
      ss = emit(w, r.b);
    } else if (len > 8) {                                           // cold: three sweeps over the 8-entry windows
      float m = -INFINITY, zs = 0.f;
      for (int p0 = r.b; p0 < r.e; p0 += 8) m = fmaxf(m, head_max8(window_logit(entry(p0, r.e), p0)));
      for (int p0 = r.b; p0 < r.e; p0 += 8) zs += head_sum8(__expf(window_logit(entry(p0, r.e), p0) - m));
      const float inv = 1.0f / (zs + 1e-16f);
      for (int p0 = r.b; p0 < r.e; p0 += 8) ss += emit(__expf(window_logit(entry(p0, r.e), p0) - m) * inv, p0);
    }
    if (a.smax && lane < GH) a.smax[(uint64_t)r.i * GH + lane] = ss;          // lanes 0..3: head = lane
    if (!r.shift()) break;
    cl = cl2;
  }
}

// ------------------------------------------------------------------------------------------ backward, source side
// Row j of the TRANSPOSED CSR: entries (i = col[t], p = perm[t] = the edge's position in the target-major CSR).
//   y_j[h] = sum_t alpha_e[p, h] * g_i,   d a_src[j, h] = sum_t de_e[p, h]
template <typename T, int VPL>
__global__ void __launch_bounds__(256, B2G_GATZ_MINB) gatz_bwd_src_kernel(const GatzArgs a) {
  constexpr int VN = Vec<T>::N;
  constexpr int BU = GatzCfg<VPL>::BU;
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  const char* xb = reinterpret_cast<const char*>(a.x) + lane * 16;
  asm volatile("" : "+l"(xb));     // keep the per-lane base in a register pair: one IMAD.WIDE.U32 per gathered row
  WarpRows r;
  if (!r.begin(a.ord, a.n_rows, wi, a.rowptr)) return;
  // perm == NULL: the weights are already in this CSR's order (position = entry index)
  auto perm_entry = [&](int p0, int e_) -> int {
    return a.perm ? window_entry(a.perm, p0, e_, lane) : max(min(p0 + lane, e_ - 1), 0);
  };
  int cl = window_entry(a.col, r.b, r.e, lane), pl = perm_entry(r.b, r.e);
  while (true) {
    const int cl2 = window_entry(a.col, r.b2, r.e2, lane), pl2 = perm_entry(r.b2, r.e2);
    r.look_ahead(a.ord, wi, a.rowptr);
    float acc[GH][VPL][VN];
#pragma unroll
    for (int h = 0; h < GH; ++h)
#pragma unroll
      for (int v = 0; v < VPL; ++v)
#pragma unroll
        for (int k = 0; k < VN; ++k) acc[h][v][k] = 0.f;
    float das[GH] = {0.f, 0.f, 0.f, 0.f};
    for (int p0 = r.b; p0 < r.e; p0 += 32) {                 // one window on meshes
      if (p0 != r.b) {
        cl = window_entry(a.col, p0, r.e, lane);
        pl = perm_entry(p0, r.e);
      }
      const int n = min(32, r.e - p0);
      const float4 al4 = ldg_f4(a.alpha_e + (uint64_t)(uint32_t)pl * GH);
      const float4 de4 = ldg_f4(a.de_e + (uint64_t)(uint32_t)pl * GH);
      const bool on = lane < n;
      float w[GH];
      auto weights = [&]() {                  // after the row loads are issued (in-order issue: see aggregate_rows.cu rows_batch)
        w[0] = on ? al4.x : 0.f; w[1] = on ? al4.y : 0.f; w[2] = on ? al4.z : 0.f; w[3] = on ? al4.w : 0.f;
        if (on) { das[0] += de4.x; das[1] += de4.y; das[2] += de4.z; das[3] += de4.w; }
      };
      if (n > BU) {
        gatz_gather_fma<T, VPL>(acc, xb, a.xrow_bytes, cl, w, 0, weights);
        for (int j = BU; j < n; j += BU) gatz_gather_fma<T, VPL>(acc, xb, a.xrow_bytes, cl, w, j, []() {});
      } else {
        switch (n) {
#define B2G_CASE(KK) case KK: if (KK <= BU) gatz_gather_fma<T, VPL, (KK <= BU ? KK : 1)>(acc, xb, a.xrow_bytes, cl, w, 0, weights); break;
          B2G_CASE(1) B2G_CASE(2) B2G_CASE(3) B2G_CASE(4) B2G_CASE(5) B2G_CASE(6) B2G_CASE(7) B2G_CASE(8)
#undef B2G_CASE
          default: break;
        }
      }
    }
    if (a.d_a) {
      warp_sum4(das[0], das[1], das[2], das[3]);
      if (lane < GH) store_da(a, r.i, lane, pick4(das, lane));
    }
    gatz_store<T, VPL>(reinterpret_cast<char*>(a.z) + (uint64_t)r.i * a.zrow_bytes, acc, lane);
    if (!r.shift()) break;
    cl = cl2; pl = pl2;
  }
}

// ------------------------------------------------------------------------------------------ weighted row sums on the tensor cores
// y[h][f] = sum_e w[e][h] * x_e[f] for the <= 8 entries of a window is a (heads x entries) . (entries x features) product:
// m16n8k8 tiles with A = the weights and B = 8 entries x 8 features.  The gathered rows arrive in the layout of
// tz_alpha_mma_kernel (lane (g, q) = entry g, 16-byte pieces q, q + 4, ...: the A-fragment layout of an entries x features
// matrix); one movmatrix (8x8 b16 transpose inside the warp) per register turns it into the B fragment: entries 2q, 2q + 1 of
// the two features that lane (., q)'s register holds, so the accumulator fragment of thread (g, q) is y[head g] at exactly the
// features of its own registers.  The weights keep fp32 accuracy: rows 0..3 of A carry their bf16 heads, rows 8..11 the bf16
// remainders (w = hi + lo to 2^-17), and the two halves of the accumulator are added.  Per 8-entry window: 32 movmatrix + 32
// HMMA + 64 FADD instead of 8 x (4 SHFL + 8 unpack + 16 packed FMA) = 224 instructions.
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  const __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&p);
}
__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t v) {
  uint32_t d;
  asm("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(v));
  return d;
}
__device__ __forceinline__ void mma_bf16_1688(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0, float c0, float c1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%7,%8,%9,%10};"
      : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
      : "r"(a0), "r"(a1), "r"(b0), "f"(c0), "f"(c1), "f"(0.f), "f"(0.f));
}
// A fragment from the packed weights wp (lane 4u + h: weight of entry u, head h; 0 past the row's end)
__device__ __forceinline__ void wsum_afrag(float wp, int g, int q, uint32_t& a0, uint32_t& a1) {
  const float w0 = __shfl_sync(0xffffffffu, wp, 8 * q + (g & 3)), w1 = __shfl_sync(0xffffffffu, wp, 8 * q + 4 + (g & 3));
  const __nv_bfloat16 h0 = __float2bfloat16_rn(w0), h1 = __float2bfloat16_rn(w1);
  const __nv_bfloat16 l0 = __float2bfloat16_rn(w0 - __bfloat162float(h0)), l1 = __float2bfloat16_rn(w1 - __bfloat162float(h1));
  a0 = g < 4 ? ((uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16)) : 0u;
  a1 = g < 4 ? ((uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16)) : 0u;
}
// single window (rows of <= 8 entries): products straight to the output row, no accumulators held
__device__ __forceinline__ void wsum_mma8_store(char* yrow, const uint4 (&X)[8], uint32_t a0, uint32_t a1, int g, int q) {
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const uint32_t xr[4] = {X[m].x, X[m].y, X[m].z, X[m].w};
    uint32_t o[4];
#pragma unroll
    for (int r4 = 0; r4 < 4; ++r4) {
      float d[4];
      mma_bf16_1688(d, a0, a1, movmatrix_trans(xr[r4]), 0.f, 0.f);
      o[r4] = pack_bf16x2(d[0] + d[2], d[1] + d[3]);
    }
    if (g < 4) __stcs(reinterpret_cast<uint4*>(yrow + g * 512 + (q + 4 * m) * 16), make_uint4(o[0], o[1], o[2], o[3]));
  }
}

// gatz_bwd_dst for bf16 rows of 512 bytes, rows of <= 8 entries (every mesh row; longer rows are left to the SIMT kernel's
// only_long pass): d alpha_eh = dz_ih . x_e as in tz_alpha_mma_kernel (16 HMMAs instead of ~380 instructions of unpack /
// packed FMA / transposed reduction), the softmax backward in the packed (entry, head) = lane layout, and - TransformerConv -
// du_i = sum_e de_eh x_e from the SAME registers through movmatrix + m16n8k8 (wsum_mma8_store).
template <bool kT>
__global__ void __launch_bounds__(256, 2) gatz_bwd_dst_mma_kernel(const GatzArgs a) {
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  const int g = lane >> 2, q = lane & 3;
  const char* xb = reinterpret_cast<const char*>(a.x) + q * 16;
  const char* ub = reinterpret_cast<const char*>(a.dz) + (g & 3) * 512 + q * 16;
  asm volatile("" : "+l"(xb));
  asm volatile("" : "+l"(ub));
  WarpRows r;
  if (!r.begin(a.ord, a.n_rows, wi, a.rowptr)) return;
  auto entry = [&](int b_, int e_) -> int { return ldg_i32_ordered(a.col + max(min(b_ + g, e_ - 1), 0)); };
  int cl = entry(r.b, r.e);
  while (true) {
    const int cl2 = entry(r.b2, r.e2);
    r.look_ahead(a.ord, wi, a.rowptr);
    const int len = r.e - r.b;
    if (len == 0) {
      if (!kT) {
        if (lane < GH) store_da(a, r.i, GH + lane, 0.f);
      } else if (a.z && g < 4) {
        char* zrow = reinterpret_cast<char*>(a.z) + (uint64_t)r.i * a.zrow_bytes;
#pragma unroll
        for (int m = 0; m < 8; ++m) __stcs(reinterpret_cast<uint4*>(zrow + g * 512 + (q + 4 * m) * 16), make_uint4(0u, 0u, 0u, 0u));
      }
    } else if (len <= 8) {
      uint4 X[8], U[8];
      const char* px = xb + (uint64_t)(uint32_t)cl * a.xrow_bytes;
#pragma unroll
      for (int m = 0; m < 8; ++m) X[m] = ldg_row16(px + 64 * m);
      const char* pu = ub + (uint64_t)r.i * a.dzrow_bytes;
#pragma unroll
      for (int m = 0; m < 8; ++m)
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(U[m].x), "=r"(U[m].y), "=r"(U[m].z), "=r"(U[m].w) : "l"(pu + 64 * m));
      // per-(entry, head) side inputs, requested before anything consumes the gathers
      const bool in = g < len;
      const uint64_t pos = (uint64_t)(r.b + min(g, len - 1));
      float in0 = 0.f, in1 = 0.f, in2 = 0.f, in3 = 0.f, eb = 0.f;
      if (kT) {
        asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(in0) : "l"(a.alpha_in + pos * GH + q));                 // alpha (forward)
        // d s_alpha of the row: bf16 columns H*F .. H*F+3 of dz_aug
        unsigned short raw;
        asm volatile("ld.global.nc.u16 %0, [%1];" : "=h"(raw)
                     : "l"(reinterpret_cast<const char*>(a.dz) + (uint64_t)r.i * a.dzrow_bytes + GH * 512 + q * 2));
        in1 = __uint_as_float((uint32_t)raw << 16);
      } else {
        asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(in0) : "l"(a.a + (uint64_t)(uint32_t)cl * a.lda + q));  // a_src
        asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(in1) : "l"(a.a + (uint64_t)r.i * a.lda + GH + q));      // a_dst
        asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(in2) : "l"(a.smax + (uint64_t)r.i * GH + q));
        asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(in3) : "l"(a.ssum + (uint64_t)r.i * GH + q));
      }
      if (a.ebias) asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(eb) : "l"(a.ebias + pos * GH + q));
      float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        mma_bf16_16816(c0, X[m].x, X[m].y, U[m].x, U[m].y);
        mma_bf16_16816(c1, X[m].z, X[m].w, U[m].z, U[m].w);
      }
      const float l0 = c0[0] + c1[0], l1 = c0[1] + c1[1];
      const int src = (lane & ~3) | (q >> 1);
      const float v0 = __shfl_sync(0xffffffffu, l0, src), v1 = __shfl_sync(0xffffffffu, l1, src);
      float dal = (q & 1) ? v1 : v0;                               // lane 4g + h: dz_ih . x_g
      float alpha, sraw = 1.0f;
      if (kT) {
        dal += in1;
        if (in) dal += eb;
        alpha = in ? in0 : 0.f;
      } else {
        sraw = in0 + in1;
        if (in) sraw += eb;
        alpha = in ? __expf(lrelu(sraw, a.slope) - in2) * (1.0f / in3) : 0.f;
      }
      const float mask = a.p_drop > 0.f ? packed_keep_scale(mix_epoch(a.seed, a.epoch), (uint64_t)(r.b + g), a.p_drop, q) : 1.0f;
      dal *= mask;
      const float t = head_sum8(alpha * dal);
      const float de = alpha * (dal - t) * (sraw > 0.f ? 1.0f : a.slope);
      if (in) {
        a.alpha_e[(uint64_t)(r.b + g) * GH + q] = alpha * mask;
        a.de_e[(uint64_t)(r.b + g) * GH + q] = de;
      }
      if (!kT) {
        const float dad = head_sum8(de);
        if (lane < GH) store_da(a, r.i, GH + lane, dad);
      } else if (a.z) {
        uint32_t a0, a1;
        wsum_afrag(de, g, q, a0, a1);                               // de = 0 past the row's end (alpha = 0)
        wsum_mma8_store(reinterpret_cast<char*>(a.z) + (uint64_t)r.i * a.zrow_bytes, X, a0, a1, g, q);
      }
    }
    if (!r.shift()) break;
    cl = cl2;
  }
}

// gatz_bwd_src for bf16 rows of 512 bytes: y_j = [sum_t alpha_e[p_t, h] g_{i_t}]_h, d a_src[j, h] = sum_t de_e[p_t, h]
__global__ void __launch_bounds__(256, 2) gatz_bwd_src_mma_kernel(const GatzArgs a) {
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  const int g = lane >> 2, q = lane & 3;
  const char* xb = reinterpret_cast<const char*>(a.x) + q * 16;
  asm volatile("" : "+l"(xb));
  WarpRows r;
  if (!r.begin(a.ord, a.n_rows, wi, a.rowptr)) return;
  auto entry_c = [&](int p0, int e_) -> int { return ldg_i32_ordered(a.col + max(min(p0 + g, e_ - 1), 0)); };
  auto entry_p = [&](int p0, int e_) -> int {                  // perm == NULL: the weights are in this CSR's order
    const int t = max(min(p0 + g, e_ - 1), 0);
    return a.perm ? ldg_i32_ordered(a.perm + t) : t;
  };
  int cl = entry_c(r.b, r.e), pl = entry_p(r.b, r.e);
  while (true) {
    const int cl2 = entry_c(r.b2, r.e2), pl2 = entry_p(r.b2, r.e2);
    r.look_ahead(a.ord, wi, a.rowptr);
    char* yrow = reinterpret_cast<char*>(a.z) + (uint64_t)r.i * a.zrow_bytes;
    float das = 0.f;
    // the window at p0 with entry-g indices (c, p): gathered rows, A fragment of the weights, d a_src partial sum
    auto window = [&](int c, int p, int p0, uint4 (&X)[8], uint32_t& a0, uint32_t& a1) {
      const char* px = xb + (uint64_t)(uint32_t)c * a.xrow_bytes;
#pragma unroll
      for (int m = 0; m < 8; ++m) X[m] = ldg_row16(px + 64 * m);
      float wv, dv = 0.f;
      asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(wv) : "l"(a.alpha_e + (uint64_t)(uint32_t)p * GH + q));
      if (a.d_a) asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(dv) : "l"(a.de_e + (uint64_t)(uint32_t)p * GH + q));
      const bool in = p0 + g < r.e;
      wsum_afrag(in ? wv : 0.f, g, q, a0, a1);
      if (a.d_a) das += head_sum8(in ? dv : 0.f);
    };
    if (r.e == r.b) {                                               // no entries: zeros (and no gather through an index
      if (g < 4) {                                                  // that belongs to no row: col may be a dummy when nnz = 0)
#pragma unroll
        for (int m = 0; m < 8; ++m) __stcs(reinterpret_cast<uint4*>(yrow + g * 512 + (q + 4 * m) * 16), make_uint4(0u, 0u, 0u, 0u));
      }
    } else if (r.e - r.b <= 8) {                                    // every mesh row
      uint4 X[8];
      uint32_t a0, a1;
      window(cl, pl, r.b, X, a0, a1);
      wsum_mma8_store(yrow, X, a0, a1, g, q);
    } else {                                                        // cold: piece by piece, accumulating over the 8-entry
#pragma unroll 1                                                    // windows (8 accumulators instead of 64: no spills)
      for (int m = 0; m < 8; ++m) {
        float acc[4][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
#pragma unroll 1
        for (int p0 = r.b; p0 < r.e; p0 += 8) {
          const int c = entry_c(p0, r.e), p = entry_p(p0, r.e);
          const uint4 x4 = ldg_row16(xb + (uint64_t)(uint32_t)c * a.xrow_bytes + 64 * m);
          const bool in = p0 + g < r.e;
          const float wv = __ldg(a.alpha_e + (uint64_t)(uint32_t)p * GH + q);
          if (m == 0 && a.d_a) das += head_sum8(in ? __ldg(a.de_e + (uint64_t)(uint32_t)p * GH + q) : 0.f);
          uint32_t a0, a1;
          wsum_afrag(in ? wv : 0.f, g, q, a0, a1);
          const uint32_t xr[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
          for (int r4 = 0; r4 < 4; ++r4) {
            float d[4];
            mma_bf16_1688(d, a0, a1, movmatrix_trans(xr[r4]), acc[r4][0], acc[r4][1]);
            acc[r4][0] = d[0] + d[2];
            acc[r4][1] = d[1] + d[3];
          }
        }
        if (g < 4)
          __stcs(reinterpret_cast<uint4*>(yrow + g * 512 + (q + 4 * m) * 16),
                 make_uint4(pack_bf16x2(acc[0][0], acc[0][1]), pack_bf16x2(acc[1][0], acc[1][1]),
                            pack_bf16x2(acc[2][0], acc[2][1]), pack_bf16x2(acc[3][0], acc[3][1])));
      }
    }
    if (a.d_a && lane < GH) store_da(a, r.i, lane, das);
    if (!r.shift()) break;
    cl = cl2; pl = pl2;
  }
}

// ------------------------------------------------------------------------------------------ a = x V^T
// Warp per row, 4 consecutive rows per step: 32 partial dot products reduced with one transposed reduction; lane
// 8r + m ends up with a[row r, m] -> one coalesced 128-byte store per step when lda = 8.
struct RowdotArgs {
  const void* x; uint32_t xrow_bytes;
  const float* V; int ldv;                   // fp32 [8, F]
  float* out; uint32_t ldo;                  // fp32 [N, >= 8]
  uint32_t n_rows;
};
template <typename T, int VPL>
__global__ void __launch_bounds__(256, 2) rowdot8_kernel(const RowdotArgs a) {
  constexpr int VN = Vec<T>::N;
  const int lane = threadIdx.x & 31;
  float vr[8][VPL][VN];                      // this lane's slice of the 8 vectors
#pragma unroll
  for (int m = 0; m < 8; ++m)
#pragma unroll
    for (int v = 0; v < VPL; ++v)
#pragma unroll
      for (int k = 0; k < VN; ++k) vr[m][v][k] = __ldg(a.V + (int64_t)m * a.ldv + (v * 32 + lane) * VN + k);
  const uint32_t warps = gridDim.x * 8u, w = blockIdx.x * 8u + (threadIdx.x >> 5);
  const char* xb = reinterpret_cast<const char*>(a.x) + lane * 16;
  asm volatile("" : "+l"(xb));     // keep the per-lane base in a register pair: one IMAD.WIDE.U32 per gathered row
  for (uint32_t i0 = w * 4u; i0 < a.n_rows; i0 += warps * 4u) {
    uint4 buf[4][VPL];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const uint32_t i = min(i0 + r, a.n_rows - 1u);
#pragma unroll
      for (int v = 0; v < VPL; ++v)
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(buf[r][v].x), "=r"(buf[r][v].y), "=r"(buf[r][v].z), "=r"(buf[r][v].w)
                     : "l"(xb + (uint64_t)i * a.xrow_bytes + 512 * v));
    }
    float part[32];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float p0[8], p1[8];
#pragma unroll
      for (int m = 0; m < 8; ++m) { p0[m] = 0.f; p1[m] = 0.f; }
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        float f[VN];
        unpack_row16(buf[r][v], f, T());
#pragma unroll
        for (int m = 0; m < 8; ++m)
#pragma unroll
          for (int k = 0; k < VN; k += 2) ffma2_mul(p0[m], p1[m], vr[m][v][k], vr[m][v][k + 1], f[k], f[k + 1]);
      }
#pragma unroll
      for (int m = 0; m < 8; ++m) part[r * 8 + m] = p0[m] + p1[m];
    }
    const float red = warp_transpose_sum32(part);
    const uint32_t i = i0 + (lane >> 3);
    if (i < a.n_rows) a.out[(uint64_t)i * a.ldo + (lane & 7)] = red;
  }
}

// a = x V^T for bf16 rows of 512 bytes on mma.sync: 8 rows x 8 vectors per warp step = one m16n8k16 chain over K = 256 with the
// rows loaded straight into A-fragment layout (lane (g, q): row g, pieces q, q + 4, ...; see tz_alpha_mma_kernel) and V resident
// in registers as B fragments, split into bf16 head + bf16 remainder so that the fp32 vectors keep their accuracy (2^-17).
// 16 loads + 32 HMMAs per 8 rows instead of ~70 instructions per row: the kernel streams x at the HBM rate.
__global__ void __launch_bounds__(256, 2) rowdot8_mma_kernel(const RowdotArgs a) {
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  uint32_t bh[16][2], bl[16][2];               // k-step s = 2m + half: features (q + 4m) * 8 + 4 * half + {0, 1 | 2, 3} of vector g
#pragma unroll
  for (int s = 0; s < 16; ++s) {
    const float4 v4 = __ldg(reinterpret_cast<const float4*>(a.V + (int64_t)g * a.ldv + (q + 4 * (s >> 1)) * 8 + 4 * (s & 1)));
    const float f[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const __nv_bfloat16 h0 = __float2bfloat16_rn(f[2 * j]), h1 = __float2bfloat16_rn(f[2 * j + 1]);
      bh[s][j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
      bl[s][j] = pack_bf16x2(f[2 * j] - __bfloat162float(h0), f[2 * j + 1] - __bfloat162float(h1));
    }
  }
  const uint32_t warps = gridDim.x * 8u, w = blockIdx.x * 8u + (threadIdx.x >> 5);
  const char* xb = reinterpret_cast<const char*>(a.x) + q * 16;
  asm volatile("" : "+l"(xb));
  for (uint32_t i0 = w * 8u; i0 < a.n_rows; i0 += warps * 8u) {
    const uint32_t i = min(i0 + (uint32_t)g, a.n_rows - 1u);
    uint4 X[8];
    const char* px = xb + (uint64_t)i * a.xrow_bytes;
#pragma unroll
    for (int m = 0; m < 8; ++m)
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(X[m].x), "=r"(X[m].y), "=r"(X[m].z), "=r"(X[m].w) : "l"(px + 64 * m));
    float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      mma_bf16_16816(c0, X[m].x, X[m].y, bh[2 * m][0], bh[2 * m][1]);
      mma_bf16_16816(c1, X[m].z, X[m].w, bh[2 * m + 1][0], bh[2 * m + 1][1]);
      mma_bf16_16816(c0, X[m].x, X[m].y, bl[2 * m][0], bl[2 * m][1]);
      mma_bf16_16816(c1, X[m].z, X[m].w, bl[2 * m + 1][0], bl[2 * m + 1][1]);
    }
    if (i0 + (uint32_t)g < a.n_rows)          // lane (g, q): a[row g][2q, 2q + 1]
      *reinterpret_cast<float2*>(a.out + (uint64_t)(i0 + g) * a.ldo + 2 * q) = make_float2(c0[0] + c1[0], c0[1] + c1[1]);
  }
}

// ------------------------------------------------------------------------------------------ launchers
template <typename K>
static inline int64_t gatz_blocks(K kernel, const RowSched& ord) {
  const int64_t cap = resident_ctas(kernel, 256);
  return ord.n_chunks < cap ? (int64_t)ord.n_chunks : cap;
}

template <typename T, int VPL>
static int gatz_launch(int which, const GatzArgs& a, cudaStream_t st) {
  if (which == 0) gatz_fwd_kernel<T, VPL><<<(unsigned)gatz_blocks(gatz_fwd_kernel<T, VPL>, a.ord), 256, 0, st>>>(a);
  else if (which == 1) gatz_bwd_dst_kernel<T, VPL, false><<<(unsigned)gatz_blocks(gatz_bwd_dst_kernel<T, VPL, false>, a.ord), 256, 0, st>>>(a);
  else if (which == 2) gatz_bwd_src_kernel<T, VPL><<<(unsigned)gatz_blocks(gatz_bwd_src_kernel<T, VPL>, a.ord), 256, 0, st>>>(a);
  else if (which == 3) tz_fwd_kernel<T, VPL><<<(unsigned)gatz_blocks(tz_fwd_kernel<T, VPL>, a.ord), 256, 0, st>>>(a);
  else gatz_bwd_dst_kernel<T, VPL, true><<<(unsigned)gatz_blocks(gatz_bwd_dst_kernel<T, VPL, true>, a.ord), 256, 0, st>>>(a);
  count_launch();
  return cuda_status();
}

static int gatz_dispatch(int which, int dt, int row_bytes, const GatzArgs& a, cudaStream_t st) {
  if (dt == B2G_F32) return row_bytes == 512 ? gatz_launch<float, 1>(which, a, st) : gatz_launch<float, 2>(which, a, st);
  return row_bytes == 512 ? gatz_launch<__nv_bfloat16, 1>(which, a, st) : gatz_launch<__nv_bfloat16, 2>(which, a, st);
}

static inline int esz(int dt) { return dt == B2G_F32 ? 4 : 2; }
static inline bool fits32(int64_t v) { return v >= 0 && v < (1ll << 32); }
// B2G_ATTN_MMA=0 in the environment at load time keeps the SIMT kernels for bf16 F = 256 (A/B runs); read once, never written
static const int g_attn_mma = [] { const char* e = getenv("B2G_ATTN_MMA"); return (e && e[0] == '0') ? 0 : 1; }();

// ------------------------------------------------------------------------------------------ edge features (edge_dim)
// TransformerConv(edge_dim = 4) in the aggregate-first form: with e_ijh = We_h a_ij (lin_edge, no bias) PyG
// adds e to the keys and to the values.  By linearity  q_ih . e_ijh / sqrt(C) = r_ih . a_ij  with r_ih = We_h^T q_ih / sqrt(C)
// (4 numbers per node and head: 16 more columns of the u GEMM) and  sum_j alpha'_ijh e_ijh = We_h m_ih  with
// m_ih = sum_j alpha'_ijh a_ij (16 more columns of z_aug), so the [E, H*C] edge embedding never exists.  Two thin kernels,
// one thread per target row over its CSR entries (16 bytes of edge attributes per entry, target-major order):
//   edge_dot4 : out[p, h] = v[i(p), 4h .. 4h+3] . ea[p]      (logit term from r; d alpha' term from dm)
//   edge_wsum4: out[i, 4h + c] = sum_p w[p, h] keep(p, h) ea[p, c]   (m from alpha and the dropout mask; dr from de)
__global__ void __launch_bounds__(256) edge_dot4_kernel(const float* __restrict__ v, int64_t ldv,
                                                        const float* __restrict__ ea, const int32_t* __restrict__ rowptr,
                                                        int64_t n, float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = rowptr[i], e = rowptr[i + 1];
    if (b == e) continue;
    float4 r[GH];
#pragma unroll
    for (int h = 0; h < GH; ++h) r[h] = ldg_f4(v + i * ldv + 4 * h);
    for (int p = b; p < e; ++p) {
      const float4 a4 = ldg_f4(ea + (uint64_t)p * 4);
      float o[GH];
#pragma unroll
      for (int h = 0; h < GH; ++h) o[h] = r[h].x * a4.x + r[h].y * a4.y + r[h].z * a4.z + r[h].w * a4.w;
      *reinterpret_cast<float4*>(out + (uint64_t)p * GH) = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) edge_wsum4_kernel(const float* __restrict__ w, const float* __restrict__ ea,
                                                         const int32_t* __restrict__ rowptr, int64_t n, float p_drop,
                                                         uint64_t seed, const uint64_t* epoch, char* __restrict__ out,
                                                         int64_t orow_bytes) {
  constexpr int VN = Vec<T>::N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float m[GH * 4];
#pragma unroll
    for (int k = 0; k < GH * 4; ++k) m[k] = 0.f;
    const int b = rowptr[i], e = rowptr[i + 1];
    for (int p = b; p < e; ++p) {
      const float4 w4 = ldg_f4(w + (uint64_t)p * GH);
      const float4 a4 = ldg_f4(ea + (uint64_t)p * 4);
      float ws[GH] = {w4.x, w4.y, w4.z, w4.w};
      if (p_drop > 0.f) {
        float sc[4];
        dropout_scale4(mix_epoch(seed, epoch), (uint64_t)p, p_drop, sc);
#pragma unroll
        for (int h = 0; h < GH; ++h) ws[h] *= sc[h];
      }
#pragma unroll
      for (int h = 0; h < GH; ++h) {
        m[4 * h + 0] += ws[h] * a4.x; m[4 * h + 1] += ws[h] * a4.y;
        m[4 * h + 2] += ws[h] * a4.z; m[4 * h + 3] += ws[h] * a4.w;
      }
    }
    char* o = out + i * orow_bytes;
#pragma unroll
    for (int k = 0; k < GH * 4; k += VN) {
      Vec<T> ov;
      ov.from_float(m + k);
      *reinterpret_cast<uint4*>(o + k * sizeof(T)) = *reinterpret_cast<uint4*>(&ov.v);
    }
  }
}

}  // namespace b2g

using namespace b2g;

extern "C" {

int b2g_gatz_supported(int64_t n, int H, int F, int dt) {
  if (dt != B2G_F32 && dt != B2G_BF16) return 0;
  const int64_t rb = (int64_t)F * esz(dt);       // bf16 rows of 1 KB would need 64 accumulator registers per head set: not built
  const bool shape = (dt == B2G_BF16) ? rb == 512 : (rb == 512 || rb == 1024);
  return (H == GH && shape && n >= 1 && n < (1ll << 32) - (1ll << 25)) ? 1 : 0;
}

int b2g_rowdot8(const void* x, int64_t ldx, const float* V, int64_t ldv, float* out, int64_t ldo, int64_t n, int F,
                int dt, void* stream) {
  if (n < 0 || !b2g_gatz_supported(n > 0 ? n : 1, GH, F, dt)) return n < 0 ? B2G_E_ARG : B2G_E_UNSUPPORTED;
  if (n == 0) return B2G_OK;
  if (!x || !V || !out || !aligned16(x) || (ldx * esz(dt)) % 16 || !fits32(ldx * esz(dt)) || ldo < 8 || !fits32(ldo)) return B2G_E_ARG;
  RowdotArgs a{x, (uint32_t)(ldx * esz(dt)), V, (int)ldv, out, (uint32_t)ldo, (uint32_t)n};
  const int rb = F * esz(dt);
  cudaStream_t st = (cudaStream_t)stream;
  if (g_attn_mma && dt == B2G_BF16 && rb == 512 && ldv % 4 == 0 && aligned16(V) && ldo % 2 == 0 && ((uintptr_t)out % 8) == 0) {
    int64_t blocks = resident_ctas(rowdot8_mma_kernel, 256);
    const int64_t want8 = ceil_div(n, 64);
    if (want8 < blocks) blocks = want8;
    rowdot8_mma_kernel<<<(unsigned)blocks, 256, 0, st>>>(a);
    count_launch();
    return cuda_status();
  }
  const int64_t want = ceil_div(n, 32);
#define B2G_RD(T, VPL)                                                                   \
  {                                                                                      \
    int64_t blocks = resident_ctas(rowdot8_kernel<T, VPL>, 256);                         \
    if (want < blocks) blocks = want;                                                    \
    rowdot8_kernel<T, VPL><<<(unsigned)blocks, 256, 0, st>>>(a);                         \
  }
  if (dt == B2G_F32) { if (rb == 512) B2G_RD(float, 1) else B2G_RD(float, 2) }
  else { if (rb == 512) B2G_RD(__nv_bfloat16, 1) else B2G_RD(__nv_bfloat16, 2) }
#undef B2G_RD
  count_launch();
  return cuda_status();
}

static int gatz_common(GatzArgs& a, int64_t n, int F, int dt, int64_t band) {
  if (!b2g_gatz_supported(n, GH, F, dt)) return B2G_E_UNSUPPORTED;
  if (!make_row_sched(n, band, a.ord)) return B2G_E_UNSUPPORTED;
  a.n_rows = (uint32_t)n;
  return B2G_OK;
}

int b2g_gatz_fwd(const void* x, int64_t ldx, const float* a_srcdst, int64_t lda, void* z, int64_t ldz, int64_t n, int H,
                 int F, int dt, float slope, const int32_t* rowptr, const int32_t* col, float* smax, float* ssum,
                 float p_drop, uint64_t seed, int64_t band, void* stream) {
  if (n < 0 || H != GH) return n < 0 ? B2G_E_ARG : B2G_E_UNSUPPORTED;
  if (n == 0) return B2G_OK;
  const int es = esz(dt);
  if (!x || !a_srcdst || !z || !rowptr || !col || (smax == nullptr) != (ssum == nullptr)) return B2G_E_ARG;
  if (!aligned16(x) || !aligned16(z) || !aligned16(a_srcdst) || (ldx * es) % 16 || (ldz * es) % 16 || lda % 4 || lda < 2 * GH)
    return B2G_E_ALIGN;
  if (!fits32(ldx * es) || !fits32(ldz * es) || !fits32(lda)) return B2G_E_SHAPE;
  GatzArgs a{};
  const int rc = gatz_common(a, n, F, dt, band);
  if (rc) return rc;
  a.x = x; a.xrow_bytes = (uint32_t)(ldx * es); a.a = a_srcdst; a.lda = (uint32_t)lda; a.z = z; a.zrow_bytes = (uint32_t)(ldz * es);
  a.rowptr = rowptr; a.col = col; a.smax = smax; a.ssum = ssum; a.slope = slope; a.p_drop = p_drop; a.seed = seed; a.epoch = dropout_epoch_ptr();
  return gatz_dispatch(0, dt, F * es, a, (cudaStream_t)stream);
}

int b2g_gatz_bwd_dst(const void* x, int64_t ldx, const float* a_srcdst, int64_t lda, const void* dz, int64_t lddz,
                     int64_t n, int H, int F, int dt, float slope, const int32_t* rowptr, const int32_t* col,
                     const float* smax, const float* ssum, float p_drop, uint64_t seed, float* alpha_e, float* de_e,
                     void* d_a, int64_t ldda, int d_a_dt, const float* edge_bias, int64_t band, int64_t max_row_len,
                     void* stream) {
  if (n < 0 || H != GH) return n < 0 ? B2G_E_ARG : B2G_E_UNSUPPORTED;
  if (n == 0) return B2G_OK;
  const int es = esz(dt);
  if (d_a_dt != B2G_F32 && d_a_dt != B2G_BF16) return B2G_E_ARG;
  if (edge_bias && !aligned16(edge_bias)) return B2G_E_ALIGN;
  if (!x || !a_srcdst || !dz || !rowptr || !col || !smax || !ssum || !alpha_e || !de_e || !d_a) return B2G_E_ARG;
  if (!aligned16(x) || !aligned16(dz) || !aligned16(a_srcdst) || !aligned16(smax) || !aligned16(ssum) || !aligned16(alpha_e) ||
      !aligned16(de_e) || (ldx * es) % 16 || (lddz * es) % 16 || lda % 4 || lda < 2 * GH || ldda < 2 * GH)
    return B2G_E_ALIGN;
  if (!fits32(ldx * es) || !fits32(lddz * es) || !fits32(lda) || !fits32(ldda)) return B2G_E_SHAPE;
  GatzArgs a{};
  const int rc = gatz_common(a, n, F, dt, band);
  if (rc) return rc;
  a.x = x; a.xrow_bytes = (uint32_t)(ldx * es); a.a = a_srcdst; a.lda = (uint32_t)lda; a.dz = dz; a.dzrow_bytes = (uint32_t)(lddz * es);
  a.rowptr = rowptr; a.col = col; a.smax = const_cast<float*>(smax); a.ssum = const_cast<float*>(ssum);
  a.slope = slope; a.p_drop = p_drop; a.seed = seed; a.epoch = dropout_epoch_ptr(); a.alpha_e = alpha_e; a.de_e = de_e; a.d_a = static_cast<float*>(d_a); a.ldda = (uint32_t)ldda;
  a.da_bf16 = d_a_dt == B2G_BF16;
  a.ebias = edge_bias;
  if (g_attn_mma && dt == B2G_BF16 && F * es == 512) {              // rows of <= 8 entries on the tensor cores, the rest after
    gatz_bwd_dst_mma_kernel<false><<<(unsigned)gatz_blocks(gatz_bwd_dst_mma_kernel<false>, a.ord), 256, 0, (cudaStream_t)stream>>>(a);
    count_launch();
    if (max_row_len > 0 && max_row_len <= 8) return cuda_status();
    a.only_long = 1;
  }
  return gatz_dispatch(1, dt, F * es, a, (cudaStream_t)stream);
}

int b2g_gatz_bwd_src(const void* g, int64_t ldg, const float* alpha_e, const float* de_e, void* y, int64_t ldy,
                     void* d_a, int64_t ldda, int d_a_dt, int64_t n, int H, int C, int dt, const int32_t* rowptr_t,
                     const int32_t* col_t, const int32_t* perm, int64_t band, void* stream) {
  if (n < 0 || H != GH) return n < 0 ? B2G_E_ARG : B2G_E_UNSUPPORTED;
  if (n == 0) return B2G_OK;
  const int es = esz(dt);
  if (d_a && d_a_dt != B2G_F32 && d_a_dt != B2G_BF16) return B2G_E_ARG;
  if (!g || !alpha_e || !de_e || !y || !rowptr_t || !col_t) return B2G_E_ARG;     // perm NULL = identity, d_a NULL = not wanted
  if (!aligned16(g) || !aligned16(y) || !aligned16(alpha_e) || !aligned16(de_e) || (ldg * es) % 16 || (ldy * es) % 16 ||
      (d_a && ldda < GH))
    return B2G_E_ALIGN;
  if (!fits32(ldg * es) || !fits32(ldy * es) || !fits32(ldda)) return B2G_E_SHAPE;
  GatzArgs a{};
  const int rc = gatz_common(a, n, C, dt, band);
  if (rc) return rc;
  a.x = g; a.xrow_bytes = (uint32_t)(ldg * es); a.z = y; a.zrow_bytes = (uint32_t)(ldy * es);
  a.rowptr = rowptr_t; a.col = col_t; a.perm = perm; a.alpha_e = const_cast<float*>(alpha_e); a.de_e = const_cast<float*>(de_e);
  a.d_a = static_cast<float*>(d_a); a.ldda = (uint32_t)ldda; a.da_bf16 = d_a && d_a_dt == B2G_BF16;
  if (g_attn_mma && dt == B2G_BF16 && C * es == 512) {              // weighted sums as m16n8k8 tile products
    gatz_bwd_src_mma_kernel<<<(unsigned)gatz_blocks(gatz_bwd_src_mma_kernel, a.ord), 256, 0, (cudaStream_t)stream>>>(a);
    count_launch();
    return cuda_status();
  }
  return gatz_dispatch(2, dt, C * es, a, (cudaStream_t)stream);
}

/* Edge-feature terms of TransformerConv(edge_dim = 4), aggregate-first (see edge_dot4_kernel).  ea_csr: fp32 [nnz, 4], the
 * edge attributes in target-major CSR order; v: fp32 [n, ldv >= 4H]; out of edge_dot4: fp32 [nnz, H]; out of edge_wsum4:
 * [n, 4H] of dtype dt with row stride ldo elements (w: fp32 [nnz, H]; p_drop > 0 applies the attention-dropout mask of
 * (seed, device epoch), the one b2g_tz_fwd draws). */
int b2g_edge_dot4(const float* v, int64_t ldv, const float* ea_csr, const int32_t* rowptr, int64_t n, int H, float* out,
                  void* stream) {
  if (n < 0 || H != GH) return n < 0 ? B2G_E_ARG : B2G_E_UNSUPPORTED;
  if (n == 0) return B2G_OK;
  if (!v || !ea_csr || !rowptr || !out) return B2G_E_ARG;
  if (!aligned16(v) || !aligned16(ea_csr) || !aligned16(out) || ldv % 4 || (ldv != 0 && ldv < 4 * GH)) return B2G_E_ALIGN;
  const int64_t blocks = std::min<int64_t>(ceil_div(n, 256), (int64_t)B2G_NUM_SMS * 16);   // ldv == 0: one v row for all nodes
  edge_dot4_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(v, ldv, ea_csr, rowptr, n, out);
  count_launch();
  return cuda_status();
}

int b2g_edge_wsum4(const float* w, const float* ea_csr, const int32_t* rowptr, int64_t n, int H, float p_drop, uint64_t seed,
                   void* out, int64_t ldo, int dt, void* stream) {
  if (n < 0 || H != GH) return n < 0 ? B2G_E_ARG : B2G_E_UNSUPPORTED;
  if (n == 0) return B2G_OK;
  const int es = esz(dt);
  if (!w || !ea_csr || !rowptr || !out) return B2G_E_ARG;
  if (!aligned16(w) || !aligned16(ea_csr) || !aligned16(out) || (ldo * es) % 16 || ldo < 4 * GH) return B2G_E_ALIGN;
  const int64_t blocks = std::min<int64_t>(ceil_div(n, 256), (int64_t)B2G_NUM_SMS * 16);
  cudaStream_t st = (cudaStream_t)stream;
  if (dt == B2G_F32)
    edge_wsum4_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(w, ea_csr, rowptr, n, p_drop, seed, dropout_epoch_ptr(),
                                                                static_cast<char*>(out), ldo * es);
  else
    edge_wsum4_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(w, ea_csr, rowptr, n, p_drop, seed, dropout_epoch_ptr(),
                                                                        static_cast<char*>(out), ldo * es);
  count_launch();
  return cuda_status();
}

/* TransformerConv, aggregate-first (see tz_fwd_kernel).  u [n, H*F] = x Mq + cq; z_aug [n, H*F + 8 + F]. */
int b2g_tz_fwd(const void* x, int64_t ldx, const void* x_self, const void* u, int64_t ldu, void* z_aug, int64_t ldz, int64_t n,
               int H, int F, int dt, const int32_t* rowptr, const int32_t* col, float* alpha_e, const float* edge_bias,
               float p_drop, uint64_t seed, int64_t band, void* stream) {
  if (n < 0 || H != GH) return n < 0 ? B2G_E_ARG : B2G_E_UNSUPPORTED;
  if (n == 0) return B2G_OK;
  const int es = esz(dt);
  if (!x || !u || !z_aug || !rowptr || !col) return B2G_E_ARG;
  if (!aligned16(x) || !aligned16(u) || !aligned16(z_aug) || (alpha_e && !aligned16(alpha_e)) ||
      (edge_bias && !aligned16(edge_bias)) || (ldx * es) % 16 ||
      (ldu * es) % 16 || (ldz * es) % 16 || ldz < (int64_t)H * F + 8 + F)
    return B2G_E_ALIGN;
  if (!fits32(ldx * es) || !fits32(ldu * es) || !fits32(ldz * es)) return B2G_E_SHAPE;
  GatzArgs a{};
  const int rc = gatz_common(a, n, F, dt, band);
  if (rc) return rc;
  if (x_self && !aligned16(x_self)) return B2G_E_ALIGN;
  a.x = x; a.x_self = x_self ? x_self : x; a.xrow_bytes = (uint32_t)(ldx * es); a.dz = u; a.dzrow_bytes = (uint32_t)(ldu * es); a.z = z_aug;
  a.zrow_bytes = (uint32_t)(ldz * es); a.rowptr = rowptr; a.col = col; a.alpha_e = alpha_e; a.p_drop = p_drop; a.seed = seed; a.epoch = dropout_epoch_ptr();
  a.ebias = edge_bias;
  return gatz_dispatch(3, dt, F * es, a, (cudaStream_t)stream);
}

/* Attention weights only (fused TransformerConv forward, gat_fused.cu): alpha_pre [nnz, H] = softmax_j(u_ih . x_j (+ edge_bias)),
 * alpha_post (NULL when p_drop == 0: identical) = alpha_pre with the attention-dropout keep scale, ssum [n, H] = per-head sums of
 * the post-dropout weights.  impl 0: bf16 rows of 512 bytes take tz_alpha_mma_kernel (logits by mma.sync), everything else
 * b2g_tz_fwd's kernel without the weighted sums; impl 1: always the latter (A/B runs, tests). */
int b2g_tz_alpha(const void* x, int64_t ldx, const void* u, int64_t ldu, int64_t n, int H, int F, int dt, const int32_t* rowptr,
                 const int32_t* col, float* alpha_pre, float* alpha_post, float* ssum, const float* edge_bias, float p_drop,
                 uint64_t seed, int64_t band, int impl, void* stream) {
  if (n < 0 || H != GH) return n < 0 ? B2G_E_ARG : B2G_E_UNSUPPORTED;
  if (n == 0) return B2G_OK;
  const int es = esz(dt);
  if (!x || !u || !rowptr || !col || !alpha_pre || !ssum || (p_drop > 0.f && !alpha_post)) return B2G_E_ARG;
  if (!aligned16(x) || !aligned16(u) || !aligned16(alpha_pre) || (alpha_post && !aligned16(alpha_post)) || !aligned16(ssum) ||
      (edge_bias && !aligned16(edge_bias)) || (ldx * es) % 16 || (ldu * es) % 16)
    return B2G_E_ALIGN;
  if (!fits32(ldx * es) || !fits32(ldu * es)) return B2G_E_SHAPE;
  GatzArgs a{};
  const int rc = gatz_common(a, n, F, dt, band);
  if (rc) return rc;
  a.x = x; a.x_self = x; a.xrow_bytes = (uint32_t)(ldx * es); a.dz = u; a.dzrow_bytes = (uint32_t)(ldu * es);
  a.rowptr = rowptr; a.col = col; a.alpha_e = alpha_pre; a.de_e = p_drop > 0.f ? alpha_post : nullptr; a.smax = ssum;
  a.p_drop = p_drop; a.seed = seed; a.epoch = dropout_epoch_ptr(); a.ebias = edge_bias; a.alpha_only = 1;
  if (impl != 1 && dt == B2G_BF16 && F * es == 512) {               // logits on the tensor cores (tz_alpha_mma_kernel)
    tz_alpha_mma_kernel<<<(unsigned)gatz_blocks(tz_alpha_mma_kernel, a.ord), 256, 0, (cudaStream_t)stream>>>(a);
    count_launch();
    return cuda_status();
  }
  return gatz_dispatch(3, dt, F * es, a, (cudaStream_t)stream);
}

/* Target side of the TransformerConv backward pass: dz_aug [n, >= H*F + 8] (columns H*F..H*F+3 = d s_alpha), alpha_in =
 * the forward pass's alpha [nnz,H]; writes alpha_e (after dropout) and de_e [nnz,H] in target-major CSR order. */
int b2g_tz_bwd_dst(const void* x, int64_t ldx, const void* dz_aug, int64_t lddz, const float* alpha_in, int64_t n, int H,
                   int F, int dt, const int32_t* rowptr, const int32_t* col, float p_drop, uint64_t seed, float* alpha_e,
                   float* de_e, void* du, int64_t lddu, const float* edge_bias, int64_t band, int64_t max_row_len, void* stream) {
  if (n < 0 || H != GH) return n < 0 ? B2G_E_ARG : B2G_E_UNSUPPORTED;
  if (n == 0) return B2G_OK;
  const int es = esz(dt);
  if (!x || !dz_aug || !alpha_in || !rowptr || !col || !alpha_e || !de_e) return B2G_E_ARG;
  if (!aligned16(x) || !aligned16(dz_aug) || !aligned16(alpha_in) || !aligned16(alpha_e) || !aligned16(de_e) ||
      (edge_bias && !aligned16(edge_bias)) || (ldx * es) % 16 || (lddz * es) % 16 || lddz < (int64_t)H * F + 8)
    return B2G_E_ALIGN;
  if (!fits32(ldx * es) || !fits32(lddz * es)) return B2G_E_SHAPE;
  GatzArgs a{};
  const int rc = gatz_common(a, n, F, dt, band);
  if (rc) return rc;
  a.x = x; a.xrow_bytes = (uint32_t)(ldx * es); a.dz = dz_aug; a.dzrow_bytes = (uint32_t)(lddz * es); a.alpha_in = alpha_in;
  a.rowptr = rowptr; a.col = col; a.p_drop = p_drop; a.seed = seed; a.epoch = dropout_epoch_ptr(); a.alpha_e = alpha_e; a.de_e = de_e;
  a.ebias = edge_bias;
  if (du) {
    if (!aligned16(du) || (lddu * es) % 16 || lddu < (int64_t)H * F) return B2G_E_ALIGN;
    if (!fits32(lddu * es)) return B2G_E_SHAPE;
    a.z = du; a.zrow_bytes = (uint32_t)(lddu * es);
  }
  if (g_attn_mma && dt == B2G_BF16 && F * es == 512) {
    gatz_bwd_dst_mma_kernel<true><<<(unsigned)gatz_blocks(gatz_bwd_dst_mma_kernel<true>, a.ord), 256, 0, (cudaStream_t)stream>>>(a);
    count_launch();
    if (max_row_len > 0 && max_row_len <= 8) return cuda_status();
    a.only_long = 1;
  }
  return gatz_dispatch(4, dt, F * es, a, (cudaStream_t)stream);
}

}  // extern "C"
