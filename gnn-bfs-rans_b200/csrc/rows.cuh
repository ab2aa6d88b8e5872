// rows.cuh — pieces shared by the warp-per-row gather kernels (aggregate_rows.cu, gat_rows.cu): the row schedule
// (narrow front, panel order on band-structured meshes), 16-byte row loads and the accumulate primitives.
#pragma once
#include "common.cuh"

namespace b2g {

// Row schedule.  The co-resident CTAs take chunks of `chunk_rows` consecutive rows round-robin, so the chip works on
// ONE front of grid x chunk_rows rows.  A mesh numbered plane by plane has neighbours at index distance ~B (the
// "band": 50 000 rows = 25.6 MB of bf16 features at cfg4); in a linear sweep the three uses of a row are 2B rows apart
// and - with the streamed output and the far-die copies - do not meet in L2.  Panel order: split every band-sized
// block [kB, (k+1)B) into panels of `panel` rows and sweep panel p of ALL blocks before panel p+1; the +-B
// neighbours of a row are then `panel` rows away in processing order.  band = 0 selects the linear order.
struct RowSched {
  uint32_t n_chunks, chunk_rows;
  uint32_t band, panel, per_panel, cpp_shift;        // band, panel: multiples of chunk_rows; per_panel = blocks << cpp_shift
  // first row and row count of chunk q; n_rows < 2^32 - 2^25 (checked by the launcher) keeps everything in 32 bits
  __device__ __forceinline__ uint32_t chunk(uint32_t q, uint32_t n_rows, uint32_t& rows) const {
    if (band == 0) {
      const uint32_t c0 = q * chunk_rows;
      rows = min(chunk_rows, n_rows - c0);
      return c0;
    }
    const uint32_t p = q / per_panel, rem = q - p * per_panel;
    const uint32_t k = rem >> cpp_shift, tc = rem & ((1u << cpp_shift) - 1u);
    const uint32_t off = p * panel + tc * chunk_rows;          // offset inside the block
    const uint32_t c0 = k * band + off;
    rows = 0;
    if (off < band && c0 < n_rows) rows = min(min(chunk_rows, band - off), n_rows - c0);
    return c0;
  }
};
constexpr int ROWS_CHUNK_DEFAULT = 32;      // rows per CTA step
constexpr int ROWS_PANEL_DEFAULT = 8192;    // rows per panel of the band order (power-of-two multiple of the chunk)

// chunk_rows / panel_rows <= 0 select the defaults; per call (b2g_seg_sum_tuned), no library-wide state
static inline bool make_row_sched(int64_t n_rows, int64_t band, RowSched& o, int chunk_rows = 0, int panel_rows = 0) {
  o = RowSched{};
  if (chunk_rows <= 0) chunk_rows = ROWS_CHUNK_DEFAULT;
  if (panel_rows <= 0) panel_rows = ROWS_PANEL_DEFAULT;
  if (chunk_rows < 8 || (chunk_rows & (chunk_rows - 1)) || chunk_rows > 4096) return false;
  if (panel_rows < chunk_rows || (panel_rows & (panel_rows - 1)) || panel_rows > (1 << 24)) return false;
  o.chunk_rows = (uint32_t)chunk_rows;
  const int64_t panel = panel_rows;
  if (band < 4 * panel || band * 2 > n_rows) {               // narrow band (already L2 friendly) or no band structure
    const int64_t nc = ceil_div(n_rows, o.chunk_rows);
    if (nc >= (1ll << 31)) return false;
    o.n_chunks = (uint32_t)nc;
    return true;
  }
  const int64_t bandr = ceil_div(band, o.chunk_rows) * o.chunk_rows;   // block >= band keeps +-band neighbours in adjacent blocks
  const int64_t blocks = ceil_div(n_rows, bandr);
  uint32_t sh = 0;
  while (((int64_t)o.chunk_rows << sh) < panel) ++sh;
  const int64_t per_panel = blocks << sh;
  const int64_t nc = ceil_div(bandr, panel) * per_panel;
  if (nc >= (1ll << 31) || bandr >= (1ll << 31)) return false;
  o.band = (uint32_t)bandr;
  o.panel = (uint32_t)panel;
  o.cpp_shift = sh;
  o.per_panel = (uint32_t)per_panel;
  o.n_chunks = (uint32_t)nc;
  return true;
}

__device__ __forceinline__ uint4 ldg_row16(const char* p) {   // gathered rows: allocate in L1 (x+-1 / self reuse inside a CTA)
  uint4 u;
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "l"(p));
  return u;
}
// acc += v, unweighted: bf16 -> sm_100 mixed-precision add (SASS FHADD.BF16; exact: bf16 -> fp32 is exact, one fp32
// rounding per add), fp32 -> packed add.rn.f32x2
__device__ __forceinline__ void add_row16(float* acc, const uint4& u, __nv_bfloat16) {
  asm volatile(
      "{\n"
      ".reg .b16 l0, h0, l1, h1, l2, h2, l3, h3;\n"
      "mov.b32 {l0, h0}, %8;\n mov.b32 {l1, h1}, %9;\n mov.b32 {l2, h2}, %10;\n mov.b32 {l3, h3}, %11;\n"
      "add.rn.f32.bf16 %0, l0, %0;\n add.rn.f32.bf16 %1, h0, %1;\n"
      "add.rn.f32.bf16 %2, l1, %2;\n add.rn.f32.bf16 %3, h1, %3;\n"
      "add.rn.f32.bf16 %4, l2, %4;\n add.rn.f32.bf16 %5, h2, %5;\n"
      "add.rn.f32.bf16 %6, l3, %6;\n add.rn.f32.bf16 %7, h3, %7;\n"
      "}\n"
      : "+f"(acc[0]), "+f"(acc[1]), "+f"(acc[2]), "+f"(acc[3]), "+f"(acc[4]), "+f"(acc[5]), "+f"(acc[6]), "+f"(acc[7])
      : "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w));
}
__device__ __forceinline__ void add_row16(float* acc, const uint4& u, float) {
  Vec<float> t;
  t.v = *reinterpret_cast<const float4*>(&u);
  add_vec(acc, t);
}
template <typename T>
__device__ __forceinline__ void fma_row16(float* acc, float w, const uint4& u) {
  Vec<T> t;
  t.v = *reinterpret_cast<const decltype(t.v)*>(&u);
  fma_vec(acc, w, t);
}


// lane l receives sum over lanes of v[l]: 31 shuffles for 32 values (fold the lane space in halves while halving the
// number of values each lane carries) instead of 32 x 5.
__device__ __forceinline__ float warp_transpose_sum32(float (&v)[32]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const bool hi = lane & 16;
    const float send = hi ? v[k] : v[k + 16], keep = hi ? v[k + 16] : v[k];
    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const bool hi = lane & 8;
    const float send = hi ? v[k] : v[k + 8], keep = hi ? v[k + 8] : v[k];
    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const bool hi = lane & 4;
    const float send = hi ? v[k] : v[k + 4], keep = hi ? v[k + 4] : v[k];
    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const bool hi = lane & 2;
    const float send = hi ? v[k] : v[k + 2], keep = hi ? v[k + 2] : v[k];
    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  {
    const bool hi = lane & 1;
    const float send = hi ? v[0] : v[1], keep = hi ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
  }
  return v[0];
}

// 16 bytes of T as fp32 lanes
__device__ __forceinline__ void unpack_row16(const uint4& u, float (&f)[8], __nv_bfloat16) {
  f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
  f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
  f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xffff0000u);
  f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xffff0000u);
}
__device__ __forceinline__ void unpack_row16(const uint4& u, float (&f)[4], float) {
  f[0] = __uint_as_float(u.x); f[1] = __uint_as_float(u.y); f[2] = __uint_as_float(u.z); f[3] = __uint_as_float(u.w);
}
// (a0, a1) += w * (f0, f1): one packed FFMA2
__device__ __forceinline__ void ffma2_acc(float& a0, float& a1, float w, float f0, float f1) {
  asm("{\n .reg .b64 a, f, ww;\n mov.b64 a, {%0, %1};\n mov.b64 f, {%3, %4};\n mov.b64 ww, {%2, %2};\n"
      " fma.rn.f32x2 a, ww, f, a;\n mov.b64 {%0, %1}, a;\n}"
      : "+f"(a0), "+f"(a1) : "f"(w), "f"(f0), "f"(f1));
}
// (a0, a1) += (x0, x1) * (y0, y1)
__device__ __forceinline__ void ffma2_mul(float& a0, float& a1, float x0, float x1, float y0, float y1) {
  asm("{\n .reg .b64 a, x, y;\n mov.b64 a, {%0, %1};\n mov.b64 x, {%2, %3};\n mov.b64 y, {%4, %5};\n"
      " fma.rn.f32x2 a, x, y, a;\n mov.b64 {%0, %1}, a;\n}"
      : "+f"(a0), "+f"(a1) : "f"(x0), "f"(x1), "f"(y0), "f"(y1));
}

}  // namespace b2g
