// common.cuh — shared device/host helpers for libb2g (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/b2g.h"

#define B2G_NUM_SMS 148  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

namespace b2g {

extern std::atomic<int64_t> g_launches;  // defined in api.cu: a diagnostic counter (b2g_launch_count), no effect on results
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

inline int cuda_status() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

// slot of the calling thread's current device in per-device lazily-initialised tables (kernel attribute opt-ins)
inline int current_device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { (void)cudaGetLastError(); dev = 0; }
  return dev;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Grid for a grid-stride kernel whose CTAs must ALL be co-resident: 148 SMs x (CTAs that fit per SM).
// A larger grid would run in waves, each wave sweeping the whole matrix on its own, and the
// neighbour-row reuse the gather kernels get from L2 would be lost (measured: 3.5x DRAM re-reads).
template <typename Kernel>
inline int64_t resident_ctas(Kernel kernel, int threads, size_t dyn_smem = 0) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, dyn_smem) != cudaSuccess || per_sm < 1) {
    (void)cudaGetLastError();
    per_sm = 1;
  }
  return (int64_t)per_sm * B2G_NUM_SMS;
}

// ---------------------------------------------------------------- 16-byte vector of features
template <typename T>
struct Vec;  // 16 bytes of T, convertible to/from fp32 lanes
template <>
struct Vec<float> {
  static constexpr int N = 4;
  float4 v;
  __device__ __forceinline__ void to_float(float* f) const { f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w; }
  __device__ __forceinline__ void from_float(const float* f) { v = make_float4(f[0], f[1], f[2], f[3]); }
};
template <>
struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  uint4 v;
  __device__ __forceinline__ void to_float(float* f) const {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // bf16 -> fp32 is a 16-bit shift
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ __forceinline__ void from_float(const float* f) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 p = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&p);
    }
    v = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

// Gather loads and their consumers (fma_vec) are BOTH `asm volatile`: volatile asm statements keep their
// relative order, so a kernel that writes "U loads, then U fma_vec" really gets U rows in flight per lane.
// (With plain C++ consumers the optimiser sank each load next to its use: load -> use -> load -> use, one
// row in flight per warp — measured with ncu as long-scoreboard stalls on every buffer and a 3x slowdown.)
// ldg_vec   : streaming data, read once (bypass L1 allocation)
// ldg_vec_l1: gathered neighbour rows; adjacent targets of a CTA share x+-1 / self rows -> allocate in L1
template <typename T>
__device__ __forceinline__ Vec<T> ldg_vec(const T* p) {
  Vec<T> r;
  uint4 u;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w)
               : "l"(p));
  r.v = *reinterpret_cast<decltype(r.v)*>(&u);
  return r;
}
template <typename T>
__device__ __forceinline__ Vec<T> ldg_vec_l1(const T* p) {
  Vec<T> r;
  uint4 u;
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w)
               : "l"(p));
  r.v = *reinterpret_cast<decltype(r.v)*>(&u);
  return r;
}
__device__ __forceinline__ uint32_t first_word(const Vec<float>& v) { return __float_as_uint(v.v.x); }
__device__ __forceinline__ uint32_t first_word(const Vec<__nv_bfloat16>& v) { return v.v.x; }
__device__ __forceinline__ int ldg_i32_ordered(const int32_t* p) {   // index load that stays in program order
  int v;
  asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
template <typename T>
__device__ __forceinline__ void stg_vec(T* p, const Vec<T>& r) {
  const uint4 u = *reinterpret_cast<const uint4*>(&r.v);
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(u.x), "r"(u.y),
               "r"(u.z), "r"(u.w)
               : "memory");
}

// Packed fp32 FMA (sm_100 fma.rn.f32x2): (a0,a1) += w * (f0,f1) in one issue slot.  Same rounding as two FFMAs.
__device__ __forceinline__ void ffma2(float& a0, float& a1, float w, float f0, float f1) {
  asm("{\n"
      ".reg .b64 ra, rw, rf;\n"
      "mov.b64 ra, {%0, %1};\n"
      "mov.b64 rw, {%2, %2};\n"
      "mov.b64 rf, {%3, %4};\n"
      "fma.rn.f32x2 ra, rw, rf, ra;\n"
      "mov.b64 {%0, %1}, ra;\n"
      "}\n"
      : "+f"(a0), "+f"(a1)
      : "f"(w), "f"(f0), "f"(f1));
}
// acc[0..N) += w * (16-byte vector of T), fp32 accumulation, packed fma.rn.f32x2.  One volatile asm block
// (conversion included) so it cannot be hoisted in between the gather loads that precede it.
__device__ __forceinline__ void fma_vec(float* acc, float w, const Vec<float>& v) {
  asm volatile(
      "{\n"
      ".reg .b64 ww, f0, f1, a0, a1;\n"
      "mov.b64 ww, {%4, %4};\n"
      "mov.b64 f0, {%5, %6};\n"
      "mov.b64 f1, {%7, %8};\n"
      "mov.b64 a0, {%0, %1};\n"
      "mov.b64 a1, {%2, %3};\n"
      "fma.rn.f32x2 a0, ww, f0, a0;\n"
      "fma.rn.f32x2 a1, ww, f1, a1;\n"
      "mov.b64 {%0, %1}, a0;\n"
      "mov.b64 {%2, %3}, a1;\n"
      "}\n"
      : "+f"(acc[0]), "+f"(acc[1]), "+f"(acc[2]), "+f"(acc[3])
      : "f"(w), "f"(v.v.x), "f"(v.v.y), "f"(v.v.z), "f"(v.v.w));
}
__device__ __forceinline__ void fma_vec(float* acc, float w, const Vec<__nv_bfloat16>& v) {
  // bf16 pair -> fp32 pair: low half shifted up, high half masked
  asm volatile(
      "{\n"
      ".reg .b32 l0, h0, l1, h1, l2, h2, l3, h3;\n"
      ".reg .b64 ww, f0, f1, f2, f3, a0, a1, a2, a3;\n"
      "shl.b32 l0, %9, 16;\n  and.b32 h0, %9, 0xffff0000;\n"
      "shl.b32 l1, %10, 16;\n and.b32 h1, %10, 0xffff0000;\n"
      "shl.b32 l2, %11, 16;\n and.b32 h2, %11, 0xffff0000;\n"
      "shl.b32 l3, %12, 16;\n and.b32 h3, %12, 0xffff0000;\n"
      "mov.b64 ww, {%8, %8};\n"
      "mov.b64 f0, {l0, h0};\n mov.b64 f1, {l1, h1};\n mov.b64 f2, {l2, h2};\n mov.b64 f3, {l3, h3};\n"
      "mov.b64 a0, {%0, %1};\n mov.b64 a1, {%2, %3};\n mov.b64 a2, {%4, %5};\n mov.b64 a3, {%6, %7};\n"
      "fma.rn.f32x2 a0, ww, f0, a0;\n"
      "fma.rn.f32x2 a1, ww, f1, a1;\n"
      "fma.rn.f32x2 a2, ww, f2, a2;\n"
      "fma.rn.f32x2 a3, ww, f3, a3;\n"
      "mov.b64 {%0, %1}, a0;\n mov.b64 {%2, %3}, a1;\n mov.b64 {%4, %5}, a2;\n mov.b64 {%6, %7}, a3;\n"
      "}\n"
      : "+f"(acc[0]), "+f"(acc[1]), "+f"(acc[2]), "+f"(acc[3]), "+f"(acc[4]), "+f"(acc[5]), "+f"(acc[6]), "+f"(acc[7])
      : "f"(w), "r"(v.v.x), "r"(v.v.y), "r"(v.v.z), "r"(v.v.w));
}

// acc += v (no weight): packed add.rn.f32x2, same volatile-asm ordering contract as fma_vec.
__device__ __forceinline__ void add_vec(float* acc, const Vec<float>& v) {
  asm volatile(
      "{\n"
      ".reg .b64 f0, f1, a0, a1;\n"
      "mov.b64 f0, {%4, %5};\n mov.b64 f1, {%6, %7};\n"
      "mov.b64 a0, {%0, %1};\n mov.b64 a1, {%2, %3};\n"
      "add.rn.f32x2 a0, a0, f0;\n add.rn.f32x2 a1, a1, f1;\n"
      "mov.b64 {%0, %1}, a0;\n mov.b64 {%2, %3}, a1;\n"
      "}\n"
      : "+f"(acc[0]), "+f"(acc[1]), "+f"(acc[2]), "+f"(acc[3])
      : "f"(v.v.x), "f"(v.v.y), "f"(v.v.z), "f"(v.v.w));
}
__device__ __forceinline__ void add_vec(float* acc, const Vec<__nv_bfloat16>& v) {
  asm volatile(
      "{\n"
      ".reg .b32 l0, h0, l1, h1, l2, h2, l3, h3;\n"
      ".reg .b64 f0, f1, f2, f3, a0, a1, a2, a3;\n"
      "shl.b32 l0, %8, 16;\n  and.b32 h0, %8, 0xffff0000;\n"
      "shl.b32 l1, %9, 16;\n  and.b32 h1, %9, 0xffff0000;\n"
      "shl.b32 l2, %10, 16;\n and.b32 h2, %10, 0xffff0000;\n"
      "shl.b32 l3, %11, 16;\n and.b32 h3, %11, 0xffff0000;\n"
      "mov.b64 f0, {l0, h0};\n mov.b64 f1, {l1, h1};\n mov.b64 f2, {l2, h2};\n mov.b64 f3, {l3, h3};\n"
      "mov.b64 a0, {%0, %1};\n mov.b64 a1, {%2, %3};\n mov.b64 a2, {%4, %5};\n mov.b64 a3, {%6, %7};\n"
      "add.rn.f32x2 a0, a0, f0;\n add.rn.f32x2 a1, a1, f1;\n add.rn.f32x2 a2, a2, f2;\n add.rn.f32x2 a3, a3, f3;\n"
      "mov.b64 {%0, %1}, a0;\n mov.b64 {%2, %3}, a1;\n mov.b64 {%4, %5}, a2;\n mov.b64 {%6, %7}, a3;\n"
      "}\n"
      : "+f"(acc[0]), "+f"(acc[1]), "+f"(acc[2]), "+f"(acc[3]), "+f"(acc[4]), "+f"(acc[5]), "+f"(acc[6]), "+f"(acc[7])
      : "r"(v.v.x), "r"(v.v.y), "r"(v.v.z), "r"(v.v.w));
}
// Predicated packed add: acc += v iff pred != 0.  The predicate lives inside the asm block, so the unrolled
// gather body has no branches (no BSSY/BSYNC) and needs no per-slot weight registers.
__device__ __forceinline__ void add_vec_if(float* acc, const Vec<float>& v, uint32_t pred) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 f0, f1, a0, a1;\n"
      "setp.ne.b32 p, %8, 0;\n"
      "mov.b64 f0, {%4, %5};\n mov.b64 f1, {%6, %7};\n"
      "mov.b64 a0, {%0, %1};\n mov.b64 a1, {%2, %3};\n"
      "@p add.rn.f32x2 a0, a0, f0;\n @p add.rn.f32x2 a1, a1, f1;\n"
      "mov.b64 {%0, %1}, a0;\n mov.b64 {%2, %3}, a1;\n"
      "}\n"
      : "+f"(acc[0]), "+f"(acc[1]), "+f"(acc[2]), "+f"(acc[3])
      : "f"(v.v.x), "f"(v.v.y), "f"(v.v.z), "f"(v.v.w), "r"(pred));
}
__device__ __forceinline__ void add_vec_if(float* acc, const Vec<__nv_bfloat16>& v, uint32_t pred) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b32 l0, h0, l1, h1, l2, h2, l3, h3;\n"
      ".reg .b64 f0, f1, f2, f3, a0, a1, a2, a3;\n"
      "setp.ne.b32 p, %12, 0;\n"
      "shl.b32 l0, %8, 16;\n  and.b32 h0, %8, 0xffff0000;\n"
      "shl.b32 l1, %9, 16;\n  and.b32 h1, %9, 0xffff0000;\n"
      "shl.b32 l2, %10, 16;\n and.b32 h2, %10, 0xffff0000;\n"
      "shl.b32 l3, %11, 16;\n and.b32 h3, %11, 0xffff0000;\n"
      "mov.b64 f0, {l0, h0};\n mov.b64 f1, {l1, h1};\n mov.b64 f2, {l2, h2};\n mov.b64 f3, {l3, h3};\n"
      "mov.b64 a0, {%0, %1};\n mov.b64 a1, {%2, %3};\n mov.b64 a2, {%4, %5};\n mov.b64 a3, {%6, %7};\n"
      "@p add.rn.f32x2 a0, a0, f0;\n @p add.rn.f32x2 a1, a1, f1;\n @p add.rn.f32x2 a2, a2, f2;\n @p add.rn.f32x2 a3, a3, f3;\n"
      "mov.b64 {%0, %1}, a0;\n mov.b64 {%2, %3}, a1;\n mov.b64 {%4, %5}, a2;\n mov.b64 {%6, %7}, a3;\n"
      "}\n"
      : "+f"(acc[0]), "+f"(acc[1]), "+f"(acc[2]), "+f"(acc[3]), "+f"(acc[4]), "+f"(acc[5]), "+f"(acc[6]), "+f"(acc[7])
      : "r"(v.v.x), "r"(v.v.y), "r"(v.v.z), "r"(v.v.w), "r"(pred));
}
// bf16 rows, unweighted: sm_100 mixed-precision add (add.rn.f32.bf16 -> SASS FHADD.BF16 Rd, Ra.H0|H1, Rc) adds a
// bf16 half straight into an fp32 accumulator: 8 instructions per 16 bytes, no unpacking.  Exact (bf16 -> fp32 is
// exact, one fp32 rounding per add), i.e. bit-identical to convert-then-add.
__device__ __forceinline__ void add_vec_mixed_if(float* acc, const Vec<__nv_bfloat16>& v, uint32_t pred) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b16 l0, h0, l1, h1, l2, h2, l3, h3;\n"
      "setp.ne.b32 p, %12, 0;\n"
      "mov.b32 {l0, h0}, %8;\n mov.b32 {l1, h1}, %9;\n mov.b32 {l2, h2}, %10;\n mov.b32 {l3, h3}, %11;\n"
      "@p add.rn.f32.bf16 %0, l0, %0;\n @p add.rn.f32.bf16 %1, h0, %1;\n"
      "@p add.rn.f32.bf16 %2, l1, %2;\n @p add.rn.f32.bf16 %3, h1, %3;\n"
      "@p add.rn.f32.bf16 %4, l2, %4;\n @p add.rn.f32.bf16 %5, h2, %5;\n"
      "@p add.rn.f32.bf16 %6, l3, %6;\n @p add.rn.f32.bf16 %7, h3, %7;\n"
      "}\n"
      : "+f"(acc[0]), "+f"(acc[1]), "+f"(acc[2]), "+f"(acc[3]), "+f"(acc[4]), "+f"(acc[5]), "+f"(acc[6]), "+f"(acc[7])
      : "r"(v.v.x), "r"(v.v.y), "r"(v.v.z), "r"(v.v.w), "r"(pred));
}
__device__ __forceinline__ void add_vec_mixed_if(float* acc, const Vec<float>& v, uint32_t pred) { add_vec_if(acc, v, pred); }
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// L2 eviction-priority policies (createpolicy) for streaming traffic.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
template <typename T>
__device__ __forceinline__ void stg_vec_hint(T* p, const Vec<T>& r, uint64_t pol) {
  const uint4 u = *reinterpret_cast<const uint4*>(&r.v);
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(u.x),
               "r"(u.y), "r"(u.z), "r"(u.w), "l"(pol)
               : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Reduce four per-lane partials (one per attention head) across the warp with 9 shuffles instead
// of 20: fold the lane space in halves while packing heads, then finish with a butterfly.
// Returns, on every lane, the full sums {s0,s1,s2,s3}.
__device__ __forceinline__ void warp_sum4(float& a, float& b, float& c, float& d) {
  const int lane = threadIdx.x & 31;
  // step 1: lanes <16 keep (a,b), lanes >=16 keep (c,d)
  {
    const bool hi = lane & 16;
    float s0 = hi ? a : c, s1 = hi ? b : d;  // what we send away
    float k0 = hi ? c : a, k1 = hi ? d : b;  // what we keep
    k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
    k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
    a = k0; b = k1;  // lanes<16: (a,b)   lanes>=16: (c,d)
  }
  // step 2: within each half, lanes with bit3 clear keep first, set keep second
  {
    const bool hi = lane & 8;
    float s = hi ? a : b, k = hi ? b : a;
    k += __shfl_xor_sync(0xffffffffu, s, 8);
    a = k;  // bit4,bit3 = 00:a 01:b 10:c 11:d
  }
  a += __shfl_xor_sync(0xffffffffu, a, 4);
  a += __shfl_xor_sync(0xffffffffu, a, 2);
  a += __shfl_xor_sync(0xffffffffu, a, 1);
  // broadcast back: head h lives on lanes with (lane>>3)==h
  const float r0 = __shfl_sync(0xffffffffu, a, 0);
  const float r1 = __shfl_sync(0xffffffffu, a, 8);
  const float r2 = __shfl_sync(0xffffffffu, a, 16);
  const float r3 = __shfl_sync(0xffffffffu, a, 24);
  a = r0; b = r1; c = r2; d = r3;
}

// ---------------------------------------------------------------- dropout epoch (CUDA-graph replay)
// Dropout seeds are kernel ARGUMENTS: a captured CUDA graph would replay the same masks forever.  The library therefore
// keeps one 64-bit epoch per device in device memory; every dropout kernel mixes it into its seed, and
// b2g_dropout_epoch_advance (one tiny launch, captured at the top of a training-step graph) increments it.  The epoch is 0
// until advanced, so eager runs draw exactly the masks their seeds define; forward and backward of one step see the
// same epoch.
const void* zero_row_ptr();                   // host: 1 KB of zeros on this device (allocated on first use), api.cu
const uint64_t* dropout_epoch_ptr();          // host: this device's epoch word (allocated on first use), api.cu
__device__ __forceinline__ uint64_t mix_epoch(uint64_t seed, const uint64_t* epoch) {
  return epoch ? seed ^ (__ldg(epoch) * 0xD1B54A32D192ED03ull) : seed;
}

// ---------------------------------------------------------------- counter-based RNG for dropout
// Philox-4x32-10 keyed by (seed), counter (element index).  One call -> 4 uniform floats in [0,1).
__device__ __forceinline__ uint4 philox4x32(uint64_t seed, uint64_t ctr) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = 0x243F6A88u, c3 = 0x85A308D3u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
// keep-mask scale for attention dropout on edge position p (4 heads at once)
__device__ __forceinline__ void dropout_scale4(uint64_t seed, uint64_t p, float p_drop, float* s) {
  const uint4 r = philox4x32(seed, p);
  const float inv = 1.0f / (1.0f - p_drop);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int h = 0; h < 4; ++h) s[h] = ((w[h] >> 8) * (1.0f / 16777216.0f) >= p_drop) ? inv : 0.0f;
}

}  // namespace b2g
