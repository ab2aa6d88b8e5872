// scan.cuh — ordered compaction building block: a 3-phase exclusive scan over a *functor* of
// per-item counts.  Phase 1 (tile sums) and phase 3 (re-scan + consume) recompute the functor
// instead of materialising flags, so a compaction reads its source twice and writes only output.
#pragma once
#include "common.cuh"

namespace b2g {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;                          // consecutive items per thread
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;    // 4096 items per CTA

inline int64_t scan_num_tiles(int64_t n) { return n > 0 ? ceil_div(n, SCAN_TILE) : 1; }
// workspace: int64 tile_offsets[num_tiles + 1]
inline int64_t scan_ws_bytes(int64_t n) { return (scan_num_tiles(n) + 1) * (int64_t)sizeof(int64_t); }

// Exclusive scan of one int per thread across the CTA; returns the thread's offset, total in *total.
__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
  __shared__ int warp_tot[SCAN_THREADS / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[w] = inc;
  __syncthreads();
  int base = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < SCAN_THREADS / 32; ++i) {
    const int t = warp_tot[i];
    if (i < w) base += t;
    tot += t;
  }
  __syncthreads();  // warp_tot reusable
  *total = tot;
  return base + inc - v;
}

// Phase 1: tile_sums[tile] = sum_{i in tile} f(i).  The functor may have side effects (marking).
template <typename F>
__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums_kernel(int64_t n, int64_t* tile_sums, F f) {
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    const int64_t i = base + k;
    if (i < n) s += f(i);
  }
  int tot;
  block_exclusive_scan(s, &tot);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

// Phase 2: in-place exclusive scan of tile sums (single CTA), grand total appended at [num_tiles]
// and optionally copied to total_out.
static __global__ void __launch_bounds__(1024) scan_tiles_kernel(int64_t* tile_sums, int64_t num_tiles, int64_t* total_out) {
  __shared__ int64_t warp_tot[32];
  __shared__ int64_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int64_t start = 0; start < num_tiles; start += 1024) {
    const int64_t i = start + threadIdx.x;
    const int64_t v = i < num_tiles ? tile_sums[i] : 0;
    int64_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[w] = inc;
    __syncthreads();
    int64_t base = 0, tot = 0;
    for (int j = 0; j < 32; ++j) {
      const int64_t t = warp_tot[j];
      if (j < w) base += t;
      tot += t;
    }
    const int64_t carry = carry_s;
    if (i < num_tiles) tile_sums[i] = carry + base + inc - v;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    tile_sums[num_tiles] = carry_s;
    if (total_out) *total_out = carry_s;
  }
}

// Phase 3: for every item with f(i) != 0 (or every item when kAll) call g(i, exclusive_prefix(i)).
template <bool kAll, typename F, typename G>
__global__ void __launch_bounds__(SCAN_THREADS) scan_consume_kernel(int64_t n, const int64_t* tile_offsets, F f, G g) {
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int c[SCAN_ITEMS];
  int s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    const int64_t i = base + k;
    c[k] = (i < n) ? f(i) : 0;
    s += c[k];
  }
  int tot;
  int64_t off = tile_offsets[blockIdx.x] + block_exclusive_scan(s, &tot);
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    if (kAll ? (base + k < n) : (c[k] != 0)) g(base + k, off);
    off += c[k];
  }
}

// Host drivers.  `tiles` = workspace of scan_ws_bytes(n).
template <typename F>
inline void scan_count(int64_t n, int64_t* tiles, int64_t* total_out, F f, cudaStream_t st) {
  const int64_t nt = scan_num_tiles(n);
  scan_tile_sums_kernel<<<(unsigned)nt, SCAN_THREADS, 0, st>>>(n, tiles, f);
  scan_tiles_kernel<<<1, 1024, 0, st>>>(tiles, nt, total_out);
  count_launch(2);
}
template <bool kAll = false, typename F, typename G>
inline void scan_consume(int64_t n, const int64_t* tiles, F f, G g, cudaStream_t st) {
  const int64_t nt = scan_num_tiles(n);
  scan_consume_kernel<kAll><<<(unsigned)nt, SCAN_THREADS, 0, st>>>(n, tiles, f, g);
  count_launch(1);
}

}  // namespace b2g
