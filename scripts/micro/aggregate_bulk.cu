// aggregate_bulk.cu — K2/K3, bulk-async gather variant of the CSR segment-sum (same contract as
// seg_sum_kernel in aggregate.cu, same fp32 summation order, bit-identical results).
//
// Why: the register-gather kernel tops out near 3.3 TB/s of DRAM+L2 traffic whatever its instruction
// count or occupancy (ncu: issue slots, L2 and DRAM all far from peak) — it is limited by how many
// loads an SM's L1 can keep outstanding.  Here every neighbour row (a contiguous F*s-byte run) is fetched
// by ONE cp.async.bulk (the TMA engine, global -> shared, mbarrier complete_tx) instead of 32 lanes x
// LDG.128, so bytes in flight are bounded by shared memory (~100 KB per SM), not by L1 request slots.
//
// Structure: no producer warp; each warp runs its own STAGES-deep ring.  An "item" is up to SLOTS (8)
// neighbours of one target row.  To issue an item, lane u loads col[j+u] and fires the bulk copy of that
// row into slot u of the stage; lane 0 arms the stage's mbarrier with the expected byte count.  The warp
// keeps STAGES-1 items in flight ahead of the one it is summing out of shared memory (conflict-free
// 16-byte reads), so the rowptr -> col -> row dependency chain is fully overlapped.
#include "common.cuh"

namespace b2g {

constexpr int BK_SLOTS = 8;    // neighbour rows per stage (mesh rows have 7 incl. the self loop)
constexpr int BK_STAGES = 4;
constexpr int BK_ITERS = 4;    // rows per warp per CTA chunk (consecutive rows stay in one CTA)

__device__ __forceinline__ uint32_t bk_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bk_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void bk_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bk_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bk_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "BK_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra BK_DONE;\n"
      "bra BK_WAIT;\n"
      "BK_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bk_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// Deterministic walk over the items of one warp: rows c0 + it*W + wi for it < BK_ITERS, then the next
// chunk c0 += grid*CHUNK; every row contributes max(1, ceil(deg / BK_SLOTS)) items.
struct BkCursor {
  int64_t c0, row;
  int it, b, e, j;
  bool valid;
};

template <typename T, int VPL, int kScale>
__global__ void __launch_bounds__(1024)
seg_sum_bulk_kernel(const T* __restrict__ x, int64_t ldx, const T* __restrict__ x_self, int64_t ldxs,
                    T* __restrict__ out, int64_t ldo, int64_t n_rows, int nvec,
                    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                    const float* __restrict__ row_scale, const float* __restrict__ col_scale,
                    float self_coef, const float* __restrict__ bias, int relu) {
  constexpr int VN = Vec<T>::N;
  extern __shared__ __align__(128) uint8_t bk_smem[];
  const int W = blockDim.x >> 5;                              // warps per CTA
  const int wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t row_bytes = (uint32_t)nvec * 16u;
  const uint32_t stage_bytes = BK_SLOTS * row_bytes;
  const uint32_t warp_bytes = BK_STAGES * stage_bytes;
  uint8_t* my_smem = bk_smem + (size_t)wi * warp_bytes;
  const uint32_t my_smem_u = bk_smem_u32(my_smem);
  const uint32_t bars = bk_smem_u32(bk_smem) + (uint32_t)W * warp_bytes + (uint32_t)wi * BK_STAGES * 8u;
  if (lane == 0) {
    for (int s = 0; s < BK_STAGES; ++s) bk_mbar_init(bars + 8u * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const uint64_t pol_stream = l2_policy_evict_first();
  const int64_t CHUNK = (int64_t)W * BK_ITERS;
  const int64_t gstride = (int64_t)gridDim.x * CHUNK;

  float bias_r[VPL][VN];
#pragma unroll
  for (int v = 0; v < VPL; ++v)
#pragma unroll
    for (int k = 0; k < VN; ++k) {
      const int vi = lane + v * 32;
      bias_r[v][k] = (bias && vi < nvec) ? __ldg(bias + vi * VN + k) : 0.f;
    }

  auto load_row = [&](BkCursor& c) {      // (re)load the extent of c.row; mark invalid past the end
    c.valid = c.row < n_rows;
    if (c.valid) {
      c.b = __ldg(rowptr + c.row);
      c.e = __ldg(rowptr + c.row + 1);
      c.j = c.b;
    }
  };
  auto next_row = [&](BkCursor& c) {
    if (++c.it == BK_ITERS) { c.it = 0; c.c0 += gstride; }
    c.row = c.c0 + (int64_t)c.it * W + wi;
    // a chunk whose first rows exist may still run past n_rows for later warps/iterations: skip forward
    while (c.row >= n_rows && c.c0 < n_rows) {
      if (++c.it == BK_ITERS) { c.it = 0; c.c0 += gstride; }
      c.row = c.c0 + (int64_t)c.it * W + wi;
    }
    load_row(c);
  };
  auto advance = [&](BkCursor& c) {       // next item
    c.j += BK_SLOTS;
    if (c.j >= c.e) next_row(c);
  };
  auto start = [&](BkCursor& c) {
    c.c0 = (int64_t)blockIdx.x * CHUNK;
    c.it = 0;
    c.row = c.c0 + wi;
    while (c.row >= n_rows && c.c0 < n_rows) {
      if (++c.it == BK_ITERS) { c.it = 0; c.c0 += gstride; }
      c.row = c.c0 + (int64_t)c.it * W + wi;
    }
    load_row(c);
  };
  // issue the bulk copies of item `c` into stage s
  auto issue = [&](const BkCursor& c, int s) {
    const int n = min(BK_SLOTS, c.e - c.j);                    // may be <= 0 for an empty row
    const uint32_t bar = bars + 8u * s;
    if (lane == 0) {
      if (n > 0) bk_mbar_expect_tx(bar, (uint32_t)n * row_bytes);
      else bk_mbar_arrive(bar);
    }
    __syncwarp();                                              // the barrier is armed before any copy can complete on it
    if (lane < n) {
      const int cidx = __ldg(col + c.j + lane);
      bk_bulk_g2s(my_smem_u + (uint32_t)s * stage_bytes + (uint32_t)lane * row_bytes, x + (int64_t)cidx * ldx, row_bytes, bar);
    }
  };

  BkCursor pc, cc;                                             // producer / consumer cursors over the same item stream
  start(pc);
  cc = pc;
  int p_stage = 0, c_stage = 0;
  uint32_t c_phase = 0;
  // prologue: fill STAGES-1 stages
  for (int s = 0; s < BK_STAGES - 1; ++s) {
    if (pc.valid) {
      issue(pc, p_stage);
      advance(pc);
    }
    if (++p_stage == BK_STAGES) p_stage = 0;
  }

  float acc[VPL][VN];
#pragma unroll
  for (int v = 0; v < VPL; ++v)
#pragma unroll
    for (int k = 0; k < VN; ++k) acc[v][k] = 0.f;

  while (cc.valid) {
    // keep the ring full: the stage being refilled was consumed in the previous iteration
    if (pc.valid) {
      issue(pc, p_stage);
      advance(pc);
    }
    if (++p_stage == BK_STAGES) p_stage = 0;

    const int n = min(BK_SLOTS, cc.e - cc.j);
    const float rs = (kScale && row_scale) ? __ldg(row_scale + cc.row) : 1.0f;
    float w_l = 1.0f;
    if (kScale == 2 && lane < n) w_l = (col_scale ? __ldg(col_scale + __ldg(col + cc.j + lane)) : 1.0f) * rs;
    bk_mbar_wait(bars + 8u * c_stage, c_phase);
    const uint8_t* st = my_smem + (size_t)c_stage * stage_bytes;
    for (int u = 0; u < n; ++u) {                               // fp32 accumulation in CSR (= edge) order
      const float w = kScale == 2 ? __shfl_sync(0xffffffffu, w_l, u) : 1.0f;
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int vi = lane + v * 32;
        if (vi < nvec) {
          Vec<T> t;
          t.v = *reinterpret_cast<const decltype(t.v)*>(st + (size_t)u * row_bytes + (size_t)vi * 16);
          fma_vec(acc[v], w, t);
        }
      }
    }
    __syncwarp();                                              // all lanes are done reading this stage before it is refilled
    if (cc.j + BK_SLOTS >= cc.e) {                             // last item of the row: epilogue + store
      const int64_t i = cc.row;
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int vi = lane + v * 32;
        if (vi < nvec) {
          if (kScale == 1) {
#pragma unroll
            for (int k = 0; k < VN; ++k) acc[v][k] *= rs;
          }
          if (self_coef != 0.f) {
            const Vec<T> sv = ldg_vec_l1<T>((x_self ? x_self + i * ldxs : x + i * ldx) + vi * VN);
            fma_vec(acc[v], self_coef, sv);
          }
#pragma unroll
          for (int k = 0; k < VN; ++k) acc[v][k] += bias_r[v][k];
          if (relu) {
#pragma unroll
            for (int k = 0; k < VN; ++k) acc[v][k] = fmaxf(acc[v][k], 0.f);
          }
          Vec<T> o;
          o.from_float(acc[v]);
          stg_vec_hint<T>(out + i * ldo + vi * VN, o, pol_stream);
#pragma unroll
          for (int k = 0; k < VN; ++k) acc[v][k] = 0.f;
        }
      }
    }
    advance(cc);
    if (++c_stage == BK_STAGES) { c_stage = 0; c_phase ^= 1; }
  }
}


// ------------------------------------------------------------------------------------------------
// cp.async (LDGSTS) variant: same ring, but every lane copies ITS OWN 16-byte pieces of each neighbour row
// (cp.async.ca: allocates in L1, so x+-1 / self rows shared by adjacent targets of the CTA still hit L1)
// and later reads back exactly those pieces: shared memory acts as a per-thread "in-flight buffer", no
// mbarrier, no cross-lane traffic, and the bytes in flight no longer cost registers.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename T, int VPL, int kScale>
__global__ void __launch_bounds__(1024)
seg_sum_cpasync_kernel(const T* __restrict__ x, int64_t ldx, const T* __restrict__ x_self, int64_t ldxs,
                       T* __restrict__ out, int64_t ldo, int64_t n_rows, int nvec,
                       const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                       const float* __restrict__ row_scale, const float* __restrict__ col_scale,
                       float self_coef, const float* __restrict__ bias, int relu) {
  constexpr int VN = Vec<T>::N;
  extern __shared__ __align__(128) uint8_t bk_smem[];
  const int W = blockDim.x >> 5;
  const int wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t row_bytes = (uint32_t)nvec * 16u;
  const uint32_t stage_bytes = BK_SLOTS * row_bytes;
  const uint32_t warp_bytes = BK_STAGES * stage_bytes;
  uint8_t* my_smem = bk_smem + (size_t)wi * warp_bytes;
  const uint32_t my_smem_u = bk_smem_u32(my_smem);
  const uint64_t pol_stream = l2_policy_evict_first();
  const int64_t CHUNK = (int64_t)W * BK_ITERS;
  const int64_t gstride = (int64_t)gridDim.x * CHUNK;

  auto load_row = [&](BkCursor& c) {
    c.valid = c.row < n_rows;
    if (c.valid) {
      c.b = __ldg(rowptr + c.row);
      c.e = __ldg(rowptr + c.row + 1);
      c.j = c.b;
    }
  };
  auto skip = [&](BkCursor& c) {
    while (c.row >= n_rows && c.c0 < n_rows) {
      if (++c.it == BK_ITERS) { c.it = 0; c.c0 += gstride; }
      c.row = c.c0 + (int64_t)c.it * W + wi;
    }
  };
  auto next_row = [&](BkCursor& c) {
    if (++c.it == BK_ITERS) { c.it = 0; c.c0 += gstride; }
    c.row = c.c0 + (int64_t)c.it * W + wi;
    skip(c);
    load_row(c);
  };
  auto advance = [&](BkCursor& c) {
    c.j += BK_SLOTS;
    if (c.j >= c.e) next_row(c);
  };
  auto issue = [&](const BkCursor& c, int s) {       // one commit group per item (possibly empty)
    if (c.valid) {
      const int n = min(BK_SLOTS, c.e - c.j);
      const uint32_t base = my_smem_u + (uint32_t)s * stage_bytes + (uint32_t)lane * 16u;
#pragma unroll
      for (int u = 0; u < BK_SLOTS; ++u) {
        if (u < n) {
          const int cidx = __ldg(col + c.j + u);          // same address on every lane: one broadcast transaction
          const T* row = x + (int64_t)cidx * ldx;
#pragma unroll
          for (int v = 0; v < VPL; ++v) {
            const int vi = lane + v * 32;
            if (vi < nvec) cp_async16(base + (uint32_t)u * row_bytes + (uint32_t)v * 512u, row + vi * VN);
          }
        }
      }
    }
    cp_async_commit();
  };

  BkCursor pc, cc;
  pc.c0 = (int64_t)blockIdx.x * CHUNK;
  pc.it = 0;
  pc.row = pc.c0 + wi;
  skip(pc);
  load_row(pc);
  cc = pc;
  int p_stage = 0, c_stage = 0;
  for (int s = 0; s < BK_STAGES - 1; ++s) {
    issue(pc, p_stage);
    if (pc.valid) advance(pc);
    if (++p_stage == BK_STAGES) p_stage = 0;
  }

  float acc[VPL][VN];
#pragma unroll
  for (int v = 0; v < VPL; ++v)
#pragma unroll
    for (int k = 0; k < VN; ++k) acc[v][k] = 0.f;

  while (cc.valid) {
    issue(pc, p_stage);                                  // refill the stage consumed one iteration ago
    if (pc.valid) advance(pc);
    if (++p_stage == BK_STAGES) p_stage = 0;

    const int n = min(BK_SLOTS, cc.e - cc.j);
    const float rs = (kScale && row_scale) ? __ldg(row_scale + cc.row) : 1.0f;
    cp_async_wait<BK_STAGES - 1>();                      // this lane's copies of item cc have landed
    const uint8_t* st = my_smem + (size_t)c_stage * stage_bytes + (size_t)lane * 16;
    for (int u = 0; u < n; ++u) {                         // fp32 accumulation in CSR (= edge) order
      float w = 1.0f;
      if (kScale == 2) w = (col_scale ? __ldg(col_scale + __ldg(col + cc.j + u)) : 1.0f) * rs;
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int vi = lane + v * 32;
        if (vi < nvec) {
          Vec<T> t;
          t.v = *reinterpret_cast<const decltype(t.v)*>(st + (size_t)u * row_bytes + (size_t)v * 512);
          fma_vec(acc[v], w, t);
        }
      }
    }
    if (cc.j + BK_SLOTS >= cc.e) {                        // last item of the row: epilogue + store
      const int64_t i = cc.row;
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int vi = lane + v * 32;
        if (vi < nvec) {
          if (kScale == 1) {
#pragma unroll
            for (int k = 0; k < VN; ++k) acc[v][k] *= rs;
          }
          if (self_coef != 0.f) {
            const Vec<T> sv = ldg_vec_l1<T>((x_self ? x_self + i * ldxs : x + i * ldx) + vi * VN);
            fma_vec(acc[v], self_coef, sv);
          }
          if (bias) {
#pragma unroll
            for (int k = 0; k < VN; ++k) acc[v][k] += __ldg(bias + vi * VN + k);
          }
          if (relu) {
#pragma unroll
            for (int k = 0; k < VN; ++k) acc[v][k] = fmaxf(acc[v][k], 0.f);
          }
          Vec<T> o;
          o.from_float(acc[v]);
          stg_vec_hint<T>(out + i * ldo + vi * VN, o, pol_stream);
#pragma unroll
          for (int k = 0; k < VN; ++k) acc[v][k] = 0.f;
        }
      }
    }
    advance(cc);
    if (++c_stage == BK_STAGES) c_stage = 0;
  }
  cp_async_wait<0>();
}

// host: pick warps per CTA from the shared-memory budget, one CTA per SM
template <typename T, int VPL>
static int launch_bulk(int variant, const void* x, int64_t ldx, const void* x_self, int64_t ldxs, void* out, int64_t ldo,
                       int64_t n_rows, int nvec, const int32_t* rowptr, const int32_t* col, const float* rs,
                       const float* cs, float self_coef, const float* bias, int relu, cudaStream_t st) {
  const int64_t warp_bytes = (int64_t)BK_STAGES * BK_SLOTS * nvec * 16;
  int warps = (int)((220 * 1024) / (warp_bytes + BK_STAGES * 8));
  if (warps > 16) warps = 16;
  if (warps < 2) return B2G_E_SHAPE;
  const size_t smem = (size_t)warps * (warp_bytes + BK_STAGES * 8);
  const int mode = cs ? 2 : (rs ? 1 : 0);
  int64_t blocks = ceil_div(n_rows, (int64_t)warps * BK_ITERS);
  if (blocks > B2G_NUM_SMS) blocks = B2G_NUM_SMS;
#define B2G_BK(MODE)                                                                                              \
  if (variant == 3) {                                                                                             \
    cudaError_t e = cudaFuncSetAttribute(seg_sum_cpasync_kernel<T, VPL, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         (int)smem);                                                              \
    if (e != cudaSuccess) return (int)e;                                                                          \
    seg_sum_cpasync_kernel<T, VPL, MODE><<<(unsigned)blocks, warps * 32, smem, st>>>(                             \
        (const T*)x, ldx, (const T*)x_self, ldxs, (T*)out, ldo, n_rows, nvec, rowptr, col, rs, cs, self_coef,     \
        bias, relu);                                                                                              \
  } else {                                                                                                        \
    cudaError_t e = cudaFuncSetAttribute(seg_sum_bulk_kernel<T, VPL, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         (int)smem);                                                              \
    if (e != cudaSuccess) return (int)e;                                                                          \
    seg_sum_bulk_kernel<T, VPL, MODE><<<(unsigned)blocks, warps * 32, smem, st>>>(                                \
        (const T*)x, ldx, (const T*)x_self, ldxs, (T*)out, ldo, n_rows, nvec, rowptr, col, rs, cs, self_coef,     \
        bias, relu);                                                                                              \
  }
  if (mode == 2) B2G_BK(2) else if (mode == 1) B2G_BK(1) else B2G_BK(0)
#undef B2G_BK
  count_launch();
  return cuda_status();
}

bool bulk_seg_sum_supported(int nvec, int64_t n_rows) { return nvec >= 16 && nvec <= 128 && n_rows >= 1; }

int bulk_seg_sum(int variant, const void* x, int64_t ldx, const void* x_self, int64_t ldxs, void* out, int64_t ldo,
                 int64_t n_rows, int nvec, int dt, const int32_t* rowptr, const int32_t* col, const float* rs,
                 const float* cs, float self_coef, const float* bias, int relu, cudaStream_t st) {
#define B2G_D(T)                                                                                                       \
  if (nvec <= 32) return launch_bulk<T, 1>(variant, x, ldx, x_self, ldxs, out, ldo, n_rows, nvec, rowptr, col, rs, cs, self_coef, bias, relu, st); \
  if (nvec <= 64) return launch_bulk<T, 2>(variant, x, ldx, x_self, ldxs, out, ldo, n_rows, nvec, rowptr, col, rs, cs, self_coef, bias, relu, st); \
  return launch_bulk<T, 4>(variant, x, ldx, x_self, ldxs, out, ldo, n_rows, nvec, rowptr, col, rs, cs, self_coef, bias, relu, st);
  if (dt == B2G_F32) { B2G_D(float) }
  B2G_D(__nv_bfloat16)
#undef B2G_D
}

}  // namespace b2g
