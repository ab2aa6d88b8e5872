// gat_fused.cu — K4f: the attention-weighted neighbour aggregation of GATConv(heads = 4, concat = False)
// (gnn_model.py:65-68,168) FUSED with the output projection:
//
//      out_i = sum_h (sum_j alpha_ijh x_j) Wc_h^T + b = z_i Wc^T + b,      z_i = [sum_j alpha_ij1 x_j | ... | sum_j alpha_ij4 x_j]
//
// In the unfused aggregate-first path (gat_rows.cu) z [N, H*F] is written to HBM by gatz_fwd_kernel and read back by the
// K6 GEMM: 2 x 20.5 GB of the layer's 56 GB at cfg4, for 26 GB of compulsory traffic.  Here z never leaves the SM:
//
//   gather warps   compute bf16 tiles of z — 128 target rows x (4 heads x 64 features) per step — straight into the
//                  128B-swizzled K-major shared-memory layout that tcgen05.mma reads (the layout TMA would produce),
//   MMA warp       one elected thread issues tcgen05.mma (kind::f16, 128 x 256 x 16) on them against Wc k-blocks streamed
//                  by TMA (Wc is 512 KB: L2-resident, not shared-memory-resident), fp32 accumulator in TMEM,
//   epilogue warps tcgen05.ld -> + bias -> bf16 -> 128-byte row stores, overlapped with the next tile (2 accumulators).
//
// The softmax is NOT in this kernel: b2g_gat_alpha (below) writes the post-dropout attention weights alpha [nnz, 4] fp32
// once (1.1 GB at cfg4; the backward pass wants them anyway), and the gather warps read 16 bytes per entry next to the
// column index.  That keeps the gather loop to: index broadcast, one 16-byte load per lane, 8 unpacks, 16 packed FMAs.
//
// K order.  A row's 4 heads x 256 features do not fit shared memory for 128 rows (256 KB), so the reduction is cut by
// FEATURE: chunk kc = features [64 kc, 64 kc + 64) of all four heads = 256 K-columns = 4 swizzle atoms (one per head).
// A warp gathers 128-byte pieces of 4 neighbour rows per load instruction (lane = 4 rows x 8 pieces), the same bytes in
// total as gathering 512-byte rows once.  The host permutes the columns of Wc accordingly:
//      Wp[c, kc*256 + h*64 + f] = Wc[c, h*256 + kc*64 + f].
//
// Also the transposed use (backward): y_j = [sum_i alpha_ijh g_i]_h over the source-major CSR followed by the dgrad GEMM is
// the same computation (perm maps a transposed-CSR position to its alpha entry).
#include "rows.cuh"
#include <cstdlib>
#include "tc_ptx.cuh"

namespace b2g {

constexpr int GF_BM = 128;                       // target rows per tile == UMMA_M
constexpr int GF_BN = 256;                       // UMMA_N; C_out <= 256
constexpr int GF_H = 4;                          // heads
constexpr int GF_KCH = 4;                        // feature chunks of 64 (F = 256 bf16 = 512-byte rows)
constexpr int GF_GW = 16;                        // gather warps
constexpr int GF_THREADS = 32 * (8 + GF_GW);     // warp 0 TMA, 1 MMA, 2-3 idle (warpgroup padding), 4-7 epilogue, 8.. gather
constexpr int GF_A_REGION = GF_BM * 128;         // 16 KB: 128 rows x 64 bf16 (one head of one chunk), SWIZZLE_128B K-major
constexpr int GF_A_CHUNK = GF_H * GF_A_REGION;   // 64 KB
constexpr int GF_W_STAGE = GF_BN * 128;          // 32 KB: 256 rows of Wp x 64 k
constexpr int GF_W_STAGES = 2;
// CTA-pair variant (kPair, tcgen05 cta_group::2): each CTA of a cluster of two gathers its own 128-row tile, the leader issues ONE
// M = 256 MMA per k-step, and each CTA holds only HALF of every Wp k-block (128 of its 256 rows): the TMA writes of Wp and the
// tensor core's B-operand reads per SM halve — the two biggest items of the shared-memory traffic that bounds this kernel
// (DESIGN §4 point 7).  Same 64 KB ring: 4 stages of 16 KB.
constexpr int GF_W_HALF = (GF_BN / 2) * 128;     // 16 KB
constexpr int GF_W_STAGES_PAIR = 4;
constexpr int GF_STG = 4 * 32 * 128;             // epilogue staging: 32 rows x 128 B per epilogue warp
constexpr int GF_WST_WARP = 2 * 4 * 8 * 16;      // per gather warp: 2 units x 4 rows x 8 entries x float4 weights = 1 KB
constexpr int GF_WST = GF_GW * GF_WST_WARP;
constexpr int GF_BAR = 256;
constexpr int GF_BVH = GF_H * GF_BN * 2;         // TransformerConv: bv_h / H as bf16 [4][256] (exact: the model's weights are bf16)
constexpr int GF_SMEM = 2 * GF_A_CHUNK + GF_W_STAGES * GF_W_STAGE + GF_STG + GF_WST + GF_BVH + GF_BAR;
// no static shared memory in this kernel: the dynamic window starts 1024-byte aligned (checked at run time with a trap)
static_assert(GF_SMEM <= 232448, "fused GAT shared-memory plan exceeds 227 KB");

struct GfArgs {
  const char* x; uint32_t xrow_bytes;            // gathered rows, bf16 [*, 256]
  const int32_t* rowptr; const int32_t* col; const int32_t* perm;
  const float* alpha;                            // fp32 [nnz, 4]: weight of (entry, head); entry = perm[pos] or pos
  const float* bias;                             // fp32 [m] or NULL
  const char* zero;                              // >= 512 bytes of zeros (padding lanes load from here)
  // TransformerConv (out = sum_h (z_h Wv_h^T + s_h bv_h) / H + x Ws^T + bs): optional epilogue terms
  const float* srow;                             // fp32 [n_rows, 4]: per-head weight sums s_ih (post-dropout) or NULL
  const float* bvh;                              // fp32 [4, m]: bv_h / H; out += sum_h s_ih bvh[h, :]  (needs srow) or NULL
  const __nv_bfloat16* addend; int64_t ldadd;    // [n_rows, m]: out += addend (the skip projection) or NULL
  __nv_bfloat16* out; int64_t ldo;
  uint32_t n_rows; int m;
  int tma_store;                                 // full 32-row x 64-column blocks leave through map_o (one bulk tensor store)
  // softmax inside the kernel (alpha == NULL; every row has <= 8 entries): the gather warps derive the attention weights of
  // their rows in the per-tile prologue from a_src / a_dst — no [nnz, 4] alpha round trip through HBM, no separate kernel
  const float* a_src; const float* a_dst; uint32_t lda;   // fp32: a_src[c * lda + h] (global node c), a_dst[r * lda + h] (row r of this call)
  const float* ebias;                            // fp32 [nnz, 4] added to the logits (edge features) or NULL
  float* smax; float* ssum;                      // fp32 [n_rows, 4] softmax statistics for the backward pass, or NULL
  float slope, p_drop;
  uint64_t seed; const uint64_t* epoch;
  RowSched ord;                                  // chunk_rows = GF_BM
};

// Load of entry t of a row, or of zeros when the row has no entry t.  NOT a predicated load: ptxas turns `@p ld` with a
// zero-initialised destination into load-to-temporary + predicated move, and that move waits for the load right behind it
// (ncu on the first versions: every gather load was followed by a long-scoreboard stall = one load in flight per warp).
// The padding lanes load from a zero row instead (always an L1 hit), so the load is unconditional and nothing depends on it
// until the FMAs consume it.
__device__ __forceinline__ uint4 gf_ldg_sel(const char* p, const char* zero, bool valid) {
  const char* q = valid ? p : zero;
  uint4 u;
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "l"(q));
  return u;
}

__device__ __forceinline__ float gf_lrelu(float s, float slope) { return s > 0.f ? s : s * slope; }

// acc[h][0..8) += w[h] * (8 bf16 of u)
__device__ __forceinline__ void gf_fma(float (&acc)[GF_H][8], const float4& w, const uint4& u) {
  float f[8];
  unpack_row16(u, f, __nv_bfloat16());
  const float wv[GF_H] = {w.x, w.y, w.z, w.w};
#pragma unroll
  for (int h = 0; h < GF_H; ++h)
#pragma unroll
    for (int k = 0; k < 8; k += 2) ffma2_acc(acc[h][k], acc[h][k + 1], wv[h], f[k], f[k + 1]);
}

// the first K (<= 8) entries of the 4 rows of a unit: K loads in flight per lane, then K x (LDS.128 weights, FMA)
template <int K>
__device__ __forceinline__ void gf_gather(float (&acc)[GF_H][8], const char* xk, uint32_t xrow_bytes, int cl, int len,
                                          int grp_lane0, const float4* wrow, const char* zero) {
  uint4 buf[K];
#pragma unroll
  for (int t = 0; t < K; ++t) {
    const uint32_t c = (uint32_t)__shfl_sync(0xffffffffu, cl, grp_lane0 + t);
    buf[t] = gf_ldg_sel(xk + (uint64_t)c * xrow_bytes, zero, t < len);
  }
#pragma unroll
  for (int t = 0; t < K; ++t) gf_fma(acc, wrow[t * 4], buf[t]);   // staged [entry][row]: the 4 rows of a warp read 64 contiguous bytes
}

// kSm: softmax inside (a.alpha == NULL); kEx: the TransformerConv epilogue terms (srow / bvh / addend) may be present.  Compile-time
// so that the GATConv instantiation carries neither the other mode's prologue nor the extra epilogue code (the tcgen05 Linear
// lost 10 % to an unused epilogue option; this kernel sits at its 80-register ceiling).
template <bool kPair, bool kSm, bool kEx>
__global__ void __launch_bounds__(GF_THREADS, 1)
gatw_gemm_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_o, const GfArgs a) {
  constexpr int W_STAGES = kPair ? GF_W_STAGES_PAIR : GF_W_STAGES;
  constexpr int W_BYTES = kPair ? GF_W_HALF : GF_W_STAGE;
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;                // CTA of the pair; rank 0 = leader (issues the MMAs)
  const uint32_t ngrp = kPair ? gridDim.x / 2 : gridDim.x;             // tile groups in flight: pairs or single CTAs
  const uint32_t gid = kPair ? blockIdx.x / 2 : blockIdx.x;
  // step t of this CTA's group: chunk q = 2t + rank (pair) or t; a pair runs the step when chunk 2t exists, a CTA whose own chunk
  // does not exist (or is empty) takes part with zero rows so that the pair's barriers stay in step
  auto my_tile = [&](uint32_t t, uint32_t& c0, uint32_t& rows) -> int {    // -1: done, 0: skip (single CTA, empty chunk), 1: run
    if ((kPair ? 2 * t : t) >= a.ord.n_chunks) return -1;
    const uint32_t q = kPair ? 2 * t + rank : t;
    rows = 0; c0 = 0;
    if (q < a.ord.n_chunks) c0 = a.ord.chunk(q, a.n_rows, rows);
    return (!kPair && rows == 0) ? 0 : 1;
  };
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if (smem_base & 1023u) __trap();
  const uint32_t abuf0 = smem_base;                                   // 2 x 64 KB z chunks
  const uint32_t wring = abuf0 + 2 * GF_A_CHUNK;                      // 2 x 32 KB W stages
  const uint32_t stg = wring + GF_W_STAGES * GF_W_STAGE;              // epilogue staging
  const uint32_t wst = stg + GF_STG;                                  // gather-weight staging
  const uint32_t sbv = wst + GF_WST;                                  // bv_h / H (bf16 [4][256]) or zeros
  const uint32_t bars = sbv + GF_BVH;
  auto full_w = [&](int s) { return bars + 8u * s; };
  auto empty_w = [&](int s) { return bars + 8u * (4 + s); };
  auto full_a = [&](int b) { return bars + 8u * (8 + b); };
  auto empty_a = [&](int b) { return bars + 8u * (10 + b); };
  auto tfull = [&](int t) { return bars + 8u * (12 + t); };
  auto tempty = [&](int t) { return bars + 8u * (14 + t); };
  const uint32_t tmem_slot = bars + 8u * 16;
  // pair: full_w / full_a / tempty are waited on by the leader's MMA thread and receive arrivals from both CTAs (the leader's copy
  // is the one in use); empty_w / empty_a / tfull are per CTA and get the leader's multicast commits
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (kEx && a.bvh) {                                                   // visible to the epilogue warps after the __syncthreads below
    __nv_bfloat16* sb = reinterpret_cast<__nv_bfloat16*>(smem_raw + (sbv - smem_base));
    for (int t = threadIdx.x; t < GF_H * GF_BN; t += GF_THREADS) {
      const int h = t / GF_BN, c = t - h * GF_BN;
      sb[t] = __float2bfloat16_rn(c < a.m ? a.bvh[h * a.m + c] : 0.f);
    }
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    for (int s = 0; s < W_STAGES; ++s) {
      mbar_init(full_w(s), 1);
      mbar_init(empty_w(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(full_a(s), (kPair ? 2 : 1) * GF_GW);       // one arrive per gather warp (of both CTAs)
      mbar_init(empty_a(s), 1);
      mbar_init(tfull(s), 1);
      mbar_init(tempty(s), (kPair ? 2 : 1) * 4);           // one arrive per epilogue warp (of both CTAs)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (kPair) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();           // the peer's barriers exist before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
    if (warp == 0 && lane == 0) {
      // ===================================================== TMA producer: Wp k-blocks, 16 per tile
      int stage = 0;
      uint32_t phase = 0;
      for (uint32_t t = gid;; t += ngrp) {
        uint32_t c0, rows;
        const int run = my_tile(t, c0, rows);
        if (run < 0) break;
        if (run == 0) continue;
        for (int kb = 0; kb < GF_KCH * GF_H; ++kb) {
          mbar_wait(empty_w(stage), phase ^ 1);
          if (kPair) {
            // both halves are counted on the leader's barrier (a complete_tx that lands before the leader's expect_tx is legal: the
            // phase cannot complete without the leader's arrival)
            if (rank == 0) mbar_expect_tx(full_w(stage), 2 * GF_W_HALF);
            tma_load_2d_pair(wring + stage * GF_W_HALF, &map_w, mapa_u32(full_w(stage), 0), kb * 64, (int)rank * (GF_BN / 2));
          } else {
            mbar_expect_tx(full_w(stage), GF_W_STAGE);
            tma_load_2d(wring + stage * GF_W_STAGE, &map_w, full_w(stage), kb * 64, 0);
          }
          if (++stage == W_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1 && lane == 0 && rank == 0) {
      // ===================================================== MMA issuer (pair: the leader CTA only, M = 256 over both CTAs)
      constexpr uint32_t idesc = make_idesc_bf16(kPair ? 2 * GF_BM : GF_BM, GF_BN);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0, g = 0;          // g = z chunks consumed so far (buffer g & 1, phase (g >> 1) & 1)
      for (uint32_t t = gid;; t += ngrp) {
        uint32_t c0, rows;
        const int run = my_tile(t, c0, rows);
        if (run < 0) break;
        if (run == 0) continue;
        if (kPair) mbar_wait_cluster(tempty(acc), acc_phase ^ 1); else mbar_wait(tempty(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * GF_BN);
        for (int kc = 0; kc < GF_KCH; ++kc, ++g) {
          const int ab = g & 1;
          if (kPair) mbar_wait_cluster(full_a(ab), (g >> 1) & 1); else mbar_wait(full_a(ab), (g >> 1) & 1);
          tc_fence_after();
          for (int h = 0; h < GF_H; ++h) {
            mbar_wait(full_w(stage), phase);
            tc_fence_after();
            const uint32_t sa = abuf0 + ab * GF_A_CHUNK + h * GF_A_REGION;
            const uint32_t sb = wring + stage * W_BYTES;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              if (kPair) tc_mma_bf16_pair(d_tmem, make_smem_desc(sa + ks * 32), make_smem_desc(sb + ks * 32), idesc, (kc | h | ks) ? 1u : 0u);
              else tc_mma_bf16(d_tmem, make_smem_desc(sa + ks * 32), make_smem_desc(sb + ks * 32), idesc, (kc | h | ks) ? 1u : 0u);
            }
            if (kPair) tc_commit_pair(empty_w(stage)); else tc_commit(empty_w(stage));
            if (++stage == W_STAGES) { stage = 0; phase ^= 1; }
          }
          if (kPair) tc_commit_pair(empty_a(ab)); else tc_commit(empty_a(ab));   // the z chunk may be overwritten once these MMAs retire
        }
        if (kPair) tc_commit_pair(tfull(acc)); else tc_commit(tfull(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp < 8) {
    // ===================================================== epilogue warps 4..7 (TMEM lane quadrant = warp & 3)
    const int qd = warp & 3;
    uint8_t* my_stg = smem_raw + (stg - smem_base) + qd * 32 * 128;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (uint32_t t = gid;; t += ngrp) {
      uint32_t c0, rows;
      const int run = my_tile(t, c0, rows);
      if (run < 0) break;
      if (run == 0) continue;
      mbar_wait(tfull(acc), acc_phase);
      tc_fence_after();
      const uint32_t row0 = c0 + qd * 32;
      const uint32_t rend = c0 + rows;
      const uint32_t my_row = row0 + lane;                 // the accumulator row this lane holds
      float4 srow4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (kEx && a.srow && my_row < rend) srow4 = __ldg(reinterpret_cast<const float4*>(a.srow) + my_row);
      // a block whose 32 rows all belong to this chunk leaves as ONE TMA tile store from the swizzled staging buffer (as in
      // gemm_tc.cu); partial blocks (chunk ends) keep the row stores — the rows past `rend` belong to another chunk
      const bool blk_tma = a.tma_store && row0 + 32 <= rend;
#pragma unroll 1
      for (int c = 0; c < a.m; c += 64) {
        const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(acc * GF_BN + c);
        if (a.tma_store) {
          if (lane == 0) tma_store_wait_read();             // the previous block's store has read the staging buffer
          __syncwarp();
        }
#pragma unroll
        for (int hlf = 0; hlf < 2; ++hlf) {
          uint4 adv[4];                                      // this row's 32 addend values, requested before the TMEM read waits
          if (kEx && a.addend) {
#pragma unroll
            for (int k8 = 0; k8 < 4; ++k8) {
              adv[k8] = make_uint4(0u, 0u, 0u, 0u);
              if (my_row < rend)
                adv[k8] = __ldg(reinterpret_cast<const uint4*>(a.addend + (int64_t)my_row * a.ldadd + c + hlf * 32 + k8 * 8));
            }
          }
          uint32_t r[32];
          tc_ld32(taddr + hlf * 32, r);
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = __uint_as_float(r[j + k]);
            const int cg = c + hlf * 32 + j;               // m % 64 == 0 (launcher): always in range
            if (a.bias) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(a.bias + cg));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(a.bias + cg + 4));
              v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
              v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
            }
            if (kEx && a.bvh) {                            // + sum_h s_ih bv_h / H: 4 broadcast LDS.128 from shared memory
              const uint8_t* sb = smem_raw + (sbv - smem_base) + cg * 2;
#pragma unroll
              for (int h = 0; h < GF_H; ++h) {
                const float sh = h == 0 ? srow4.x : (h == 1 ? srow4.y : (h == 2 ? srow4.z : srow4.w));
                float f[8];
                unpack_row16(*reinterpret_cast<const uint4*>(sb + h * GF_BN * 2), f, __nv_bfloat16());
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] += sh * f[k];
              }
            }
            if (kEx && a.addend) {                         // + the skip projection of this row (prefetched above; zeros past the end)
              float f[8];
              unpack_row16(adv[j >> 3], f, __nv_bfloat16());
#pragma unroll
              for (int k = 0; k < 8; ++k) v[k] += f[k];
            }
            Vec<__nv_bfloat16> o;
            o.from_float(v);
            *reinterpret_cast<uint4*>(my_stg + lane * 128 + (((hlf * 4 + (j >> 3)) ^ (lane & 7)) << 4)) = o.v;
          }
        }
        if (blk_tma) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) tma_store_2d(&map_o, smem_u32(my_stg), c, (int)row0);
          continue;
        }
        __syncwarp();
        const int piece = lane & 7;
#pragma unroll
        for (int r4 = 0; r4 < 32; r4 += 4) {
          const int rr = r4 + (lane >> 3);
          if (row0 + rr < rend) {
            const uint4 val = *reinterpret_cast<const uint4*>(my_stg + rr * 128 + ((piece ^ (rr & 7)) << 4));
            __nv_bfloat16* dst = a.out + (int64_t)(row0 + rr) * a.ldo + c + piece * 8;
            asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "r"(val.x), "r"(val.y),
                         "r"(val.z), "r"(val.w)
                         : "memory");
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {                                       // pair: the leader's MMA thread waits for both CTAs' epilogues
        if (kPair) mbar_arrive_cluster(mapa_u32(tempty(acc), 0)); else mbar_arrive(tempty(acc));
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (a.tma_store && lane == 0) tma_store_wait_all();      // shared memory stays valid until the last store has read it
  } else {
    // ===================================================== gather warps 8..: z chunks into the swizzled A operand
    asm volatile("setmaxnreg.inc.sync.aligned.u32 88;");
    const int gw = warp - 8;
    const int r4 = lane >> 3, p = lane & 7;
    float4* wrow_base = reinterpret_cast<float4*>(smem_raw + (wst - smem_base) + gw * GF_WST_WARP);   // [unit][entry][r4]
    const char* xl = a.x + p * 16;
    const char* zl = a.zero + p * 16;
    uint32_t g = 0;
    for (uint32_t t = gid;; t += ngrp) {
      uint32_t c0, rows;
      const int run = my_tile(t, c0, rows);
      if (run < 0) break;
      if (run == 0) continue;
      // ---- per tile: this lane's entry (row r4 of the unit, entry p) of both units: column index + 4 head weights
      int cl[2], len[2], b0[2], mlen[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const uint32_t rl = (uint32_t)(gw + GF_GW * u) * 4u + r4;          // row inside the tile
        const bool valid = rl < rows;
        const uint32_t row = valid ? c0 + rl : c0;
        const int b = __ldg(a.rowptr + row), e = __ldg(a.rowptr + row + 1);
        len[u] = valid ? e - b : 0;
        b0[u] = b;
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const bool has = p < len[u];
        const int pos = b0[u] + (has ? p : 0);
        int c = 0, pi = pos;
        float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!kSm) {
          if (has) {
            c = __ldg(a.col + pos);
            if (a.perm) pi = __ldg(a.perm + pos);
            w4 = __ldg(reinterpret_cast<const float4*>(a.alpha) + pi);
          }
        } else {
          // exact max-subtracted softmax over the (<= 8) entries of row r4: the 8 lanes of a row are lanes r4 * 8 .. r4 * 8 + 7.
          // (Same box: 9.91 ms layer forward against 10.32 with the separate b2g_gat_alpha kernel.  A software pipeline that
          // requested rowptr two tiles and col one tile ahead to shorten this prologue's rowptr -> col -> a_src chain made the
          // kernel 1 ms SLOWER — 8 more live registers spilled at the 80-register ceiling of a 768-thread CTA — and was dropped.)
          const uint32_t rl_ = (uint32_t)(gw + GF_GW * u) * 4u + r4;
          const uint32_t row = rl_ < rows ? c0 + rl_ : c0;
          float s4[GF_H] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
          if (has) {
            c = __ldg(a.col + pos);
            const float4 as4 = __ldg(reinterpret_cast<const float4*>(a.a_src + (uint64_t)(uint32_t)c * a.lda));
            const float4 ad4 = __ldg(reinterpret_cast<const float4*>(a.a_dst + (uint64_t)row * a.lda));
            float4 eb4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (a.ebias) eb4 = __ldg(reinterpret_cast<const float4*>(a.ebias) + pos);
            s4[0] = gf_lrelu(as4.x + ad4.x + eb4.x, a.slope); s4[1] = gf_lrelu(as4.y + ad4.y + eb4.y, a.slope);
            s4[2] = gf_lrelu(as4.z + ad4.z + eb4.z, a.slope); s4[3] = gf_lrelu(as4.w + ad4.w + eb4.w, a.slope);
          }
          float m4[GF_H], z4[GF_H], wv[GF_H];
#pragma unroll
          for (int h = 0; h < GF_H; ++h) {
            float m = s4[h];
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
            const float ex = has ? __expf(s4[h] - m) : 0.f;
            float z = ex;
            z += __shfl_xor_sync(0xffffffffu, z, 1);
            z += __shfl_xor_sync(0xffffffffu, z, 2);
            z += __shfl_xor_sync(0xffffffffu, z, 4);
            z += 1e-16f;
            m4[h] = m; z4[h] = z;
            wv[h] = ex * (1.0f / z);
          }
          if (a.p_drop > 0.f && has) {
            float sc[4];
            dropout_scale4(mix_epoch(a.seed, a.epoch), (uint64_t)pos, a.p_drop, sc);
#pragma unroll
            for (int h = 0; h < GF_H; ++h) wv[h] *= sc[h];
          }
          w4 = make_float4(wv[0], wv[1], wv[2], wv[3]);
          if (a.smax && p == 0 && rl_ < rows) {                  // same statistics as gat_alpha_kernel (empty row: 0 / 1e-16)
            const bool any = len[u] > 0;
            *reinterpret_cast<float4*>(a.smax + (uint64_t)row * GF_H) =
                any ? make_float4(m4[0], m4[1], m4[2], m4[3]) : make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(a.ssum + (uint64_t)row * GF_H) =
                any ? make_float4(z4[0], z4[1], z4[2], z4[3]) : make_float4(1e-16f, 1e-16f, 1e-16f, 1e-16f);
          }
        }
        cl[u] = c;
        wrow_base[(u * 8 + p) * 4 + r4] = w4;             // [unit][entry][row]: conflict-free LDS.128 in the FMA loop
        int ml = len[u];
        ml = max(ml, __shfl_xor_sync(0xffffffffu, ml, 8));
        ml = max(ml, __shfl_xor_sync(0xffffffffu, ml, 16));
        mlen[u] = ml;
      }
      __syncwarp();
      for (int kc = 0; kc < GF_KCH; ++kc, ++g) {
        const int ab = g & 1;
        mbar_wait(empty_a(ab), ((g >> 1) & 1) ^ 1);
        const char* xk = xl + kc * 128;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          float acc[GF_H][8];
#pragma unroll
          for (int h = 0; h < GF_H; ++h)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[h][k] = 0.f;
          const float4* wrow = wrow_base + u * 32 + r4;
          const int g0 = lane & 24;
          switch (min(mlen[u], 8)) {                                      // warp-uniform
#define B2G_CASE(KK) case KK: gf_gather<KK>(acc, xk, a.xrow_bytes, cl[u], len[u], g0, wrow, zl); break;
            B2G_CASE(1) B2G_CASE(2) B2G_CASE(3) B2G_CASE(4) B2G_CASE(5) B2G_CASE(6) B2G_CASE(7) B2G_CASE(8)
#undef B2G_CASE
            default: break;
          }
          for (int t = 8; !kSm && t < mlen[u]; ++t) {                     // rows longer than 8 entries (cold on meshes; never with kSm)
            const bool has = t < len[u];
            float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f);
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (has) {
              const int pos = b0[u] + t;
              const int c = __ldg(a.col + pos);
              const int pi = a.perm ? __ldg(a.perm + pos) : pos;
              w4 = __ldg(reinterpret_cast<const float4*>(a.alpha) + pi);
              v = ldg_row16(xk + (uint64_t)(uint32_t)c * a.xrow_bytes);
            }
            gf_fma(acc, w4, v);
          }
          const uint32_t rl = (uint32_t)(gw + GF_GW * u) * 4u + r4;
          const uint32_t dst = abuf0 + ab * GF_A_CHUNK + rl * 128u + (uint32_t)((p ^ (rl & 7)) << 4);
#pragma unroll
          for (int h = 0; h < GF_H; ++h) {
            Vec<__nv_bfloat16> o;
            o.from_float(acc[h]);
            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(dst + h * GF_A_REGION), "r"(o.v.x), "r"(o.v.y),
                         "r"(o.v.z), "r"(o.v.w)
                         : "memory");
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to tcgen05.mma
        __syncwarp();
        if (lane == 0) {                                                   // pair: release at cluster scope to the leader's MMA thread
          if (kPair) mbar_arrive_cluster(mapa_u32(full_a(ab), 0)); else mbar_arrive(full_a(ab));
        }
      }
      __syncwarp();                                                        // weight staging is rewritten by the next tile
    }
  }

  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();           // no CTA leaves (or frees TMEM) while its peer may still signal or read it
  if (warp == 1) {
    tc_fence_after();
    if (kPair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ------------------------------------------------------------------------------------------ attention weights
// alpha[p, h] for every entry p of the target-major CSR: exact max-subtracted segment softmax of
// leaky_relu(a_src[col_p, h] + a_dst[i, h]) (+ edge term) over the entries of row i (PyG softmax: denominator + 1e-16),
// times the attention-dropout keep scale.  One thread per (row, head); rows of <= 8 entries keep their scores in registers.
struct AlphaArgs {
  const float* a; uint32_t lda;              // fp32 [N, >= 8]: a_src | a_dst
  const int32_t* rowptr; const int32_t* col;
  const float* ebias;                        // fp32 [nnz, 4] added to the logits before the LeakyReLU (edge features) or NULL
  float* alpha;                              // [nnz, 4]
  float* smax; float* ssum;                  // [N, 4] or NULL
  uint32_t n_rows, row0;                     // rows [row0, row0 + n_rows) of the problem (all arrays are indexed globally)
  float slope, p_drop;
  uint64_t seed; const uint64_t* epoch;
};

__global__ void __launch_bounds__(256) gat_alpha_kernel(const AlphaArgs a) {
  const uint64_t total = (uint64_t)a.n_rows * GF_H;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t i = a.row0 + (uint32_t)(t >> 2);
    const int h = (int)(t & 3);
    const int b = __ldg(a.rowptr + i), e = __ldg(a.rowptr + i + 1);
    const int len = e - b;
    const float ad = __ldg(a.a + (uint64_t)i * a.lda + GF_H + h);
    float m = -INFINITY, zs = 0.f;
    if (len <= 8) {
      float s[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        s[k] = -INFINITY;
        if (k < len) {
          const uint32_t c = (uint32_t)__ldg(a.col + b + k);
          float v = __ldg(a.a + (uint64_t)c * a.lda + h) + ad;
          if (a.ebias) v += __ldg(a.ebias + (uint64_t)(b + k) * GF_H + h);
          s[k] = gf_lrelu(v, a.slope);
          m = fmaxf(m, s[k]);
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        s[k] = k < len ? __expf(s[k] - m) : 0.f;
        zs += s[k];
      }
      zs += 1e-16f;
      const float inv = 1.0f / zs;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (k < len) {
          float w = s[k] * inv;
          if (a.p_drop > 0.f) {
            float sc[4];
            dropout_scale4(mix_epoch(a.seed, a.epoch), (uint64_t)(b + k), a.p_drop, sc);
            w *= (h == 0 ? sc[0] : (h == 1 ? sc[1] : (h == 2 ? sc[2] : sc[3])));
          }
          a.alpha[(uint64_t)(b + k) * GF_H + h] = w;
        }
    } else {
      auto score = [&](int pos) {
        const uint32_t c = (uint32_t)__ldg(a.col + pos);
        float v = __ldg(a.a + (uint64_t)c * a.lda + h) + ad;
        if (a.ebias) v += __ldg(a.ebias + (uint64_t)pos * GF_H + h);
        return gf_lrelu(v, a.slope);
      };
      for (int pos = b; pos < e; ++pos) m = fmaxf(m, score(pos));
      for (int pos = b; pos < e; ++pos) zs += __expf(score(pos) - m);
      zs += 1e-16f;
      const float inv = 1.0f / zs;
      for (int pos = b; pos < e; ++pos) {
        float w = __expf(score(pos) - m) * inv;
        if (a.p_drop > 0.f) {
          float sc[4];
          dropout_scale4(mix_epoch(a.seed, a.epoch), (uint64_t)pos, a.p_drop, sc);
          w *= (h == 0 ? sc[0] : (h == 1 ? sc[1] : (h == 2 ? sc[2] : sc[3])));
        }
        a.alpha[(uint64_t)pos * GF_H + h] = w;
      }
    }
    if (a.smax) {
      a.smax[(uint64_t)i * GF_H + h] = len > 0 ? m : 0.f;
      a.ssum[(uint64_t)i * GF_H + h] = len > 0 ? zs : 1e-16f;
    }
  }
}

// ------------------------------------------------------------------------------------------ GATConv(edge_dim) edge rows
// PyG GATConv with edge_dim: remove_self_loops drops the given loops WITH their attributes, add_self_loops(fill_value =
// 'mean') gives every node's new loop the mean attribute of its incoming (non-loop) edges, 0 without any.  Output: the
// attributes in the order of the self-loop-replaced target-major CSR.  eid[p] < E: input edge eid[p]; eid[p] >= E: the new
// loop of this row.  The mean is sum / count in fp32 in edge order (the CSR is a stable sort: PyG's CPU scatter order).
__global__ void __launch_bounds__(256) edge_rows_sl_kernel(const float* __restrict__ ea, const int32_t* __restrict__ eid,
                                                           const int32_t* __restrict__ rowptr, int64_t n, int64_t E,
                                                           float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = rowptr[i], e = rowptr[i + 1];
    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
    int cnt = 0, loop_pos = -1;
    for (int p = b; p < e; ++p) {
      const int64_t id = eid[p];
      if (id < E) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(ea) + id);
        reinterpret_cast<float4*>(out)[p] = v;
        sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
        ++cnt;
      } else {
        loop_pos = p;
      }
    }
    if (loop_pos >= 0) {
      const float c = (float)(cnt > 0 ? cnt : 1);
      reinterpret_cast<float4*>(out)[loop_pos] = make_float4(sum.x / c, sum.y / c, sum.z / c, sum.w / c);
    }
  }
}

}  // namespace b2g

using namespace b2g;

extern "C" {

int b2g_gatw_gemm_supported(int64_t n, int H, int F, int C, int dt) {
  return (dt == B2G_BF16 && H == GF_H && F == 256 && C >= 64 && C <= GF_BN && (C % 64) == 0 && n >= 1 &&
          n < (1ll << 32) - (1ll << 25)) ? 1 : 0;
}

// B2G_GATW_PAIR in the environment at load time: 1 = CTA pairs whenever there are two tiles, 0 = never, unset = large problems
static const int g_gatw_tma_store = [] { const char* e = getenv("B2G_GATW_TMA_STORE"); return (e && e[0] == '0') ? 0 : 1; }();   // A/B
static const int g_gatw_pair = [] { const char* e = getenv("B2G_GATW_PAIR"); return !e ? -1 : (e[0] == '0' ? 0 : 1); }();

int b2g_gatw_gemm_ex(const void* x, int64_t ldx, const int32_t* rowptr, const int32_t* col, const int32_t* perm,
                     const float* alpha, const void* wp, int64_t ldw, const float* bias, const float* srow, const float* bvh,
                     const void* addend, int64_t ldadd, void* out, int64_t ldo, int64_t n_rows, int H, int F, int C, int dt,
                     int64_t band, void* stream);

int b2g_gatw_gemm(const void* x, int64_t ldx, const int32_t* rowptr, const int32_t* col, const int32_t* perm,
                  const float* alpha, const void* wp, int64_t ldw, const float* bias, void* out, int64_t ldo, int64_t n_rows,
                  int H, int F, int C, int dt, int64_t band, void* stream) {
  return b2g_gatw_gemm_ex(x, ldx, rowptr, col, perm, alpha, wp, ldw, bias, nullptr, nullptr, nullptr, 0, out, ldo, n_rows, H, F, C,
                          dt, band, stream);
}

// softmax inside the kernel (sm != NULL, alpha == NULL): see GfArgs
struct GfSoftmax {
  const float* a_src; const float* a_dst; int64_t lda; const float* ebias; float* smax; float* ssum;
  float slope, p_drop; uint64_t seed;
};

static int gatw_launch(const void* x, int64_t ldx, const int32_t* rowptr, const int32_t* col, const int32_t* perm,
                       const float* alpha, const GfSoftmax* sm, const void* wp, int64_t ldw, const float* bias, const float* srow,
                       const float* bvh, const void* addend, int64_t ldadd, void* out, int64_t ldo, int64_t n_rows, int H, int F,
                       int C, int dt, int64_t band, void* stream);

int b2g_gatw_gemm_ex(const void* x, int64_t ldx, const int32_t* rowptr, const int32_t* col, const int32_t* perm,
                     const float* alpha, const void* wp, int64_t ldw, const float* bias, const float* srow, const float* bvh,
                     const void* addend, int64_t ldadd, void* out, int64_t ldo, int64_t n_rows, int H, int F, int C, int dt,
                     int64_t band, void* stream) {
  if (!alpha && n_rows > 0) return B2G_E_ARG;
  return gatw_launch(x, ldx, rowptr, col, perm, alpha, nullptr, wp, ldw, bias, srow, bvh, addend, ldadd, out, ldo, n_rows, H, F, C,
                     dt, band, stream);
}

/* GATConv forward with the segment softmax INSIDE the fused kernel (rows of <= 8 entries): a_src fp32 indexed by global source
 * node (row stride lda), a_dst fp32 indexed by the rows of this call (same stride), edge_bias fp32 [nnz, 4] or NULL, smax / ssum
 * fp32 [n_rows, 4] (both or neither: statistics for the backward pass); dropout as b2g_gat_alpha.  max_row_len must be the longest
 * row of (rowptr, col) and <= 8, else B2G_E_UNSUPPORTED (use b2g_gat_alpha + b2g_gatw_gemm). */
int b2g_gatw_gemm_sm(const void* x, int64_t ldx, const int32_t* rowptr, const int32_t* col, const float* a_src, const float* a_dst,
                     int64_t lda, const float* edge_bias, float slope, float p_drop, uint64_t seed, float* smax, float* ssum,
                     const void* wp, int64_t ldw, const float* bias, void* out, int64_t ldo, int64_t n_rows, int64_t max_row_len,
                     int H, int F, int C, int dt, int64_t band, void* stream) {
  if (n_rows < 0 || (smax == nullptr) != (ssum == nullptr)) return B2G_E_ARG;
  if (n_rows == 0) return B2G_OK;
  if (max_row_len < 1 || max_row_len > 8) return B2G_E_UNSUPPORTED;
  if (!a_src || !a_dst) return B2G_E_ARG;
  if (!aligned16(a_src) || !aligned16(a_dst) || lda % 4 || lda < GF_H || lda >= (1ll << 32) || (edge_bias && !aligned16(edge_bias)) ||
      (smax && (!aligned16(smax) || !aligned16(ssum))))
    return B2G_E_ALIGN;
  const GfSoftmax sm{a_src, a_dst, lda, edge_bias, smax, ssum, slope, p_drop, seed};
  return gatw_launch(x, ldx, rowptr, col, nullptr, nullptr, &sm, wp, ldw, bias, nullptr, nullptr, nullptr, 0, out, ldo, n_rows, H, F, C,
                     dt, band, stream);
}

static int gatw_launch(const void* x, int64_t ldx, const int32_t* rowptr, const int32_t* col, const int32_t* perm,
                       const float* alpha, const GfSoftmax* sm, const void* wp, int64_t ldw, const float* bias, const float* srow,
                       const float* bvh, const void* addend, int64_t ldadd, void* out, int64_t ldo, int64_t n_rows, int H, int F,
                       int C, int dt, int64_t band, void* stream) {
  if (n_rows < 0) return B2G_E_ARG;
  if ((bvh != nullptr) != (srow != nullptr)) return B2G_E_ARG;
  if ((srow && !aligned16(srow)) || (bvh && !aligned16(bvh)) || (addend && (!aligned16(addend) || (ldadd * 2) % 16))) return B2G_E_ALIGN;
  if (n_rows == 0) return B2G_OK;
  if (!b2g_gatw_gemm_supported(n_rows, H, F, C, dt)) return B2G_E_UNSUPPORTED;
  if (!x || !rowptr || !col || (!alpha && !sm) || !wp || !out) return B2G_E_ARG;
  if (!aligned16(x) || (alpha && !aligned16(alpha)) || !aligned16(wp) || !aligned16(out) || (bias && !aligned16(bias)) || (ldx * 2) % 16 ||
      (ldw * 2) % 16 || (ldo * 2) % 16 || ldx * 2 >= (1ll << 32))
    return B2G_E_ALIGN;
  static bool attr_set[64] = {false};
  const int dev = current_device_slot();
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(gatw_gemm_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, GF_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gatw_gemm_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, GF_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gatw_gemm_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GF_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gatw_gemm_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GF_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gatw_gemm_kernel<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, GF_SMEM);
    if (e != cudaSuccess) return (int)e;
    attr_set[dev] = true;
  }
  GfArgs a{};
  if (!make_row_sched(n_rows, band, a.ord, GF_BM, 8192)) return B2G_E_UNSUPPORTED;
  // CTA pairs for problems with at least two tiles per SM (B2G_GATW_PAIR=0/1 in the environment at load time overrides: A/B runs)
  const bool pair = g_gatw_pair >= 0 ? (g_gatw_pair == 1 && a.ord.n_chunks >= 2) : false;
  CUtensorMap map_w;
  if (!tc_make_map_bf16(&map_w, wp, C, (int64_t)H * F, ldw, pair ? GF_BN / 2 : GF_BN)) return B2G_E_UNSUPPORTED;
  a.x = static_cast<const char*>(x); a.xrow_bytes = (uint32_t)(ldx * 2);
  a.rowptr = rowptr; a.col = col; a.perm = perm; a.alpha = alpha; a.bias = bias;
  a.zero = static_cast<const char*>(zero_row_ptr());
  if (!a.zero) return B2G_E_UNSUPPORTED;
  if (sm) {
    a.a_src = sm->a_src; a.a_dst = sm->a_dst; a.lda = (uint32_t)sm->lda; a.ebias = sm->ebias; a.smax = sm->smax; a.ssum = sm->ssum;
    a.slope = sm->slope; a.p_drop = sm->p_drop; a.seed = sm->seed; a.epoch = dropout_epoch_ptr();
  }
  a.srow = srow; a.bvh = bvh; a.addend = static_cast<const __nv_bfloat16*>(addend); a.ldadd = ldadd;
  a.out = static_cast<__nv_bfloat16*>(out); a.ldo = ldo; a.n_rows = (uint32_t)n_rows; a.m = C;
  CUtensorMap map_o = map_w;                                   // placeholder when the row-store epilogue is used
  a.tma_store = (g_gatw_tma_store && tc_make_map_bf16(&map_o, out, n_rows, C, ldo, 32)) ? 1 : 0;
  if (pair) {
    const unsigned pairs_wanted = (a.ord.n_chunks + 1) / 2;
    const unsigned pairs = pairs_wanted < (unsigned)(B2G_NUM_SMS / 2) ? pairs_wanted : (unsigned)(B2G_NUM_SMS / 2);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(GF_THREADS);
    cfg.dynamicSmemBytes = GF_SMEM;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const cudaError_t e = sm ? cudaLaunchKernelEx(&cfg, gatw_gemm_kernel<true, true, false>, map_w, map_o, a)
                             : cudaLaunchKernelEx(&cfg, gatw_gemm_kernel<true, false, true>, map_w, map_o, a);
    if (e != cudaSuccess) return (int)e;
  } else {
    const unsigned grid = a.ord.n_chunks < (uint32_t)B2G_NUM_SMS ? a.ord.n_chunks : (unsigned)B2G_NUM_SMS;
    const bool ex = srow || bvh || addend;
    if (sm) gatw_gemm_kernel<false, true, false><<<grid, GF_THREADS, GF_SMEM, (cudaStream_t)stream>>>(map_w, map_o, a);
    else if (ex) gatw_gemm_kernel<false, false, true><<<grid, GF_THREADS, GF_SMEM, (cudaStream_t)stream>>>(map_w, map_o, a);
    else gatw_gemm_kernel<false, false, false><<<grid, GF_THREADS, GF_SMEM, (cudaStream_t)stream>>>(map_w, map_o, a);
  }
  count_launch();
  return cuda_status();
}

int b2g_gat_alpha(const float* a_srcdst, int64_t lda, const int32_t* rowptr, const int32_t* col, const float* edge_bias,
                  int64_t row0, int64_t n, int H, float slope, float p_drop, uint64_t seed, float* alpha, float* smax, float* ssum,
                  void* stream) {
  if (n < 0 || row0 < 0 || row0 + n >= (1ll << 32) || H != GF_H) return (n < 0 || row0 < 0) ? B2G_E_ARG : B2G_E_UNSUPPORTED;
  if (n == 0) return B2G_OK;
  if (!a_srcdst || !rowptr || !col || !alpha || (smax == nullptr) != (ssum == nullptr)) return B2G_E_ARG;
  if (lda < 2 * GF_H || lda >= (1ll << 32) || n >= (1ll << 32)) return B2G_E_SHAPE;
  AlphaArgs a{a_srcdst, (uint32_t)lda, rowptr, col, edge_bias, alpha, smax, ssum, (uint32_t)n, (uint32_t)row0, slope, p_drop, seed,
              dropout_epoch_ptr()};
  const int64_t want = ceil_div(n * GF_H, 256);
  const int64_t cap = (int64_t)B2G_NUM_SMS * 16;
  gat_alpha_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(a);
  count_launch();
  return cuda_status();
}

int b2g_edge_rows_sl(const float* edge_attr, int64_t E, const int32_t* eid, const int32_t* rowptr, int64_t n, float* out,
                     void* stream) {
  if (n < 0 || E < 0) return B2G_E_ARG;
  if (n == 0) return B2G_OK;
  if (!eid || !rowptr || !out || (E > 0 && !edge_attr)) return B2G_E_ARG;
  if ((edge_attr && !aligned16(edge_attr)) || !aligned16(out)) return B2G_E_ALIGN;
  const int64_t want = ceil_div(n, 256);
  const int64_t cap = (int64_t)B2G_NUM_SMS * 16;
  edge_rows_sl_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(edge_attr, eid, rowptr, n, E, out);
  count_launch();
  return cuda_status();
}

}  // extern "C"
