// gemm_tc.cu — K6 (tensor-core path): the per-node Linear  Y[n,m] = act(rs[n] * X[n,k] W[m,k]^T + b[m])
// as a hand-written sm_100a kernel: TMA (cp.async.bulk.tensor, 128B swizzle) stages bf16 operand
// tiles in shared memory, ONE elected thread issues tcgen05.mma (kind::f16, 128 x 256 x 16, fp32
// accumulate) into TMEM, tcgen05.commit hands completion to mbarriers, four epilogue warps read the
// accumulator back with tcgen05.ld and apply row-scale / bias / ReLU / the fp32 aux split.
// Replaces F.linear inside GCNConv / GATConv / GINConv / TransformerConv.
//
// Shape of the problem: n = 10^6..10^8 rows, k = m = 256 (..3328 fused outputs): arithmetic
// intensity at bf16 is ~k/2 flop/B < the B200 ridge (~210), so the kernel is HBM-bound and is laid
// out to stream X once and Y once: persistent CTAs (one per SM), a 4-stage TMA ring, two TMEM
// accumulators (2 x 256 columns) so the epilogue of tile t overlaps the MMAs of tile t+1; W (<= 1.7 MB)
// stays L2-resident.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner,
// warps 2..9 = epilogue (TMEM lane quadrant = warp_idx % 4; the two warps of a quadrant take alternate
// 64-column chunks).  The epilogue is instruction-bound with one warp per scheduler, hence eight warps
// and a specialised chunk body per (row-scale, bias, ReLU) combination.
#include <cstdlib>
#include "rows.cuh"
#include "tc_ptx.cuh"

namespace b2g {

constexpr int TC_BM = 128;        // rows of X per tile == UMMA_M (cta_group::1)
constexpr int TC_BN = 256;        // output columns per tile == UMMA_N
constexpr int TC_BK = 64;         // bf16 elements per k-block = 128 bytes = one swizzle atom row
constexpr int TC_THREADS = 320;    // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two per TMEM lane quadrant)
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;   // 16 KB
constexpr int TC_B_BYTES = TC_BN * TC_BK * 2;   // 32 KB
constexpr int TC_TMEM_COLS = 512; // two 256-column fp32 accumulators
constexpr int TC_STG_PITCH = 128; // epilogue staging: 32 rows x 128 B per warp, 16-byte pieces XOR-swizzled by row
constexpr int TC_STG_BYTES = TC_EPI_WARPS * 32 * TC_STG_PITCH;   // 32 KB: a private 32-row x 128-byte buffer per epilogue warp
                                                                  // (the streaming plan holds two per warp: the TMA store of one
                                                                  // chunk reads its buffer while the next chunk is packed)
constexpr int TC_BAR_BYTES = 128;                     // (2*5 + 6) x 8 bytes
// Two shared-memory plans (227 KB = 232448 B per CTA on sm_100):
//  kBRes = true  (k <= 256: GCNConv / GINConv / plain Linear, and the wide outputs of the attention layers, one 256-row
//                block of W per CTA group): the W block is loaded ONCE per CTA and stays resident (k/64 x 32 KB); only X
//                tiles stream through a 4-deep ring of 16 KB stages.
//  kBRes = false (wide fused outputs: GAT 4F+8, Transformer 13F): a 4-deep ring of (X 16 KB + W 32 KB) stages.
constexpr int TC_RES_STAGES = 4;
constexpr int TC_STR_STAGES = 3;
constexpr int TC_RES_KB_MAX = 4;
constexpr int TC_SMEM_RES = TC_RES_KB_MAX * TC_B_BYTES + TC_RES_STAGES * TC_A_BYTES + TC_STG_BYTES + TC_BAR_BYTES;   // 229504
constexpr int TC_SMEM_STR = TC_STR_STAGES * (TC_A_BYTES + TC_B_BYTES) + 2 * TC_STG_BYTES + TC_BAR_BYTES;              // 213120
constexpr int TC_SMEM_DUAL = TC_STR_STAGES * (2 * TC_A_BYTES + TC_B_BYTES) + TC_STG_BYTES + TC_BAR_BYTES;                // 229504
static_assert(TC_SMEM_DUAL + 1024 <= 232448, "shared-memory plan 2 exceeds 227 KB");
static_assert(TC_SMEM_RES + 1024 <= 232448 && TC_SMEM_STR + 1024 <= 232448, "shared-memory plan exceeds 227 KB (1 KB is charged for the 1024-byte alignment)");

// One 32-column slice of the accumulator row held by this lane: fused row-scale / bias / ReLU, pack to
// bf16 and park it in the warp's staging buffer (16-byte pieces XOR-swizzled by row: conflict-free).
template <bool kRS, bool kBias, bool kRelu>
__device__ __forceinline__ void epi_pack(const uint32_t (&r)[32], float rs, const float* __restrict__ bias, int cg0,
                                         uint8_t* stg, int lane, int hlf, const uint4* mk4) {
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __uint_as_float(r[j + k]);
    if (kRS) {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] *= rs;
    }
    if (kBias) {   // same address on every lane: two broadcast 16-byte loads
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + cg0 + j));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + cg0 + j + 4));
      v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
      v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
    }
    if (kRelu) {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = fmaxf(v[k], 0.f);
    }
    if (mk4) {                                   // warp-uniform: this row's 8 mask values (prefetched), zero where mask <= 0
      float mk[8];
      unpack_row16(mk4[j >> 3], mk, __nv_bfloat16());
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = mk[k] > 0.f ? v[k] : 0.f;
    }
    Vec<__nv_bfloat16> o;
    o.from_float(v);
    *reinterpret_cast<uint4*>(stg + lane * TC_STG_PITCH + (((hlf * 4 + (j >> 3)) ^ (lane & 7)) << 4)) = o.v;
  }
}

struct TcParams {
  int64_t n;
  int m, m_main, k;
  const float* bias;
  const float* row_scale;
  __nv_bfloat16* Y;
  int64_t ldy;
  float* aux;
  int64_t ldaux;
  int act;
  int tma_store;   // full 64-column chunks leave through map_y (cp.async.bulk.tensor stores) instead of LDS + 128-byte row stores
  const __nv_bfloat16* mask;   // [n, m_main] or NULL: Y = 0 where mask <= 0 (ReLU backward fused into the dgrad GEMM: mask = the
  int64_t ldmask;              // ReLU's output), applied after row scale / bias / activation
};

// kPlan: 0 = streaming (X + W k-blocks through the ring), 1 = resident W block (k <= 256), 2 = streaming with TWO row tiles per
// W k-block (m <= 256, long k: the dgrad GEMMs of the attention layers, k = 1032 .. 3336).  Plan 0 re-reads the whole W (up to
// 1.7 MB) from L2 for every 128 rows: ncu on the k = 3336 dgrad showed 69 GB of DRAM reads at 3.9 TB/s with ~200 GB of L2 -> SM
// traffic.  Plan 2 multiplies each W k-block with two X tiles into the two TMEM accumulators (no accumulator double buffering:
// the epilogue of a pair is 3 us against a 40 us main loop), which halves the W traffic per row.
// kMask: the ReLU-backward mask epilogue (b2g_linear_fwd_masked).  A separate instantiation: merely carrying the (unused) mask
// code made the plain kernel 10 % slower (same box: 256 -> 256 1.79-1.86 ms against 1.62-1.67).
template <int kPlan, bool kMask = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_linear_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ CUtensorMap map_y, const TcParams p) {
  constexpr bool kBRes = kPlan == 1, kDual = kPlan == 2;
  constexpr int STAGES = kBRes ? TC_RES_STAGES : TC_STR_STAGES;
  constexpr int STAGE_BYTES = kBRes ? TC_A_BYTES : ((kDual ? 2 : 1) * TC_A_BYTES + TC_B_BYTES);
  constexpr int B_OFF = (kDual ? 2 : 1) * TC_A_BYTES;       // W block inside a streaming stage
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if (smem_base & 1023u) __trap();                          // SWIZZLE_128B tiles need 1024-byte alignment
  const uint32_t bres = smem_base;                          // resident W (kBRes only)
  const uint32_t ring = smem_base + (kBRes ? TC_RES_KB_MAX * TC_B_BYTES : 0);
  const uint32_t stg = ring + STAGES * STAGE_BYTES;
  constexpr int NBUF = kPlan == 0 ? 2 : 1;                  // staging buffers per epilogue warp
  const uint32_t bars = stg + NBUF * TC_STG_BYTES;
  // barrier slots (8 bytes each): full[S], empty[S], tmem_full[2], tmem_empty[2], b_full; then the TMEM base word
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * STAGES + 2 + a); };
  const uint32_t bfull_bar = bars + 8u * (2 * STAGES + 4);
  const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 5);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row_tiles = (p.n + TC_BM - 1) / TC_BM;
  const int col_tiles = (p.m + TC_BN - 1) / TC_BN;
  const int64_t tiles = row_tiles * col_tiles;
  const int kblocks = (p.k + TC_BK - 1) / TC_BK;
  // j-th tile of this CTA.  With at least one row tile per CTA a CTA takes ALL column tiles of a row tile back to back: the
  // X tile is then read from DRAM once and from L2 for the other column tiles (interleaved over CTAs the 256 -> 1024 GEMM
  // read x twice: ncu 9.9 GB for 5.1 GB).  Small problems keep the interleaved order so that every SM gets a tile.
  const bool row_major = row_tiles >= (int64_t)gridDim.x;
  // Resident plan (k <= 256): the grid is col_tiles groups of CTAs; a CTA keeps ONE 256-row block of W in shared memory and
  // sweeps the row tiles of its group, so a tile costs 64 KB of TMA traffic (X) instead of 192 KB (X + the W block again):
  // the streamed 256 -> 1024 GEMM moved 60 GB from L2 to the SMs for 25.6 GB of HBM traffic.  The groups walk the rows in
  // step, so the col_tiles reads of an X tile meet in L2.
  const int my_ct = kBRes ? (int)(blockIdx.x % (unsigned)col_tiles) : 0;
  auto tile_at = [&](int64_t j, int64_t& rt, int& ct) -> bool {
    if (kBRes) {
      rt = (int64_t)(blockIdx.x / (unsigned)col_tiles) + j * (int64_t)(gridDim.x / (unsigned)col_tiles);
      ct = my_ct;
      return rt < row_tiles;
    }
    if (kDual) {                                             // tiles 2P, 2P + 1 of pair P = blockIdx.x + (j / 2) * gridDim.x
      rt = 2 * ((int64_t)blockIdx.x + (j >> 1) * (int64_t)gridDim.x) + (j & 1);
      ct = 0;
      return rt < row_tiles;
    }
    if (row_major) {
      const int64_t g = j / col_tiles;
      rt = blockIdx.x + g * gridDim.x;
      ct = (int)(j - g * col_tiles);
      return rt < row_tiles;
    }
    const int64_t t = blockIdx.x + j * gridDim.x;
    rt = t / col_tiles;
    ct = (int)(t - rt * col_tiles);
    return t < tiles;
  };

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), TC_EPI_WARPS);   // one arrive per epilogue warp
    }
    mbar_init(bfull_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM allocation is warp-collective; the same warp frees it
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================================================== TMA producer (one elected lane)
    if (lane == 0) {
      if (kBRes) {                                           // W: loaded once, resident for every tile of this CTA
        mbar_expect_tx(bfull_bar, (uint32_t)kblocks * TC_B_BYTES);
        for (int kb = 0; kb < kblocks; ++kb) tma_load_2d(bres + kb * TC_B_BYTES, &map_b, bfull_bar, kb * TC_BK, my_ct * TC_BN);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t j = 0;; j += kDual ? 2 : 1) {
        int64_t rt;
        int ct;
        if (!tile_at(j, rt, ct)) break;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = ring + stage * STAGE_BYTES;
          mbar_expect_tx(full_bar(stage), STAGE_BYTES);
          tma_load_2d(sa, &map_a, full_bar(stage), kb * TC_BK, (int)(rt * TC_BM));
          if (kDual) tma_load_2d(sa + TC_A_BYTES, &map_a, full_bar(stage), kb * TC_BK, (int)((rt + 1) * TC_BM));   // rows >= n: zeros
          if (!kBRes) tma_load_2d(sa + B_OFF, &map_b, full_bar(stage), kb * TC_BK, ct * TC_BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (one elected lane)
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(TC_BM, TC_BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      if (kBRes) {
        mbar_wait(bfull_bar, 0);
        tc_fence_after();
      }
      for (int64_t j = 0;; j += kDual ? 2 : 1) {
        int64_t rt;
        int ct;
        if (!tile_at(j, rt, ct)) break;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);          // epilogue has drained this accumulator
        if (kDual) mbar_wait(tempty_bar(1), acc_phase ^ 1); // (acc == 0 here: a pair takes both)
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * TC_BN);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(full_bar(stage), phase);                 // TMA bytes have landed
          tc_fence_after();
          const uint32_t sa = ring + stage * STAGE_BYTES;
          const uint32_t sb = kBRes ? (bres + kb * TC_B_BYTES) : (sa + B_OFF);
#pragma unroll
          for (int ks = 0; ks < TC_BK / 16; ++ks) {          // UMMA_K = 16 bf16 = 32 bytes inside the swizzle row
            const uint64_t ad = make_smem_desc(sa + ks * 32);
            const uint64_t bd = make_smem_desc(sb + ks * 32);
            tc_mma_bf16(d_tmem, ad, bd, idesc, (kb | ks) ? 1u : 0u);
          }
          if (kDual) {
#pragma unroll
            for (int ks = 0; ks < TC_BK / 16; ++ks) {        // the second row tile against the same W k-block
              const uint64_t ad = make_smem_desc(sa + TC_A_BYTES + ks * 32);
              const uint64_t bd = make_smem_desc(sb + ks * 32);
              tc_mma_bf16(tmem_base + (uint32_t)TC_BN, ad, bd, idesc, (kb | ks) ? 1u : 0u);
            }
          }
          tc_commit(empty_bar(stage));                       // frees the smem slot when these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(tfull_bar(acc));                           // accumulator complete -> epilogue
        if (kDual) {
          tc_commit(tfull_bar(1));
          acc_phase ^= 1;
        } else if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================================================== epilogue warps 2..9
    const int q = warp & 3;                                   // TMEM lane quadrant this warp may read
    const int half = (warp - 2) >> 2;                         // which of the quadrant's two warps
    uint8_t* const my_stg0 = smem_raw + (stg - smem_base) + (warp - 2) * NBUF * 32 * TC_STG_PITCH;
    uint32_t sbuf = 0;                                        // staging buffer of the next chunk
    const bool vec_ok = (p.m_main % 8) == 0;                  // 16-byte pieces never straddle the Y / aux split
    const int flags = (p.row_scale ? 1 : 0) | (p.bias ? 2 : 0) | (p.act == 1 ? 4 : 0);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int64_t j = 0;; ++j) {
      int64_t rt;
      int ct;
      if (!tile_at(j, rt, ct)) break;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const int64_t row0 = rt * TC_BM + q * 32;
      const int64_t row = row0 + lane;
      const bool row_ok = row < p.n;
      const float rs = (p.row_scale && row_ok) ? __ldg(p.row_scale + row) : 1.0f;
      const __nv_bfloat16* mrow = (kMask && p.mask) ? p.mask + (row_ok ? row : 0) * p.ldmask : nullptr;   // rows >= n: any valid row
      const int col0 = ct * TC_BN;
      const int ncols = min(TC_BN, p.m - col0);
#pragma unroll 1
      for (int c = half * 64; c < ncols; c += 128) {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * TC_BN + c);
        const int cg0 = col0 + c;
        if (vec_ok && cg0 + 64 <= p.m_main) {
          // ---- fast path: 64 full columns -> bf16 -> swizzled staging -> one TMA tile store (or 128-byte row stores).
          // ncu on the 256 -> 1024 GEMM (write-heavy): the LSU data pipe was 81 % busy with STS + LDS + STG of every output
          // byte; the staging buffer IS the SWIZZLE_128B image of a 32-row x 64-column box, so the copy-out is one bulk store.
          uint8_t* const my_stg = my_stg0 + sbuf * (32 * TC_STG_PITCH);
          uint32_t r2[2][32];
          tc_ld32_issue(taddr, r2[0]);                        // both halves in flight: one TMEM latency per chunk
          tc_ld32_issue(taddr + 32, r2[1]);
          uint4 mkv[2][4];                                    // the row's 64 mask values, requested before anything waits: a load
          if (kMask && mrow) {                                // per 8 columns inside epi_pack ran the epilogue at one L2 latency each
#pragma unroll
            for (int t8 = 0; t8 < 8; ++t8) mkv[t8 >> 2][t8 & 3] = __ldg(reinterpret_cast<const uint4*>(mrow + cg0 + t8 * 8));
          }
          if (p.tma_store) {
            if (lane == 0) {                                  // the store that last read THIS buffer has finished reading
              if (NBUF == 2) tma_store_wait_read1(); else tma_store_wait_read();
            }
            __syncwarp();
          }
          tc_wait_ld();
#pragma unroll
          for (int hlf = 0; hlf < 2; ++hlf) {
            const uint32_t (&r)[32] = r2[hlf];
            switch (flags) {   // warp-uniform; each case is a straight-line body without per-element predicates
              case 0: epi_pack<false, false, false>(r, rs, p.bias, cg0 + hlf * 32, my_stg, lane, hlf, (kMask && mrow) ? mkv[hlf] : nullptr); break;
              case 1: epi_pack<true, false, false>(r, rs, p.bias, cg0 + hlf * 32, my_stg, lane, hlf, (kMask && mrow) ? mkv[hlf] : nullptr); break;
              case 2: epi_pack<false, true, false>(r, rs, p.bias, cg0 + hlf * 32, my_stg, lane, hlf, (kMask && mrow) ? mkv[hlf] : nullptr); break;
              case 3: epi_pack<true, true, false>(r, rs, p.bias, cg0 + hlf * 32, my_stg, lane, hlf, (kMask && mrow) ? mkv[hlf] : nullptr); break;
              case 4: epi_pack<false, false, true>(r, rs, p.bias, cg0 + hlf * 32, my_stg, lane, hlf, (kMask && mrow) ? mkv[hlf] : nullptr); break;
              case 5: epi_pack<true, false, true>(r, rs, p.bias, cg0 + hlf * 32, my_stg, lane, hlf, (kMask && mrow) ? mkv[hlf] : nullptr); break;
              case 6: epi_pack<false, true, true>(r, rs, p.bias, cg0 + hlf * 32, my_stg, lane, hlf, (kMask && mrow) ? mkv[hlf] : nullptr); break;
              default: epi_pack<true, true, true>(r, rs, p.bias, cg0 + hlf * 32, my_stg, lane, hlf, (kMask && mrow) ? mkv[hlf] : nullptr); break;
            }
          }
          if (p.tma_store) {
            fence_proxy_async_smem();                         // generic-proxy STS -> visible to the async proxy
            __syncwarp();
            if (lane == 0 && row0 < p.n) tma_store_2d(&map_y, smem_u32(my_stg), cg0, (int)row0);   // rows >= n are clipped
            sbuf = (sbuf + 1) & (NBUF - 1);
            continue;
          }
          __syncwarp();
          const int piece = lane & 7;                         // 16-byte piece inside the 128-byte row segment
#pragma unroll
          for (int r4 = 0; r4 < 32; r4 += 4) {
            const int rr = r4 + (lane >> 3);
            if (row0 + rr < p.n) {
              const uint4 val = *reinterpret_cast<const uint4*>(my_stg + rr * TC_STG_PITCH + ((piece ^ (rr & 7)) << 4));
              __nv_bfloat16* dst = p.Y + (row0 + rr) * p.ldy + cg0 + piece * 8;
              asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "r"(val.x), "r"(val.y),
                           "r"(val.z), "r"(val.w)
                           : "memory");
            }
          }
          __syncwarp();
        } else {
          // ---- ragged / split chunk (tail columns, fp32 aux columns): element-wise
          for (int hlf = 0; hlf < 2; ++hlf) {
            const int cc = c + hlf * 32;
            if (cc >= ncols) break;                           // warp-uniform
            uint32_t r[32];
            tc_ld32(taddr + hlf * 32, r);
            if (row_ok) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {                    // static indices keep r[] in registers
                const int cg = col0 + cc + j;
                if (cg < p.m) {
                  float x = __uint_as_float(r[j]);
                  if (p.row_scale) x *= rs;
                  if (p.bias) x += __ldg(p.bias + cg);
                  if (p.act == 1) x = fmaxf(x, 0.f);
                  if (kMask && mrow && cg < p.m_main && !(__bfloat162float(mrow[cg]) > 0.f)) x = 0.f;
                  if (cg < p.m_main) p.Y[row * p.ldy + cg] = __float2bfloat16_rn(x);
                  else p.aux[row * p.ldaux + (cg - p.m_main)] = x;
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (p.tma_store && lane == 0) tma_store_wait_all();       // shared memory stays valid until the last store has read it
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
  }
}

// =====================================================================================================
// fp32 Linear on the tensor cores: 3 x TF32 split ("3xTF32").  The parity gate for fp32 is 1e-5, which a single
// TF32 pass (10-bit mantissa, ~1e-3) cannot meet and which costs 35 ms per cfg4 layer on the FP32 pipe.  Every fp32
// operand is split exactly into hi = the TF32-representable head (low 13 mantissa bits cleared) and lo = x - hi
// (exact in fp32), and D += A_hi*B_hi + A_lo*B_hi + A_hi*B_lo with fp32 accumulation in TMEM; the dropped term
// A_lo*B_lo and the truncation of lo to TF32 are ~2^-22 relative (tests: <= 1e-6 of the max-abs reference).
// X tiles are split in shared memory by four converter warps between the TMA load and the MMAs (in place for hi,
// a second buffer for lo, then fence.proxy.async so the tensor core's async proxy sees the generic-proxy writes);
// W is split once per call by a small kernel into a [2m, k] workspace and streamed through TMA like X.
constexpr int T3_BK = 32;                               // fp32 elements per k-block = 128 bytes
constexpr int T3_A = TC_BM * T3_BK * 4;                 // 16 KB
constexpr int T3_B = TC_BN * T3_BK * 4;                 // 32 KB
constexpr int T3_STAGE = 2 * T3_A + 2 * T3_B;           // A_hi, A_lo, B_hi, B_lo = 96 KB
constexpr int T3_STAGES = 2;
constexpr int T3_KCHUNK = 256;                           // longest reduction accumulated in TMEM in one go (fp32 path)
constexpr int T3_SMEM = T3_STAGES * T3_STAGE + 128;
static_assert(T3_SMEM + 1024 <= 232448, "3xTF32 shared-memory plan exceeds 227 KB");

__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}

__global__ void __launch_bounds__(256) split_tf32_kernel(const float* __restrict__ w, int64_t ldw, int m, int k,
                                                         float* __restrict__ out /*[2m,k]*/) {
  const int64_t total = (int64_t)m * k;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(t / k), c = (int)(t - (int64_t)r * k);
    const float v = w[(int64_t)r * ldw + c];
    const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    out[t] = hi;
    out[total + t] = v - hi;
  }
}

struct T3Params {
  int64_t n;
  int m, m_main, k;
  const float* bias;
  const float* row_scale;
  float* Y;
  int64_t ldy;
  float* aux;
  int64_t ldaux;
  int act;
  int accum;                 // Y += result (k-chunked reductions: fp32 adds between chunks, see tc_linear_fwd_tf32x3)
};

__global__ void __launch_bounds__(TC_THREADS, 1)
tc_linear_tf32x3_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const T3Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if (smem_base & 1023u) __trap();
  const uint32_t bars = smem_base + T3_STAGES * T3_STAGE;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto conv_bar = [&](int s) { return bars + 8u * (T3_STAGES + s); };
  auto empty_bar = [&](int s) { return bars + 8u * (2 * T3_STAGES + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (3 * T3_STAGES + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (3 * T3_STAGES + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (3 * T3_STAGES + 4);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_base));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row_tiles = (p.n + TC_BM - 1) / TC_BM;
  const int col_tiles = (p.m + TC_BN - 1) / TC_BN;
  const int64_t tiles = row_tiles * col_tiles;
  const int kblocks = (p.k + T3_BK - 1) / T3_BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    for (int s = 0; s < T3_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(conv_bar(s), 4);      // one arrive per converter warp
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);    // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================================================== TMA producer: A (raw fp32), B_hi, B_lo
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int64_t rt = t / col_tiles;
        const int ct = (int)(t - rt * col_tiles);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * T3_STAGE;
          mbar_expect_tx(full_bar(stage), T3_A + 2 * T3_B);
          tma_load_2d(sa, &map_a, full_bar(stage), kb * T3_BK, (int)(rt * TC_BM));
          tma_load_2d(sa + 2 * T3_A, &map_b, full_bar(stage), kb * T3_BK, ct * TC_BN);                 // W_hi rows
          tma_load_2d(sa + 2 * T3_A + T3_B, &map_b, full_bar(stage), kb * T3_BK, p.m + ct * TC_BN);   // W_lo rows
          if (++stage == T3_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer: 3 products per k-block
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(TC_BM, TC_BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * TC_BN);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(conv_bar(stage), phase);                 // X tile split into hi / lo by the converter warps
          tc_fence_after();
          const uint32_t a_hi = smem_base + stage * T3_STAGE, a_lo = a_hi + T3_A;
          const uint32_t b_hi = a_hi + 2 * T3_A, b_lo = b_hi + T3_B;
#pragma unroll
          for (int ks = 0; ks < T3_BK / 8; ++ks) {           // UMMA_K = 8 tf32 = 32 bytes
            const uint64_t dah = make_smem_desc(a_hi + ks * 32), dal = make_smem_desc(a_lo + ks * 32);
            const uint64_t dbh = make_smem_desc(b_hi + ks * 32), dbl = make_smem_desc(b_lo + ks * 32);
            tc_mma_tf32(d_tmem, dal, dbh, idesc, (kb | ks) ? 1u : 0u);   // small terms first
            tc_mma_tf32(d_tmem, dah, dbl, idesc, 1u);
            tc_mma_tf32(d_tmem, dah, dbh, idesc, 1u);
          }
          tc_commit(empty_bar(stage));
          if (++stage == T3_STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(tfull_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp < 6) {
    // ===================================================== converter warps 2..5: X -> (hi in place, lo)
    const int tid = threadIdx.x - 64;                        // 0..127
    int stage = 0;
    uint32_t phase = 0;
    for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(full_bar(stage), phase);
        float4* hi = reinterpret_cast<float4*>(smem_raw + (size_t)stage * T3_STAGE);
        float4* lo = reinterpret_cast<float4*>(smem_raw + (size_t)stage * T3_STAGE + T3_A);
#pragma unroll
        for (int i = 0; i < T3_A / 16 / 128; ++i) {          // 8 float4 per thread; elementwise, layout-agnostic
          const int idx = i * 128 + tid;
          const float4 v = hi[idx];
          float4 h, l;
          h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); l.x = v.x - h.x;
          h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); l.y = v.y - h.y;
          h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); l.z = v.z - h.z;
          h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); l.w = v.w - h.w;
          hi[idx] = h;
          lo[idx] = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to tcgen05.mma
        __syncwarp();
        if (lane == 0) mbar_arrive(conv_bar(stage));
        if (++stage == T3_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================================================== epilogue warps 6..9 (TMEM lane quadrant = warp % 4)
    const int q = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
      const int64_t rt = t / col_tiles;
      const int ct = (int)(t - rt * col_tiles);
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const int64_t row = rt * TC_BM + q * 32 + lane;
      const bool row_ok = row < p.n;
      const float rs = (p.row_scale && row_ok) ? __ldg(p.row_scale + row) : 1.0f;
      const int col0 = ct * TC_BN;
      const int ncols = min(TC_BN, p.m - col0);
#pragma unroll 1
      for (int c = 0; c < ncols; c += 32) {
        uint32_t r[32];
        tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * TC_BN + c), r);
        if (row_ok) {
          const int cg0 = col0 + c;
          if (cg0 + 32 <= p.m_main) {                          // a full 128-byte line of this row
            float* dst = p.Y + row * p.ldy + cg0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 o;
              float* of = reinterpret_cast<float*>(&o);
              float4 old4 = make_float4(0.f, 0.f, 0.f, 0.f);
              if (p.accum) old4 = *reinterpret_cast<const float4*>(dst + j);
              const float* oldf = reinterpret_cast<const float*>(&old4);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float x = __uint_as_float(r[j + e]);
                if (p.row_scale) x *= rs;
                x += oldf[e];
                if (p.bias) x += __ldg(p.bias + cg0 + j + e);
                if (p.act == 1) x = fmaxf(x, 0.f);
                of[e] = x;
              }
              *reinterpret_cast<float4*>(dst + j) = o;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int cg = cg0 + j;
              if (cg < p.m) {
                float x = __uint_as_float(r[j]);
                if (p.row_scale) x *= rs;
                if (p.accum && cg < p.m_main) x += p.Y[row * p.ldy + cg];
                if (p.bias) x += __ldg(p.bias + cg);
                if (p.act == 1) x = fmaxf(x, 0.f);
                if (cg < p.m_main) p.Y[row * p.ldy + cg] = x;
                else p.aux[row * p.ldaux + (cg - p.m_main)] = x;
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
  }
}

// =====================================================================================================
// wgrad on the tensor cores:  dW[m,k] = sum_n dY[n,m] X[n,k]   (bf16 in, fp32 out)
// Both operands are "MN-major" for tcgen05: the reduction index n is the SLOW index of dY [n,m] and X [n,k].
// TMA boxes of {64 contiguous columns (128 B), 64 rows of n} land in shared memory as 64 rows x 128 B with the
// 128-byte swizzle = the canonical MN-major SW128 layout ((T,8,m),(8,k)):((1,T,LBO),(8T,SBO)) with SBO = 1024 B
// (8 n-rows) and LBO = 8192 B (the next 64-column block).  One UMMA (K = 16 n-rows) advances the start address
// by 2 x SBO.  A CTA owns one 256(m) x 256(k) output tile for a contiguous n-range: two 128 x 256 fp32
// accumulators fill all 512 TMEM columns; partial tiles go to a workspace and are summed in a fixed order.
constexpr int WG_BN = 64;                          // n rows (reduction) per stage
constexpr int WG_SUB = 64 * WG_BN * 2;             // one {64 cols x 64 rows} bf16 box = 8 KB
constexpr int WG_OPER = 4 * WG_SUB;                // 256 columns of one operand = 32 KB
constexpr int WG_STAGE = 2 * WG_OPER;              // dY block + X block = 64 KB
constexpr int WG_STAGES = 3;
constexpr int WG_SMEM = WG_STAGES * WG_STAGE + 128;
static_assert(WG_SMEM + 1024 <= 232448, "wgrad shared-memory plan exceeds 227 KB");

__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(WG_SUB >> 4) << 16;            // LBO: next 64-element block along M/N
  d |= (uint64_t)(1024 >> 4) << 32;              // SBO: next group of 8 rows along K
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc_bf16_mn(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct WgParams {
  int64_t n;
  int m, k, m_chunks, k_chunks, splits;
  int64_t rows_per_split;                         // multiple of WG_BN
  float* partial;                                 // [splits][m][k]
};

__global__ void __launch_bounds__(TC_THREADS, 1)
tc_wgrad_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x, const WgParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if (smem_base & 1023u) __trap();
  const uint32_t bars = smem_base + WG_STAGES * WG_STAGE;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (WG_STAGES + s); };
  const uint32_t tfull_bar = bars + 8u * (2 * WG_STAGES);
  const uint32_t tmem_slot = bars + 8u * (2 * WG_STAGES + 1);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_base));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // work item of this CTA: (m chunk, k chunk, split of n)
  const int tile = blockIdx.x / p.splits, split = blockIdx.x % p.splits;
  const int mc = tile / p.k_chunks, kc = tile % p.k_chunks;
  const int64_t n0 = (int64_t)split * p.rows_per_split;
  const int64_t n1 = min(p.n, n0 + p.rows_per_split);
  const int nblocks = n1 > n0 ? (int)((n1 - n0 + WG_BN - 1) / WG_BN) : 0;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_dy) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    for (int s = 0; s < WG_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tfull_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int nb = 0; nb < nblocks; ++nb) {
        mbar_wait(empty_bar(stage), phase ^ 1);
        const uint32_t sa = smem_base + stage * WG_STAGE;
        mbar_expect_tx(full_bar(stage), WG_STAGE);
        const int row = (int)(n0 + (int64_t)nb * WG_BN);
        // rows past n1 but inside the tensor belong to the next split: they must not be counted twice, so the
        // split boundaries are multiples of WG_BN (rows_per_split) and only the global tail is zero-filled.
#pragma unroll
        for (int sub = 0; sub < 4; ++sub) {
          tma_load_2d(sa + sub * WG_SUB, &map_dy, full_bar(stage), mc * 256 + sub * 64, row);
          tma_load_2d(sa + WG_OPER + sub * WG_SUB, &map_x, full_bar(stage), kc * 256 + sub * 64, row);
        }
        if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16_mn(128, 256);
      int stage = 0;
      uint32_t phase = 0;
      for (int nb = 0; nb < nblocks; ++nb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * WG_STAGE;
        const uint32_t sb = sa + WG_OPER;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {                      // two 128-row halves of the 256 m rows
#pragma unroll
          for (int ks = 0; ks < WG_BN / 16; ++ks) {           // K = 16 n-rows per UMMA = 2 x SBO
            const uint64_t ad = make_smem_desc_mn(sa + mt * 2 * WG_SUB + ks * 2048);
            const uint64_t bd = make_smem_desc_mn(sb + ks * 2048);
            tc_mma_bf16(tmem_base + (uint32_t)(mt * 256), ad, bd, idesc, (nb | ks) ? 1u : 0u);
          }
        }
        tc_commit(empty_bar(stage));
        if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
      }
      tc_commit(tfull_bar);
    }
  } else {
    // epilogue (once): TMEM -> fp32 partial tile
    const int q = warp & 3, half = (warp - 2) >> 2;           // half = which 128-row accumulator
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    const int mrow = mc * 256 + half * 128 + q * 32 + lane;
    float* dst = p.partial + ((int64_t)split * p.m + mrow) * p.k + kc * 256;
#pragma unroll 1
    for (int c = 0; c < 256; c += 32) {
      uint32_t r[32];
      if (nblocks > 0) tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 256 + c), r);
      if (mrow < p.m) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int kk = kc * 256 + c + j;
          if (kk < p.k) dst[c + j] = nblocks > 0 ? __uint_as_float(r[j]) : 0.f;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// 2-D bf16 tensor [rows, cols] with row stride ld (elements); box = [box_rows, 64 cols], 128B swizzle, zero OOB fill
static bool make_map(CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                     bool f32 = false) {
  if (rows >= ((int64_t)1 << 31)) return false;
  PFN_encodeTiled enc = get_encode();
  if (!enc) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * (f32 ? 4 : 2)};
  cuuint32_t box[2] = {(cuuint32_t)(f32 ? T3_BK : TC_BK), (cuuint32_t)box_rows};     // 128 bytes wide either way
  cuuint32_t estr[2] = {1, 1};
  return enc(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool tc_make_map_bf16(CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  return make_map(map, ptr, rows, cols, ld, box_rows, false);
}

void wgrad_reduce_launch(const float* partial, int splits, int m, int k, float* dW, int64_t lddw, cudaStream_t st);

static void wg_plan(int64_t n, int m, int k, WgParams& p) {
  p.n = n; p.m = m; p.k = k;
  p.m_chunks = (int)ceil_div(m, 256);
  p.k_chunks = (int)ceil_div(k, 256);
  const int tiles = p.m_chunks * p.k_chunks;
  int splits = B2G_NUM_SMS / tiles;
  if (splits < 1) splits = 1;
  const int64_t blocks = ceil_div(n, WG_BN);
  if (splits > blocks) splits = (int)(blocks > 0 ? blocks : 1);
  p.splits = splits;
  p.rows_per_split = ceil_div(blocks, splits) * WG_BN;
}

int64_t tc_wgrad_ws_bytes(int64_t n, int m, int k) {
  WgParams p;
  wg_plan(n, m, k, p);
  return (int64_t)p.splits * m * k * 4 + 256;
}

int tc_linear_wgrad(const void* dY, int64_t lddy, const void* X, int64_t ldx, float* dW, int64_t lddw, int64_t n,
                    int m, int k, void* ws, cudaStream_t st) {
  if (!aligned16(dY) || !aligned16(X) || (lddy * 2) % 16 || (ldx * 2) % 16) return B2G_E_ALIGN;
  static bool attr_set[64] = {false};                       // the opt-in is per device
  const int dev = current_device_slot();
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM);
    if (e != cudaSuccess) return (int)e;
    attr_set[dev] = true;
  }
  WgParams p;
  wg_plan(n, m, k, p);
  p.partial = static_cast<float*>(ws);
  CUtensorMap map_dy, map_x;
  if (!make_map(&map_dy, dY, n, m, lddy, WG_BN) || !make_map(&map_x, X, n, k, ldx, WG_BN)) return B2G_E_UNSUPPORTED;
  const unsigned grid = (unsigned)(p.m_chunks * p.k_chunks * p.splits);
  tc_wgrad_kernel<<<grid, TC_THREADS, WG_SMEM, st>>>(map_dy, map_x, p);
  count_launch();
  int rc = cuda_status();
  if (rc) return rc;
  wgrad_reduce_launch(p.partial, p.splits, m, k, dW, lddw, st);
  return cuda_status();
}

static int g_tf32x3 = 1;   // fp32 Linear on tensor cores via the 3xTF32 split (0 = exact-fp32 SIMT only)
void tc_set_tf32x3(int on) { g_tf32x3 = on ? 1 : 0; }

bool tc_linear_supported(int64_t n, int m, int k, int dt, int which) {
  // fp32: the TMEM accumulation truncates, so the error grows ~linearly with the reduction length (measured: 3e-6 at
  // k = 256, 2.6e-5 at k = 3328 against the 1e-5 gate) -> reductions longer than T3_KCHUNK are accumulated chunk by
  // chunk with IEEE fp32 adds in the epilogue (tc_linear_fwd_tf32x3).
  if (dt == B2G_F32) return g_tf32x3 && which == 0 && n >= 1 && k >= 4 && k <= 16384 && (k % 4) == 0 && m >= 1 && get_encode() != nullptr;
  if (dt != B2G_BF16) return false;
  if (which == 2) return n >= 1 && m % 8 == 0 && k % 8 == 0 && m >= 8 && k >= 8 && get_encode() != nullptr;
  if (which != 0) return false;
  if (n < 1 || k < 8 || (k % 8) != 0 || m < 1) return false;
  return get_encode() != nullptr;
}

int64_t tc_linear_ws_bytes(int64_t, int m, int k, int dt, int) { return dt == B2G_F32 ? (int64_t)2 * m * k * 4 + 256 : 256; }

static int tc_linear_fwd_tf32x3(const void* X, int64_t ldx, const void* W, int64_t ldw, const float* bias,
                                const float* row_scale, void* Y, int64_t ldy, float* aux, int64_t ldaux, int64_t n,
                                int m, int m_main, int k, int act, void* ws, cudaStream_t st) {
  if (!ws) return B2G_E_ARG;
  if (!aligned16(X) || (ldx * 4) % 16) return B2G_E_ALIGN;
  if (m_main > 0 && (!aligned16(Y) || (ldy * 4) % 16)) return B2G_E_ALIGN;
  static bool attr_set[64] = {false};
  const int dev = current_device_slot();
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(tc_linear_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T3_SMEM);
    if (e != cudaSuccess) return (int)e;
    attr_set[dev] = true;
  }
  // The TMEM accumulation truncates: its error grows ~linearly with the reduction length (3e-6 at k = 256, 2.6e-5 at
  // k = 3328 against the 1e-5 gate).  Longer reductions are therefore cut into chunks of T3_KCHUNK columns: one launch per
  // chunk, each adding its fp32 partial product into Y with IEEE adds (bias / activation ride on the last chunk).
  float* wsplit = static_cast<float*>(ws);                     // [2m, kc]: rows 0..m-1 = hi, m..2m-1 = lo
  const int nchunks = (int)ceil_div(k, T3_KCHUNK);
  if (nchunks > 1 && m_main != m) return B2G_E_UNSUPPORTED;    // the fp32 aux split only exists for single-chunk GEMMs
  for (int c = 0; c < nchunks; ++c) {
    const int c0 = c * T3_KCHUNK, kc = (k - c0 < T3_KCHUNK) ? k - c0 : T3_KCHUNK;
    const bool last = c == nchunks - 1;
    const float* Xc = static_cast<const float*>(X) + c0;
    const float* Wc = static_cast<const float*>(W) + c0;
    const int64_t total = (int64_t)m * kc;
    int64_t blocks = ceil_div(total, 256);
    if (blocks > B2G_NUM_SMS * 4) blocks = B2G_NUM_SMS * 4;
    split_tf32_kernel<<<(unsigned)blocks, 256, 0, st>>>(Wc, ldw, m, kc, wsplit);
    CUtensorMap map_a, map_b;
    if (!make_map(&map_a, Xc, n, kc, ldx, TC_BM, true) || !make_map(&map_b, wsplit, 2 * (int64_t)m, kc, kc, TC_BN, true))
      return B2G_E_UNSUPPORTED;
    T3Params p;
    p.n = n; p.m = m; p.m_main = m_main; p.k = kc; p.bias = last ? bias : nullptr; p.row_scale = row_scale;
    p.Y = static_cast<float*>(Y); p.ldy = ldy; p.aux = aux; p.ldaux = ldaux; p.act = last ? act : 0; p.accum = c > 0 ? 1 : 0;
    const int64_t tiles = ceil_div(n, TC_BM) * ceil_div(m, TC_BN);
    const unsigned grid = (unsigned)(tiles < B2G_NUM_SMS ? tiles : B2G_NUM_SMS);
    tc_linear_tf32x3_kernel<<<grid, TC_THREADS, T3_SMEM, st>>>(map_a, map_b, p);
    count_launch(2);
  }
  return cuda_status();
}

// B2G_TC_TMA_STORE=0 in the environment at first use selects the LDS + row-store epilogue (A/B runs); read once, never written
static const int g_dual_plan = [] { const char* e = getenv("B2G_TC_DUAL"); return (e && e[0] == '0') ? 0 : 1; }();   // same, plan 2
static const int g_tma_store = [] { const char* e = getenv("B2G_TC_TMA_STORE"); return (e && e[0] == '0') ? 0 : 1; }();

int tc_linear_fwd(const void* X, int64_t ldx, const void* W, int64_t ldw, const float* bias,
                  const float* row_scale, void* Y, int64_t ldy, float* aux, int64_t ldaux, int64_t n, int m,
                  int m_main, int k, int dt, int act, int reserve_sms, const void* mask, int64_t ldmask, void* ws, cudaStream_t st) {
  if (mask && (dt != B2G_BF16 || !aligned16(mask) || (ldmask * 2) % 16 || m_main != m)) return B2G_E_UNSUPPORTED;
  if (dt == B2G_F32)
    return tc_linear_fwd_tf32x3(X, ldx, W, ldw, bias, row_scale, Y, ldy, aux, ldaux, n, m, m_main, k, act, ws, st);
  if (dt != B2G_BF16) return B2G_E_UNSUPPORTED;
  if (!aligned16(X) || !aligned16(W) || (ldx * 2) % 16 || (ldw * 2) % 16) return B2G_E_ALIGN;
  if (m_main > 0 && (!aligned16(Y) || (ldy * 2) % 16)) return B2G_E_ALIGN;
  static bool attr_set[64] = {false};
  const int dev = current_device_slot();
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(tc_linear_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_RES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_linear_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_STR);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_linear_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_DUAL);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_linear_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_RES);
    if (e != cudaSuccess) return (int)e;
    attr_set[dev] = true;
  }
  CUtensorMap map_a, map_b;
  if (!make_map(&map_a, X, n, k, ldx, TC_BM) || !make_map(&map_b, W, m, k, ldw, TC_BN)) return B2G_E_UNSUPPORTED;
  TcParams p;
  p.n = n; p.m = m; p.m_main = m_main; p.k = k; p.bias = bias; p.row_scale = row_scale;
  p.Y = static_cast<__nv_bfloat16*>(Y); p.ldy = ldy; p.aux = aux; p.ldaux = ldaux; p.act = act;
  p.mask = static_cast<const __nv_bfloat16*>(mask); p.ldmask = ldmask;
  CUtensorMap map_y = map_a;                                  // placeholder when the row-store epilogue is used
  p.tma_store = (g_tma_store && m_main >= 64 && (m_main % 8) == 0 && make_map(&map_y, Y, n, m_main, ldy, 32)) ? 1 : 0;
  const int64_t row_tiles = ceil_div(n, TC_BM), col_tiles = ceil_div(m, TC_BN);
  const int64_t tiles = row_tiles * col_tiles;
  int sms = B2G_NUM_SMS - reserve_sms;                        // bf16 kernels only; the fp32 path ignores the hint
  if (sms < 1) sms = 1;
  if (k <= TC_RES_KB_MAX * TC_BK && col_tiles <= sms) {
    int64_t per = sms / col_tiles;                            // CTAs per column group
    if (per > row_tiles) per = row_tiles;
    if (mask) tc_linear_kernel<1, true><<<(unsigned)(per * col_tiles), TC_THREADS, TC_SMEM_RES, st>>>(map_a, map_b, map_y, p);
    else tc_linear_kernel<1><<<(unsigned)(per * col_tiles), TC_THREADS, TC_SMEM_RES, st>>>(map_a, map_b, map_y, p);
  } else if (mask) {
    return B2G_E_UNSUPPORTED;                                 // the mask epilogue exists for the resident plan (k <= 256) only
  } else if (g_dual_plan && col_tiles == 1 && row_tiles >= 2 * (int64_t)sms) {
    tc_linear_kernel<2><<<(unsigned)sms, TC_THREADS, TC_SMEM_DUAL, st>>>(map_a, map_b, map_y, p);
  } else {
    const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
    tc_linear_kernel<0><<<grid, TC_THREADS, TC_SMEM_STR, st>>>(map_a, map_b, map_y, p);
  }
  count_launch();
  return cuda_status();
}

}  // namespace b2g
