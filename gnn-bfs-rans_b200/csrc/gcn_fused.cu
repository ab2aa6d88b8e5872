// gcn_fused.cu — K2f: the CSR segment-sum of GCNConv / GINConv FUSED with the layer's (first) Linear:
//
//      GCNConv:  out_i = dinv_i * (sum_j dinv_j x_j) W^T + b            (gnn_model.py:63,166)
//      GINConv:  h1_i  = relu((sum_j x_j + (1 + eps) x_i) W1^T + b1)    (gnn_model.py:70-75,166; rows 6, 9, 10; inference)
//
// PyG (and the unfused path here: K6 GEMM then K2 seg_rows) projects first and aggregates the projected rows: the [N, F]
// intermediate is written and read back (10.2 GB of the layer's 21.5 GB at cfg4).  The sum over j is linear, so the
// aggregation can run on the INPUT rows and feed the GEMM from shared memory, as K4f does for GATConv (gat_fused.cu):
//
//   gather warps   z tiles — 128 target rows x 64 features per step, 4 rows x 8 sixteen-byte pieces per warp-load, both
//                  units of a warp in flight together (14 loads) — written as bf16 into the 128B-swizzled K-major A operand,
//   MMA warp       tcgen05.mma (kind::f16, 128 x 256 x 16) against W, which is RESIDENT in shared memory (k <= 256: 128 KB,
//                  loaded once per CTA by TMA; the z chunk ring is only 4 x 16 KB here), fp32 accumulators in TMEM (2),
//   epilogue warps tcgen05.ld -> row scale (dinv_i) -> bias -> ReLU -> bf16 -> 128-byte row stores.
//
// HBM traffic of the layer forward: x once + out once + indices = 10.6 GB instead of 21.5 GB.  Deterministic: fp32
// accumulation in CSR order per row, one bf16 rounding of z before the tensor core (the unfused path rounds x W^T instead).
#include "rows.cuh"
#include "tc_ptx.cuh"

namespace b2g {

constexpr int SF_BM = 128;                       // target rows per tile == UMMA_M
constexpr int SF_BN = 256;                       // UMMA_N; C_out <= 256
constexpr int SF_KCH = 4;                        // 64-feature chunks (F = 256 bf16 = 512-byte rows) == k-blocks of W
constexpr int SF_GW = 16;                        // gather warps
constexpr int SF_THREADS = 32 * (8 + SF_GW);     // warp 0 TMA, 1 MMA, 2-3 idle, 4-7 epilogue, 8.. gather
constexpr int SF_A_CHUNK = SF_BM * 128;          // 16 KB: 128 rows x 64 bf16, SWIZZLE_128B K-major
constexpr int SF_NBUF = 2;                       // z chunk ring
constexpr int SF_W_KB = SF_BN * 128;             // 32 KB: 256 rows of W x 64 k
constexpr int SF_W_STAGES = 2;                   // W k-blocks are STREAMED (128 KB per tile from L2), not resident: see below
constexpr int SF_STG = 4 * 32 * 128;             // epilogue staging
constexpr int SF_BAR = 256;
// 64 + 32 + 16 KB = 112 KB of shared memory -> the 132 KB carve-out leaves ~120 KB of L1.  With W resident (128 KB, 208 KB in
// total, ~45 KB of L1) the kernel ran 5.9-7.0 ms: the gathered lines of 16 warps x 7-14 loads in flight did not fit the L1.
constexpr int SF_SMEM = SF_W_STAGES * SF_W_KB + SF_NBUF * SF_A_CHUNK + SF_STG + SF_BAR;
static_assert(SF_SMEM <= 232448, "fused GCN shared-memory plan exceeds 227 KB");

struct SfArgs {
  const char* x; uint32_t xrow_bytes;            // gathered rows, bf16 [*, 256]
  const int32_t* rowptr; const int32_t* col;
  const float* col_scale;                        // fp32 [n_src]: weight of source j (GCN: deg^-1/2) or NULL (= 1)
  const float* row_scale;                        // fp32 [n_rows]: scale of target i, applied in the epilogue, or NULL
  const float* bias;                             // fp32 [m] or NULL
  float self_coef;                               // + self_coef * x_i (GIN: 1 + eps); 0 = none
  int relu;
  __nv_bfloat16* out; int64_t ldo;
  uint32_t n_rows; int m;
  const char* zero;                              // >= 512 bytes of zeros (padding lanes load from here)
  RowSched ord;                                  // chunk_rows = SF_BM
};

__device__ __forceinline__ uint4 sf_ldg_sel(const char* p, const char* zero, bool valid) {
  const char* q = valid ? p : zero;
  uint4 u;
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "l"(q));
  return u;
}
__device__ __forceinline__ void sf_fma(float (&acc)[8], float w, const uint4& u) {
  float f[8];
  unpack_row16(u, f, __nv_bfloat16());
#pragma unroll
  for (int k = 0; k < 8; k += 2) ffma2_acc(acc[k], acc[k + 1], w, f[k], f[k + 1]);
}
// loads of the first K entries of the 4 rows of a unit (one 16-byte piece per lane and entry)
template <int K>
__device__ __forceinline__ void sf_issue(uint4 (&buf)[8], const char* xk, uint32_t xrow_bytes, int cl, int len, int g0,
                                         const char* zero) {
#pragma unroll
  for (int t = 0; t < K; ++t) {
    const uint32_t c = (uint32_t)__shfl_sync(0xffffffffu, cl, g0 + t);
    buf[t] = sf_ldg_sel(xk + (uint64_t)c * xrow_bytes, zero, t < len);
  }
}
template <int K>
__device__ __forceinline__ void sf_consume(float (&acc)[8], const uint4 (&buf)[8], float wl, int g0) {
#pragma unroll
  for (int t = 0; t < K; ++t) sf_fma(acc, __shfl_sync(0xffffffffu, wl, g0 + t), buf[t]);
}

__global__ void __launch_bounds__(SF_THREADS, 1)
segw_gemm_kernel(const __grid_constant__ CUtensorMap map_w, const SfArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if (smem_base & 1023u) __trap();
  const uint32_t wring = smem_base;                                   // W k-block ring: 2 x 32 KB
  const uint32_t aring = wring + SF_W_STAGES * SF_W_KB;               // 2 x 16 KB z chunks
  const uint32_t stg = aring + SF_NBUF * SF_A_CHUNK;
  const uint32_t bars = stg + SF_STG;
  auto full_a = [&](int b) { return bars + 8u * b; };
  auto empty_a = [&](int b) { return bars + 8u * (SF_NBUF + b); };
  auto tfull = [&](int t) { return bars + 8u * (2 * SF_NBUF + t); };
  auto tempty = [&](int t) { return bars + 8u * (2 * SF_NBUF + 2 + t); };
  auto full_w = [&](int s_) { return bars + 8u * (2 * SF_NBUF + 4 + s_); };
  auto empty_w = [&](int s_) { return bars + 8u * (2 * SF_NBUF + 4 + SF_W_STAGES + s_); };
  const uint32_t tmem_slot = bars + 8u * (2 * SF_NBUF + 4 + 2 * SF_W_STAGES);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_base));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    for (int s = 0; s < SF_NBUF; ++s) {
      mbar_init(full_a(s), SF_GW);
      mbar_init(empty_a(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull(s), 1);
      mbar_init(tempty(s), 4);
    }
    for (int s = 0; s < SF_W_STAGES; ++s) {
      mbar_init(full_w(s), 1);
      mbar_init(empty_w(s), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
    if (warp == 0 && lane == 0) {
      // ===================================================== TMA producer: the 4 k-blocks of W per tile (L2-resident)
      int stage = 0;
      uint32_t phase = 0;
      for (uint32_t q = blockIdx.x; q < a.ord.n_chunks; q += gridDim.x) {
        uint32_t rows;
        a.ord.chunk(q, a.n_rows, rows);
        if (rows == 0) continue;
        for (int kb = 0; kb < SF_KCH; ++kb) {
          mbar_wait(empty_w(stage), phase ^ 1);
          mbar_expect_tx(full_w(stage), SF_W_KB);
          tma_load_2d(wring + stage * SF_W_KB, &map_w, full_w(stage), kb * 64, 0);
          if (++stage == SF_W_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1 && lane == 0) {
      // ===================================================== MMA issuer
      constexpr uint32_t idesc = make_idesc_bf16(SF_BM, SF_BN);
      int acc = 0, stage = 0;
      uint32_t acc_phase = 0, g = 0, phase = 0;            // g = z chunks consumed (buffer g % NBUF, phase (g / NBUF) & 1)
      for (uint32_t q = blockIdx.x; q < a.ord.n_chunks; q += gridDim.x) {
        uint32_t rows;
        a.ord.chunk(q, a.n_rows, rows);
        if (rows == 0) continue;
        mbar_wait(tempty(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * SF_BN);
        for (int kc = 0; kc < SF_KCH; ++kc, ++g) {
          const int ab = g & (SF_NBUF - 1);
          mbar_wait(full_w(stage), phase);
          mbar_wait(full_a(ab), (g / SF_NBUF) & 1);
          tc_fence_after();
          const uint32_t sa = aring + ab * SF_A_CHUNK, sb = wring + stage * SF_W_KB;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            tc_mma_bf16(d_tmem, make_smem_desc(sa + ks * 32), make_smem_desc(sb + ks * 32), idesc, (kc | ks) ? 1u : 0u);
          tc_commit(empty_w(stage));
          tc_commit(empty_a(ab));
          if (++stage == SF_W_STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(tfull(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp < 8) {
    // ===================================================== epilogue warps 4..7 (TMEM lane quadrant = warp & 3)
    const int qd = warp & 3;
    uint8_t* my_stg = smem_raw + (stg - smem_base) + qd * 32 * 128;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (uint32_t q = blockIdx.x; q < a.ord.n_chunks; q += gridDim.x) {
      uint32_t rows;
      const uint32_t c0 = a.ord.chunk(q, a.n_rows, rows);
      if (rows == 0) continue;
      const uint32_t row0 = c0 + qd * 32, rend = c0 + rows;
      const uint32_t my_row = row0 + lane;
      float rs = 1.0f;
      if (a.row_scale && my_row < rend) rs = __ldg(a.row_scale + my_row);     // requested before the wait
      mbar_wait(tfull(acc), acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < a.m; c += 64) {
        const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(acc * SF_BN + c);
#pragma unroll
        for (int hlf = 0; hlf < 2; ++hlf) {
          uint32_t r[32];
          tc_ld32(taddr + hlf * 32, r);
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = __uint_as_float(r[j + k]) * rs;
            if (a.bias) {
              const int cg = c + hlf * 32 + j;
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(a.bias + cg));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(a.bias + cg + 4));
              v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
              v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
            }
            if (a.relu) {
#pragma unroll
              for (int k = 0; k < 8; ++k) v[k] = fmaxf(v[k], 0.f);
            }
            Vec<__nv_bfloat16> o;
            o.from_float(v);
            *reinterpret_cast<uint4*>(my_stg + lane * 128 + (((hlf * 4 + (j >> 3)) ^ (lane & 7)) << 4)) = o.v;
          }
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
      if (lane == 0) mbar_arrive(tempty(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================================================== gather warps 8..: z chunks into the swizzled A operand
    asm volatile("setmaxnreg.inc.sync.aligned.u32 88;");
    const int gw = warp - 8;
    const int r4 = lane >> 3, p = lane & 7, g0 = lane & 24;
    const char* xl = a.x + p * 16;
    const char* zl = a.zero + p * 16;
    uint32_t g = 0;
    for (uint32_t q = blockIdx.x; q < a.ord.n_chunks; q += gridDim.x) {
      uint32_t rows;
      const uint32_t c0 = a.ord.chunk(q, a.n_rows, rows);
      if (rows == 0) continue;
      // ---- per tile: this lane's entry (row r4 of the unit, entry p) of both units: column index + weight
      int cl[2], len[2], b0[2], mlen[2];
      float wl[2];
      uint32_t rl[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        rl[u] = (uint32_t)(gw + SF_GW * u) * 4u + r4;
        const bool valid = rl[u] < rows;
        const uint32_t row = valid ? c0 + rl[u] : c0;
        const int b = __ldg(a.rowptr + row), e = __ldg(a.rowptr + row + 1);
        len[u] = valid ? e - b : 0;
        b0[u] = b;
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const bool has = p < len[u];
        int c = 0;
        float w = 0.f;
        if (has) {
          c = __ldg(a.col + b0[u] + p);
          w = a.col_scale ? __ldg(a.col_scale + c) : 1.0f;
        }
        cl[u] = c;
        wl[u] = w;
        int ml = len[u];
        ml = max(ml, __shfl_xor_sync(0xffffffffu, ml, 8));
        ml = max(ml, __shfl_xor_sync(0xffffffffu, ml, 16));
        mlen[u] = ml;
      }
      for (int kc = 0; kc < SF_KCH; ++kc, ++g) {
        const int ab = g & (SF_NBUF - 1);
        const char* xk = xl + kc * 128;
        // one unit at a time: <= 8 loads in flight per lane (the shared-memory plan leaves ~40 KB of L1: 16 warps x 14
        // loads x 512 B in flight measured 2.5x SLOWER than 16 x 7 — the L1 cannot hold the lines of that many misses)
        bool waited = false;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          uint4 buf[8];
          const int ku = min(mlen[u], 8);
          switch (ku) {
#define B2G_CASE(KK) case KK: sf_issue<KK>(buf, xk, a.xrow_bytes, cl[u], len[u], g0, zl); break;
            B2G_CASE(1) B2G_CASE(2) B2G_CASE(3) B2G_CASE(4) B2G_CASE(5) B2G_CASE(6) B2G_CASE(7) B2G_CASE(8)
#undef B2G_CASE
            default: break;
          }
          uint4 selfv = make_uint4(0u, 0u, 0u, 0u);
          if (a.self_coef != 0.f) selfv = sf_ldg_sel(xk + (uint64_t)(c0 + rl[u]) * a.xrow_bytes, zl, rl[u] < rows);
          float acc[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k] = 0.f;
          switch (ku) {
#define B2G_CASE(KK) case KK: sf_consume<KK>(acc, buf, wl[u], g0); break;
            B2G_CASE(1) B2G_CASE(2) B2G_CASE(3) B2G_CASE(4) B2G_CASE(5) B2G_CASE(6) B2G_CASE(7) B2G_CASE(8)
#undef B2G_CASE
            default: break;
          }
          for (int t = 8; t < mlen[u]; ++t) {                          // rows longer than 8 entries (cold on meshes)
            float w = 0.f;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (t < len[u]) {
              const int c = __ldg(a.col + b0[u] + t);
              w = a.col_scale ? __ldg(a.col_scale + c) : 1.0f;
              v = ldg_row16(xk + (uint64_t)(uint32_t)c * a.xrow_bytes);
            }
            sf_fma(acc, w, v);
          }
          if (a.self_coef != 0.f) sf_fma(acc, a.self_coef, selfv);
          if (!waited) {
            mbar_wait(empty_a(ab), ((g / SF_NBUF) & 1) ^ 1);           // the MMAs that read this buffer NBUF chunks ago have retired
            waited = true;
          }
          const uint32_t dst = aring + ab * SF_A_CHUNK + rl[u] * 128u + (uint32_t)((p ^ (rl[u] & 7)) << 4);
          Vec<__nv_bfloat16> o;
          o.from_float(acc);
          asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(o.v.x), "r"(o.v.y), "r"(o.v.z), "r"(o.v.w) : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(full_a(ab));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

}  // namespace b2g

using namespace b2g;

extern "C" {

int b2g_segw_gemm_supported(int64_t n, int F, int C, int dt) {
  return (dt == B2G_BF16 && F == 256 && C >= 64 && C <= SF_BN && (C % 64) == 0 && n >= 1 && n < (1ll << 32) - (1ll << 25)) ? 1 : 0;
}

int b2g_segw_gemm(const void* x, int64_t ldx, const int32_t* rowptr, const int32_t* col, const float* col_scale,
                  const float* row_scale, float self_coef, const void* w, int64_t ldw, const float* bias, int relu, void* out,
                  int64_t ldo, int64_t n_rows, int F, int C, int dt, int64_t band, void* stream) {
  if (n_rows < 0) return B2G_E_ARG;
  if (n_rows == 0) return B2G_OK;
  if (!b2g_segw_gemm_supported(n_rows, F, C, dt)) return B2G_E_UNSUPPORTED;
  if (!x || !rowptr || !col || !w || !out) return B2G_E_ARG;
  if (!aligned16(x) || !aligned16(w) || !aligned16(out) || (bias && !aligned16(bias)) || (ldx * 2) % 16 || (ldw * 2) % 16 ||
      (ldo * 2) % 16 || ldx * 2 >= (1ll << 32))
    return B2G_E_ALIGN;
  static bool attr_set[64] = {false};
  const int dev = current_device_slot();
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(segw_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SF_SMEM);
    // the smallest carve-out that holds the 112 KB: everything else of the 256 KB stays L1 for the gathered rows
    if (e == cudaSuccess) e = cudaFuncSetAttribute(segw_gemm_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 50);
    if (e != cudaSuccess) return (int)e;
    attr_set[dev] = true;
  }
  SfArgs a{};
  if (!make_row_sched(n_rows, band, a.ord, SF_BM, 8192)) return B2G_E_UNSUPPORTED;
  CUtensorMap map_w;
  if (!tc_make_map_bf16(&map_w, w, C, F, ldw, SF_BN)) return B2G_E_UNSUPPORTED;
  a.x = static_cast<const char*>(x); a.xrow_bytes = (uint32_t)(ldx * 2);
  a.rowptr = rowptr; a.col = col; a.col_scale = col_scale; a.row_scale = row_scale; a.bias = bias; a.self_coef = self_coef;
  a.relu = relu;
  a.zero = static_cast<const char*>(zero_row_ptr());
  if (!a.zero) return B2G_E_UNSUPPORTED;
  a.out = static_cast<__nv_bfloat16*>(out); a.ldo = ldo; a.n_rows = (uint32_t)n_rows; a.m = C;
  const unsigned grid = a.ord.n_chunks < (uint32_t)B2G_NUM_SMS ? a.ord.n_chunks : (unsigned)B2G_NUM_SMS;
  segw_gemm_kernel<<<grid, SF_THREADS, SF_SMEM, (cudaStream_t)stream>>>(map_w, a);
  count_launch();
  return cuda_status();
}

}  // extern "C"
