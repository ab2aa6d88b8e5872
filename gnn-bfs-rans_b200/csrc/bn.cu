// bn.cu — the step either side of every layer call in the reference's FlowGNN (gnn_model.py:184-192):
//     h = h + h_new;  h = BatchNorm(h);  h = relu(h);  h = dropout(h)
// torch_geometric.nn.BatchNorm is part of the drop-in surface (gnn_model.py:9,87), so BatchNorm over [N, C] node
// features is a kernel of this library; the residual add, ReLU and dropout around it can be fused in by the caller
// (flow_model.FlowGNN(fused_glue=True)).  Measured before (2.5 M cells, C = 256, bf16): torch's channels-last batch-norm
// kernels + the elementwise glue were 55 % of the train step (33 of 60 ms), 5-10x off the HBM time of their traffic.
//
// HBM-bound.  Training forward = two passes: column statistics of s = x (+ r), then y = act(gamma (s - mean) rstd + beta)
// (* dropout); backward = two passes (column sums of dz and dz * xhat, then ds).  Algorithmic bytes per element:
// forward 2 reads (+1 with residual) + 1 write (+1 when s is kept), backward 3 reads + 2 reads + 1 write.
// Deterministic: per-CTA partial sums over a fixed row assignment, combined in a fixed order (double) by one CTA.
// Variance uses sums shifted by row 0 (no cancellation when |mean| >> std).
#include "common.cuh"

namespace b2g {

constexpr int BN_BLOCKS = B2G_NUM_SMS * 4;
constexpr int BN_THREADS = 256;

template <typename T>
__device__ __forceinline__ void ld16(const T* p, float (&f)[Vec<T>::N]) {
  ldg_vec<T>(p).to_float(f);
}

struct BnArgs {
  const void* x; int64_t ldx;            // input rows
  const void* r; int64_t ldr;            // optional residual (s = x + r)
  void* y; int64_t ldy;                  // output / forward output (backward: for the ReLU / dropout mask)
  void* s_out; int64_t lds;              // forward: optional copy of s;  backward: s
  const void* dy; int64_t lddy;
  void* ds; int64_t ldds;
  int64_t n; int C; int nvec;
  const float* mean; const float* rstd;  // [C]
  const float* gamma; const float* beta; // [C] (may be null: 1 / 0)
  const float* sums;                     // backward apply: [2, C] = sum dz, sum dz * xhat
  float* partial;                        // [BN_BLOCKS, 2, C]
  int relu; float p_drop; uint64_t seed; float drop_scale; int training; const uint64_t* epoch;
};

// ---- column partial sums of (a, b) over this CTA's rows; kind 0: a = s - K, b = (s - K)^2;  kind 1: a = dz, b = dz * xhat
template <typename T, int KIND>
__global__ void __launch_bounds__(BN_THREADS) bn_partial_kernel(const BnArgs a) {
  constexpr int VN = Vec<T>::N;
  __shared__ float red[2][BN_THREADS * VN];
  const int rpi = BN_THREADS / a.nvec;                       // rows per CTA iteration
  const int r_in = threadIdx.x / a.nvec, v = threadIdx.x % a.nvec;
  float s0[VN], s1[VN];
#pragma unroll
  for (int k = 0; k < VN; ++k) { s0[k] = 0.f; s1[k] = 0.f; }
  if (r_in < rpi) {
    const int c = v * VN;
    float kk[VN], mu[VN], rs[VN];
    if (KIND == 0) {
      ld16<T>((const T*)a.x + c, kk);
      if (a.r) {
        float t[VN];
        ld16<T>((const T*)a.r + c, t);
#pragma unroll
        for (int k = 0; k < VN; ++k) kk[k] += t[k];
        Vec<T> sv;
        sv.from_float(kk);
        sv.to_float(kk);
      }
    } else {
#pragma unroll
      for (int k = 0; k < VN; ++k) { mu[k] = __ldg(a.mean + c + k); rs[k] = __ldg(a.rstd + c + k); }
    }
    for (int64_t row = (int64_t)blockIdx.x * rpi + r_in; row < a.n; row += (int64_t)gridDim.x * rpi) {
      if (KIND == 0) {
        float f[VN];
        ld16<T>((const T*)a.x + row * a.ldx + c, f);
        if (a.r) {
          float t[VN];
          ld16<T>((const T*)a.r + row * a.ldr + c, t);
#pragma unroll
          for (int k = 0; k < VN; ++k) f[k] += t[k];
          Vec<T> sv;                                           // statistics of s as stored (rounded to T), see bn_apply_kernel
          sv.from_float(f);
          sv.to_float(f);
        }
#pragma unroll
        for (int k = 0; k < VN; ++k) {
          const float d = f[k] - kk[k];
          s0[k] += d;
          s1[k] = fmaf(d, d, s1[k]);
        }
      } else {
        float g[VN], sv[VN];
        ld16<T>((const T*)a.dy + row * a.lddy + c, g);
        ld16<T>((const T*)a.s_out + row * a.lds + c, sv);
        if (a.relu) {
          float yv[VN];
          ld16<T>((const T*)a.y + row * a.ldy + c, yv);
#pragma unroll
          for (int k = 0; k < VN; ++k) g[k] = yv[k] > 0.f ? g[k] * a.drop_scale : 0.f;
        }
#pragma unroll
        for (int k = 0; k < VN; ++k) {
          const float xh = (sv[k] - mu[k]) * rs[k];
          s0[k] += g[k];
          s1[k] = fmaf(g[k], xh, s1[k]);
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < VN; ++k) {
    red[0][threadIdx.x * VN + k] = s0[k];
    red[1][threadIdx.x * VN + k] = s1[k];
  }
  __syncthreads();
  if ((int)threadIdx.x < a.nvec) {
#pragma unroll
    for (int k = 0; k < VN; ++k) {
      float t0 = 0.f, t1 = 0.f;
      for (int rr = 0; rr < rpi; ++rr) {                       // fixed order
        t0 += red[0][(rr * a.nvec + threadIdx.x) * VN + k];
        t1 += red[1][(rr * a.nvec + threadIdx.x) * VN + k];
      }
      const int64_t o = (int64_t)blockIdx.x * 2 * a.C + threadIdx.x * VN + k;
      a.partial[o] = t0;
      a.partial[o + a.C] = t1;
    }
  }
}

// kind 0: out[0,c] = mean, out[1,c] = rstd, out[2,c] = biased variance;   kind 1: out[0,c] = sum dz, out[1,c] = sum dz xhat
template <typename T>
__global__ void bn_finalize_kernel(const float* __restrict__ partial, int nb, int C, int64_t n, int kind, float eps,
                                   const T* __restrict__ x0, const T* __restrict__ r0, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double t0 = 0.0, t1 = 0.0;
  for (int b = 0; b < nb; ++b) {
    t0 += (double)partial[(int64_t)b * 2 * C + c];
    t1 += (double)partial[(int64_t)b * 2 * C + C + c];
  }
  if (kind == 0) {
    double K = (double)(float)x0[c];
    if (r0) K = (double)(float)(T)((float)x0[c] + (float)r0[c]);   // the same rounded sum the partial kernel shifted by
    const double md = t0 / (double)n;
    double var = t1 / (double)n - md * md;
    if (var < 0.0) var = 0.0;
    out[c] = (float)(K + md);
    out[C + c] = (float)(1.0 / sqrt(var + (double)eps));
    out[2 * C + c] = (float)var;
  } else {
    out[c] = (float)t0;
    out[C + c] = (float)t1;
  }
}

// The two elementwise passes: a thread owns ONE 16-byte column group (its mean / rstd / gamma / beta live in registers
// for the whole kernel) and walks the rows with a CTA-grid stride, BN_U rows per step so that BN_U x (2-3) 16-byte loads
// are in flight per thread.  (A first version indexed (row, column) from a flat element counter: a 64-bit division and
// 4 parameter loads per element made it 3x slower than its HBM time.)
constexpr int BN_U = 4;

// y = act(gamma (s - mean) rstd + beta) [* dropout];  optionally keeps s = x + r for the backward pass
template <typename T>
__global__ void __launch_bounds__(BN_THREADS) bn_apply_kernel(const BnArgs a) {
  constexpr int VN = Vec<T>::N;
  const int rpi = BN_THREADS / a.nvec;
  const int r_in = threadIdx.x / a.nvec, v = threadIdx.x % a.nvec;
  if (r_in >= rpi) return;
  const int c = v * VN;
  float mu[VN], sc[VN], sh[VN];                                  // y = (s - mu) * sc + sh
#pragma unroll
  for (int k = 0; k < VN; ++k) {
    mu[k] = __ldg(a.mean + c + k);
    sc[k] = __ldg(a.rstd + c + k) * (a.gamma ? __ldg(a.gamma + c + k) : 1.0f);
    sh[k] = a.beta ? __ldg(a.beta + c + k) : 0.0f;
  }
  const float keep_scale = a.p_drop > 0.f ? 1.0f / (1.0f - a.p_drop) : 1.0f;
  const int64_t stride = (int64_t)gridDim.x * rpi;
  for (int64_t row0 = (int64_t)blockIdx.x * rpi + r_in; row0 < a.n; row0 += stride * BN_U) {
    Vec<T> xv[BN_U], rv[BN_U];
#pragma unroll
    for (int u = 0; u < BN_U; ++u) {
      const int64_t row = row0 + u * stride;
      if (row < a.n) {
        xv[u] = ldg_vec<T>((const T*)a.x + row * a.ldx + c);
        if (a.r) rv[u] = ldg_vec<T>((const T*)a.r + row * a.ldr + c);
      }
    }
#pragma unroll
    for (int u = 0; u < BN_U; ++u) {
      const int64_t row = row0 + u * stride;
      if (row >= a.n) break;
      float f[VN];
      xv[u].to_float(f);
      if (a.r) {
        float q[VN];
        rv[u].to_float(q);
#pragma unroll
        for (int k = 0; k < VN; ++k) f[k] += q[k];
        Vec<T> sv;                                             // s as stored: what the statistics saw and backward reads
        sv.from_float(f);
        if (a.s_out) stg_vec<T>((T*)a.s_out + row * a.lds + c, sv);
        sv.to_float(f);
      }
      float o[VN];
#pragma unroll
      for (int k = 0; k < VN; ++k) {
        o[k] = fmaf(f[k] - mu[k], sc[k], sh[k]);
        if (a.relu) o[k] = fmaxf(o[k], 0.f);
      }
      if (a.p_drop > 0.f) {
        const uint64_t t = (uint64_t)row * a.nvec + v;
#pragma unroll
        for (int k4 = 0; k4 < VN; k4 += 4) {
          const uint4 rnd = philox4x32(mix_epoch(a.seed, a.epoch), t * (VN / 4) + (k4 >> 2));
          const uint32_t w[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) o[k4 + k] = ((w[k] >> 8) * (1.0f / 16777216.0f) >= a.p_drop) ? o[k4 + k] * keep_scale : 0.f;
        }
      }
      Vec<T> ov;
      ov.from_float(o);
      stg_vec<T>((T*)a.y + row * a.ldy + c, ov);
    }
  }
}

// ds = gamma rstd (dz - sum(dz)/n - xhat sum(dz xhat)/n)   (training);   ds = gamma rstd dz   (eval)
template <typename T>
__global__ void __launch_bounds__(BN_THREADS) bn_bwd_apply_kernel(const BnArgs a) {
  constexpr int VN = Vec<T>::N;
  const int rpi = BN_THREADS / a.nvec;
  const int r_in = threadIdx.x / a.nvec, v = threadIdx.x % a.nvec;
  if (r_in >= rpi) return;
  const int c = v * VN;
  const float inv_n = 1.0f / (float)a.n;
  float mu[VN], rs[VN], gs[VN], m0[VN], m1[VN];                  // ds = (dz - m0 - xhat * m1) * gs
#pragma unroll
  for (int k = 0; k < VN; ++k) {
    mu[k] = __ldg(a.mean + c + k);
    rs[k] = __ldg(a.rstd + c + k);
    gs[k] = rs[k] * (a.gamma ? __ldg(a.gamma + c + k) : 1.0f);
    m0[k] = a.training ? __ldg(a.sums + c + k) * inv_n : 0.f;
    m1[k] = a.training ? __ldg(a.sums + a.C + c + k) * inv_n : 0.f;
  }
  const int64_t stride = (int64_t)gridDim.x * rpi;
  for (int64_t row0 = (int64_t)blockIdx.x * rpi + r_in; row0 < a.n; row0 += stride * BN_U) {
    Vec<T> gv[BN_U], sv[BN_U], yv[BN_U];
#pragma unroll
    for (int u = 0; u < BN_U; ++u) {
      const int64_t row = row0 + u * stride;
      if (row < a.n) {
        gv[u] = ldg_vec<T>((const T*)a.dy + row * a.lddy + c);
        sv[u] = ldg_vec<T>((const T*)a.s_out + row * a.lds + c);
        if (a.relu) yv[u] = ldg_vec<T>((const T*)a.y + row * a.ldy + c);
      }
    }
#pragma unroll
    for (int u = 0; u < BN_U; ++u) {
      const int64_t row = row0 + u * stride;
      if (row >= a.n) break;
      float g[VN], sf[VN];
      gv[u].to_float(g);
      sv[u].to_float(sf);
      if (a.relu) {
        float yf[VN];
        yv[u].to_float(yf);
#pragma unroll
        for (int k = 0; k < VN; ++k) g[k] = yf[k] > 0.f ? g[k] * a.drop_scale : 0.f;
      }
      float o[VN];
#pragma unroll
      for (int k = 0; k < VN; ++k) {
        const float xh = (sf[k] - mu[k]) * rs[k];
        o[k] = (g[k] - m0[k] - xh * m1[k]) * gs[k];
      }
      Vec<T> ov;
      ov.from_float(o);
      stg_vec<T>((T*)a.ds + row * a.ldds + c, ov);
    }
  }
}

static inline unsigned bn_grid(int64_t n, int nvec) {       // CTAs for the row-strided elementwise passes
  const int rpi = BN_THREADS / nvec;
  int64_t b = ceil_div(n > 0 ? n : 1, (int64_t)rpi * BN_U);
  const int64_t cap = (int64_t)B2G_NUM_SMS * 8;
  return (unsigned)(b < cap ? b : cap);
}
static inline int esz(int dt) { return dt == B2G_F32 ? 4 : 2; }
static inline bool rows_ok(const void* p, int64_t ld, int dt) { return p && aligned16(p) && (ld * esz(dt)) % 16 == 0; }
static inline int bn_shape(int C, int dt) {
  const int vn = 16 / esz(dt);
  if (C <= 0 || C % vn) return 0;
  const int nvec = C / vn;
  return nvec <= BN_THREADS ? nvec : 0;
}

}  // namespace b2g

using namespace b2g;

extern "C" {

int64_t b2g_bn_workspace_bytes(int C) { return C > 0 ? (int64_t)BN_BLOCKS * 2 * C * 4 : B2G_E_ARG; }

int b2g_bn_stats(const void* x, int64_t ldx, const void* r, int64_t ldr, int64_t n, int C, int dt, float eps,
                 float* stats, void* ws, void* stream) {
  if (n <= 0 || (dt != B2G_F32 && dt != B2G_BF16) || !stats || !ws) return B2G_E_ARG;
  const int nvec = bn_shape(C, dt);
  if (!nvec) return B2G_E_SHAPE;
  if (!rows_ok(x, ldx, dt) || (r && !rows_ok(r, ldr, dt))) return B2G_E_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  BnArgs a{};
  a.x = x; a.ldx = ldx; a.r = r; a.ldr = ldr; a.n = n; a.C = C; a.nvec = nvec; a.partial = (float*)ws;
  const int rpi = BN_THREADS / nvec;
  const int64_t nbw = ceil_div(n, rpi);
  const int nb = (int)(nbw < BN_BLOCKS ? nbw : BN_BLOCKS);
  if (dt == B2G_F32) {
    bn_partial_kernel<float, 0><<<nb, BN_THREADS, 0, st>>>(a);
    bn_finalize_kernel<float><<<(unsigned)ceil_div(C, 128), 128, 0, st>>>((const float*)ws, nb, C, n, 0, eps, (const float*)x, (const float*)r, stats);
  } else {
    bn_partial_kernel<__nv_bfloat16, 0><<<nb, BN_THREADS, 0, st>>>(a);
    bn_finalize_kernel<__nv_bfloat16><<<(unsigned)ceil_div(C, 128), 128, 0, st>>>((const float*)ws, nb, C, n, 0, eps, (const __nv_bfloat16*)x, (const __nv_bfloat16*)r, stats);
  }
  count_launch(2);
  return cuda_status();
}

int b2g_bn_apply(const void* x, int64_t ldx, const void* r, int64_t ldr, void* y, int64_t ldy, void* s_out, int64_t lds,
                 int64_t n, int C, int dt, const float* mean, const float* rstd, const float* gamma, const float* beta,
                 int relu, float p_drop, uint64_t seed, void* stream) {
  if (n < 0 || (dt != B2G_F32 && dt != B2G_BF16) || !mean || !rstd || p_drop < 0.f || p_drop >= 1.f) return B2G_E_ARG;
  if (n == 0) return B2G_OK;
  const int nvec = bn_shape(C, dt);
  if (!nvec) return B2G_E_SHAPE;
  if (!rows_ok(x, ldx, dt) || !rows_ok(y, ldy, dt) || (r && !rows_ok(r, ldr, dt)) || (s_out && !rows_ok(s_out, lds, dt))) return B2G_E_ALIGN;
  BnArgs a{};
  a.x = x; a.ldx = ldx; a.r = r; a.ldr = ldr; a.y = y; a.ldy = ldy; a.s_out = s_out; a.lds = lds; a.n = n; a.C = C; a.nvec = nvec;
  a.mean = mean; a.rstd = rstd; a.gamma = gamma; a.beta = beta; a.relu = relu; a.p_drop = p_drop; a.seed = seed; a.epoch = dropout_epoch_ptr();
  cudaStream_t st = (cudaStream_t)stream;
  if (dt == B2G_F32) bn_apply_kernel<float><<<bn_grid(n, nvec), BN_THREADS, 0, st>>>(a);
  else bn_apply_kernel<__nv_bfloat16><<<bn_grid(n, nvec), BN_THREADS, 0, st>>>(a);
  count_launch();
  return cuda_status();
}

int b2g_bn_bwd_stats(const void* dy, int64_t lddy, const void* y, int64_t ldy, const void* s, int64_t lds, int64_t n,
                     int C, int dt, const float* mean, const float* rstd, int relu, float drop_scale, float* sums,
                     void* ws, void* stream) {
  if (n <= 0 || (dt != B2G_F32 && dt != B2G_BF16) || !mean || !rstd || !sums || !ws) return B2G_E_ARG;
  const int nvec = bn_shape(C, dt);
  if (!nvec) return B2G_E_SHAPE;
  if (!rows_ok(dy, lddy, dt) || !rows_ok(s, lds, dt) || (relu && !rows_ok(y, ldy, dt))) return B2G_E_ALIGN;
  BnArgs a{};
  a.dy = dy; a.lddy = lddy; a.y = const_cast<void*>(y); a.ldy = ldy; a.s_out = const_cast<void*>(s); a.lds = lds;
  a.n = n; a.C = C; a.nvec = nvec; a.mean = mean; a.rstd = rstd; a.relu = relu; a.drop_scale = drop_scale; a.partial = (float*)ws;
  cudaStream_t st = (cudaStream_t)stream;
  const int rpi = BN_THREADS / nvec;
  const int64_t nbw = ceil_div(n, rpi);
  const int nb = (int)(nbw < BN_BLOCKS ? nbw : BN_BLOCKS);
  if (dt == B2G_F32) {
    bn_partial_kernel<float, 1><<<nb, BN_THREADS, 0, st>>>(a);
    bn_finalize_kernel<float><<<(unsigned)ceil_div(C, 128), 128, 0, st>>>((const float*)ws, nb, C, n, 1, 0.f, nullptr, nullptr, sums);
  } else {
    bn_partial_kernel<__nv_bfloat16, 1><<<nb, BN_THREADS, 0, st>>>(a);
    bn_finalize_kernel<__nv_bfloat16><<<(unsigned)ceil_div(C, 128), 128, 0, st>>>((const float*)ws, nb, C, n, 1, 0.f, nullptr, nullptr, sums);
  }
  count_launch(2);
  return cuda_status();
}

int b2g_bn_bwd_apply(const void* dy, int64_t lddy, const void* y, int64_t ldy, const void* s, int64_t lds, void* ds,
                     int64_t ldds, int64_t n, int C, int dt, const float* mean, const float* rstd, const float* gamma,
                     const float* sums, int relu, float drop_scale, int training, void* stream) {
  if (n < 0 || (dt != B2G_F32 && dt != B2G_BF16) || !mean || !rstd || (training && !sums)) return B2G_E_ARG;
  if (n == 0) return B2G_OK;
  const int nvec = bn_shape(C, dt);
  if (!nvec) return B2G_E_SHAPE;
  if (!rows_ok(dy, lddy, dt) || !rows_ok(s, lds, dt) || !rows_ok(ds, ldds, dt) || (relu && !rows_ok(y, ldy, dt))) return B2G_E_ALIGN;
  BnArgs a{};
  a.dy = dy; a.lddy = lddy; a.y = const_cast<void*>(y); a.ldy = ldy; a.s_out = const_cast<void*>(s); a.lds = lds; a.ds = ds; a.ldds = ldds;
  a.n = n; a.C = C; a.nvec = nvec; a.mean = mean; a.rstd = rstd; a.gamma = gamma; a.sums = sums; a.relu = relu;
  a.drop_scale = drop_scale; a.training = training;
  cudaStream_t st = (cudaStream_t)stream;
  if (dt == B2G_F32) bn_bwd_apply_kernel<float><<<bn_grid(n, nvec), BN_THREADS, 0, st>>>(a);
  else bn_bwd_apply_kernel<__nv_bfloat16><<<bn_grid(n, nvec), BN_THREADS, 0, st>>>(a);
  count_launch();
  return cuda_status();
}

}  // extern "C"
