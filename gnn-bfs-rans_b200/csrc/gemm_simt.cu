// gemm_simt.cu — K6 (SIMT path): the dense per-node Linear transforms inside the conv layers
// (PyG Linear / torch nn.Linear) as fp32-accumulating FFMA tile kernels.
// This is the exact-fp32 path (the 1e-5 parity gate rules out single-pass TF32) and the path for
// shapes the tcgen05 kernel (gemm_tc.cu) does not cover.  128x128x16 CTA tile, 8x8 per thread.
//   fwd  : Y[n,m]  = act(rs[n] * sum_k X[n,k] W[m,k] + b[m])       A row-major, B row-major (NT)
//   dgrad: dX[n,k] = sum_m dY[n,m] W[m,k]                           A row-major, B "k-major" (NN)
//   wgrad: dW[m,k] = sum_n dY[n,m] X[n,k], db[m] = sum_n dY[n,m]    split over n, ordered reduce
#include "common.cuh"

namespace b2g {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4;

template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ float from_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// Load 8 consecutive elements p[0..8) (guarded by `valid` elements, rest = 0) as floats.
template <typename T>
__device__ __forceinline__ void load8(const T* __restrict__ p, int valid, bool vec_ok, float* f) {
  if (valid >= 8 && vec_ok) {
    if (sizeof(T) == 4) {
      const float4 a = *reinterpret_cast<const float4*>(p);
      const float4 b = *reinterpret_cast<const float4*>(p + 4);
      f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    } else {
      Vec<__nv_bfloat16> v;
      v.v = *reinterpret_cast<const uint4*>(p);
      v.to_float(f);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = i < valid ? to_f<T>(p[i]) : 0.f;
  }
}

// C[M x N] tile kernel.  A: [rowsA, K] row-major (reduction contiguous).
// kBT = false: B is [colsN, K] row-major (reduction contiguous)      -> NT
// kBT = true : B is [K, colsN] row-major (output column contiguous)  -> NN
template <typename T, bool kBT>
__global__ void __launch_bounds__(256)
gemm_rowA_kernel(const T* __restrict__ A, int64_t lda, const T* __restrict__ B, int64_t ldb,
                 T* __restrict__ C, int64_t ldc, float* __restrict__ aux, int64_t ldaux, int n_main,
                 int64_t M, int N, int K,
                 const float* __restrict__ bias, const float* __restrict__ row_scale, int act, bool vec_ok) {
  __shared__ float As[BK][BM + PAD];
  __shared__ float Bs[BK][BN + PAD];
  const int t = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int ty = t >> 4, tx = t & 15;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    {  // A tile: 128 rows x 16 k; thread -> row t/2, k-half (t&1)*8
      const int r = t >> 1, kh = (t & 1) * 8;
      float f[8];
      const int64_t row = m0 + r;
      const int valid = row < M ? max(0, min(8, K - (k0 + kh))) : 0;
      if (valid > 0) load8<T>(A + row * lda + k0 + kh, valid, vec_ok, f);
      else {
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) As[kh + i][r] = f[i];
    }
    if (!kBT) {  // B tile from [N,K]: row = output column
      const int r = t >> 1, kh = (t & 1) * 8;
      float f[8];
      const int colg = n0 + r;
      const int valid = colg < N ? max(0, min(8, K - (k0 + kh))) : 0;
      if (valid > 0) load8<T>(B + (int64_t)colg * ldb + k0 + kh, valid, vec_ok, f);
      else {
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) Bs[kh + i][r] = f[i];
    } else {     // B tile from [K,N]: 16 k-rows x 128 cols; thread -> k-row t/16, col (t&15)*8
      const int kr = t >> 4, c8 = (t & 15) * 8;
      float f[8];
      const int valid = (k0 + kr) < K ? max(0, min(8, N - (n0 + c8))) : 0;
      if (valid > 0) load8<T>(B + (int64_t)(k0 + kr) * ldb + n0 + c8, valid, vec_ok, f);
      else {
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) Bs[kr][c8 + i] = f[i];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[8], b[8];
      *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
      *reinterpret_cast<float4*>(a + 4) = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(&Bs[kk][tx * 8]);
      *reinterpret_cast<float4*>(b + 4) = *reinterpret_cast<const float4*>(&Bs[kk][tx * 8 + 4]);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  // epilogue
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t row = m0 + ty * 8 + i;
    if (row >= M) continue;
    const float rs = row_scale ? __ldg(row_scale + row) : 1.0f;
    const int c0 = n0 + tx * 8;
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = acc[i][j];
      if (row_scale) v *= rs;
      if (bias && c0 + j < N) v += __ldg(bias + c0 + j);
      if (act == 1) v = fmaxf(v, 0.f);
      o[j] = v;
    }
    if (c0 >= n_main) {  // tail columns -> fp32 aux (n_main is a multiple of 8 or == N)
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (c0 + j < N) aux[row * ldaux + (c0 + j - n_main)] = o[j];
      continue;
    }
    T* dst = C + row * ldc + c0;
    if (c0 + 8 <= n_main && vec_ok) {
      if (sizeof(T) == 4) {
        *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
      } else {
        Vec<__nv_bfloat16> v;
        v.from_float(o);
        *reinterpret_cast<uint4*>(dst) = v.v;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (c0 + j < n_main) dst[j] = from_f<T>(o[j]);
        else if (c0 + j < N) aux[row * ldaux + (c0 + j - n_main)] = o[j];
      }
    }
  }
}

// wgrad: partial[s][m][k] = sum_{n in split s} dY[n,m] X[n,k];  pdb[s][m] = sum dY[n,m]
template <typename T>
__global__ void __launch_bounds__(256)
gemm_wgrad_kernel(const T* __restrict__ dY, int64_t lddy, const T* __restrict__ X, int64_t ldx,
                  float* __restrict__ partial, float* __restrict__ pdb, int64_t N, int M, int K,
                  int64_t rows_per_split, bool vec_ok) {
  __shared__ float As[BK][BM + PAD];  // [n][m]
  __shared__ float Bs[BK][BN + PAD];  // [n][k]
  const int t = threadIdx.x;
  const int m0 = blockIdx.x * BM, k0 = blockIdx.y * BN;
  const int split = blockIdx.z;
  const int64_t nb = (int64_t)split * rows_per_split;
  const int64_t ne = min(N, nb + rows_per_split);
  const int ty = t >> 4, tx = t & 15;
  float acc[8][8], dbacc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    dbacc[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  }
  const int nr = t >> 4, c8 = (t & 15) * 8;
  for (int64_t n0 = nb; n0 < ne; n0 += BK) {
    float f[8];
    const int64_t row = n0 + nr;
    int valid = row < ne ? max(0, min(8, M - (m0 + c8))) : 0;
    if (valid > 0) load8<T>(dY + row * lddy + m0 + c8, valid, vec_ok, f);
    else {
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) As[nr][c8 + i] = f[i];
    valid = row < ne ? max(0, min(8, K - (k0 + c8))) : 0;
    if (valid > 0) load8<T>(X + row * ldx + k0 + c8, valid, vec_ok, f);
    else {
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) Bs[nr][c8 + i] = f[i];
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[8], b[8];
      *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
      *reinterpret_cast<float4*>(a + 4) = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(&Bs[kk][tx * 8]);
      *reinterpret_cast<float4*>(b + 4) = *reinterpret_cast<const float4*>(&Bs[kk][tx * 8 + 4]);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        dbacc[i] += a[i];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }
    __syncthreads();
  }
  float* P = partial + (int64_t)split * M * K;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty * 8 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = k0 + tx * 8 + j;
      if (k < K) P[(int64_t)m * K + k] = acc[i][j];
    }
    if (pdb && blockIdx.y == 0 && tx == 0) pdb[(int64_t)split * M + m] = dbacc[i];
  }
}

__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ partial, const float* __restrict__ pdb, int splits,
                    int M, int K, float* __restrict__ dW, int64_t lddw, float* __restrict__ db) {
  const int64_t total = (int64_t)M * K;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total + M;
       idx += (int64_t)gridDim.x * blockDim.x) {
    if (idx < total) {
      float s = 0.f;
      for (int sp = 0; sp < splits; ++sp) s += partial[(int64_t)sp * total + idx];  // fixed order
      dW[(idx / K) * lddw + (idx % K)] = s;
    } else if (db) {
      const int m = (int)(idx - total);
      float s = 0.f;
      for (int sp = 0; sp < splits; ++sp) s += pdb[(int64_t)sp * M + m];
      db[m] = s;
    }
  }
}

void wgrad_reduce_launch(const float* partial, int splits, int m, int k, float* dW, int64_t lddw, cudaStream_t st) {
  const int64_t total = (int64_t)m * k;
  int64_t blocks = ceil_div(total, 256);
  if (blocks > B2G_NUM_SMS * 8) blocks = B2G_NUM_SMS * 8;
  wgrad_reduce_kernel<<<(unsigned)blocks, 256, 0, st>>>(partial, nullptr, splits, m, k, dW, lddw, nullptr);
  count_launch();
}

int wgrad_splits(int64_t n, int m, int k) {
  const int64_t tiles = ceil_div(m, BM) * ceil_div(k, BN);
  int64_t s = (B2G_NUM_SMS * 4) / tiles;
  const int64_t max_s = ceil_div(n, 4 * BK);
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  return (int)s;
}

template <typename T>
static int simt_fwd(const void* X, int64_t ldx, const void* W, int64_t ldw, const float* bias,
                    const float* rs, void* Y, int64_t ldy, float* aux, int64_t ldaux, int m_main,
                    int64_t n, int m, int k, int act, bool bt, cudaStream_t st) {
  const bool vec_ok = aligned16(X) && aligned16(W) && aligned16(Y) && (ldx * sizeof(T)) % 16 == 0 &&
                      (ldw * sizeof(T)) % 16 == 0 && (ldy * sizeof(T)) % 16 == 0;
  dim3 grid((unsigned)ceil_div(n, BM), (unsigned)ceil_div(m, BN));
  if (bt)
    gemm_rowA_kernel<T, true><<<grid, 256, 0, st>>>((const T*)X, ldx, (const T*)W, ldw, (T*)Y, ldy, aux, ldaux, m_main, n, m, k, bias, rs, act, vec_ok);
  else
    gemm_rowA_kernel<T, false><<<grid, 256, 0, st>>>((const T*)X, ldx, (const T*)W, ldw, (T*)Y, ldy, aux, ldaux, m_main, n, m, k, bias, rs, act, vec_ok);
  count_launch();
  return cuda_status();
}

int simt_linear_fwd(const void* X, int64_t ldx, const void* W, int64_t ldw, const float* bias,
                    const float* rs, void* Y, int64_t ldy, float* aux, int64_t ldaux, int64_t n, int m,
                    int m_main, int k, int dt, int act, cudaStream_t st) {
  if (dt == B2G_F32) return simt_fwd<float>(X, ldx, W, ldw, bias, rs, Y, ldy, aux, ldaux, m_main, n, m, k, act, false, st);
  return simt_fwd<__nv_bfloat16>(X, ldx, W, ldw, bias, rs, Y, ldy, aux, ldaux, m_main, n, m, k, act, false, st);
}
int simt_linear_dgrad(const void* dY, int64_t lddy, const void* W, int64_t ldw, void* dX, int64_t lddx,
                      int64_t n, int m, int k, int dt, cudaStream_t st) {
  // dX[n,k] = dY[n,m] W[m,k]: output columns = k, reduction = m, B = W as [m(red), k(out)]
  if (dt == B2G_F32) return simt_fwd<float>(dY, lddy, W, ldw, nullptr, nullptr, dX, lddx, nullptr, 0, k, n, k, m, 0, true, st);
  return simt_fwd<__nv_bfloat16>(dY, lddy, W, ldw, nullptr, nullptr, dX, lddx, nullptr, 0, k, n, k, m, 0, true, st);
}
int64_t simt_wgrad_ws_bytes(int64_t n, int m, int k) {
  const int s = wgrad_splits(n, m, k);
  return (int64_t)s * ((int64_t)m * k + m) * (int64_t)sizeof(float);
}
int simt_linear_wgrad(const void* dY, int64_t lddy, const void* X, int64_t ldx, float* dW, int64_t lddw,
                      float* db, int64_t n, int m, int k, int dt, void* ws, cudaStream_t st) {
  const int s = wgrad_splits(n, m, k);
  const int64_t rps = ceil_div(ceil_div(n, s), BK) * BK;
  float* partial = static_cast<float*>(ws);
  float* pdb = partial + (int64_t)s * m * k;
  dim3 grid((unsigned)ceil_div(m, BM), (unsigned)ceil_div(k, BN), (unsigned)s);
  if (dt == B2G_F32) {
    const bool vec_ok = aligned16(dY) && aligned16(X) && (lddy * 4) % 16 == 0 && (ldx * 4) % 16 == 0;
    gemm_wgrad_kernel<float><<<grid, 256, 0, st>>>((const float*)dY, lddy, (const float*)X, ldx, partial, pdb, n, m, k, rps, vec_ok);
  } else {
    const bool vec_ok = aligned16(dY) && aligned16(X) && (lddy * 2) % 16 == 0 && (ldx * 2) % 16 == 0;
    gemm_wgrad_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)dY, lddy, (const __nv_bfloat16*)X, ldx, partial, pdb, n, m, k, rps, vec_ok);
  }
  const int64_t total = (int64_t)m * k + m;
  int64_t blocks = ceil_div(total, 256);
  if (blocks > B2G_NUM_SMS * 8) blocks = B2G_NUM_SMS * 8;
  wgrad_reduce_kernel<<<(unsigned)blocks, 256, 0, st>>>(partial, pdb, s, m, k, dW, lddw, db);
  count_launch(2);
  return cuda_status();
}

}  // namespace b2g
