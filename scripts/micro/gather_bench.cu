// Microbenchmark: what can the B200 memory system deliver for the mesh-gather access pattern?
// out[i] = sum_{u} x[i + off_u] over rows of ROWB bytes (bf16 F=256 -> 512 B), offsets known arithmetically
// (no index loads), all loads batched.  Variants: which offsets, warps per row, CTAs per SM.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint4 ld_pol(const uint4* p, uint64_t pol) {
  uint4 r;
  asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
  return r;
}
template <int NOFF, int ROWV /*16B vectors per row / 32*/, int POL = 0>
__global__ void __launch_bounds__(256) gather_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, int64_t n_rows,
                                                     const int* __restrict__ offs, int chunk_rows) {
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  int off[NOFF];
#pragma unroll
  for (int u = 0; u < NOFF; ++u) off[u] = offs[u];
  const int64_t rowv = 32 * ROWV;
  uint64_t pol = 0;
  if (POL == 1) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  if (POL == 2) asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_unchanged.b64 %0, 0.5;" : "=l"(pol));
  for (int64_t c0 = (int64_t)blockIdx.x * chunk_rows; c0 < n_rows; c0 += (int64_t)gridDim.x * chunk_rows)
    for (int it = 0; it * 8 < chunk_rows; ++it) {
      const int64_t i = c0 + it * 8 + wi;
      if (i >= n_rows) break;
      uint4 buf[NOFF][ROWV];
#pragma unroll
      for (int u = 0; u < NOFF; ++u) {
        int64_t r = i + off[u];
        r = r < 0 ? i : (r >= n_rows ? i : r);
#pragma unroll
        for (int v = 0; v < ROWV; ++v) buf[u][v] = POL ? ld_pol(x + r * rowv + lane + 32 * v, pol) : __ldg(x + r * rowv + lane + 32 * v);
      }
      uint4 acc[ROWV];
#pragma unroll
      for (int v = 0; v < ROWV; ++v) acc[v] = make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int u = 0; u < NOFF; ++u)
#pragma unroll
        for (int v = 0; v < ROWV; ++v) {
          acc[v].x += buf[u][v].x; acc[v].y += buf[u][v].y; acc[v].z += buf[u][v].z; acc[v].w += buf[u][v].w;
        }
#pragma unroll
      for (int v = 0; v < ROWV; ++v) __stcs(out + i * rowv + lane + 32 * v, acc[v]);
    }
}

template <int NOFF, int ROWV, int POL = 0>
void run(const char* name, const uint4* x, uint4* out, int64_t n, const int* h_off, int ctas_per_sm, int chunk_rows) {
  int* d_off;
  CK(cudaMalloc(&d_off, NOFF * 4));
  CK(cudaMemcpy(d_off, h_off, NOFF * 4, cudaMemcpyHostToDevice));
  int maxb = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&maxb, gather_kernel<NOFF, ROWV, POL>, 256, 0));
  int per = ctas_per_sm < maxb ? ctas_per_sm : maxb;
  int grid = 148 * per;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int w = 0; w < 2; ++w) gather_kernel<NOFF, ROWV, POL><<<grid, 256>>>(x, out, n, d_off, chunk_rows);
  CK(cudaEventRecord(e0));
  const int reps = 5;
  for (int r = 0; r < reps; ++r) gather_kernel<NOFF, ROWV, POL><<<grid, 256>>>(x, out, n, d_off, chunk_rows);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
  double rowb = 512.0 * ROWV;
  printf("%-34s noff=%d rowB=%4.0f ctas/sm=%d(max %d) chunk=%3d: %7.3f ms  alg %6.0f GB/s  gather %6.0f GB/s\n", name, NOFF, rowb, per, maxb,
         chunk_rows, ms, 2 * n * rowb / ms / 1e6, (NOFF + 1) * n * rowb / ms / 1e6);
  CK(cudaFree(d_off));
}

int main() {
  const int64_t n = 10000000;
  uint4 *x, *out;
  CK(cudaMalloc(&x, n * 1024)); CK(cudaMalloc(&out, n * 1024));
  CK(cudaMemset(x, 1, n * 1024));
  const int mesh7[7] = {-50000, -250, -1, 0, 1, 250, 50000};
  const int near5[5] = {-250, -1, 0, 1, 250};
  const int near3[3] = {-1, 0, 1};
  const int self1[1] = {0};
  const int far3[3] = {-50000, 0, 50000};
  const int farfar3[3] = {-2000000, 0, 2000000};
  // reuse-distance sweep: {self, +-D rows}; with perfect L2 reuse DRAM traffic is 10.2 GB (1.6 ms), without 20.5 GB
  for (int D : {500, 2000, 5000, 10000, 20000, 35000, 50000, 100000}) {
    int o[3] = {-D, 0, D};
    char nm[64];
    snprintf(nm, 64, "far3 D=%d (%.1f MB apart)", D, D * 512e-6);
    run<3, 1, 0>(nm, x, out, n, o, 8, 128);
  }
  for (int D : {20000, 50000}) {
    int o[3] = {-D, 0, D};
    char nm[64];
    snprintf(nm, 64, "far3 D=%d evict_last", D);
    run<3, 1, 1>(nm, x, out, n, o, 8, 128);
    snprintf(nm, 64, "far3 D=%d evict_last 50%%", D);
    run<3, 1, 2>(nm, x, out, n, o, 8, 128);
  }
  run<7, 1, 0>("mesh7 default", x, out, n, mesh7, 8, 128);
  run<7, 1, 1>("mesh7 evict_last", x, out, n, mesh7, 8, 128);
  run<7, 1, 2>("mesh7 evict_last 50%", x, out, n, mesh7, 8, 128);
  for (int chunk : {256, 512, 2048}) run<7, 1, 0>("mesh7 big chunks", x, out, n, mesh7, 8, chunk);
  return 0;
}
