// Microbenchmark 2: start from the arithmetic-offset gather (2.4 ms) and add the real kernel's ingredients one
// at a time: (F) bf16->fp32 packed FMA accumulation, (I) indices from CSR arrays, (D) the opaque-dependency trick.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void fma8(float* acc, float w, uint4 v) {
  asm volatile(
      "{\n.reg .b32 l0,h0,l1,h1,l2,h2,l3,h3;\n.reg .b64 ww,f0,f1,f2,f3,a0,a1,a2,a3;\n"
      "shl.b32 l0,%9,16;\n and.b32 h0,%9,0xffff0000;\n shl.b32 l1,%10,16;\n and.b32 h1,%10,0xffff0000;\n"
      "shl.b32 l2,%11,16;\n and.b32 h2,%11,0xffff0000;\n shl.b32 l3,%12,16;\n and.b32 h3,%12,0xffff0000;\n"
      "mov.b64 ww,{%8,%8};\n mov.b64 f0,{l0,h0};\n mov.b64 f1,{l1,h1};\n mov.b64 f2,{l2,h2};\n mov.b64 f3,{l3,h3};\n"
      "mov.b64 a0,{%0,%1};\n mov.b64 a1,{%2,%3};\n mov.b64 a2,{%4,%5};\n mov.b64 a3,{%6,%7};\n"
      "fma.rn.f32x2 a0,ww,f0,a0;\n fma.rn.f32x2 a1,ww,f1,a1;\n fma.rn.f32x2 a2,ww,f2,a2;\n fma.rn.f32x2 a3,ww,f3,a3;\n"
      "mov.b64 {%0,%1},a0;\n mov.b64 {%2,%3},a1;\n mov.b64 {%4,%5},a2;\n mov.b64 {%6,%7},a3;\n}\n"
      : "+f"(acc[0]), "+f"(acc[1]), "+f"(acc[2]), "+f"(acc[3]), "+f"(acc[4]), "+f"(acc[5]), "+f"(acc[6]), "+f"(acc[7])
      : "f"(w), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
}

// FLT: fp32 FMA accumulate; IDX: 0 arithmetic offsets, 1 CSR (rowptr/col) ; DEP: opaque dependency; MINB: launch-bounds min blocks
template <int FLT, int IDX, int DEP, int MINB>
__global__ void __launch_bounds__(256, MINB) k(const uint4* __restrict__ x, uint4* __restrict__ out, int64_t n_rows,
                                               const int* __restrict__ rowptr, const int* __restrict__ col, int chunk_rows) {
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  constexpr int U = 8;
  for (int64_t c0 = (int64_t)blockIdx.x * chunk_rows; c0 < n_rows; c0 += (int64_t)gridDim.x * chunk_rows)
    for (int it = 0; it * 8 < chunk_rows; ++it) {
      const int64_t i = c0 + it * 8 + wi;
      if (i >= n_rows) break;
      int c[U];
      float w[U];
      if (IDX) {
        const int b = __ldg(rowptr + i), e = __ldg(rowptr + i + 1);
#pragma unroll
        for (int u = 0; u < U; ++u) { c[u] = (b + u < e) ? __ldg(col + b + u) : (int)i; w[u] = (b + u < e) ? 1.f : 0.f; }
      } else {
        const int off[U] = {-50000, -250, -1, 1, 250, 50000, 0, 0};
#pragma unroll
        for (int u = 0; u < U; ++u) { int64_t r = i + off[u]; c[u] = (int)((r < 0 || r >= n_rows) ? i : r); w[u] = u < 7 ? 1.f : 0.f; }
      }
      uint4 buf[U];
#pragma unroll
      for (int u = 0; u < U; ++u) buf[u] = __ldg(x + (int64_t)c[u] * 32 + lane);
      if (DEP) {
        uint32_t dep = 0;
#pragma unroll
        for (int u = 0; u < U; ++u) dep ^= buf[u].x;
        dep ^= __shfl_sync(0xffffffffu, dep, lane);
#pragma unroll
        for (int u = 0; u < U; ++u) w[u] = __uint_as_float(__float_as_uint(w[u]) | dep);
      }
      if (FLT) {
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int u = 0; u < U; ++u) fma8(acc, w[u], buf[u]);
        uint4 o;
        o.x = __float_as_uint(acc[0]) >> 16 | (__float_as_uint(acc[1]) & 0xffff0000u);
        o.y = __float_as_uint(acc[2]) >> 16 | (__float_as_uint(acc[3]) & 0xffff0000u);
        o.z = __float_as_uint(acc[4]) >> 16 | (__float_as_uint(acc[5]) & 0xffff0000u);
        o.w = __float_as_uint(acc[6]) >> 16 | (__float_as_uint(acc[7]) & 0xffff0000u);
        __stcs(out + i * 32 + lane, o);
      } else {
        uint4 a = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < U; ++u) if (w[u] != 0.f) { a.x += buf[u].x; a.y += buf[u].y; a.z += buf[u].z; a.w += buf[u].w; }
        __stcs(out + i * 32 + lane, a);
      }
    }
}

template <int FLT, int IDX, int DEP, int MINB>
void run(const char* name, const uint4* x, uint4* out, int64_t n, const int* rowptr, const int* col, int chunk) {
  int maxb = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&maxb, k<FLT, IDX, DEP, MINB>, 256, 0));
  cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k<FLT, IDX, DEP, MINB>));
  int grid = 148 * maxb;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int w = 0; w < 2; ++w) k<FLT, IDX, DEP, MINB><<<grid, 256>>>(x, out, n, rowptr, col, chunk);
  CK(cudaEventRecord(e0));
  for (int r = 0; r < 5; ++r) k<FLT, IDX, DEP, MINB><<<grid, 256>>>(x, out, n, rowptr, col, chunk);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 5;
  printf("%-44s flt=%d idx=%d dep=%d regs=%3d local=%3zu ctas/sm=%d chunk=%3d: %7.3f ms  alg %5.0f GB/s (%4.1f%%)\n", name, FLT, IDX, DEP,
         fa.numRegs, fa.localSizeBytes, maxb, chunk, ms, 2.0 * n * 512 / ms / 1e6, 2.0 * n * 512 / ms / 1e6 / 65.53);
}

int main() {
  const int nx = 250, ny = 200, nz = 200;
  const int64_t n = (int64_t)nx * ny * nz;
  uint4 *x, *out;
  CK(cudaMalloc(&x, n * 512)); CK(cudaMalloc(&out, n * 512)); CK(cudaMemset(x, 1, n * 512));
  std::vector<int> rp(n + 1), cl; cl.reserve(7 * n);
  for (int64_t i = 0; i < n; ++i) {
    rp[i] = (int)cl.size();
    int ix = i % nx, iy = (i / nx) % ny, iz = i / (nx * ny);
    if (iz > 0) cl.push_back(i - nx * ny);
    if (iy > 0) cl.push_back(i - nx);
    if (ix > 0) cl.push_back(i - 1);
    if (ix < nx - 1) cl.push_back(i + 1);
    if (iy < ny - 1) cl.push_back(i + nx);
    if (iz < nz - 1) cl.push_back(i + nx * ny);
    cl.push_back(i);
  }
  rp[n] = (int)cl.size();
  int *d_rp, *d_cl;
  CK(cudaMalloc(&d_rp, (n + 1) * 4)); CK(cudaMalloc(&d_cl, cl.size() * 4));
  CK(cudaMemcpy(d_rp, rp.data(), (n + 1) * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_cl, cl.data(), cl.size() * 4, cudaMemcpyHostToDevice));
  for (int chunk : {32, 128}) {
    run<0, 0, 0, 1>("M0 arithmetic offsets, int add", x, out, n, d_rp, d_cl, chunk);
    run<0, 0, 1, 1>("M0+D", x, out, n, d_rp, d_cl, chunk);
    run<1, 0, 0, 1>("M1 +fp32 FMA2 accumulate", x, out, n, d_rp, d_cl, chunk);
    run<1, 0, 1, 1>("M1+D", x, out, n, d_rp, d_cl, chunk);
    run<0, 1, 0, 1>("M2 CSR indices, int add", x, out, n, d_rp, d_cl, chunk);
    run<0, 1, 1, 1>("M2+D", x, out, n, d_rp, d_cl, chunk);
    run<1, 1, 0, 1>("M3 CSR indices + fp32 FMA2", x, out, n, d_rp, d_cl, chunk);
    run<1, 1, 1, 1>("M3+D (= the real kernel's structure)", x, out, n, d_rp, d_cl, chunk);
    run<1, 1, 1, 4>("M3+D minBlocks=4", x, out, n, d_rp, d_cl, chunk);
    run<1, 1, 1, 5>("M3+D minBlocks=5", x, out, n, d_rp, d_cl, chunk);
    run<1, 1, 1, 6>("M3+D minBlocks=6", x, out, n, d_rp, d_cl, chunk);
  }
  return 0;
}
