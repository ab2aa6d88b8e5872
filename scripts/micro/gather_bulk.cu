// Microbenchmark (NOT part of libb2g.so): the cfg4 mesh gather with the neighbour rows
// staged in shared memory by cp.async.bulk instead of held in registers by LDG (option (b)).
//   out[i] = sum_u x[i + off_u],  7 offsets of the hex stencil (0, +-1, +-nx, +-nx*ny), rows of 512 bytes (bf16 F = 256),
//   offsets arithmetic (no index loads) like gather_bench.cu, so the two programs bound the same access pattern.
// Per warp: a ring of STAGES row slots (7 x 512 B each).  Issue of row k: lane 0 arms the slot's mbarrier with the byte
// count and launches seven 512-byte bulk copies (UBLKCP) that signal it; consumption: wait on the mbarrier, 7
// LDS.128 per lane, integer adds (the memory-system ceiling, not the bf16 arithmetic), 16-byte streaming store.
// Question it answers: does the TMA unit sustain one 512-byte request per ~8 cycles per SM, i.e. does the kernel reach
// the ~1.9 ms shared-memory bound (HBM floor 1.67 ms) where the LDG version measures 2.5-2.7 ms?
// Measured on a B200: 2.42 ms at best (16 warps x 3 slots, 96-row chunks), 2.5-2.7 ms for 8 x 4 / 16 x 2 -> no.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o gather_bulk gather_bulk.cu && ./gather_bulk
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int NOFF = 7;
constexpr int ROWB = 512;
constexpr int SLOT = 4096;                       // 7 x 512 B, padded to a power of two

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "elect.sync _|P, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n" : "=r"(pred));
  return pred != 0;
}

struct Offs { int v[NOFF]; };

template <int WARPS, int STAGES>
__global__ void __launch_bounds__(WARPS * 32, 1) gather_bulk_kernel(const char* __restrict__ x, char* __restrict__ out,
                                                                    int64_t n_rows, Offs offs, int chunk_rows) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int lane = threadIdx.x & 31;
  // warp id through a shuffle: provably warp-uniform for the compiler, so everything derived from it (row ids, addresses,
  // barrier slots) goes to uniform registers with a plain R2UR instead of an ELECT / R2UR.BROADCAST loop per bulk copy
  const int wi = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  uint8_t* ring = smem + (size_t)wi * STAGES * SLOT;
  const uint32_t ring_s = smem_u32(ring);
  const uint32_t bars = smem_u32(smem + (size_t)WARPS * STAGES * SLOT) + wi * STAGES * 8;
  if (lane == 0)
    for (int s = 0; s < STAGES; ++s) mbar_init(bars + 8 * s, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();

  // this warp's rows: wi, wi + WARPS, ... of chunk blockIdx.x, then of chunk blockIdx.x + gridDim.x, ...  Two cursors
  // (issue runs STAGES rows ahead of consumption) advance incrementally: no division in the loop.
  struct Cursor {
    int64_t c0;
    int j;
  };
  const int per_chunk = chunk_rows / WARPS;      // chunk_rows is a multiple of WARPS (checked by the host)
  auto row_at = [&](const Cursor& c) -> int64_t {
    const int64_t i = c.c0 + (int64_t)c.j * WARPS + wi;
    return (c.c0 < n_rows && i < n_rows) ? i : -1;
  };
  auto advance = [&](Cursor& c) {
    if (++c.j == per_chunk) { c.j = 0; c.c0 += (int64_t)gridDim.x * chunk_rows; }
  };
  // UBLKCP is a uniform-datapath instruction (one per warp, operands in uniform registers): per-lane issue would be
  // serialised by the compiler anyway, so one elected lane (elect.sync) arms the barrier and launches the 7 copies back to back
  auto issue = [&](int64_t i, int s) {
    if (elect_one()) {
      mbar_expect_tx(bars + 8 * s, NOFF * ROWB);
#pragma unroll
      for (int u = 0; u < NOFF; ++u) {
        uint32_t r = (uint32_t)i + (uint32_t)offs.v[u];                 // 32-bit row ids, like the CSR columns
        r = r < (uint32_t)n_rows ? r : (uint32_t)i;                      // a negative row wraps to a huge value
        bulk_g2s(ring_s + s * SLOT + u * ROWB, x + (uint64_t)r * ROWB, ROWB, bars + 8 * s);
      }
    }
  };

  Cursor ci{(int64_t)blockIdx.x * chunk_rows, 0}, cc = ci;
  for (int k = 0; k < STAGES; ++k) {
    const int64_t i = row_at(ci);
    if (i < 0) break;
    issue(i, k);
    advance(ci);
  }
  uint32_t phase = 0;
  int s = 0;
  while (true) {
    const int64_t i = row_at(cc);
    if (i < 0) break;
    advance(cc);
    mbar_wait(bars + 8 * s, phase);
    uint4 acc = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int u = 0; u < NOFF; ++u) {
      const uint4 v = *reinterpret_cast<const uint4*>(ring + s * SLOT + u * ROWB + lane * 16);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    __syncwarp();                                // every lane has its data in registers: the slot may be refilled
    const int64_t inext = row_at(ci);
    if (inext >= 0) {
      issue(inext, s);
      advance(ci);
    }
    __stcs(reinterpret_cast<uint4*>(out + i * ROWB + lane * 16), acc);
    if (++s == STAGES) { s = 0; phase ^= 1; }
  }
}

template <int WARPS, int STAGES>
void run(const char* x, char* out, int64_t n, const Offs& offs, int chunk_rows) {
  const size_t smem = (size_t)WARPS * STAGES * SLOT + WARPS * STAGES * 8;
  if (chunk_rows % WARPS) { printf("chunk_rows must be a multiple of WARPS\n"); exit(1); }
  CK(cudaFuncSetAttribute(gather_bulk_kernel<WARPS, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = 148;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int it = 0; it < 3; ++it) gather_bulk_kernel<WARPS, STAGES><<<grid, WARPS * 32, smem>>>(x, out, n, offs, chunk_rows);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  const int iters = 10;
  for (int it = 0; it < iters; ++it) gather_bulk_kernel<WARPS, STAGES><<<grid, WARPS * 32, smem>>>(x, out, n, offs, chunk_rows);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= iters;
  printf("warps %2d stages %d chunk %5d smem %6zu B: %.3f ms  %.0f GB/s algorithmic (read once + write once)\n", WARPS, STAGES,
         chunk_rows, smem, ms, 2.0 * n * ROWB / (ms * 1e-3) / 1e9);
}

int main() {
  const int nx = 250, ny = 200, nz = 200;
  const int64_t n = (int64_t)nx * ny * nz;
  char *x, *out;
  CK(cudaMalloc(&x, n * ROWB));
  CK(cudaMalloc(&out, n * ROWB));
  CK(cudaMemset(x, 1, n * ROWB));
  Offs offs{{0, -1, 1, -nx, nx, -nx * ny, nx * ny}};
  for (int chunk : {48, 96, 384}) {            // multiples of every WARPS below
    run<8, 4>(x, out, n, offs, chunk);
    run<16, 3>(x, out, n, offs, chunk);
    run<16, 2>(x, out, n, offs, chunk);
    run<24, 2>(x, out, n, offs, chunk);
  }
  // correctness spot check of the last run: every interior byte of x is 1 -> each 32-bit word sums to 7 * 0x01010101
  uint32_t h[4];
  CK(cudaMemcpy(h, out + (n / 2) * ROWB, 16, cudaMemcpyDeviceToHost));
  printf("check: %08x (expect %08x)\n", h[0], 7u * 0x01010101u);
  return 0;
}
