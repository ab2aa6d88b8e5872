// api.cu — library-wide entry points of libb2g.so.
#include "common.cuh"

namespace b2g {
std::atomic<int64_t> g_launches{0};

static uint64_t* g_epoch[64] = {nullptr};
const uint64_t* dropout_epoch_ptr() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!g_epoch[dev]) {
    uint64_t* p = nullptr;
    if (cudaMalloc(&p, sizeof(uint64_t)) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
    cudaMemset(p, 0, sizeof(uint64_t));
    g_epoch[dev] = p;
  }
  return g_epoch[dev];
}
// 1 KB of zeros per device: the address padding lanes of the gather kernels load from (an unconditional load of zeros
// instead of a predicated load, which ptxas implements as load-to-temporary + predicated move = a wait on every load)
static const void* g_zero_row[64] = {nullptr};
const void* zero_row_ptr() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!g_zero_row[dev]) {
    void* p = nullptr;
    if (cudaMalloc(&p, 1024) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
    cudaMemset(p, 0, 1024);
    g_zero_row[dev] = p;
  }
  return g_zero_row[dev];
}
__global__ void epoch_advance_kernel(uint64_t* e) { *e += 1; }
__global__ void epoch_set_kernel(uint64_t* e, uint64_t v) { *e = v; }
}

extern "C" {

int b2g_version(void) { return B2G_VERSION; }

int64_t b2g_launch_count(void) { return b2g::g_launches.load(std::memory_order_relaxed); }
void b2g_launch_count_reset(void) { b2g::g_launches.store(0, std::memory_order_relaxed); }

int b2g_dropout_epoch_advance(void* stream) {
  uint64_t* e = const_cast<uint64_t*>(b2g::dropout_epoch_ptr());
  if (!e) return B2G_E_UNSUPPORTED;
  b2g::epoch_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(e);
  b2g::count_launch();
  return b2g::cuda_status();
}

int b2g_dropout_epoch_set(uint64_t value, void* stream) {
  uint64_t* e = const_cast<uint64_t*>(b2g::dropout_epoch_ptr());
  if (!e) return B2G_E_UNSUPPORTED;
  b2g::epoch_set_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(e, value);
  b2g::count_launch();
  return b2g::cuda_status();
}

const char* b2g_error_string(int code) {
  switch (code) {
    case B2G_OK: return "ok";
    case B2G_E_ARG: return "b2g: invalid argument (null pointer, negative size or bad enum)";
    case B2G_E_ALIGN: return "b2g: pointer or row stride not 16-byte aligned";
    case B2G_E_SHAPE: return "b2g: unsupported feature width / heads combination";
    case B2G_E_RANGE: return "b2g: size does not fit the int32 CSR index type";
    case B2G_E_UNSUPPORTED: return "b2g: no kernel for this request in this build";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "b2g: unknown error";
}

}  // extern "C"
