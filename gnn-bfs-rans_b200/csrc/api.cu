// api.cu — library-wide entry points of libb2g.so.
#include "common.cuh"

namespace b2g {
int64_t g_launches = 0;
}

extern "C" {

int b2g_version(void) { return B2G_VERSION; }

int64_t b2g_launch_count(void) { return b2g::g_launches; }
void b2g_launch_count_reset(void) { b2g::g_launches = 0; }

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
