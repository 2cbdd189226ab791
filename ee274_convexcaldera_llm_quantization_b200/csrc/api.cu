// Version / status / bookkeeping entry points of libcaldera_b200.
#include "common.cuh"

namespace cb {
long long g_launch_count = 0;
}

extern "C" int cb_version(void) { return CB_VERSION; }

extern "C" int64_t cb_kernel_launch_count(void) {
  return (int64_t)__atomic_load_n(&cb::g_launch_count, __ATOMIC_RELAXED);
}

extern "C" void cb_note_launches(int64_t n) {
  if (n > 0) cb::note_launch((int)n);
}

extern "C" const char* cb_status_string(int status) {
  switch (status) {
    case CB_OK: return "ok";
    case CB_ERR_ARG: return "invalid argument (null pointer, negative size or bad enum)";
    case CB_ERR_BITS: return "Bit-width not supported!";
    case CB_ERR_BLOCK: return "number of elements is not divisible by the block size";
    case CB_ERR_WORKSPACE: return "workspace too small";
    case CB_ERR_UNSUPPORTED: return "configuration not supported by this build";
    default: break;
  }
  if (status >= CB_ERR_NUMERIC) return "numerical failure";
  if (status >= CB_ERR_CUDA_BASE) return cudaGetErrorString((cudaError_t)(status - CB_ERR_CUDA_BASE));
  return "unknown status";
}
