// api.cu — library-level entry points: version, error string, device check.
#include "common.cuh"
#include <string.h>

namespace {
thread_local char g_err[512] = "";
}

int mc_set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

extern "C" int mc_version(void) { return MCB200_VERSION; }

extern "C" const char* mc_last_error_string(void) { return g_err; }

extern "C" int mc_device_ok(void) {
  int dev = 0;
  MC_CUDA(cudaGetDevice(&dev));
  int major = 0;
  MC_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) {
    mc_set_error(MC_ERR_SHAPE, "libmcb200 is built for sm_100a only; current device has compute capability major %d", major);
    return 0;
  }
  return 1;
}
