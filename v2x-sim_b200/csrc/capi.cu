// C-ABI plumbing: version, thread-local error string, device capability probe.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace v2x {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return V2X_ERR_CUDA;
}

bool pdl_enabled() {
  static const bool on = getenv("V2X_NO_PDL") == nullptr;
  return on;
}

}  // namespace v2x

extern "C" int v2x_version(void) { return 1; }

extern "C" const char* v2x_last_error(void) { return v2x::g_err; }

extern "C" int v2x_device_ok(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return major == 10 ? 1 : 0;
}
