// ABI glue: version, per-thread error string, device check.
#include <stdarg.h>
#include <stdio.h>
#include "common.cuh"

static thread_local char g_err[512] = "no error";

void sg2im_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" int sg2im_abi_version(void) { return 1; }
extern "C" const char* sg2im_last_error_string(void) { return g_err; }

extern "C" int sg2im_device_ok(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}
