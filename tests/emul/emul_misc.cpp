// TEST INFRASTRUCTURE: the library's error sink for the emulated translation units.
#define SG2IM_EMUL 1
#include <cstdarg>
#include <cstdio>
#include "../../sg2im_b200/csrc/common.cuh"

static char g_err[512];
void sg2im_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char* emul_last_error() { return g_err; }
extern "C" unsigned long long emul_blocks_run() { return emul::blocks_run(); }
#ifndef SG2IM_EMUL_THREADS
extern "C" unsigned long long emul_cluster_blocks_run() { return emul::cluster_blocks_run(); }
#endif
// the few abi.cu entry points the Python binding expects from a loaded library
extern "C" const char* sg2im_last_error_string() { return g_err; }
extern "C" int sg2im_abi_version() { return 1; }
extern "C" int sg2im_device_ok() { return 1; }
