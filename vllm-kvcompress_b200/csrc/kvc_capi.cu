// libkvc_b200 housekeeping: error strings, version, arch.
#include <stdarg.h>

#include "kvc_common.cuh"

namespace kvc {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace kvc

extern "C" const char* kvc_last_error(void) { return kvc::g_err; }
extern "C" int kvc_abi_version(void) { return 1; }
extern "C" const char* kvc_arch(void) { return "sm_100a"; }
