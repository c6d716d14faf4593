// capi.cu -- version / error plumbing of the C ABI (include/b200det.h).
#include <stdarg.h>

#include "common.cuh"

namespace b200 {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace b200

extern "C" int b200_version(void) { return B200DET_VERSION; }
extern "C" const char* b200_last_error_string(void) { return b200::g_err; }
