// Error plumbing + version for the C ABI (include/segvlad.h).
#include <stdarg.h>

#include "common.cuh"

namespace segvlad {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace segvlad

extern "C" int segvlad_version(void) { return 100; }
extern "C" const char* segvlad_last_error(void) { return segvlad::g_err; }
