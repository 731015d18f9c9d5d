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
static unsigned long long g_launches = 0;
void count_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }
unsigned long long launches() { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }
}  // namespace segvlad

// ---- optional per-kernel timing (bench.py's roofline leg): CUDA events recorded on the launch stream ----
namespace segvlad {
static int g_prof_on = 0;
struct ProfSlot { cudaEvent_t a, b; int tag; };
static ProfSlot g_prof[4096];
static int g_prof_n = 0;
int prof_begin(int tag, cudaStream_t st) {
  if (!g_prof_on || g_prof_n >= 4096) return -1;
  ProfSlot& s = g_prof[g_prof_n];
  if (cudaEventCreate(&s.a) != cudaSuccess || cudaEventCreate(&s.b) != cudaSuccess) return -1;
  s.tag = tag;
  cudaEventRecord(s.a, st);
  return g_prof_n++;
}
void prof_end(int slot, cudaStream_t st) {
  if (slot >= 0) cudaEventRecord(g_prof[slot].b, st);
}
}  // namespace segvlad

extern "C" void segvlad_profile_enable(int on) { segvlad::g_prof_on = on; }
extern "C" int segvlad_profile_read(int tag, double* total_ms, int* launches) {
  using namespace segvlad;
  double tot = 0.0;
  int n = 0;
  for (int i = 0; i < g_prof_n; ++i) {
    if (g_prof[i].tag != tag) continue;
    float ms = 0.f;
    if (cudaEventSynchronize(g_prof[i].b) != cudaSuccess || cudaEventElapsedTime(&ms, g_prof[i].a, g_prof[i].b) != cudaSuccess) {
      set_error("profile_read: event error");
      return SEGVLAD_ECUDA;
    }
    tot += ms;
    ++n;
  }
  if (total_ms) *total_ms = tot;
  if (launches) *launches = n;
  return SEGVLAD_OK;
}
extern "C" void segvlad_profile_reset(void) {
  using namespace segvlad;
  for (int i = 0; i < g_prof_n; ++i) { cudaEventDestroy(g_prof[i].a); cudaEventDestroy(g_prof[i].b); }
  g_prof_n = 0;
}

extern "C" int segvlad_version(void) { return 100; }
namespace segvlad { unsigned long long launches(); }
extern "C" uint64_t segvlad_launch_count(void) { return segvlad::launches(); }
extern "C" const char* segvlad_last_error(void) { return segvlad::g_err; }
