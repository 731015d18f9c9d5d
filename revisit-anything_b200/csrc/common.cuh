// Shared helpers for libsegvlad (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/segvlad.h"

namespace segvlad {

void set_error(const char* fmt, ...);
void count_launch();
int prof_begin(int tag, cudaStream_t st);  // no-ops unless segvlad_profile_enable(1)
void prof_end(int slot, cudaStream_t st);  // every kernel launch of the library is counted (segvlad_launch_count)

#define SV_CHECK_CUDA(expr)                                                                   \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      ::segvlad::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, cudaGetErrorName(_e), \
                           cudaGetErrorString(_e));                                           \
      return SEGVLAD_ECUDA;                                                                   \
    }                                                                                         \
  } while (0)

#define SV_REQUIRE(cond, ...)              \
  do {                                     \
    if (!(cond)) {                         \
      ::segvlad::set_error(__VA_ARGS__);   \
      return SEGVLAD_EINVAL;               \
    }                                      \
  } while (0)

#define SV_CHECK_LAUNCH()            \
  do {                               \
    ::segvlad::count_launch();       \
    SV_CHECK_CUDA(cudaGetLastError()); \
  } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over a caller-provided workspace.
struct Carver {
  char* base;
  size_t off;
  explicit Carver(void* p) : base(reinterpret_cast<char*>(p)), off(0) {}
  template <typename T>
  T* take(size_t n) {
    off = align_up(off, 256);
    T* p = reinterpret_cast<T*>(base ? base + off : nullptr);
    off += n * sizeof(T);
    return p;
  }
  size_t total() const { return align_up(off, 256); }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// order-preserving float <-> uint mapping (for atomic min/max and radix/bitonic keys)
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

}  // namespace segvlad
