// Micro-benchmark: issue cost (cycles per warp instruction per SMSP) of FFMA, FFMA2, DADD, F2F.F64.F32 on sm_100a.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_rates tools/pipe_rates.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(float* out, long long* cyc, int iters) {
  float a[16]; float2 b[16]; double c[16];
  for (int i = 0; i < 16; ++i) { a[i] = threadIdx.x + i; b[i] = make_float2(a[i], a[i] + 1); c[i] = a[i]; }
  float m = 1.0000001f; float2 m2 = make_float2(m, m); double dm = 1e-9;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (OP == 0) a[i] = fmaf(a[i], m, m);
      if (OP == 1) b[i] = __ffma2_rn(b[i], m2, m2);
      if (OP == 2) c[i] += dm;
      if (OP == 3) c[i] += (double)a[i];          // F2F + DADD
      if (OP == 4) { a[i] = fmaf(a[i], m, m); c[i] += dm; }   // both pipes
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 16; ++i) s += a[i] + b[i].x + b[i].y + (float)c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 148 * 8);
  const char* names[] = {"FFMA", "FFMA2", "DADD", "F2F+DADD", "FFMA+DADD"};
  for (int warps : {4, 12}) for (int op = 0; op < 5; ++op) {
    int iters = 2000; long long h[148];
    if (op == 0) k<0><<<148, warps * 32>>>(out, cyc, iters); if (op == 1) k<1><<<148, warps * 32>>>(out, cyc, iters);
    if (op == 2) k<2><<<148, warps * 32>>>(out, cyc, iters); if (op == 3) k<3><<<148, warps * 32>>>(out, cyc, iters);
    if (op == 4) k<4><<<148, warps * 32>>>(out, cyc, iters);
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = (double)h[0] / (iters * 16.0);   // cycles per op-group per warp
    printf("%-10s warps/SM=%2d : %.2f cycles per instruction-group per warp  => %.2f cycles per warp-instr per SMSP\n", names[op], warps, c, c / (warps / 4.0));
  }
  return 0;
}
