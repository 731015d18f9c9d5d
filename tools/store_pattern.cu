// Dev microbenchmark: achievable write bandwidth of the aggregation kernel's output pattern (fp64 [S][K*D], one CTA per
// SM writing 128-row x 128-channel passes of (image, cluster) blocks) with plain coalesced LSU stores, against variants.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_pattern store_pattern.cu && ./store_pattern
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int S_IMG = 128, K = 64, D = 1536, IMGS = 16;

// mode 0: exact pattern, instruction = 4 rows x 128 B      mode 1: exact pattern, instruction = 1 row x 512 B
// mode 2: linear (each warp a private contiguous stream)    mode 3: exact pattern, but pass-major over a 256-column pass
template <int MODE>
__global__ void k_store(double* __restrict__ out, int n_items, int warps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint4 v = make_uint4(lane, warp, blockIdx.x, 7);
  size_t lin = ((size_t)blockIdx.x * warps + warp) * (size_t)(7 * 12 * 8 / (warps / 8)) * 4096 / 16 * 0;  // unused
  (void)lin;
  size_t nlin = 0;
  for (int id = blockIdx.x; id < n_items; id += gridDim.x) {
    const int img = id / K, k = id % K;
    for (int pass = 0; pass < D / 128; ++pass) {
      // the pass = 128 rows x 128 doubles (1 KB per row); split over the warps
      // unit of work: (row group of 32 rows, 16-double column block): 4 x 8 = 32 units per pass
      for (int u = warp; u < 32; u += warps) {
        const int rg = u & 3, cb = u >> 2;   // rows 32 rg .., columns 16 cb ..
        if (MODE == 0) {
          for (int i = 0; i < 8; ++i) {
            const int r = rg * 32 + 4 * i + (lane >> 3), c = lane & 7;
            char* p = reinterpret_cast<char*>(out + ((size_t)(img * S_IMG + r) * K + k) * D + pass * 128 + cb * 16) + c * 16;
            asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
          }
        } else if (MODE == 1) {
          // unit = (8 rows, 64 doubles): u -> rows 8 (u & 15).., columns 64 (u >> 4)..; instruction = one row x 512 B
          const int r0 = (u & 15) * 8, c0 = (u >> 4) * 64;
          for (int i = 0; i < 8; ++i) {
            char* p = reinterpret_cast<char*>(out + ((size_t)(img * S_IMG + r0 + i) * K + k) * D + pass * 128 + c0) + lane * 16;
            asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
          }
        } else if (MODE == 2) {
          char* base = reinterpret_cast<char*>(out) + ((size_t)blockIdx.x * warps + warp) * ((size_t)IMGS * S_IMG * K * D * 8 / (148 * warps) / 4096 * 4096);
          for (int i = 0; i < 8; ++i) {
            char* p = base + (nlin % ((size_t)IMGS * S_IMG * K * D * 8 / (148 * warps) / 4096)) * 4096 + i * 512 + lane * 16;
            asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
          }
          ++nlin;
        }
      }
    }
  }
}

template <int MODE>
static void run(const char* name, double* out, int warps) {
  const int n_items = IMGS * K;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 2; ++i) k_store<MODE><<<148, warps * 32>>>(out, n_items, warps);
  cudaEventRecord(e0);
  for (int i = 0; i < 5; ++i) k_store<MODE><<<148, warps * 32>>>(out, n_items, warps);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  const double bytes = (double)IMGS * S_IMG * K * D * 8;
  printf("%-44s warps %2d  %.4f ms  %.0f GB/s  err=%s\n", name, warps, ms, bytes / ms * 1e-6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  double* out;
  const size_t bytes = (size_t)IMGS * S_IMG * K * D * 8;
  cudaMalloc(&out, bytes);
  cudaMemset(out, 0, bytes);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); for (int i = 0; i < 5; ++i) cudaMemsetAsync(out, 0, bytes); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); printf("memset %.4f ms %.0f GB/s\n", ms / 5, bytes / (ms / 5) * 1e-6);
  for (int w : {8, 16, 32}) run<0>("exact pattern, 4 rows x 128 B per instruction", out, w);
  for (int w : {8, 16, 32}) run<1>("exact pattern, 1 row x 512 B per instruction", out, w);
  for (int w : {8, 16, 32}) run<2>("linear private streams", out, w);
  return 0;
}
