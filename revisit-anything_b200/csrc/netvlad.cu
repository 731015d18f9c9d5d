// NetVLAD + anti-burst aggregation (BASELINE config 5), sm_100a, fp32.
//
// Replaces VLAD-BuFF/models/aggregators/aggregation.py:266-361 NetVLAD.forward with antiburst=True (defaults
// ab_relu/ab_inv/ab_soft False, eval.py:384-386) and getWeights :148-162:
//   x_hat = normalize(x) over D;  a = softmax_k(W x_hat);  selfDis = -2 + 2 x_hat^T x_hat  [N,N];
//   w_p = (sum_q sigmoid(ab_w*selfDis_pq + ab_b))^ab_p;  a_kp /= w_p;
//   V_k = sum_p a_kp (x_hat_p - c_k);  intra-norm over D;  flatten;  L2.
// Pure data parallel over images (no collective).  Round-1 implementation: correctness-first fp32 FFMA tiles
// (64x64x16 shared-memory tiles, 4x4 micro-tiles); the N x N self-similarity never touches HBM (sigmoid row sums
// are reduced inside the tile, per-tile partials are summed in a fixed order => deterministic).
#include "common.cuh"

namespace segvlad {

// tokens [B][D][N] -> x_hat [B][N][D] (token-major, unit rows)
__global__ void __launch_bounds__(256)
nv_normalize_kernel(const float* __restrict__ x, int N, int D, float* __restrict__ xh) {
  const int b = blockIdx.y, p0 = blockIdx.x * 32;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int p = p0 + lane;
  const bool ok = p < N;
  const float* xb = x + (size_t)b * D * N;
  __shared__ float s_ss[8][32];
  __shared__ float s_t[8][32][33];
  float ss = 0.f;
  for (int d = w; d < D; d += 8) {
    float v = ok ? xb[(size_t)d * N + p] : 0.f;
    ss = fmaf(v, v, ss);
  }
  s_ss[w][lane] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int ww = 0; ww < 8; ++ww) tot += s_ss[ww][lane];
  const float nrm = fmaxf(sqrtf(tot), 1e-12f);
  const int dchunk = ((D + 7) / 8 + 31) / 32 * 32;
  const int d0 = w * dchunk, d1 = min(D, d0 + dchunk);
  for (int dd = d0; dd < d1; dd += 32) {
    for (int r = 0; r < 32; ++r) {
      int d = dd + r;
      s_t[w][r][lane] = (d < d1 && ok) ? xb[(size_t)d * N + p] / nrm : 0.f;
    }
    __syncwarp();
    const int d = dd + lane;
    if (d < d1)
      for (int t = 0; t < 32; ++t)
        if (p0 + t < N) xh[((size_t)b * N + p0 + t) * D + d] = s_t[w][lane][t];
    __syncwarp();
  }
}

// C[i][j] = sum_d A[i][d] * Bm[j][d] for a 64x64 tile; A rows i0.., B rows j0.. (both row-major with pitch D)
__device__ __forceinline__ void nv_tile_nt(const float* __restrict__ A, int na, int i0, const float* __restrict__ Bm, int nb,
                                           int j0, int D, float (&acc)[4][4], float (*As)[65], float (*Bs)[65]) {
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < D; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      const int rr = i >> 4, kk = i & 15;
      As[kk][rr] = (i0 + rr < na && k0 + kk < D) ? A[(size_t)(i0 + rr) * D + k0 + kk] : 0.f;
      Bs[kk][rr] = (j0 + rr < nb && k0 + kk < D) ? Bm[(size_t)(j0 + rr) * D + k0 + kk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; bb[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
}

// burst weights, partial row sums: part[b][qt][p] = sum_{q in tile qt} sigmoid(ab_w*(2<x_p,x_q> - 2) + ab_b)
__global__ void __launch_bounds__(256)
nv_burst_kernel(const float* __restrict__ xh, int N, int D, float ab_w, float ab_b, float* __restrict__ part, int n_qt) {
  __shared__ float As[16][65], Bs[16][65];
  __shared__ float s_row[64][17];
  const int b = blockIdx.z, i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
  const float* X = xh + (size_t)b * N * D;
  float acc[4][4];
  nv_tile_nt(X, N, i0, X, N, j0, D, acc, As, Bs);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j0 + tx * 4 + j < N) {
        const float dis = -2.f + 2.f * acc[i][j];
        const float z = dis * ab_w + ab_b;
        s += 1.f / (1.f + expf(-z));
      }
    }
    s_row[ty * 4 + i][tx] = s;
  }
  __syncthreads();
  if (threadIdx.x < 64 && i0 + threadIdx.x < N) {
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < 16; ++t) s += s_row[threadIdx.x][t];
    part[((size_t)b * n_qt + blockIdx.x) * N + i0 + threadIdx.x] = s;
  }
}

// logits[b][k][p] = <W_k, x_hat_p>
__global__ void __launch_bounds__(256)
nv_logits_kernel(const float* __restrict__ xh, const float* __restrict__ W, int N, int D, int K,
                 float* __restrict__ logits) {
  __shared__ float As[16][65], Bs[16][65];
  const int b = blockIdx.z, i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;   // i: cluster, j: token
  float acc[4][4];
  nv_tile_nt(W, K, i0, xh + (size_t)b * N * D, N, j0, D, acc, As, Bs);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = i0 + ty * 4 + i, p = j0 + tx * 4 + j;
      if (k < K && p < N) logits[((size_t)b * K + k) * N + p] = acc[i][j];
    }
}

// in place: logits -> a = softmax_k(logits) / w_p,  w_p = (sum_qt part)^ab_p
__global__ void nv_softassign_kernel(float* __restrict__ a, const float* __restrict__ part, int N, int K, int n_qt,
                                     float ab_p) {
  const int b = blockIdx.y, p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= N) return;
  float w = 0.f;
  for (int t = 0; t < n_qt; ++t) w += part[((size_t)b * n_qt + t) * N + p];
  w = (ab_p == 1.f) ? w : powf(w, ab_p);
  float* col = a + (size_t)b * K * N + p;
  float mx = -INFINITY;
  for (int k = 0; k < K; ++k) mx = fmaxf(mx, col[(size_t)k * N]);
  float sum = 0.f;
  for (int k = 0; k < K; ++k) sum += expf(col[(size_t)k * N] - mx);
  for (int k = 0; k < K; ++k) col[(size_t)k * N] = expf(col[(size_t)k * N] - mx) / sum / w;
}

// V[b][k][d] = sum_p a[k][p] * (x_hat[p][d] - c[k][d])   (64 clusters x 64 channels per CTA)
__global__ void __launch_bounds__(256)
nv_vlad_kernel(const float* __restrict__ xh, const float* __restrict__ a, const float* __restrict__ cent, int N, int D,
               int K, float* __restrict__ V) {
  __shared__ float As[16][65], Bs[16][65];
  const int b = blockIdx.z, k0 = blockIdx.y * 64, d0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const float* X = xh + (size_t)b * N * D;
  const float* Ab = a + (size_t)b * K * N;
  float c[4][4], acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + ty * 4 + i, d = d0 + tx * 4 + j;
      c[i][j] = (k < K && d < D) ? cent[(size_t)k * D + d] : 0.f;
      acc[i][j] = 0.f;
    }
  for (int p0 = 0; p0 < N; p0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      const int rr = i & 63, pp = i >> 6;
      As[pp][rr] = (k0 + rr < K && p0 + pp < N) ? Ab[(size_t)(k0 + rr) * N + p0 + pp] : 0.f;
      Bs[pp][rr] = (d0 + rr < D && p0 + pp < N) ? X[(size_t)(p0 + pp) * D + d0 + rr] : 0.f;
    }
    __syncthreads();
    const int pmax = min(16, N - p0);
    for (int pp = 0; pp < pmax; ++pp) {
      float av[4], xv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { av[i] = As[pp][ty * 4 + i]; xv[i] = Bs[pp][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], xv[j] - c[i][j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + ty * 4 + i, d = d0 + tx * 4 + j;
      if (k < K && d < D) V[((size_t)b * K + k) * D + d] = acc[i][j];
    }
}

// intra-norm per (b,k) over D, then L2 over the flattened K*D row; one CTA per image
__global__ void __launch_bounds__(1024)
nv_finalize_kernel(const float* __restrict__ V, int D, int K, float* __restrict__ out) {
  const int b = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __shared__ float s_nk[256];
  __shared__ float s_tot;
  const float* Vb = V + (size_t)b * K * D;
  for (int k = w; k < K; k += 32) {
    float ss = 0.f;
    for (int d = lane; d < D; d += 32) { float v = Vb[(size_t)k * D + d]; ss = fmaf(v, v, ss); }
    ss = warp_sum(ss);
    if (lane == 0) s_nk[k] = fmaxf(sqrtf(ss), 1e-12f);
  }
  __syncthreads();
  // ||flattened|| after intra-normalisation, evaluated on the normalised values like the reference
  float part = 0.f;
  for (int i = threadIdx.x; i < K * D; i += 1024) {
    float v = Vb[i] / s_nk[i / D];
    part = fmaf(v, v, part);
  }
  part = warp_sum(part);
  __shared__ float s_w[32];
  if (lane == 0) s_w[w] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 32; ++i) t += s_w[i];
    s_tot = fmaxf(sqrtf(t), 1e-12f);
  }
  __syncthreads();
  const float tot = s_tot;
  for (int i = threadIdx.x; i < K * D; i += 1024) out[(size_t)b * K * D + i] = Vb[i] / s_nk[i / D] / tot;
}

void nv_finalize_launch(const float* V, int B, int D, int K, float* out, cudaStream_t st) {
  nv_finalize_kernel<<<B, 1024, 0, st>>>(V, D, K, out);
}

// tensor-core path (netvlad_tc.cu)
bool nv_tc_supported(int N, int D, int K);
size_t nv_tc_workspace_bytes(int B, int N, int D, int K);
int nv_tc_run(const float* x, int B, int N, int D, const float* centroids, const float* conv_weight, int K, float ab_w,
              float ab_b, float ab_p, float* out, void* workspace, cudaStream_t st);

struct NvLayout { float* xh; float* part; float* a; float* V; size_t total; };
static NvLayout carve_nv(void* ws, int B, int N, int D, int K) {
  Carver c(ws);
  NvLayout L;
  const int n_qt = (N + 63) / 64;
  L.xh = c.take<float>((size_t)B * N * D);
  L.part = c.take<float>((size_t)B * n_qt * N);
  L.a = c.take<float>((size_t)B * K * N);
  L.V = c.take<float>((size_t)B * K * D);
  L.total = c.total();
  return L;
}

}  // namespace segvlad

using namespace segvlad;

extern "C" size_t segvlad_netvlad_workspace_bytes(int B, int N, int D, int K) {
  if (B <= 0 || N <= 0 || D <= 0 || K <= 0) return 0;
  const size_t simt = carve_nv(nullptr, B, N, D, K).total;
  const size_t tc = (K <= 128 && K % 16 == 0) ? nv_tc_workspace_bytes(B, N, D, K) : 0;   // (independent of the env switch)
  return simt > tc ? simt : tc;
}

extern "C" int segvlad_netvlad_antiburst(const float* x, int B, int N, int D, const float* centroids,
                                         const float* conv_weight, int K, float ab_w, float ab_b, float ab_p, float* out,
                                         void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  SV_REQUIRE(x && centroids && conv_weight && out, "netvlad: null pointer");
  SV_REQUIRE(B > 0 && N > 0 && D > 0 && K > 0 && K <= 256, "netvlad: bad shape (K <= 256)");
  const size_t need = segvlad_netvlad_workspace_bytes(B, N, D, K);
  if (!workspace || workspace_bytes < need) {
    set_error("netvlad: workspace %zu < required %zu", workspace_bytes, need);
    return SEGVLAD_EWORKSPACE;
  }
  if (nv_tc_supported(N, D, K))   // tcgen05: fp16 (hi, lo) split operands, three products; the FFMA kernels below stay as the
    return nv_tc_run(x, B, N, D, centroids, conv_weight, K, ab_w, ab_b, ab_p, out, workspace, st);   // cross-check / fallback
  NvLayout L = carve_nv(workspace, B, N, D, K);
  const int n_qt = (N + 63) / 64;
  nv_normalize_kernel<<<dim3((N + 31) / 32, B), 256, 0, st>>>(x, N, D, L.xh);
  SV_CHECK_LAUNCH();
  nv_burst_kernel<<<dim3(n_qt, n_qt, B), 256, 0, st>>>(L.xh, N, D, ab_w, ab_b, L.part, n_qt);
  SV_CHECK_LAUNCH();
  nv_logits_kernel<<<dim3(n_qt, (K + 63) / 64, B), 256, 0, st>>>(L.xh, conv_weight, N, D, K, L.a);
  SV_CHECK_LAUNCH();
  nv_softassign_kernel<<<dim3((N + 127) / 128, B), 128, 0, st>>>(L.a, L.part, N, K, n_qt, ab_p);
  SV_CHECK_LAUNCH();
  nv_vlad_kernel<<<dim3((D + 63) / 64, (K + 63) / 64, B), 256, 0, st>>>(L.xh, L.a, centroids, N, D, K, L.V);
  SV_CHECK_LAUNCH();
  nv_finalize_kernel<<<B, 1024, 0, st>>>(L.V, D, K, out);
  SV_CHECK_LAUNCH();
  return SEGVLAD_OK;
}
