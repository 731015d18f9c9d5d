// PCA-whitening projection of segment descriptors, sm_100a.
//
// Replaces func_vpr.py:1419-1443 apply_pca_transform_from_pkl (sklearn 1.3.2 PCA.transform with whiten=True, fitted at
// place_rec_pca.py:339-342) followed, optionally, by func_vpr.py:1673-1676 normalizeFeat:
//     Y = ((X - mean) @ components^T) / sqrt(explained_variance)          X [S, D_in] fp64, components [D_out, D_in] fp32
// Round-1 implementation, correctness-first: an fp64 FMA tile GEMM on the CUDA cores (the reference computes this in fp64
// on the CPU; 64 x 64 x 16 shared-memory tiles, 4 x 4 fp64 micro-tiles, the mean is subtracted and the fp32 components are
// widened while staging), split over K = D_in so that a single image (S ~ 150 rows) still fills the machine; the split
// partials are summed in a fixed order (deterministic), scaled, and optionally row-normalised WITHOUT eps like the
// reference.  The tensor-core version (split-bf16 mainloop of knn.cu with a two-level accumulation) is the next step.
#include "common.cuh"

namespace segvlad {

constexpr int kPjTile = 64, kPjK = 16;

// part[split][S][Dout] = sum_{d in split range} (X[s][d] - mean[d]) * W[o][d]
__global__ void __launch_bounds__(256)
pca_partial_kernel(const double* __restrict__ X, const float* __restrict__ W, const double* __restrict__ mean, int S,
                   int Din, int Dout, int k_per_split, double* __restrict__ part) {
  __shared__ double As[kPjK][kPjTile + 1], Bs[kPjK][kPjTile + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int row0 = blockIdx.y * kPjTile, col0 = blockIdx.x * kPjTile;
  const int kb = blockIdx.z * k_per_split, ke = min(Din, kb + k_per_split);
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  for (int k0 = kb; k0 < ke; k0 += kPjK) {
    for (int i = threadIdx.x; i < kPjTile * kPjK; i += 256) {
      const int rr = i >> 4, kk = i & 15;
      const int gk = k0 + kk;
      As[kk][rr] = (row0 + rr < S && gk < ke) ? X[(size_t)(row0 + rr) * Din + gk] - mean[gk] : 0.0;
      Bs[kk][rr] = (col0 + rr < Dout && gk < ke) ? (double)W[(size_t)(col0 + rr) * Din + gk] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kPjK; ++kk) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  double* p = part + (size_t)blockIdx.z * S * Dout;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = row0 + ty * 4 + i, c = col0 + tx * 4 + j;
      if (r < S && c < Dout) p[(size_t)r * Dout + c] = acc[i][j];
    }
}

// Y[s][o] = (sum_split part) / sqrt(ev[o]); optional row L2 normalisation without eps.  One CTA per row.
__global__ void __launch_bounds__(256)
pca_finalize_kernel(const double* __restrict__ part, int n_split, const float* __restrict__ ev, int S, int Dout,
                    int normalize_rows, double* __restrict__ Y) {
  const int s = blockIdx.x;
  __shared__ double s_w[8];
  double ss = 0.0;
  for (int o = threadIdx.x; o < Dout; o += 256) {
    double v = 0.0;
    for (int z = 0; z < n_split; ++z) v += part[((size_t)z * S + s) * Dout + o];
    v /= (double)sqrtf(ev[o]);   // sklearn: np.sqrt(explained_variance_) is evaluated in the array's fp32, then divides fp64
    Y[(size_t)s * Dout + o] = v;
    ss += v * v;
  }
  if (!normalize_rows) return;
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = ss;
  __syncthreads();
  double tot = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += s_w[i];
  const double nrm = sqrt(tot);
  for (int o = threadIdx.x; o < Dout; o += 256) Y[(size_t)s * Dout + o] /= nrm;   // zero row -> NaN, like normalizeFeat
}

// shared with project_tc.cu: sum of the K-split partials in a fixed order, 1/sqrt(ev), optional row normalisation
void pca_finalize_launch(const double* part, int n_split, const float* ev, int S, int Dout, int normalize_rows, double* Y,
                         cudaStream_t st) {
  pca_finalize_kernel<<<S, 256, 0, st>>>(part, n_split, ev, S, Dout, normalize_rows, Y);
}

static int pca_splits(int S, int Din, int Dout) {
  const long long tiles = (long long)((S + kPjTile - 1) / kPjTile) * ((Dout + kPjTile - 1) / kPjTile);
  int z = (int)((4 * 148 + tiles - 1) / tiles);   // ~4 CTAs per SM
  const int zmax = (Din + 1023) / 1024;           // keep >= 1024 channels per split
  if (z > zmax) z = zmax;
  if (z < 1) z = 1;
  return z;
}

}  // namespace segvlad

using namespace segvlad;

extern "C" size_t segvlad_pca_workspace_bytes(int S, int D_in, int D_out) {
  if (S <= 0 || D_in <= 0 || D_out <= 0) return 256;
  return align_up((size_t)pca_splits(S, D_in, D_out) * S * D_out * sizeof(double), 256) + 256;
}

extern "C" int segvlad_pca_project(const double* X, int S, int D_in, const float* components, const double* mean,
                                   const float* explained_variance, int D_out, int normalize_rows, double* Y,
                                   void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  SV_REQUIRE(X && components && mean && explained_variance && Y, "pca_project: null pointer");
  SV_REQUIRE(S >= 0 && D_in > 0 && D_out > 0, "pca_project: bad shape");
  if (S == 0) return SEGVLAD_OK;
  const size_t need = segvlad_pca_workspace_bytes(S, D_in, D_out);
  if (!workspace || workspace_bytes < need) {
    set_error("pca_project: workspace %zu < required %zu", workspace_bytes, need);
    return SEGVLAD_EWORKSPACE;
  }
  const int z = pca_splits(S, D_in, D_out);
  int kps = (D_in + z - 1) / z;
  kps = (kps + kPjK - 1) / kPjK * kPjK;
  const int zz = (D_in + kps - 1) / kps;
  double* part = reinterpret_cast<double*>(workspace);
  dim3 grid((D_out + kPjTile - 1) / kPjTile, (S + kPjTile - 1) / kPjTile, zz);
  pca_partial_kernel<<<grid, 256, 0, st>>>(X, components, mean, S, D_in, D_out, kps, part);
  SV_CHECK_LAUNCH();
  pca_finalize_kernel<<<S, 256, 0, st>>>(part, zz, explained_variance, S, D_out, normalize_rows, Y);
  SV_CHECK_LAUNCH();
  return SEGVLAD_OK;
}
