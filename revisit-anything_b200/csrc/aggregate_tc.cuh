// Tensor-core (tcgen05) formulation of the masked residual aggregation -- interface between aggregate.cu (driver,
// membership / cluster-list preparation) and aggregate_tc.cu (kernels).
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace segvlad {

constexpr int kTcSegTile = 128;   // segments per MMA tile (M)
constexpr int kTcTokChunk = 64;   // tokens per K chunk (one 128-byte swizzle row of bf16)
constexpr int kTcPassN = 128;     // descriptor channels per accumulator pass (N)
// TMEM accumulation truncates: every accumulating MMA can cost the running sum up to one ulp, coherently (measured ~6e-8
// relative per MMA).  One accumulator chain therefore covers at most kTcSubChunks token chunks (128 tokens = 24 MMAs,
// <= 2.9e-6 relative worst case, ~1.4e-6 typical); the chains of a longer cluster land in successive TMEM buffers and the
// epilogue adds them with fp32 round-to-nearest adds in registers (<= 12 of them for a 1530-token image).
constexpr int kTcSubChunks = 2;

struct AggTcArgs {
  const float* R;             // [B][N][D] fp32 residual rows (token-major), or null: build RT from the tokens (below)
  const float* tokens_dn;     // [B][D][N] tokens, centres [K][D], labels [B][N], ||x|| [B][N] (used when R is null)
  const float* centers;
  const int* labels;
  const float* nrm;
  const int* cl_ptr;          // [B][K+1] cluster boundaries in the label-sorted token order
  const int* cl_tok;          // [B][N]   label-sorted token ids
  const uint16_t* memS;       // [n_groups][N] membership words (bit j = segment 8*g + j), label-sorted order
  const int* cpred;           // [S_total] number of non-empty clusters per segment
  double* norms;              // [S_total][K] block norms (for the row-norm fix-up)
  const int32_t* seg_offsets_host;
  int B, N, D, K, S_total;
  void* out;                  // [S_total][K*D]
  int out_dtype;              // SEGVLAD_OUT_F64 / _F32, or SEGVLAD_OUT_PCA_PLANES: out = [3][S_total][K*D] bf16 planes of
  const float* pca_mean;      // (descriptor - pca_mean) [fp32 copy of the model mean], the A operand of the tensor-core PCA projection (row f1)
  __nv_bfloat16* RT;          // workspace: [3 planes][B][D][Np] transposed, label-sorted bf16 split of R
  int* tile_tbl;              // workspace: [n_tiles][4] = image, first group, first segment, #segments
  double* xnorm;              // workspace: [n_items][128] per-CTA partial sums of squares of the single-sweep items (sibling exchange)
  unsigned long long* probe;  // development aid: per-CTA cycle counters [grid][16] (segvlad_debug_aggregate_probe), or null
};

inline int agg_tc_np(int N) { return (int)align_up((size_t)N, kTcTokChunk); }
inline size_t agg_tc_rt_elems(int B, int N, int D) { return (size_t)3 * B * D * agg_tc_np(N); }
inline int agg_tc_max_tiles(int B, int S_total) { return S_total / kTcSegTile + B; }
inline size_t agg_tc_xnorm_elems(int B, int S_total, int D, int K) {       // channel splits J <= number of passes
  return (size_t)agg_tc_max_tiles(B, S_total) * K * ((D + kTcPassN - 1) / kTcPassN) * kTcSegTile;
}
bool agg_tc_supported(int N, int D, int K);
int agg_tc_fused_channels(int N, int K);               // > 0: RT can be built straight from [D][N] tokens
int agg_tc_run(const AggTcArgs& a, cudaStream_t st);   // returns SEGVLAD_* status

// Tensor-core cosine assignment for [D][N] tokens (assign_tc.cu): labels [B][N], nrm [B][N] = max(||x||, eps) (1 if prenorm)
bool assign_tc_supported(int N, int D, int K);
size_t assign_tc_workspace_elems(int D, int K);        // bf16 elements for the c_hat planes
int assign_tc_run(const float* tokens, int B, int N, int D, const float* centers, int K, int prenorm,
                  __nv_bfloat16* chat_planes, int* labels, float* nrm, cudaStream_t st);

}  // namespace segvlad
