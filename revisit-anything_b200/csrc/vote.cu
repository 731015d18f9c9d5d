// Segment -> image vote for sm_100a.
//
// Replaces func_vpr.py:80-243 get_matches, branch "max_seg_topk_wt_borda_Im" (:207-224) with
// weighted_borda_count (:61-77), plus the integer np.bincount of "max_seg_topk" (:118-125).
//
// Semantics reproduced exactly:
//   lo, hi = min / max over the WHOLE [Nq, k_vote] sims array (fp32);
//   value(t) = fl32( fl32(s - lo) / fl32(hi - lo) ), promoted to double;
//   per query image, hits are visited in the order t = rank * n_seg + seg (rank-major); the score of a
//   reference image is the SEQUENTIAL double sum of its hits in that order; images are ranked by score
//   descending, ties (and NaN) keep first-insertion order (Python's stable sorted(reverse=True)).
// One CTA per query image: hits are keyed (image << 32 | t), bitonic-sorted (=> grouped by image with
// t ascending), each group is summed sequentially by its head thread => bit-identical to the Python
// loop, no float atomics.  Images with more hits than fit in shared memory use the same code on a
// global scratch region.
#include "common.cuh"

namespace segvlad {

constexpr int kVoteThreads = 1024;
constexpr int kVoteSmemHits = 8192;  // 24 B per hit slot -> 192 KB dynamic shared memory

// kNN padding entries (match < 0: a bank or shard with fewer than k_vote rows) are not hits and carry d2 = +inf; they
// are left out of the global min / max (with them lo = -inf and every vote weight would be NaN)
__global__ void minmax_kernel(const long long* __restrict__ matches, const float* __restrict__ sims, long long total, int ld,
                              int kv, int is_d2, uint32_t* __restrict__ mm) {
  float lo = INFINITY, hi = -INFINITY;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long row = i / kv;
    int col = (int)(i - row * kv);
    if (matches[row * ld + col] < 0) continue;
    float v = sims[row * ld + col];
    if (is_d2) v = __fsub_rn(2.0f, v);
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(mm + 0, float_to_ordered(lo));
    atomicMax(mm + 1, float_to_ordered(hi));
  }
}

__global__ void minmax_export_kernel(const uint32_t* __restrict__ mm, float* __restrict__ out) {
  out[0] = ordered_to_float(mm[0]);
  out[1] = ordered_to_float(mm[1]);
}

__device__ __forceinline__ bool vote_better(double sa, unsigned ta, double sb, unsigned tb) {
  // (score desc, first-insertion asc); incomparable (NaN) scores count as equal
  if (sa > sb) return true;
  if (sa < sb) return false;
  return ta < tb;
}

__global__ void __launch_bounds__(kVoteThreads)
vote_kernel(const long long* __restrict__ matches, const float* __restrict__ sims, int ld, int is_d2, int kv,
            const int* __restrict__ qimg_off, const int* __restrict__ qrow, const int* __restrict__ rimg, int Nr,
            int n_rimg, int n_pred,
            const uint32_t* __restrict__ mm, int* __restrict__ preds, double* __restrict__ pred_scores,
            double* __restrict__ scores_dense, int* __restrict__ counts_dense, char* __restrict__ scratch,
            int P_cap) {
  extern __shared__ __align__(16) char smem_raw[];
  const int qi = blockIdx.x;
  const int q0 = qimg_off[qi], n = qimg_off[qi + 1] - q0;
  const int H = n * kv;
  int P = 1;
  while (P < H) P <<= 1;
  char* base = (P <= kVoteSmemHits) ? smem_raw : scratch + (size_t)qi * P_cap * 24;
  const int cap = (P <= kVoteSmemHits) ? P : P_cap;
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(base);
  double* sc = reinterpret_cast<double*>(base + (size_t)cap * 8);
  float* vals = reinterpret_cast<float*>(base + (size_t)cap * 16);
  unsigned* hc = reinterpret_cast<unsigned*>(base + (size_t)cap * 20);
  const int tid = threadIdx.x;

  const float lo = ordered_to_float(mm[0]), hi = ordered_to_float(mm[1]);
  const float rng = __fsub_rn(hi, lo);
  for (int t = tid; t < P; t += kVoteThreads) {
    unsigned long long key = ~0ull;
    if (t < H) {
      const int s = t % n, k = t / n;
      const size_t off = (size_t)(qrow ? qrow[q0 + s] : q0 + s) * ld + k;
      const long long m = matches[off];
      float v = sims[off];
      if (is_d2) v = __fsub_rn(2.0f, v);
      vals[t] = __fdiv_rn(__fsub_rn(v, lo), rng);
      if (m >= 0 && m < Nr) key = ((unsigned long long)(unsigned)rimg[m] << 32) | (unsigned)t;
    }
    keys[t] = key;
    hc[t] = 0;
  }
  __syncthreads();
  // bitonic sort, ascending; t enumerates the P / 2 compare-exchanges of a stage (i = t with a zero bit inserted at distance j)
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (P >> 1); t += kVoteThreads) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int ixj = i | j;
        const unsigned long long a = keys[i], b = keys[ixj];
        const bool up = (i & k) == 0;
        if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
      }
      __syncthreads();
    }
  }
  // group heads: sequential double sum in t order (== dict accumulation order of weighted_borda_count)
  for (int i = tid; i < P; i += kVoteThreads) {
    const unsigned long long key = keys[i];
    if (key == ~0ull) continue;
    const unsigned img = (unsigned)(key >> 32);
    if (i > 0 && (unsigned)(keys[i - 1] >> 32) == img) continue;
    double acc = (double)vals[(unsigned)key];
    unsigned cnt = 1;
    for (int j = i + 1; j < P; ++j) {
      const unsigned long long kj = keys[j];
      if ((unsigned)(kj >> 32) != img || kj == ~0ull) break;
      acc += (double)vals[(unsigned)kj];
      ++cnt;
    }
    sc[i] = acc;
    hc[i] = cnt;
    if (scores_dense && img < (unsigned)n_rimg) scores_dense[(size_t)qi * n_rimg + img] = acc;
    if (counts_dense && img < (unsigned)n_rimg) counts_dense[(size_t)qi * n_rimg + img] = (int)cnt;
  }
  __syncthreads();
  // top-n_pred by (score desc, first insertion asc)
  __shared__ double s_s[32];
  __shared__ unsigned s_t[32];
  __shared__ int s_p[32];
  __shared__ int s_win;
  for (int r = 0; r < n_pred; ++r) {
    double bs = 0.0;
    unsigned bt = 0xffffffffu;
    int bp = -1;
    for (int i = tid; i < P; i += kVoteThreads) {
      if (hc[i] == 0) continue;
      const double s = sc[i];
      const unsigned t = (unsigned)keys[i];
      if (bp < 0 || vote_better(s, t, bs, bt)) { bs = s; bt = t; bp = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double os = __shfl_xor_sync(0xffffffffu, bs, o);
      const unsigned ot = __shfl_xor_sync(0xffffffffu, bt, o);
      const int op = __shfl_xor_sync(0xffffffffu, bp, o);
      if (op >= 0 && (bp < 0 || vote_better(os, ot, bs, bt))) { bs = os; bt = ot; bp = op; }
    }
    if ((tid & 31) == 0) { s_s[tid >> 5] = bs; s_t[tid >> 5] = bt; s_p[tid >> 5] = bp; }
    __syncthreads();
    if (tid == 0) {
      double ws = 0.0; unsigned wt = 0xffffffffu; int wp = -1;
      for (int w = 0; w < kVoteThreads / 32; ++w) {
        if (s_p[w] >= 0 && (wp < 0 || vote_better(s_s[w], s_t[w], ws, wt))) { ws = s_s[w]; wt = s_t[w]; wp = s_p[w]; }
      }
      s_win = wp;
      if (wp >= 0) {
        preds[(size_t)qi * n_pred + r] = (int)(keys[wp] >> 32);
        if (pred_scores) pred_scores[(size_t)qi * n_pred + r] = sc[wp];
        hc[wp] = 0;
      } else {
        preds[(size_t)qi * n_pred + r] = -1;
        if (pred_scores) pred_scores[(size_t)qi * n_pred + r] = 0.0;
      }
    }
    __syncthreads();
    if (s_win < 0) {  // no more distinct images: pad the rest
      for (int rr = r + 1 + tid; rr < n_pred; rr += kVoteThreads) {
        preds[(size_t)qi * n_pred + rr] = -1;
        if (pred_scores) pred_scores[(size_t)qi * n_pred + rr] = 0.0;
      }
      break;
    }
  }
}

static int next_pow2(int x) { int p = 1; while (p < x) p <<= 1; return p; }

}  // namespace segvlad

using namespace segvlad;

extern "C" size_t segvlad_vote_workspace_bytes(int Nq, int k_vote, int n_qimg, int max_segs_per_qimg) {
  size_t b = 256;  // min/max words
  long long H = (long long)max_segs_per_qimg * k_vote;
  if (H > kVoteSmemHits) b += (size_t)n_qimg * (size_t)next_pow2((int)H) * 24 + 256;
  (void)Nq;
  return b;
}

extern "C" int segvlad_vote(const int64_t* matches, const float* sims, int ld, int sims_is_d2, int k_vote, int Nq,
                            const int32_t* qimg_offsets, const int32_t* qrow_index, int n_qimg,
                            int max_segs_per_qimg, const int32_t* rseg_to_rimg, int Nr, int n_rimg, int n_pred,
                            int32_t* preds, double* pred_scores, double* scores_dense, int32_t* counts_dense,
                            float* minmax_out, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  SV_REQUIRE(Nq >= 0 && k_vote > 0 && ld >= k_vote && n_qimg >= 0 && n_pred > 0 && Nr > 0 && n_rimg > 0,
             "vote: bad shape");
  SV_REQUIRE((long long)max_segs_per_qimg * k_vote < (1ll << 30), "vote: too many hits per query image");
  if (n_qimg == 0) return SEGVLAD_OK;
  const size_t need = segvlad_vote_workspace_bytes(Nq, k_vote, n_qimg, max_segs_per_qimg);
  if (!workspace || workspace_bytes < need) {
    set_error("vote: workspace %zu < required %zu", workspace_bytes, need);
    return SEGVLAD_EWORKSPACE;
  }
  uint32_t* mm = reinterpret_cast<uint32_t*>(workspace);
  char* scratch = reinterpret_cast<char*>(workspace) + 256;
  const uint32_t init[2] = {0xffffffffu, 0u};
  SV_CHECK_CUDA(cudaMemcpyAsync(mm, init, sizeof(init), cudaMemcpyHostToDevice, st));
  const long long total = (long long)Nq * k_vote;
  if (total > 0) {
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    minmax_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const long long*>(matches), sims, total, ld, k_vote, sims_is_d2,
                                          mm);
    SV_CHECK_LAUNCH();
  }
  if (minmax_out) {
    minmax_export_kernel<<<1, 1, 0, st>>>(mm, minmax_out);
    SV_CHECK_LAUNCH();
  }
  if (scores_dense) SV_CHECK_CUDA(cudaMemsetAsync(scores_dense, 0, sizeof(double) * (size_t)n_qimg * n_rimg, st));
  if (counts_dense) SV_CHECK_CUDA(cudaMemsetAsync(counts_dense, 0, sizeof(int) * (size_t)n_qimg * n_rimg, st));
  const int Pmax = next_pow2((int)((long long)max_segs_per_qimg * k_vote));
  const int P_cap = Pmax > kVoteSmemHits ? Pmax : 0;
  // images whose hits fit use shared memory even when the largest image needs the global scratch
  const size_t smem = (size_t)(Pmax > kVoteSmemHits ? kVoteSmemHits : Pmax) * 24;
  SV_CHECK_CUDA(cudaFuncSetAttribute(vote_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kVoteSmemHits * 24));
  vote_kernel<<<n_qimg, kVoteThreads, smem, st>>>(reinterpret_cast<const long long*>(matches), sims, ld, sims_is_d2,
                                                 k_vote, qimg_offsets, qrow_index, rseg_to_rimg, Nr, n_rimg, n_pred, mm, preds,
                                                 pred_scores, scores_dense, counts_dense, scratch, P_cap);
  SV_CHECK_LAUNCH();
  return SEGVLAD_OK;
}
