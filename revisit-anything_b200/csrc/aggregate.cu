// Per-(Super)Segment masked hard-assignment VLAD aggregation for sm_100a.
//
// Replaces (batched over images) the reference chain
//   seg_vlad_gpu_single  func_vpr.py:1065-1101  -> vlad_single :1140-1179 -> vlad_matmuls_per_cluster :1181-1210
//
// Pipeline (all on the caller's stream, no host sync, no float atomics => bit-reproducible):
//   normalize_centers   c_hat = c / max(||c||,eps)                       (func_vpr.py:1145)
//   assign_*            per token: ||x||, argmax_k <x, c_hat_k>, residual row r = x/max(||x||,eps) - c[label]
//                                                                      (func_vpr.py:1085, 1146, 1151)
//   cluster_lists       counting sort of the image's tokens by label (replaces torch.where per cluster :1197)
//   superseg_union      member'[s] = OR_t adj[s,t] member[t]            (func_vpr.py:1200, bitmask form)
//   group_transpose     token-major membership words for groups of SEG_GROUP segments
//   nonempty            predicted number of non-empty clusters per segment (row norm)
//   aggregate           V[s,k,:] = sum_{p in seg s, label k} (double) r_p ; intra-norm ; row-norm ; store
//                                                                      (func_vpr.py:1201-1205)
//   rownorm_fixup       exact handling of blocks with 0 < ||V|| < eps or exact cancellation
//
// HBM-bound by design: tokens are read once (+ once from L2), the [S, K*D] output is written once with
// streaming stores.  Accumulation is fp64 like the reference (the addends are fp32 values, so the sums
// are exact to ~1e-16 and independent of the summation order in practice).
#include "common.cuh"
#include "aggregate_tc.cuh"

namespace segvlad {

constexpr int kSegGroup = 8;       // segments per aggregate CTA (register tile: 8 x 4 accumulators / thread)
constexpr int kFlushEvery = 16;    // fp32 partial sums are promoted to the fp64 accumulators every 16 rows
constexpr int kRingRows = 16;      // residual rows in flight per CTA (the other half of shared memory stages the output)
constexpr int kRowsPerSlot = 4;    // rows sharing one full/empty mbarrier pair (r1 probe: per-row mbarrier traffic from
                                   // 12 consumer warps + producer, ~26 ops/row, was the consumer's bottleneck)
constexpr int kRingSlots = kRingRows / kRowsPerSlot;
constexpr float kEpsF = 1e-12f;
constexpr double kEpsD = 1e-12;

// ------------------------------------------------------------------------------------------------
__global__ void normalize_centers_kernel(const float* __restrict__ c, int K, int D,
                                         float* __restrict__ chatT, int Kp) {
  // one block per centre; chatT is [D][Kp] (k fastest) so a warp can fetch 32 consecutive k with LDG.128
  int k = blockIdx.x;
  __shared__ float s_w[32];
  float ss = 0.f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float v = c[(size_t)k * D + d];
    ss += v * v;
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = (threadIdx.x < (blockDim.x >> 5)) ? s_w[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) s_w[0] = fmaxf(sqrtf(v), kEpsF);
  }
  __syncthreads();
  float nrm = s_w[0];
  for (int d = threadIdx.x; d < D; d += blockDim.x) chatT[(size_t)d * Kp + k] = c[(size_t)k * D + d] / nrm;
}

// ------------------------------------------------------------------------------------------------
// Assignment + residual rows, tokens in the reference's [D, N] layout.
// v2 (r1: v0 streamed the centre table through L1 with 9 load instructions per 32 FFMAs and ran at ~3 TFLOP/s):
// a register-tiled fp32 GEMM.  CTA = 128 tokens x one 32-cluster chunk at a time, 128 threads; thread (tg, kg) owns
// 8 tokens x 4 clusters (32 accumulators); per channel it issues 2 LDS.128 (tokens) + 1 LDS.128 (centres) for
// 32 FFMAs.  x and c_hat tiles ([16 ch][128 tok], [16 ch][32 k]) are staged with cp.async, double buffered.
// Summation order per (token, cluster): channels ascending, fp32 FMA (the reference's SGEMM is not bit-reproducible;
// labels are compared where the fp64 margin exceeds 1e-5).
constexpr int kAsgTok = 128;   // tokens per CTA
constexpr int kAsgDc = 16;     // channels per pipeline stage

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }

constexpr int kAsgSplit = 8;   // channel splits (split-K) -> enough CTAs for small batches; summed in fixed order

// Stage 1: partial dot products over one channel range.  grid (token tiles, B, kAsgSplit), 128 threads.
//   part[((b*kAsgSplit + split)*N + p)*Kp + k] = sum_{d in range} x[d][p] * c_hat[k][d];   ssq likewise for ||x||^2
__global__ void __launch_bounds__(128)
assign_partial_kernel(const float* __restrict__ tokens, int N, int D, const float* __restrict__ chatT, int K, int Kp,
                      float* __restrict__ part, float* __restrict__ ssq) {
  __shared__ __align__(16) float s_x[2][kAsgDc][kAsgTok];
  __shared__ __align__(16) float s_c[2][kAsgDc][32];
  const int b = blockIdx.y, p0 = blockIdx.x * kAsgTok, split = blockIdx.z;
  const int tid = threadIdx.x;
  const int tg = tid & 15, kg = tid >> 4;          // token group (8 tokens), cluster group (4 clusters)
  const float* tok = tokens + (size_t)b * D * N;
  const int dper = ((D + kAsgSplit - 1) / kAsgSplit + kAsgDc - 1) / kAsgDc * kAsgDc;
  const int d_lo = min(D, split * dper), d_hi = min(D, d_lo + dper);
  const bool vec_ok = (N % 4 == 0) && (p0 + kAsgTok <= N);   // 16-byte aligned, fully inside: cp.async 16 B
  const bool vec2_ok = (N % 2 == 0);

  auto stage = [&](int buf, int d0, int kc) {
    if (vec_ok) {
      for (int i = tid; i < kAsgDc * (kAsgTok / 4); i += 128) {
        const int r = i / (kAsgTok / 4), c4 = i % (kAsgTok / 4);
        const int d = d0 + r;
        if (d < d_hi) cp_async16(&s_x[buf][r][c4 * 4], tok + (size_t)d * N + p0 + c4 * 4);
        else *reinterpret_cast<float4*>(&s_x[buf][r][c4 * 4]) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else if (vec2_ok) {    // N even (1530 = 34 x 45): rows are 8-byte aligned
      for (int i = tid; i < kAsgDc * (kAsgTok / 2); i += 128) {
        const int r = i / (kAsgTok / 2), c2 = i % (kAsgTok / 2);
        const int d = d0 + r, p = p0 + c2 * 2;
        if (d < d_hi && p + 1 < N) cp_async8(&s_x[buf][r][c2 * 2], tok + (size_t)d * N + p);
        else {
          s_x[buf][r][c2 * 2] = (d < d_hi && p < N) ? tok[(size_t)d * N + p] : 0.f;
          s_x[buf][r][c2 * 2 + 1] = 0.f;
        }
      }
    } else {
      for (int i = tid; i < kAsgDc * kAsgTok; i += 128) {
        const int r = i / kAsgTok, c = i % kAsgTok;
        const int d = d0 + r, p = p0 + c;
        if (d < d_hi && p < N) cp_async4(&s_x[buf][r][c], tok + (size_t)d * N + p);
        else s_x[buf][r][c] = 0.f;
      }
    }
    {
      const int r = tid / 8, c4 = tid % 8;
      const int d = d0 + r;
      if (d < d_hi) cp_async16(&s_c[buf][r][c4 * 4], chatT + (size_t)d * Kp + kc + c4 * 4);
      else *reinterpret_cast<float4*>(&s_c[buf][r][c4 * 4]) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    cp_async_commit();
  };

  float ss[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) ss[i] = 0.f;
  const int n_ch = (d_hi - d_lo + kAsgDc - 1) / kAsgDc;
  float* pb = part + ((size_t)b * kAsgSplit + split) * N * Kp;
  for (int kc = 0; kc < K; kc += 32) {
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    if (n_ch > 0) stage(0, d_lo, kc);
    for (int c = 0; c < n_ch; ++c) {
      const int buf = c & 1;
      if (c + 1 < n_ch) { stage(buf ^ 1, d_lo + (c + 1) * kAsgDc, kc); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
      __syncthreads();
#pragma unroll
      for (int r = 0; r < kAsgDc; ++r) {
        const float4 x0 = *reinterpret_cast<const float4*>(&s_x[buf][r][tg * 8]);
        const float4 x1 = *reinterpret_cast<const float4*>(&s_x[buf][r][tg * 8 + 4]);
        const float4 cv = *reinterpret_cast<const float4*>(&s_c[buf][r][kg * 4]);
        const float xs[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          acc[i][0] = fmaf(xs[i], cv.x, acc[i][0]);
          acc[i][1] = fmaf(xs[i], cv.y, acc[i][1]);
          acc[i][2] = fmaf(xs[i], cv.z, acc[i][2]);
          acc[i][3] = fmaf(xs[i], cv.w, acc[i][3]);
          if (kc == 0 && kg == 0) ss[i] = fmaf(xs[i], xs[i], ss[i]);
        }
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int p = p0 + tg * 8 + i;
      if (p < N) *reinterpret_cast<float4*>(pb + (size_t)p * Kp + kc + kg * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
  }
  if (kg == 0)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int p = p0 + tg * 8 + i;
      if (p < N) ssq[((size_t)b * kAsgSplit + split) * N + p] = ss[i];
    }
}

// Stage 2: one thread per token: sum the split partials (ascending split), argmax over clusters (first index on
// ties), ||x||.
__global__ void assign_finalize_kernel(const float* __restrict__ part, const float* __restrict__ ssq, int N, int K, int Kp,
                                       int prenorm, int* __restrict__ labels, float* __restrict__ nrm_out) {
  const int b = blockIdx.y, p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= N) return;
  float bestv = -INFINITY;
  int besti = 0;
  for (int k = 0; k < K; ++k) {
    float v = 0.f;
#pragma unroll
    for (int s2 = 0; s2 < kAsgSplit; ++s2) v += part[(((size_t)b * kAsgSplit + s2) * N + p) * Kp + k];
    if (v > bestv) { bestv = v; besti = k; }
  }
  float t = 0.f;
#pragma unroll
  for (int s2 = 0; s2 < kAsgSplit; ++s2) t += ssq[((size_t)b * kAsgSplit + s2) * N + p];
  labels[(size_t)b * N + p] = besti;
  nrm_out[(size_t)b * N + p] = prenorm ? 1.0f : fmaxf(sqrtf(t), kEpsF);
}

// Stage 3: residual rows r = x / ||x|| - c[label], transposed to token-major.  grid (N/32, B, D/256); 8 warps, each
// a 32-channel slab of 32 tokens through a padded smem tile.
__global__ void __launch_bounds__(256)
residual_dn_kernel(const float* __restrict__ tokens, int N, int D, const float* __restrict__ centers,
                   const int* __restrict__ labels, const float* __restrict__ nrm_in, float* __restrict__ R) {
  __shared__ float s_t[8][32][33];
  __shared__ int s_lab[32];
  const int b = blockIdx.y, p0 = blockIdx.x * 32;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int p = p0 + lane;
  const bool valid = p < N;
  const float* tok = tokens + (size_t)b * D * N;
  if (w == 0) s_lab[lane] = valid ? labels[(size_t)b * N + p] : 0;
  const float nrm = valid ? nrm_in[(size_t)b * N + p] : 1.f;
  __syncthreads();
  const int dd = blockIdx.z * 256 + w * 32;
  if (dd >= D) return;
#pragma unroll 8
  for (int r = 0; r < 32; ++r) {
    const int d = dd + r;
    const float x = (d < D && valid) ? __ldg(tok + (size_t)d * N + p) : 0.f;
    s_t[w][r][lane] = x / nrm;
  }
  __syncwarp();
  const int d = dd + lane;
  if (d < D) {
#pragma unroll 4
    for (int t = 0; t < 32; ++t) {
      const int pp = p0 + t;
      if (pp < N) R[((size_t)b * N + pp) * D + d] = s_t[w][lane][t] - __ldg(centers + (size_t)s_lab[t] * D + d);
    }
  }
}

// Token-major [N, D] input: one warp per token, lanes stride the channels, warp-shuffle reductions.
__global__ void __launch_bounds__(256)
assign_nd_kernel(const float* __restrict__ tokens, int N, int D, const float* __restrict__ centers,
                 const float* __restrict__ chatT, int K, int Kp, int prenorm, float* __restrict__ R,
                 int* __restrict__ labels) {
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= N) return;
  const float* x = tokens + ((size_t)b * N + p) * D;
  float ss = 0.f, bestv = -INFINITY;
  int besti = 0;
  for (int kc = 0; kc < K; kc += 32) {
    float acc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = 0.f;
    for (int d = lane; d < D; d += 32) {
      float v = __ldg(x + d);
      if (kc == 0) ss = fmaf(v, v, ss);
      const float4* cr = reinterpret_cast<const float4*>(chatT + (size_t)d * Kp + kc);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 cv = __ldg(cr + j);
        acc[4 * j + 0] = fmaf(v, cv.x, acc[4 * j + 0]);
        acc[4 * j + 1] = fmaf(v, cv.y, acc[4 * j + 1]);
        acc[4 * j + 2] = fmaf(v, cv.z, acc[4 * j + 2]);
        acc[4 * j + 3] = fmaf(v, cv.w, acc[4 * j + 3]);
      }
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float v = warp_sum(acc[j]);
      if (kc + j < K && v > bestv) { bestv = v; besti = kc + j; }
    }
  }
  ss = warp_sum(ss);
  const float nrm = prenorm ? 1.0f : fmaxf(sqrtf(ss), kEpsF);
  if (lane == 0) labels[(size_t)b * N + p] = besti;
  float* r = R + ((size_t)b * N + p) * D;
  const float* c = centers + (size_t)besti * D;
  for (int d = lane; d < D; d += 32) r[d] = __ldg(x + d) / nrm - __ldg(c + d);
}

// ------------------------------------------------------------------------------------------------
// Counting sort of one image's tokens by label; tokens stay in ascending order inside a cluster.
__global__ void __launch_bounds__(1024)
cluster_lists_kernel(const int* __restrict__ labels, int N, int K, int* __restrict__ cl_ptr,
                     int* __restrict__ cl_tok) {
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int* lab = labels + (size_t)b * N;
  __shared__ int s_cnt[256 + 1];
  for (int k = threadIdx.x; k <= K; k += blockDim.x) s_cnt[k] = 0;
  __syncthreads();
  for (int p = threadIdx.x; p < N; p += blockDim.x) {   // histogram (integer: order-free); caller labels outside [0, K) are ignored
    const int l = lab[p];
    if ((unsigned)l < (unsigned)K) atomicAdd(&s_cnt[l], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int k = 0; k < K; ++k) { int c = s_cnt[k]; s_cnt[k] = run; run += c; }
    s_cnt[K] = run;
  }
  __syncthreads();
  for (int k = threadIdx.x; k <= K; k += blockDim.x) cl_ptr[(size_t)b * (K + 1) + k] = s_cnt[k];
  for (int k = w; k < K; k += nw) {
    int pos = s_cnt[k];
    for (int base = 0; base < N; base += 32) {
      int p = base + lane;
      bool m = p < N && lab[p] == k;
      unsigned bal = __ballot_sync(0xffffffffu, m);
      if (m) cl_tok[(size_t)b * N + pos + __popc(bal & ((1u << lane) - 1u))] = p;
      pos += __popc(bal);
    }
  }
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int find_image(const int* __restrict__ seg_off, int B, int s) {
  int lo = 0, hi = B;  // largest b with seg_off[b] <= s
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (seg_off[mid] <= s) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void superseg_union_kernel(const uint32_t* __restrict__ mem, const uint8_t* __restrict__ adj,
                                      const int* __restrict__ seg_off, const long long* __restrict__ adj_off,
                                      int B, int W, uint32_t* __restrict__ sup) {
  const int s = blockIdx.x;
  const int b = find_image(seg_off, B, s);
  const int s0 = seg_off[b], Si = seg_off[b + 1] - s0;
  const uint8_t* arow = adj + adj_off[b] + (size_t)(s - s0) * Si;
  for (int wd = threadIdx.x; wd < W; wd += blockDim.x) {
    uint32_t acc = 0;
    for (int t = 0; t < Si; ++t)
      if (arow[t]) acc |= mem[(size_t)(s0 + t) * W + wd];
    sup[(size_t)s * W + wd] = acc;
  }
}

// memS[g][i] = membership word (bit j = segment s0+j of group g) of the i-th token of the image IN LABEL-SORTED ORDER
// (cl_tok), so the aggregate kernel's producer reads token ids and membership words as two independent, coalesced
// streams.
__global__ void group_transpose_kernel(const uint32_t* __restrict__ sup, const int* __restrict__ grp_img,
                                       const int* __restrict__ grp_seg0, const int* __restrict__ grp_nseg,
                                       const int* __restrict__ cl_tok, int N, int W, uint16_t* __restrict__ memS) {
  const int g = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int p = cl_tok[(size_t)grp_img[g] * N + i];
  const int s0 = grp_seg0[g], ns = grp_nseg[g];
  unsigned m = 0;
  for (int j = 0; j < ns; ++j) m |= ((sup[(size_t)(s0 + j) * W + (p >> 5)] >> (p & 31)) & 1u) << j;
  memS[(size_t)g * N + i] = (uint16_t)m;
}

// cnt[g][k] = number of tokens of cluster k that belong to at least one segment of group g (= ring slots the
// aggregate kernel will see for (g, k)); one warp per (g, k)
__global__ void group_counts_kernel(const uint16_t* __restrict__ memS, const int* __restrict__ grp_img,
                                    const int* __restrict__ cl_ptr, int N, int K, int* __restrict__ cnt) {
  const int g = blockIdx.x, lane = threadIdx.x & 31;
  const int b = grp_img[g];
  for (int k = threadIdx.x >> 5; k < K; k += blockDim.x >> 5) {
    const int beg = cl_ptr[(size_t)b * (K + 1) + k], end = cl_ptr[(size_t)b * (K + 1) + k + 1];
    int c = 0;
    for (int i = beg + lane; i < end; i += 32) c += memS[(size_t)g * N + i] != 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) cnt[(size_t)g * K + k] = c;
  }
}

// predicted #non-empty (segment, cluster) blocks per segment: one warp per segment
__global__ void nonempty_kernel(const uint32_t* __restrict__ sup, const int* __restrict__ labels,
                                const int* __restrict__ seg_off, int B, int N, int W,
                                int* __restrict__ cpred) {
  const int s = blockIdx.x, lane = threadIdx.x;
  const int b = find_image(seg_off, B, s);
  const int* lab = labels + (size_t)b * N;
  uint32_t mk[4] = {0, 0, 0, 0};
  for (int wd = lane; wd < W; wd += 32) {
    uint32_t bits = sup[(size_t)s * W + wd];
    while (bits) {
      int j = __ffs(bits) - 1;
      bits &= bits - 1;
      int p = wd * 32 + j;
      if (p < N) { int k = lab[p]; mk[k >> 5] |= 1u << (k & 31); }
    }
  }
  int c = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t v = mk[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, o);
    c += __popc(v);
  }
  if (lane == 0) cpred[s] = c;
}

// ------------------------------------------------------------------------------------------------
// optional timing probe (tests/agg_probe.py): per-CTA cycle counters written by the producer / first consumer warp
static unsigned long long* g_agg_dbg = nullptr;

template <typename OutT> struct Stage4;
template <> struct Stage4<double> {
  static __device__ __forceinline__ void st(double* p, double a, double b, double c, double d) {
    *reinterpret_cast<double2*>(p) = make_double2(a, b);
    *reinterpret_cast<double2*>(p + 2) = make_double2(c, d);
  }
};
template <> struct Stage4<float> {
  static __device__ __forceinline__ void st(float* p, double a, double b, double c, double d) {
    *reinterpret_cast<float4*>(p) = make_float4((float)a, (float)b, (float)c, (float)d);
  }
};

// ---- mbarrier / bulk-copy PTX (sm_90+; 1-D TMA bulk copy, SASS UBLKCP) ----
__device__ __forceinline__ uint32_t agg_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void agg_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void agg_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void agg_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void agg_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  const long long t0 = clock64();
  for (uint32_t spin = 0;; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    if ((spin & 1023u) == 1023u && clock64() - t0 > 4000000000ll) __trap();  // a bug, not a wait: do not hang the box
  }
}
__device__ __forceinline__ bool agg_mbar_test(uint32_t bar, uint32_t parity) {   // non-blocking probe
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void agg_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void agg_bulk_store(void* gdst, uint32_t ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void agg_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void agg_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void agg_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Masked residual aggregation, v2 (r1 ncu on v0: 1 CTA/SM, DRAM 10 %, latency-bound on the dependent
// index -> membership -> row loads; fp64 adds predicated instead of branched).
//   grid (n_groups_total, ceil(K / k_per_cta)); block = 32 (producer warp) + D/4 consumer threads.
//   CTA = one group of <= kSegGroup segments of one image x a range of clusters.
//   producer warp: walks the clusters' token lists 32 tokens at a time (lane = token), reads the CTA-uniform
//     membership word, and for every token that is a member of some segment of the group claims the next ring
//     slot and issues ONE bulk copy (cp.async.bulk, 6 KB residual row) that completes on the slot's mbarrier;
//     a marker slot closes each cluster.  It runs up to kRingSlots rows ahead of the consumers.
//   consumer thread t owns channels 4t..4t+3 of kSegGroup fp64 accumulators: per slot one LDS.128 and, for each
//     member segment (warp-uniform branch), 4 DADDs.  At a cluster marker: block norms by warp shuffles +
//     named barrier, scale, streaming stores, reset.
template <typename OutT>
__global__ void __launch_bounds__(416, 1)
aggregate_kernel(const float* __restrict__ R, const int* __restrict__ cl_ptr, const int* __restrict__ cl_tok,
                 const uint16_t* __restrict__ memS, const int* __restrict__ grp_cnt,
                 const int* __restrict__ grp_img, const int* __restrict__ grp_seg0, const int* __restrict__ grp_nseg,
                 const int* __restrict__ cpred, int N, int D, int K, int k_per_cta, OutT* __restrict__ out,
                 double* __restrict__ norms, unsigned long long* __restrict__ dbg) {
  extern __shared__ __align__(128) unsigned char agg_smem[];
  const long long t_start = clock64();
  const int g = blockIdx.x;
  const int k0 = blockIdx.y * k_per_cta, k1 = min(K, k0 + k_per_cta);
  const int b = grp_img[g], s0 = grp_seg0[g], ns = grp_nseg[g];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_cwarps = (blockDim.x >> 5) - 1;
  const uint32_t row_bytes = (uint32_t)D * 4u;
  float* ring = reinterpret_cast<float*>(agg_smem);
  OutT* staging = reinterpret_cast<OutT*>(agg_smem + (size_t)kRingRows * row_bytes);      // [kSegGroup][D] scaled output
  float* maskf = reinterpret_cast<float*>(agg_smem + (size_t)kRingRows * row_bytes + (size_t)kSegGroup * D * sizeof(OutT));   // [row][kSegGroup] 0/1
  unsigned* meta = reinterpret_cast<unsigned*>(maskf + kRingRows * kSegGroup);           // [row] membership word
  uint64_t* bars = reinterpret_cast<uint64_t*>(meta + kRingRows);
  double* s_red = reinterpret_cast<double*>(bars + 2 * kRingSlots);      // [12 warps][kSegGroup]
  double* s_scale = s_red + 12 * kSegGroup;                              // [kSegGroup]
  const uint32_t bar_full = agg_smem_u32(bars), bar_empty = agg_smem_u32(bars + kRingSlots);
  if (tid == 0) {
    for (int i = 0; i < kRingSlots; ++i) { agg_mbar_init(bar_full + 8 * i, kRowsPerSlot); agg_mbar_init(bar_empty + 8 * i, n_cwarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  double* s_rowinv = s_scale + kSegGroup;   // 1 / max(sqrt(#non-empty clusters of the segment), eps), loaded once
  if (tid >= 32 && tid < 32 + kSegGroup) {
    const int j = tid - 32;
    s_rowinv[j] = (j < ns) ? 1.0 / fmax(sqrt((double)cpred[s0 + j]), kEpsD) : 0.0;
  }
  __syncthreads();

  if (warp == 0) {
    // ================= producer warp =================
    const int* toks = cl_tok + (size_t)b * N;
    const uint16_t* mrow = memS + (size_t)g * N;
    const float* Rb = R + (size_t)b * N * D;
    // The CTA's tokens are the contiguous range [cl_ptr[k0], cl_ptr[k1]) of the label-sorted list: the producer walks
    // it in windows of 32 (lane = token) regardless of cluster boundaries -- the consumers know how many rows each
    // cluster contributes (grp_cnt).  Token ids and membership words are independent coalesced loads, prefetched one
    // window ahead.  A window is issued in 4 sub-batches of 8 lanes so that slot re-use is waited for at quarter-ring
    // granularity.
    const int i_beg = cl_ptr[(size_t)b * (K + 1) + k0], i_end = cl_ptr[(size_t)b * (K + 1) + k1];
    unsigned seq = 0;   // ring sequence number of the next slot (warp-uniform)
    int p_cur = 0;
    unsigned m_cur = 0;
    long long t_poll = 0, t_issue = 0;
    if (i_beg + lane < i_end) { p_cur = toks[i_beg + lane]; m_cur = mrow[i_beg + lane]; }
    for (int i0 = i_beg; i0 < i_end; i0 += 32) {
      int p_nxt = 0;
      unsigned m_nxt = 0;
      if (i0 + 32 + lane < i_end) { p_nxt = toks[i0 + 32 + lane]; m_nxt = mrow[i0 + 32 + lane]; }   // prefetch
      const unsigned act = __ballot_sync(0xffffffffu, m_cur != 0u);
      const unsigned n = seq + __popc(act & ((1u << lane) - 1u));
      const unsigned rrow = n % kRingRows;                               // ring row of this lane's token
      const unsigned slot = rrow / kRowsPerSlot, gen = n / kRingRows;    // barrier pair / ring generation
#pragma unroll 1
      for (int sub = 0; sub < 4; ++sub) {
        const bool mine = (m_cur != 0u) && ((lane >> 3) == sub);
        if (!__any_sync(0xffffffffu, mine)) continue;
        // Slots are released by the consumers in ring order, so once the slot of the LAST row of this sub-batch is
        // free every earlier one is too: that one lane blocks in mbarrier.try_wait (the warp sleeps in hardware --
        // a spinning producer steals issue slots from the consumer warps that share its SM sub-partition, r1 probe).
        const long long t0 = dbg ? clock64() : 0;
        const unsigned mm = __ballot_sync(0xffffffffu, mine);
        if (lane == 31 - __clz(mm)) agg_mbar_wait(bar_empty + 8 * slot, (gen & 1u) ^ 1u);
        __syncwarp();
        const long long t1 = dbg ? clock64() : 0;
        if (dbg) t_poll += t1 - t0;
        if (mine) {
          float4 f0, f1;
          f0.x = (m_cur & 1u) ? 1.f : 0.f;   f0.y = (m_cur & 2u) ? 1.f : 0.f;
          f0.z = (m_cur & 4u) ? 1.f : 0.f;   f0.w = (m_cur & 8u) ? 1.f : 0.f;
          f1.x = (m_cur & 16u) ? 1.f : 0.f;  f1.y = (m_cur & 32u) ? 1.f : 0.f;
          f1.z = (m_cur & 64u) ? 1.f : 0.f;  f1.w = (m_cur & 128u) ? 1.f : 0.f;
          *reinterpret_cast<float4*>(maskf + rrow * kSegGroup) = f0;
          *reinterpret_cast<float4*>(maskf + rrow * kSegGroup + 4) = f1;
          agg_mbar_expect_tx(bar_full + 8 * slot, row_bytes);
          agg_bulk_load(agg_smem_u32(ring) + rrow * row_bytes, Rb + (size_t)p_cur * D, row_bytes, bar_full + 8 * slot);
        }
        __syncwarp();
        if (dbg) t_issue += clock64() - t1;
      }
      seq += __popc(act);
      p_cur = p_nxt;
      m_cur = m_nxt;
    }
    // complete the last, partially filled slot so that its consumers are released
    if (lane == 0 && (seq % kRowsPerSlot) != 0) {
      const unsigned slot = (seq % kRingRows) / kRowsPerSlot;
      for (unsigned i = seq % kRowsPerSlot; i < kRowsPerSlot; ++i) agg_mbar_arrive(bar_full + 8 * slot);
    }
    if (dbg && lane == 0) {
      unsigned long long* o = dbg + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 16;
      o[0] = (unsigned long long)(clock64() - t_start); o[1] = (unsigned long long)t_poll; o[2] = (unsigned long long)t_issue; o[3] = seq;
    }
    return;
  }

  // ================= consumer warps =================
  const int t = tid - 32;                 // channel quad
  const int cw = warp - 1;
  const int d = 4 * t;
  const bool act_ch = d < D;              // last warp may be partially idle when D % 128 != 0
  // Accumulation: branch-free packed fp32 FMAs (fma.rn.f32x2) with the 0/1 membership as multiplier -- acc + 1*r is
  // the plain fp32 add, acc + 0*r leaves it untouched -- into fp32 partial sums that are promoted to the fp64
  // accumulators every kFlushEvery rows (and at the end of the cluster).  A partial holds <= 16 addends, so its
  // rounding error is <= 15 * 2^-24 of the partial's magnitude (~1e-6 worst case, ~2e-7 typical): inside the 1e-5
  // descriptor tolerance with a wide margin, while the half-rate fp64 pipe is touched 16x less often.
  // (r1 ncu: the per-(row, segment) branched fp64 adds executed ~128 instructions per row and warp, IPC 1.2.)
  static_assert(kSegGroup == 8, "two float4 mask loads per row");
  double acc[kSegGroup][4];
  float2 a32[kSegGroup][2];
#pragma unroll
  for (int j = 0; j < kSegGroup; ++j) {
    acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0;
    a32[j][0] = a32[j][1] = make_float2(0.f, 0.f);
  }
  int since = 0;
  unsigned seq = 0;
  long long t_wait = 0, t_epi = 0, t_e1 = 0, t_e2 = 0, t_e3 = 0, t_loop = 0;
  const bool probe = dbg != nullptr;
  // r1 profile of the row loop: ~65 instructions per row and warp (address arithmetic, generic->shared conversions,
  // predicate juggling) at SM-level IPC ~2 => issue-bound.  The loop below is kept minimal: explicit shared-space
  // loads from precomputed 32-bit addresses, 16 packed FMAs, slot hand-shake every kRowsPerSlot rows.
#define SV_FMA(J, B, RLO, RHI)                                     \
      a32[J][0] = __ffma2_rn(make_float2(B, B), RLO, a32[J][0]);   \
      a32[J][1] = __ffma2_rn(make_float2(B, B), RHI, a32[J][1]);
#define SV_ROW(R, B0, B1)                                                                   \
      {                                                                                     \
        const float2 rlo = make_float2(R.x, R.y), rhi = make_float2(R.z, R.w);              \
        SV_FMA(0, B0.x, rlo, rhi) SV_FMA(1, B0.y, rlo, rhi) SV_FMA(2, B0.z, rlo, rhi) SV_FMA(3, B0.w, rlo, rhi) \
        SV_FMA(4, B1.x, rlo, rhi) SV_FMA(5, B1.y, rlo, rhi) SV_FMA(6, B1.z, rlo, rhi) SV_FMA(7, B1.w, rlo, rhi) \
      }
#define SV_FLUSH()                                                                          \
      {                                                                                     \
        since = 0;                                                                          \
        _Pragma("unroll") for (int j = 0; j < kSegGroup; ++j) {                             \
          acc[j][0] += (double)a32[j][0].x; acc[j][1] += (double)a32[j][0].y;               \
          acc[j][2] += (double)a32[j][1].x; acc[j][3] += (double)a32[j][1].y;               \
          a32[j][0] = a32[j][1] = make_float2(0.f, 0.f);                                    \
        }                                                                                   \
      }
  const uint32_t ring_s = agg_smem_u32(ring) + (act_ch ? d : 0) * 4;   // idle channel quads read (and ignore) quad 0
  const uint32_t mask_s = agg_smem_u32(maskf);
  for (int k = k0; k < k1; ++k) {
    const int n_rows = grp_cnt[(size_t)g * K + k];
    long long tl0 = 0;
    if (probe) tl0 = clock64();
    // rows are consumed in ring order; the full barrier is waited for when a row opens a new slot (every
    // kRowsPerSlot rows) and the empty barrier is signalled when a row closes one
#pragma unroll 2
    for (int rix = 0; rix < n_rows; ++rix) {
      const unsigned rrow = seq & (kRingRows - 1);
      if ((seq & (kRowsPerSlot - 1)) == 0) {
        long long tw0 = 0;
        if (probe) tw0 = clock64();
        agg_mbar_wait(bar_full + 8 * (rrow / kRowsPerSlot), (seq / kRingRows) & 1u);
        if (probe) t_wait += clock64() - tw0;
      }
      const float4 r = lds128(ring_s + rrow * row_bytes);
      const float4 b0 = lds128(mask_s + rrow * (kSegGroup * 4));
      const float4 b1 = lds128(mask_s + rrow * (kSegGroup * 4) + 16);
      SV_ROW(r, b0, b1)
      if ((seq & (kRowsPerSlot - 1)) == kRowsPerSlot - 1) {
        __syncwarp();
        if (lane == 0) agg_mbar_arrive(bar_empty + 8 * (rrow / kRowsPerSlot));
      }
      ++seq;
      if (++since >= kFlushEvery) SV_FLUSH()
    }
    // end-of-cluster: promote the remaining partial sums
    long long te0 = 0;
    if (probe) { te0 = clock64(); t_loop += te0 - tl0; }
    SV_FLUSH()
    // ---- cluster k complete: intra-norm, row scale, store ----
#pragma unroll
    for (int j = 0; j < kSegGroup; ++j) {
      double ss = acc[j][0] * acc[j][0] + acc[j][1] * acc[j][1] + acc[j][2] * acc[j][2] + acc[j][3] * acc[j][3];
      if (!act_ch) ss = 0.0;   // idle channel quads (D % 128 != 0) shadow quad 0: keep them out of the norm
      ss = warp_sum(ss);
      if (lane == 0) s_red[cw * kSegGroup + j] = ss;
    }
    long long tp1 = 0;
    if (probe) { tp1 = clock64(); t_e1 += tp1 - te0; }
    if (tid == 32) agg_bulk_wait_read();   // the previous cluster's staged block has left shared memory
    asm volatile("bar.sync 1, %0;" ::"r"(n_cwarps * 32) : "memory");
    if (cw == 0) {
      // 32 lanes: lane = (quarter q, segment j); each sums the partials of warps q, q+4, q+8, ... in ascending order,
      // then the four quarters are combined by two xor-shuffles (fixed order => deterministic)
      const int j = lane & 7, q = lane >> 3;
      double tot = 0.0;
      for (int ww = q; ww < n_cwarps; ww += 4) tot += s_red[ww * kSegGroup + j];
      tot += __shfl_xor_sync(0xffffffffu, tot, 8);
      tot += __shfl_xor_sync(0xffffffffu, tot, 16);
      if (q == 0) {
        double sc = 0.0;
        if (j < ns) {
          const double nrm = sqrt(tot);
          norms[(size_t)(s0 + j) * K + k] = nrm;
          sc = (1.0 / fmax(nrm, kEpsD)) * s_rowinv[j];
        }
        s_scale[j] = sc;
      }
    }
    long long tp2 = 0;
    if (probe) { tp2 = clock64(); t_e2 += tp2 - tp1; }
    asm volatile("bar.sync 1, %0;" ::"r"(n_cwarps * 32) : "memory");
    if (probe) t_e3 += clock64() - tp2;
    // scaled block -> shared-memory staging -> asynchronous bulk stores (TMA): the 8 x D x 8 B of output leave the SM
    // while the consumers already accumulate the next cluster (r1 probe: direct 16-byte stores cost ~4.6 k cycles per
    // epilogue at ~21 B/clk/SM and nothing else ran meanwhile)
#pragma unroll
    for (int j = 0; j < kSegGroup; ++j) {
      if (j < ns && act_ch) {
        const double sc = s_scale[j];
        Stage4<OutT>::st(staging + (size_t)j * D + d, acc[j][0] * sc, acc[j][1] * sc, acc[j][2] * sc, acc[j][3] * sc);
      }
      acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0;
    }
    agg_fence_async();
    asm volatile("bar.sync 1, %0;" ::"r"(n_cwarps * 32) : "memory");
    if (tid == 32) {
      const uint32_t blk_bytes = (uint32_t)D * (uint32_t)sizeof(OutT);
      for (int j = 0; j < ns; ++j)
        agg_bulk_store(out + (size_t)(s0 + j) * K * D + (size_t)k * D, agg_smem_u32(staging + (size_t)j * D), blk_bytes);
      agg_bulk_commit();
    }
    if (probe) t_epi += clock64() - te0;
  }
#undef SV_FLUSH
#undef SV_ROW
#undef SV_FMA
  if (tid == 32) agg_bulk_wait_read();   // shared memory must outlive the last asynchronous stores
  if (dbg && tid == 32) {
    unsigned long long* o = dbg + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 16;
    o[4] = (unsigned long long)(clock64() - t_start); o[5] = (unsigned long long)t_wait; o[6] = (unsigned long long)t_epi; o[7] = seq;
    o[8] = (unsigned long long)t_e1; o[9] = (unsigned long long)t_e2; o[10] = (unsigned long long)t_e3; o[11] = (unsigned long long)t_loop;
  }
}

// Row norm as the reference computes it: sqrt(sum_k (||V_k|| / max(||V_k||,eps))^2).  The aggregate
// kernel used sqrt(#non-empty blocks); they differ only if some non-empty block has ||V_k|| < eps
// (exact cancellation / zero residuals).  Those rows are rescaled here (normally: none).
template <typename OutT>
__global__ void rownorm_fixup_kernel(const double* __restrict__ norms, const int* __restrict__ cpred, int K,
                                     size_t row_len, OutT* __restrict__ out) {
  const int s = blockIdx.x;
  __shared__ double s_factor;
  if (threadIdx.x < 32) {
    // q is exactly 0 or 1 for every block except the degenerate ones, so the sum is exact in any order
    double tsum = 0.0;
    for (int k = threadIdx.x; k < K; k += 32) {
      double n = norms[(size_t)s * K + k];
      double q = n / fmax(n, kEpsD);
      tsum += q * q;
    }
    tsum = warp_sum(tsum);
    if (threadIdx.x == 0) {
      double m_true = sqrt(tsum), m_pred = sqrt((double)cpred[s]);
      s_factor = (m_true == m_pred) ? 1.0 : fmax(m_pred, kEpsD) / fmax(m_true, kEpsD);
    }
  }
  __syncthreads();
  const double f = s_factor;
  if (f == 1.0) return;
  OutT* o = out + (size_t)s * row_len;
  for (size_t i = threadIdx.x; i < row_len; i += blockDim.x) o[i] = (OutT)((double)o[i] * f);
}

// The same correction on the PCA-planes output (SEGVLAD_OUT_PCA_PLANES): x = v - mean was split into three bf16 planes whose
// sum is the fp32 x exactly, so v = x + mean can be rebuilt, rescaled and split again (normally no row needs it).
__global__ void rownorm_fixup_planes_kernel(const double* __restrict__ norms, const int* __restrict__ cpred, int K, size_t row_len,
                                            size_t plane_stride, const float* __restrict__ mean, __nv_bfloat16* __restrict__ planes) {
  const int s = blockIdx.x;
  __shared__ double s_factor;
  if (threadIdx.x < 32) {
    double tsum = 0.0;
    for (int k = threadIdx.x; k < K; k += 32) {
      double n = norms[(size_t)s * K + k];
      double q = n / fmax(n, kEpsD);
      tsum += q * q;
    }
    tsum = warp_sum(tsum);
    if (threadIdx.x == 0) {
      double m_true = sqrt(tsum), m_pred = sqrt((double)cpred[s]);
      s_factor = (m_true == m_pred) ? 1.0 : fmax(m_pred, kEpsD) / fmax(m_true, kEpsD);
    }
  }
  __syncthreads();
  const double f = s_factor;
  if (f == 1.0) return;
  __nv_bfloat16* p = planes + (size_t)s * row_len;
  for (size_t i = threadIdx.x; i < row_len; i += blockDim.x) {
    const float x = (__bfloat162float(p[i]) + __bfloat162float(p[i + plane_stride])) + __bfloat162float(p[i + 2 * plane_stride]);
    const float y = (float)(((double)x + (double)mean[i]) * f - (double)mean[i]);
    const __nv_bfloat16 h = __float2bfloat16_rn(y);
    const float r = y - __bfloat162float(h);
    const __nv_bfloat16 m = __float2bfloat16_rn(r);
    p[i] = __float2bfloat16_rn(r - __bfloat162float(m));
    p[i + plane_stride] = m;
    p[i + 2 * plane_stride] = h;
  }
}

// ------------------------------------------------------------------------------------------------
// Pixel masks -> patch membership bits (func_vpr.py:1088-1092).  One thread per (segment, patch).
__global__ void mask_to_membership_kernel(const uint8_t* __restrict__ masks, int S, int Hm, int Wm, int H,
                                          int W, int patch, int dh, int dw, uint32_t* __restrict__ bits,
                                          int Wd) {
  const int s = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int N = dh * dw;
  bool any = false;
  if (p < N) {
    const int pi = p / dw, pj = p % dw;
    const int i0 = pi * patch, i1 = (pi == dh - 1) ? H : i0 + patch;
    const int j0 = pj * patch, j1 = (pj == dw - 1) ? W : j0 + patch;
    const float sh = (float)Hm / (float)H, sw = (float)Wm / (float)W;  // torch 'nearest' source index
    const uint8_t* m = masks + (size_t)s * Hm * Wm;
    for (int i = i0; i < i1 && !any; ++i) {
      int si = min((int)floorf((float)i * sh), Hm - 1);
      for (int j = j0; j < j1; ++j) {
        int sj = min((int)floorf((float)j * sw), Wm - 1);
        if (m[(size_t)si * Wm + sj]) { any = true; break; }
      }
    }
  }
  unsigned bal = __ballot_sync(0xffffffffu, any);
  if ((threadIdx.x & 31) == 0 && (p >> 5) < Wd) bits[(size_t)s * Wd + (p >> 5)] = bal;
}

// ------------------------------------------------------------------------------------------------
// Mask centroids: the input of the Delaunay neighbourhood graph (func_vpr.py:1314, `np.array(np.nonzero(m)).mean(1)[::-1]`
// = (mean column, mean row) in fp64).  The masks are on the GPU anyway (membership kernel above); the host loop over
// np.nonzero of every 240 x 320 mask was 19 ms per image in the config-1 run, 5 x the whole aggregation.  Integer sums are
// exact, so sum / count in fp64 is the same double numpy's mean returns (its pairwise fp64 sum of integers < 2^53 is exact
// too); an empty mask gives 0 / 0 = NaN like numpy.  One CTA per mask.
__global__ void __launch_bounds__(256)
mask_centroid_kernel(const uint8_t* __restrict__ masks, int Hm, int Wm, double* __restrict__ cxy) {
  const uint8_t* m = masks + (size_t)blockIdx.x * Hm * Wm;
  unsigned long long sx = 0, sy = 0;
  unsigned cnt = 0;
  for (int i = threadIdx.x; i < Hm * Wm; i += 256) {
    if (m[i]) { sy += (unsigned)(i / Wm); sx += (unsigned)(i % Wm); ++cnt; }
  }
  __shared__ unsigned long long s_x[8], s_y[8];
  __shared__ unsigned s_c[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sx += __shfl_xor_sync(0xffffffffu, sx, o);
    sy += __shfl_xor_sync(0xffffffffu, sy, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if ((threadIdx.x & 31) == 0) { s_x[threadIdx.x >> 5] = sx; s_y[threadIdx.x >> 5] = sy; s_c[threadIdx.x >> 5] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long tx = 0, ty = 0;
    unsigned tc = 0;
    for (int w = 0; w < 8; ++w) { tx += s_x[w]; ty += s_y[w]; tc += s_c[w]; }
    cxy[2 * blockIdx.x] = (double)tx / (double)tc;
    cxy[2 * blockIdx.x + 1] = (double)ty / (double)tc;
  }
}

// ------------------------------------------------------------------------------------------------
struct AggLayout {
  float* chatT; float* R; float* part; float* ssq; float* nrm; int* labels; int* cl_ptr; int* cl_tok; uint32_t* sup; uint16_t* memT; int* gcnt;
  int* cpred; double* norms; int* seg_off; long long* adj_off; int* grp_img; int* grp_seg0; int* grp_nseg;
  __nv_bfloat16* RT; int* tile_tbl;   // tensor-core path (aggregate_tc.cu)
  double* xnorm;
  __nv_bfloat16* chat_planes;         // tensor-core assignment (assign_tc.cu)
  size_t total;
};

static AggLayout carve_agg(void* ws, int B, int N, int D, int K, int S_total) {
  Carver c(ws);
  AggLayout L;
  const int Kp = (int)align_up(K, 32), W = (N + 31) / 32;
  const int max_groups = S_total / kSegGroup + B;  // sum_b ceil(S_b / G) <= S_total/G + B
  L.chatT = c.take<float>((size_t)D * Kp);
  L.R = c.take<float>((size_t)B * N * D);
  L.part = c.take<float>((size_t)B * kAsgSplit * N * Kp);
  L.ssq = c.take<float>((size_t)B * kAsgSplit * N);
  L.nrm = c.take<float>((size_t)B * N);
  L.labels = c.take<int>((size_t)B * N);
  L.cl_ptr = c.take<int>((size_t)B * (K + 1));
  L.cl_tok = c.take<int>((size_t)B * N);
  L.sup = c.take<uint32_t>((size_t)S_total * W);
  L.memT = c.take<uint16_t>((size_t)max_groups * N);
  L.gcnt = c.take<int>((size_t)max_groups * K);
  L.cpred = c.take<int>(S_total);
  L.norms = c.take<double>((size_t)S_total * K);
  L.seg_off = c.take<int>(B + 1);
  L.adj_off = c.take<long long>(B + 1);
  L.grp_img = c.take<int>(max_groups);
  L.grp_seg0 = c.take<int>(max_groups);
  L.grp_nseg = c.take<int>(max_groups);
  L.RT = c.take<__nv_bfloat16>(agg_tc_supported(N, D, K) ? agg_tc_rt_elems(B, N, D) : 0);
  L.tile_tbl = c.take<int>((size_t)4 * agg_tc_max_tiles(B, S_total));
  L.xnorm = c.take<double>(agg_tc_supported(N, D, K) ? agg_tc_xnorm_elems(B, S_total, D, K) : 0);
  L.chat_planes = c.take<__nv_bfloat16>(assign_tc_workspace_elems(D, K));
  L.total = c.total();
  return L;
}

}  // namespace segvlad

using namespace segvlad;

extern "C" size_t segvlad_aggregate_workspace_bytes(int n_images, int N, int D_t, int K, int S_total) {
  if (n_images <= 0 || N <= 0 || D_t <= 0 || K <= 0 || S_total < 0) return 0;
  return carve_agg(nullptr, n_images, N, D_t, K, S_total).total;
}

// Common driver.  residuals_in != nullptr: skip normalise/assign and aggregate the caller's residual
// rows [B*N, D] fp32 with the caller's labels (vlad_matmuls_per_cluster drop-in, func_vpr.py:1181).
static int aggregate_driver(const float* tokens, const float* residuals_in, const int32_t* labels_in, int B, int N,
                            int D, int token_layout, const float* centers, int K, const uint32_t* member_bits,
                            const int32_t* seg_offsets_host, const uint8_t* adj, void* out, int out_dtype,
                            int32_t* labels_out, void* workspace, size_t workspace_bytes, cudaStream_t st,
                            const float* pca_mean = nullptr) {
  SV_REQUIRE(B > 0 && N > 0 && D > 0 && K > 0, "aggregate: non-positive shape");
  const bool use_tc = agg_tc_supported(N, D, K);
  SV_REQUIRE(out_dtype != SEGVLAD_OUT_PCA_PLANES || (use_tc && pca_mean && D % 64 == 0),
             "aggregate: the PCA-planes output needs the tensor-core path (D_t %% 64 == 0) and the model mean");
  SV_REQUIRE(use_tc || (D % 4 == 0 && D <= 1536), "aggregate: D_t must be a multiple of 4 and <= 1536 (got %d)", D);
  SV_REQUIRE(K <= 128, "aggregate: K must be <= 128 (got %d)", K);
  const int layout = token_layout & 1, prenorm = (token_layout & SEGVLAD_TOKENS_PRENORMALIZED) ? 1 : 0;
  SV_REQUIRE((token_layout & ~3) == 0, "aggregate: bad token_layout");
  SV_REQUIRE(out_dtype == SEGVLAD_OUT_F64 || out_dtype == SEGVLAD_OUT_F32 || out_dtype == SEGVLAD_OUT_PCA_PLANES,
             "aggregate: bad out_dtype");
  SV_REQUIRE(seg_offsets_host && seg_offsets_host[0] == 0, "aggregate: seg_offsets_host[0] must be 0");
  const int S_total = seg_offsets_host[B];
  for (int b = 0; b < B; ++b)
    SV_REQUIRE(seg_offsets_host[b + 1] >= seg_offsets_host[b], "aggregate: seg_offsets not monotone");
  if (S_total == 0) return SEGVLAD_OK;
  AggLayout L = carve_agg(workspace, B, N, D, K, S_total);
  if (workspace_bytes < L.total || !workspace) {
    set_error("aggregate: workspace %zu < required %zu", workspace_bytes, L.total);
    return SEGVLAD_EWORKSPACE;
  }
  const int Kp = (int)align_up(K, 32), W = (N + 31) / 32;

  // host-side metadata: offsets and segment-group tables (one small H2D copy each)
  const int max_groups = S_total / kSegGroup + B;
  int* h = (int*)malloc(sizeof(int) * (size_t)(3 * max_groups + 2) + sizeof(long long) * (size_t)(B + 1));
  SV_REQUIRE(h, "aggregate: host malloc failed");
  int* h_img = h; int* h_s0 = h + max_groups; int* h_ns = h + 2 * max_groups;
  long long* h_adj = reinterpret_cast<long long*>(h + 3 * max_groups + ((3 * max_groups) & 1));
  int ng = 0;
  long long ao = 0;
  for (int b = 0; b < B; ++b) {
    int s0 = seg_offsets_host[b], Si = seg_offsets_host[b + 1] - s0;
    h_adj[b] = ao;
    ao += (long long)Si * Si;
    for (int g0 = 0; g0 < Si; g0 += kSegGroup) {
      h_img[ng] = b; h_s0[ng] = s0 + g0; h_ns[ng] = min(kSegGroup, Si - g0); ++ng;
    }
  }
  h_adj[B] = ao;
  cudaError_t e = cudaMemcpyAsync(L.seg_off, seg_offsets_host, sizeof(int) * (B + 1), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(L.adj_off, h_adj, sizeof(long long) * (B + 1), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(L.grp_img, h_img, sizeof(int) * ng, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(L.grp_seg0, h_s0, sizeof(int) * ng, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(L.grp_nseg, h_ns, sizeof(int) * ng, cudaMemcpyHostToDevice, st);
  free(h);  // pageable-source cudaMemcpyAsync has staged the bytes before returning
  SV_CHECK_CUDA(e);

  const float* R = L.R;
  const int* labels = L.labels;
  bool fused_rt = false;
  if (residuals_in) {
    R = residuals_in;
    labels = labels_in;
  } else {
    const bool asg_tc = layout == SEGVLAD_TOKENS_DN && assign_tc_supported(N, D, K);
    if (!asg_tc) {
      // zero padding columns of chatT (k >= K) so the 32-wide loads see zeros
      SV_CHECK_CUDA(cudaMemsetAsync(L.chatT, 0, sizeof(float) * (size_t)D * Kp, st));
      normalize_centers_kernel<<<K, 256, 0, st>>>(centers, K, D, L.chatT, Kp);
      SV_CHECK_LAUNCH();
    }
    if (layout == SEGVLAD_TOKENS_DN) {
      if (asg_tc) {
        const int rc = assign_tc_run(tokens, B, N, D, centers, K, prenorm, L.chat_planes, L.labels, L.nrm, st);
        if (rc != SEGVLAD_OK) return rc;
      } else {
        assign_partial_kernel<<<dim3((N + kAsgTok - 1) / kAsgTok, B, kAsgSplit), 128, 0, st>>>(tokens, N, D, L.chatT, K, Kp,
                                                                                            L.part, L.ssq);
        SV_CHECK_LAUNCH();
        assign_finalize_kernel<<<dim3((N + 127) / 128, B), 128, 0, st>>>(L.part, L.ssq, N, K, Kp, prenorm, L.labels, L.nrm);
        SV_CHECK_LAUNCH();
      }
      // tensor-core path: the residual planes are built straight from the tokens after the cluster lists (no R)
      fused_rt = use_tc && agg_tc_fused_channels(N, K) > 0;
      if (!fused_rt)
        residual_dn_kernel<<<dim3((N + 31) / 32, B, (D + 255) / 256), 256, 0, st>>>(tokens, N, D, centers, L.labels, L.nrm,
                                                                                 L.R);
    } else {
      dim3 grid((N + 7) / 8, B);
      assign_nd_kernel<<<grid, 256, 0, st>>>(tokens, N, D, centers, L.chatT, K, Kp, prenorm, L.R, L.labels);
    }
    SV_CHECK_LAUNCH();
  }
  cluster_lists_kernel<<<B, 1024, 0, st>>>(labels, N, K, L.cl_ptr, L.cl_tok);
  SV_CHECK_LAUNCH();
  const uint32_t* sup = member_bits;
  if (adj) {
    superseg_union_kernel<<<S_total, 64, 0, st>>>(member_bits, adj, L.seg_off, L.adj_off, B, W, L.sup);
    SV_CHECK_LAUNCH();
    sup = L.sup;
  }
  group_transpose_kernel<<<dim3((N + 255) / 256, ng), 256, 0, st>>>(sup, L.grp_img, L.grp_seg0, L.grp_nseg, L.cl_tok, N, W,
                                                                    L.memT);
  SV_CHECK_LAUNCH();
  group_counts_kernel<<<ng, 256, 0, st>>>(L.memT, L.grp_img, L.cl_ptr, N, K, L.gcnt);
  SV_CHECK_LAUNCH();
  nonempty_kernel<<<S_total, 32, 0, st>>>(sup, labels, L.seg_off, B, N, W, L.cpred);
  SV_CHECK_LAUNCH();
  if (use_tc) {
    // tensor-core path: mask tile x bf16-split residual planes per (image, 128-segment tile, cluster)
    AggTcArgs ta;
    ta.R = fused_rt ? nullptr : R; ta.tokens_dn = tokens; ta.centers = centers; ta.labels = L.labels; ta.nrm = L.nrm;
    ta.cl_ptr = L.cl_ptr; ta.cl_tok = L.cl_tok; ta.memS = L.memT; ta.cpred = L.cpred; ta.norms = L.norms;
    ta.seg_offsets_host = seg_offsets_host; ta.B = B; ta.N = N; ta.D = D; ta.K = K; ta.S_total = S_total;
    ta.out = out; ta.out_dtype = out_dtype; ta.RT = L.RT; ta.tile_tbl = L.tile_tbl; ta.xnorm = L.xnorm; ta.probe = g_agg_dbg;
    ta.pca_mean = pca_mean;
    const int rc = agg_tc_run(ta, st);
    if (rc != SEGVLAD_OK) return rc;
    if (out_dtype == SEGVLAD_OUT_PCA_PLANES)
      rownorm_fixup_planes_kernel<<<S_total, 256, 0, st>>>(L.norms, L.cpred, K, (size_t)K * D, (size_t)S_total * K * D, pca_mean,
                                                           (__nv_bfloat16*)out);
    else if (out_dtype == SEGVLAD_OUT_F64)
      rownorm_fixup_kernel<double><<<S_total, 256, 0, st>>>(L.norms, L.cpred, K, (size_t)K * D, (double*)out);
    else
      rownorm_fixup_kernel<float><<<S_total, 256, 0, st>>>(L.norms, L.cpred, K, (size_t)K * D, (float*)out);
    SV_CHECK_LAUNCH();
    if (labels_out && !residuals_in)
      SV_CHECK_CUDA(cudaMemcpyAsync(labels_out, L.labels, sizeof(int) * (size_t)B * N, cudaMemcpyDeviceToDevice, st));
    return SEGVLAD_OK;
  }
  // one warp of producers + D/4 consumer threads; clusters are split over blockIdx.y until the grid has
  // >= ~2 CTAs per SM (single-image calls) -- batched calls keep every cluster of a group in one CTA
  const int threads = 32 + (int)align_up((size_t)(D / 4), 32);
  int k_per_cta = K;
  while (k_per_cta > 1 && (long long)ng * ((K + k_per_cta - 1) / k_per_cta) < 4 * 148) k_per_cta = (k_per_cta + 1) / 2;
  const dim3 agrid(ng, (K + k_per_cta - 1) / k_per_cta);
  const size_t asmem = (size_t)kRingRows * D * 4 + (size_t)kSegGroup * D * (out_dtype == SEGVLAD_OUT_F64 ? 8 : 4) +
                       kRingRows * kSegGroup * 4 + kRingRows * 4 + 2 * kRingSlots * 8 +
                       14 * kSegGroup * 8 + 64;
  const int pslot = prof_begin(SEGVLAD_PROF_AGGREGATE, st);
  if (out_dtype == SEGVLAD_OUT_F64) {
    SV_CHECK_CUDA(cudaFuncSetAttribute(aggregate_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)asmem));
    aggregate_kernel<double><<<agrid, threads, asmem, st>>>(R, L.cl_ptr, L.cl_tok, L.memT, L.gcnt, L.grp_img, L.grp_seg0,
                                                           L.grp_nseg, L.cpred, N, D, K, k_per_cta, (double*)out, L.norms, g_agg_dbg);
    prof_end(pslot, st);
    SV_CHECK_LAUNCH();
    rownorm_fixup_kernel<double><<<S_total, 256, 0, st>>>(L.norms, L.cpred, K, (size_t)K * D, (double*)out);
  } else {
    SV_CHECK_CUDA(cudaFuncSetAttribute(aggregate_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)asmem));
    aggregate_kernel<float><<<agrid, threads, asmem, st>>>(R, L.cl_ptr, L.cl_tok, L.memT, L.gcnt, L.grp_img, L.grp_seg0,
                                                          L.grp_nseg, L.cpred, N, D, K, k_per_cta, (float*)out, L.norms, g_agg_dbg);
    prof_end(pslot, st);
    SV_CHECK_LAUNCH();
    rownorm_fixup_kernel<float><<<S_total, 256, 0, st>>>(L.norms, L.cpred, K, (size_t)K * D, (float*)out);
  }
  SV_CHECK_LAUNCH();
  if (labels_out && !residuals_in)
    SV_CHECK_CUDA(cudaMemcpyAsync(labels_out, L.labels, sizeof(int) * (size_t)B * N, cudaMemcpyDeviceToDevice, st));
  return SEGVLAD_OK;
}

extern "C" int segvlad_aggregate_batch(const float* tokens, int B, int N, int D, int token_layout,
                                       const float* centers, int K, const uint32_t* member_bits,
                                       const int32_t* seg_offsets_host, const uint8_t* adj, void* out,
                                       int out_dtype, int32_t* labels_out, void* workspace,
                                       size_t workspace_bytes, void* stream_) {
  SV_REQUIRE(tokens && centers && member_bits && out, "aggregate: null pointer");
  return aggregate_driver(tokens, nullptr, nullptr, B, N, D, token_layout, centers, K, member_bits, seg_offsets_host,
                          adj, out, out_dtype, labels_out, workspace, workspace_bytes,
                          reinterpret_cast<cudaStream_t>(stream_));
}

extern "C" int segvlad_aggregate_residuals(const float* residuals, const int32_t* labels, int B, int N, int D, int K,
                                           const uint32_t* member_bits, const int32_t* seg_offsets_host,
                                           const uint8_t* adj, void* out, int out_dtype, void* workspace,
                                           size_t workspace_bytes, void* stream_) {
  SV_REQUIRE(residuals && labels && member_bits && out, "aggregate_residuals: null pointer");
  return aggregate_driver(nullptr, residuals, labels, B, N, D, SEGVLAD_TOKENS_ND, nullptr, K, member_bits,
                          seg_offsets_host, adj, out, out_dtype, nullptr, workspace, workspace_bytes,
                          reinterpret_cast<cudaStream_t>(stream_));
}

extern "C" int segvlad_aggregate_batch_pca(const float* tokens, int B, int N, int D, int token_layout, const float* centers,
                                           int K, const uint32_t* member_bits, const int32_t* seg_offsets_host,
                                           const uint8_t* adj, const float* pca_mean_f32, void* x_planes, int32_t* labels_out,
                                           void* workspace, size_t workspace_bytes, void* stream_) {
  SV_REQUIRE(tokens && centers && member_bits && x_planes && pca_mean_f32, "aggregate_batch_pca: null pointer");
  SV_REQUIRE((reinterpret_cast<uintptr_t>(x_planes) & 15) == 0 && (reinterpret_cast<uintptr_t>(pca_mean_f32) & 15) == 0,
             "aggregate_batch_pca: x_planes and pca_mean_f32 must be 16-byte aligned");
  return aggregate_driver(tokens, nullptr, nullptr, B, N, D, token_layout, centers, K, member_bits, seg_offsets_host, adj,
                          x_planes, SEGVLAD_OUT_PCA_PLANES, labels_out, workspace, workspace_bytes,
                          reinterpret_cast<cudaStream_t>(stream_), pca_mean_f32);
}

// timing probe: buf = device array of >= 8 * (#aggregate CTAs) uint64, or NULL to disable (not part of the product API)
extern "C" void segvlad_debug_aggregate_probe(unsigned long long* buf) { segvlad::g_agg_dbg = buf; }

extern "C" int segvlad_mask_to_membership(const uint8_t* masks, int S, int Hm, int Wm, int H, int W, int patch,
                                          uint32_t* member_bits, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  SV_REQUIRE(S >= 0 && Hm > 0 && Wm > 0 && H >= patch && W >= patch && patch > 0, "mask_to_membership: bad shape");
  if (S == 0) return SEGVLAD_OK;
  const int dh = H / patch, dw = W / patch, N = dh * dw, Wd = (N + 31) / 32;
  dim3 grid((Wd * 32 + 255) / 256, S);
  mask_to_membership_kernel<<<grid, 256, 0, st>>>(masks, S, Hm, Wm, H, W, patch, dh, dw, member_bits, Wd);
  SV_CHECK_LAUNCH();
  return SEGVLAD_OK;
}

extern "C" int segvlad_mask_centroids(const uint8_t* masks, int S, int Hm, int Wm, double* centroids_xy, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  SV_REQUIRE(S >= 0 && Hm > 0 && Wm > 0 && (S == 0 || (masks && centroids_xy)), "mask_centroids: bad arguments");
  if (S == 0) return SEGVLAD_OK;
  mask_centroid_kernel<<<S, 256, 0, st>>>(masks, Hm, Wm, centroids_xy);
  SV_CHECK_LAUNCH();
  return SEGVLAD_OK;
}
