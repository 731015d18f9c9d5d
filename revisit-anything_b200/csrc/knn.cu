// Exhaustive squared-L2 kNN of query-segment descriptors against a reference-bank shard, sm_100a.
//
// Replaces faiss.IndexFlatL2.add/search at place_rec_main.py:53-61 (faiss 1.7.3 BLAS path:
// d2 = ||q||^2 + ||r||^2 - 2<q,r>, clamped at 0, k smallest ascending).
//
// Structure
//   bank_prepare        fp32 [n,D] -> ONE fp16 plane of the row scaled by a power of two (row maximum in [2^14, 2^15)),
//                       fp32 ||x||^2, and the row's exact rounding residual rho = ||x - fp16(x)||_2; the bank keeps the
//                       running maxima of rho and ||x|| (error model of the filter below)
//   knn_tc_filter       persistent tcgen05 kernel: all-pairs <q,r> in ONE fp16 MMA pass (kind::f16, fp32 accumulate in
//                       TMEM), TMA-fed shared-memory ring, double-buffered 128x256 accumulators; the epilogue turns each
//                       inner product into an approximate d2, compares it with the row's threshold and appends
//                       survivors to the row's candidate list (the all-pairs matrix is never written to HBM)
//   knn_simt_filter     same epilogue behind a plain fp32 FFMA tile kernel (cross-check path)
//   knn_refine          per row: radix-select the k-th best candidate, drop what can no longer be a result
//   knn_rescore         exact fp32 re-evaluation of the surviving candidates on the resident fp32 rows
//   merge_topk          k-way merge of per-shard lists after the all-gather
//
// Exactness of the single fp16 pass.  |<q,r> - <q^,r^>| <= ||q|| rho_r + rho_q ||r^|| (Cauchy-Schwarz on
// q.(r - r^) + (q - q^).r^), plus the fp32 accumulation error of the TMEM chain (<= c_acc ||q|| ||r||), so every
// approximate d2 of a query row is within E_row = 2 (||q|| rho_max + rho_q (n_max + rho_max) + c_acc ||q|| n_max) (+
// epilogue rounding) of the fp32 value the re-score computes.  If T is the k-th smallest APPROXIMATE d2 seen so far, the
// exact k-th is <= T + E, so a reference with approximate d2 > T + 2E can never be among the k nearest: the filter keeps
// `approx <= T + 2E`, the survivors (k + a few) are re-scored exactly and the k best of those are the exact answer.
// For unit-norm 1536-D descriptors 2E ~ 1.7e-3 against a d2 spread of ~0.05: ~10 % more candidates than k.
//
// The reference bank is scanned in rounds of geometrically growing chunks: after a chunk, T = current
// k-th best, so a later chunk of n refs leaves ~ n*k/seen survivors per row.  Buffers overflowing
// (adversarially ordered or massively duplicated banks) raise a flag; the host then re-runs that query block with
// chunks <= C - k and an exact re-score + exact (d2, idx) selection of k after every round (cannot overflow).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include <map>
#include <vector>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace segvlad {

constexpr int kCandCap = 4096;     // candidate slots per query row
constexpr int kMaxK = 1024;
constexpr int kExactFlag = (int)0x80000000;  // cand_idx sign bit: this candidate's d2 is already the exact fp32 value
constexpr int kQueryBlock = 16384; // query rows per pass (bounds the workspace)
constexpr int kTileM = 128, kTileN = 256, kTileK = 64;
// filter kernel threads = 32 * (2 + kEpi): warp0 TMA, warp1 MMA (+TMEM alloc), then kEpi = 8 epilogue warps (2 per TMEM lane
// quarter) or 16 (4 per lane quarter, 64 columns each) for shallow contractions (D <= 1024), where a tile's MMAs are too
// short to hide an 8-warp epilogue

struct SelState {
  float* tau;      // [rows] current k-th best d2 (+inf until k candidates seen)
  int* cnt;        // [rows] candidates stored
  float* cand_d2;  // [rows][kCandCap]
  int* cand_idx;   // [rows][kCandCap]  shard-local ref row
  int* overflow;   // [1]
};

__device__ __forceinline__ float make_d2(float qn, float rn, float ip) {
  float v = __fsub_rn(__fadd_rn(qn, rn), 2.0f * ip);
  return v > 0.f ? v : 0.f;  // clamp negatives (and -0, NaN) to +0 like faiss
}

// survivors of rounds >= 1: one atomic slot claim per survivor
__device__ __forceinline__ void cand_append(const SelState& s, int row, int col, float d2) {
  int pos = atomicAdd(s.cnt + row, 1);
  if (pos < kCandCap) {
    s.cand_d2[(size_t)row * kCandCap + pos] = d2;
    s.cand_idx[(size_t)row * kCandCap + pos] = col;
  } else {
    *s.overflow = 1;
  }
}

// ------------------------------------------------------------------------------------------------
// bank preparation: one warp per row
struct BankView {
  const __half* h16;      // [n][Dp] fp16 plane of the row scaled by 2^(14 - e_row) (sub-normal results flushed to 0)
  const float* norms;     // [n] ||x||^2 (fp32)
  const float4* meta;     // [n] {||x||^2, 2^(e_row - 14), rho = ||x - h16 / scale||_2 (rounded up), ||x|| (rounded up)}
  const unsigned* stats;  // [0] max rho, [1] max ||x|| over the prepared rows (bit patterns of non-negative floats)
  const float* x32;       // [n][D] fp32 rows inside the bank (exact re-scoring of the selected candidates) -- or unused when
  const float* const* x32_slot;   // the bank is a VIEW: *x32_slot (device word in the stats block) is where the rows live:
  int Dp;                 // the bank's own region or the caller's matrix (segvlad_bank_prepare_view: no 4 n D byte copy)
};
static inline int padded_dim(int D) { return (int)align_up((size_t)D, kTileK); }
static BankView bank_view(const void* bank, int n, int D) {
  Carver c(const_cast<void*>(bank));
  BankView v;
  v.Dp = padded_dim(D);
  v.h16 = c.take<__half>((size_t)n * v.Dp);
  v.norms = c.take<float>(n);
  v.meta = c.take<float4>(n);
  v.stats = c.take<unsigned>(64);
  v.x32 = c.take<float>((size_t)n * D);
  v.x32_slot = reinterpret_cast<const float* const*>(v.stats + 8);
  return v;
}
struct BankOut { __half* h16; float* norms; float4* meta; unsigned* stats; float* x32; int Dp; };
static BankOut bank_out(const BankView& v) {
  return {const_cast<__half*>(v.h16), const_cast<float*>(v.norms), const_cast<float4*>(v.meta),
          const_cast<unsigned*>(v.stats), const_cast<float*>(v.x32), v.Dp};
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Second half of the row preparation; the fp32 values of the row already sit in x32 (written by this warp's lanes at
// the same d stride, or copied there from the host).  ss / mx: this lane's partial sum of squares / maximum |x|.
__device__ __forceinline__ void bank_finish_row(const BankOut& b, int row, int D, float ss, float mx, int lane) {
  ss = warp_sum(ss);
  mx = warp_max(mx);
  int e = 0;
  if (mx > 0.f && mx < INFINITY) e = ilogbf(mx);
  e = e < -100 ? -100 : (e > 100 ? 100 : e);
  const float sc = __int_as_float((127 + 14 - e) << 23);    // 2^(14 - e): row maximum -> [2^14, 2^15) < fp16 max
  const float inv = __int_as_float((127 - 14 + e) << 23);
  const float* xr = b.x32 + (size_t)row * D;
  __half* hr = b.h16 + (size_t)row * b.Dp;
  float err2 = 0.f;
  for (int d = lane; d < b.Dp; d += 32) {
    const float v = d < D ? xr[d] : 0.f;
    __half h = __float2half_rn(v * sc);
    float hf = __half2float(h);
    if (fabsf(hf) < 6.103515625e-05f) { h = __ushort_as_half((unsigned short)0); hf = 0.f; }   // no fp16 sub-normals
    const float er = v - hf * inv;      // exact in fp32 (both on v's grid, within a factor of two)
    err2 = fmaf(er, er, err2);
    hr[d] = h;
  }
  err2 = warp_sum(err2);
  if (lane == 0) {
    float rho = sqrtf(err2) * 1.0005f, nr = sqrtf(ss) * 1.000001f;
    if (!(rho < INFINITY)) rho = 0.f;   // NaN / inf rows: their scores are NaN -> clamped to 0 in both passes
    if (!(nr < INFINITY)) nr = 0.f;
    b.norms[row] = ss;
    b.meta[row] = make_float4(ss, inv, rho, nr);
    volatile unsigned* vs = b.stats;
    if (__float_as_uint(rho) > vs[0]) atomicMax(b.stats + 0, __float_as_uint(rho));
    if (__float_as_uint(nr) > vs[1]) atomicMax(b.stats + 1, __float_as_uint(nr));
  }
}

// copy_rows = 0: b.x32 IS x (view bank): nothing to copy, the second half reads the caller's rows
__global__ void bank_prepare_kernel(const float* __restrict__ x, int n, int D, BankOut b, int copy_rows) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* xr = x + (size_t)row * D;
  float* x32r = b.x32 + (size_t)row * D;
  float ss = 0.f, mx = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float v = xr[d];
    if (copy_rows) x32r[d] = v;
    ss = fmaf(v, v, ss);
    mx = fmaxf(mx, fabsf(v));
  }
  bank_finish_row(b, row, D, ss, mx, lane);
}

// rows [row0, row0+nrows) whose fp32 values already sit in the bank's x32 region (host-streamed path)
__global__ void bank_prepare_rows_kernel(int row0, int nrows, int D, BankOut b) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= nrows) return;
  const int row = row0 + r;
  const float* xr = b.x32 + (size_t)row * D;
  float ss = 0.f, mx = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float v = xr[d];
    ss = fmaf(v, v, ss);
    mx = fmaxf(mx, fabsf(v));
  }
  bank_finish_row(b, row, D, ss, mx, lane);
}

__global__ void bank_prepare_f64_kernel(const double* __restrict__ x, int n, int D, int normalize_rows, BankOut b) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const double* xr = x + (size_t)row * D;
  float* x32r = b.x32 + (size_t)row * D;
  double nrm = 1.0;
  if (normalize_rows) {
    double s2 = 0.0;
    for (int d = lane; d < D; d += 32) s2 += xr[d] * xr[d];
    nrm = sqrt(warp_sum(s2));  // no eps: a zero row becomes NaN exactly like normalizeFeat
  }
  float ss = 0.f, mx = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float v = (float)(normalize_rows ? xr[d] / nrm : xr[d]);
    x32r[d] = v;
    ss = fmaf(v, v, ss);
    mx = fmaxf(mx, fabsf(v));
  }
  bank_finish_row(b, row, D, ss, mx, lane);
}

// Error model of the approximate (single fp16 pass) scores of one query row -- see the header comment.
struct ErrModel {
  const float4* qmeta;     // query bank meta
  const unsigned* rstats;  // reference bank maxima; nullptr: the scores are exact (SIMT path), bound = 0
  float c_acc;             // fp32 accumulation error of the TMEM chain + the fp32 re-score, relative to ||q|| ||r||
};
__device__ __forceinline__ float row_err_bound(const ErrModel& em, int qrow) {
  if (!em.rstats) return 0.f;
  const float4 m = em.qmeta[qrow];
  const float rmax = __uint_as_float(em.rstats[0]), nmax = __uint_as_float(em.rstats[1]);
  const float nq = m.w, s = nq + nmax;
  float e = 2.f * (nq * rmax + m.z * (nmax + rmax) + em.c_acc * nq * nmax);
  e = fmaf(4.8e-7f, s * s, e);          // roundings of the epilogue arithmetic (2^-21 (||q|| + ||r||)^2)
  return (e < INFINITY) ? e * 1.001f : 0.f;
}
static inline float acc_err_const(int D) {
  return (float)((D / 16 + 32) * ldexp(1.0, -22) + (D / 32 + 8) * ldexp(1.0, -23));
}

__global__ void row_norms_kernel(const float* __restrict__ x, int n, int D, float* __restrict__ norms) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  float ss = 0.f;
  for (int d = lane; d < D; d += 32) { float v = x[(size_t)row * D + d]; ss = fmaf(v, v, ss); }
  ss = warp_sum(ss);
  if (lane == 0) norms[row] = ss;
}

__global__ void sel_init_kernel(SelState s, int rows, int first_chunk) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < rows) { s.tau[r] = INFINITY; s.cnt[r] = first_chunk; }
  if (r == 0) *s.overflow = 0;
}

// ------------------------------------------------------------------------------------------------
// SIMT cross-check path: 64x64 tile, 16x16 threads, 4x4 micro-tile, fp32 FFMA in ascending d order.
__global__ void __launch_bounds__(256)
knn_simt_filter_kernel(const float* __restrict__ q, const float* __restrict__ r, const float* __restrict__ qn,
                       const float* __restrict__ rn, int rows, int D, int c0, int c1, int first_round, SelState sel) {
  __shared__ float As[16][65], Bs[16][65];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int row0 = blockIdx.y * 64, col0 = c0 + blockIdx.x * 64;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < D; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      int rr = i >> 4, kk = i & 15;
      int gr = row0 + rr, gc = col0 + rr, gk = k0 + kk;
      As[kk][rr] = (gr < rows && gk < D) ? q[(size_t)gr * D + gk] : 0.f;
      Bs[kk][rr] = (gc < c1 && gk < D) ? r[(size_t)gc * D + gk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = row0 + ty * 4 + i;
    if (row >= rows) continue;
    const float qnr = qn[row], tau = sel.tau[row];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = col0 + tx * 4 + j;
      if (col >= c1) continue;
      const float d2 = make_d2(qnr, rn[col], acc[i][j]);
      if (first_round) {
        sel.cand_d2[(size_t)row * kCandCap + (col - c0)] = d2;
        sel.cand_idx[(size_t)row * kCandCap + (col - c0)] = col;
      } else if (d2 <= tau) {
        cand_append(sel, row, col, d2);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Persistent tcgen05 all-pairs + filter kernel, templated on the CTA-group size.
//   kCtas == 1 : one CTA per 128 x 256 tile (4-stage ring of 48 KB: Q 128 rows + R 256 rows of one 64-channel block)
//   kCtas == 2 : a CTA PAIR (cluster 2x1x1, tcgen05 cta_group::2) per 256 x 256 tile; each CTA stages its own 128
//                query rows and HALF of the reference tile (6-stage ring of 32 KB), the pair's MMA (M=256) reads
//                both halves => L2->SM operand traffic per MMA drops by 1/3
// grid = min(#tiles, #SMs) CTAs (pairs: even); tile t -> (col_tile = t / n_row_tiles, row_tile = t % n_row_tiles)
// so concurrently running CTAs share a reference tile through L2 and the bank streams from HBM once.
template <int kCtas> struct TcCfg;
template <> struct TcCfg<1> { static constexpr int kStagesT = 4; static constexpr uint32_t kRRows = 256; };
template <> struct TcCfg<2> { static constexpr int kStagesT = 6; static constexpr uint32_t kRRows = 128; };
constexpr uint32_t kQTileBytes = kTileM * kTileK * 2;   // 16 KB
template <int kCtas> __host__ __device__ constexpr uint32_t tc_stage_bytes() { return (kTileM + TcCfg<kCtas>::kRRows) * kTileK * 2; }
template <int kCtas> __host__ __device__ constexpr size_t tc_smem_bytes() {
  return TcCfg<kCtas>::kStagesT * tc_stage_bytes<kCtas>() + 1024 /*align*/ + 256 /*barriers*/ + 4096 /*column constants*/;
}

constexpr uint32_t kPeerMask = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address (-> even CTA)

__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar & kPeerMask), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {  // arrive on the same barrier offset in both CTAs
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void tc_mma_f16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {  // arrive on the even CTA's barrier
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerMask) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// approximate d2 from the scaled fp16 inner product: ip = acc * inv_q * inv_r; cq = 2 inv_q
__device__ __forceinline__ float make_d2_scaled(float qn, float rn, float acc, float cq, float inv_r) {
  const float v = __fsub_rn(__fadd_rn(qn, rn), (acc * cq) * inv_r);
  return v > 0.f ? v : 0.f;
}

// kEpi: epilogue warps per CTA (8 or 16).  r2 ncu at D = 512: a 256 x 256 tile's MMAs take ~2.1 us, the 8-warp epilogue
// (128 columns per thread, latency-bound at ~1 instruction per clock and SM) ~5.6 us -> tensor pipe 35-38 %.
template <int kCtas, int kEpi>
__global__ void __launch_bounds__(64 + 32 * kEpi, 1)
knn_tc_filter_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_r,
                     const float4* __restrict__ rmeta, ErrModel em, int q_row0, int rows, int c0, int c1, int num_kb,
                     int first_round, SelState sel) {
  constexpr int kSt = TcCfg<kCtas>::kStagesT;
  constexpr uint32_t kStageB = tc_stage_bytes<kCtas>();
  constexpr int kPairM = kTileM * kCtas;                            // query rows per (pair) tile
  // kind::f16 instruction descriptor: D=f32 (bit4), A=B=f16 (format 0), K-major both, N>>3 @17, M>>4 @24
  constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(kTileN >> 3) << 17) | ((uint32_t)(kPairM >> 4) << 24);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment for SWIZZLE_128B tiles.  stage layout: [Q 16K][R kRRows x 128 B]
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSt * kStageB);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
  float* cc_inv = reinterpret_cast<float*>(bars + 32);   // [2 tile parities][kTileN] 2^(e_r - 14) of the tile's columns
  float* cc_nrn = cc_inv + 2 * kTileN;                   // [2][kTileN] -||r||^2
  const uint32_t bar_full = smem_u32(bars + 0);        // [kSt <= 8]  (pairs: only the leader's copies are used)
  const uint32_t bar_empty = smem_u32(bars + 8);       // [kSt <= 8]
  const uint32_t bar_tfull = smem_u32(bars + 16);      // [2] accumulator ready
  const uint32_t bar_tempty = smem_u32(bars + 18);     // [2] accumulator drained (pairs: leader's copies)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t cta_rank = 0;
  if (kCtas == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
  const bool leader = cta_rank == 0;
  const int n_row_tiles = (rows + kPairM - 1) / kPairM;
  const int n_col_tiles = (c1 - c0 + kTileN - 1) / kTileN;
  const int n_tiles = n_row_tiles * n_col_tiles;
  const int worker = blockIdx.x / kCtas, n_workers = gridDim.x / kCtas;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kSt; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, kEpi * kCtas); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM: all 512 columns = two 128x256 fp32 accumulators (per CTA)
    if (kCtas == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    }
  }
  tc_fence_before();
  if (kCtas == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (one thread per CTA) =====
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_r) : "memory");
      uint32_t stage = 0, phase = 0;
      for (int t = worker; t < n_tiles; t += n_workers) {
        const int ct = t / n_row_tiles, rt = t - ct * n_row_tiles;
        const int qy = q_row0 + rt * kPairM + (int)cta_rank * kTileM;
        const int ry = c0 + ct * kTileN + (int)cta_rank * (int)TcCfg<kCtas>::kRRows * (kCtas - 1);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t sbase = smem_u32(smem + stage * kStageB);
          const uint32_t fb = bar_full + 8 * stage;
          if (kCtas == 1) {
            mbar_arrive_expect_tx(fb, kStageB);
            tma_load_2d(sbase, &map_q, fb, kb * kTileK, qy);
            tma_load_2d(sbase + kQTileBytes, &map_r, fb, kb * kTileK, ry);
          } else {
            if (leader) mbar_arrive_expect_tx(fb, 2 * kStageB);   // both CTAs' bytes land on the leader's barrier
            tma_load_2d_pair(sbase, &map_q, fb, kb * kTileK, qy);
            tma_load_2d_pair(sbase + kQTileBytes, &map_r, fb, kb * kTileK, ry);
          }
          if (++stage == kSt) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread; pairs: the leader CTA only) =====
    if (lane == 0 && leader) {
      uint32_t stage = 0, phase = 0, it = 0;
      for (int t = worker; t < n_tiles; t += n_workers, ++it) {
        const uint32_t buf = it & 1, use = it >> 1;
        mbar_wait(bar_tempty + 8 * buf, (use & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * kTileN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sbase = smem_u32(smem + stage * kStageB);
          const uint64_t qd = umma_desc_sw128(sbase), rd = umma_desc_sw128(sbase + kQTileBytes);
#pragma unroll
          for (int kk = 0; kk < kTileK / 16; ++kk) {
            const uint64_t adv = (uint64_t)(kk * 32 >> 4);  // 16 fp16 = 32 B along K inside the swizzle row
            if (kCtas == 1) tc_mma_bf16(d_tmem, qd + adv, rd + adv, kIdesc, (kb | kk) != 0);
            else tc_mma_f16_pair(d_tmem, qd + adv, rd + adv, kIdesc, (kb | kk) != 0);
          }
          // frees the smem stage (in both CTAs) when these MMAs retire
          if (kCtas == 1) tc_commit(bar_empty + 8 * stage); else tc_commit_pair(bar_empty + 8 * stage);
          if (++stage == kSt) { stage = 0; phase ^= 1; }
        }
        if (kCtas == 1) tc_commit(bar_tfull + 8 * buf); else tc_commit_pair(bar_tfull + 8 * buf);  // accumulator done
      }
    }
  } else {
    // ===== epilogue warps: TMEM lane quarter = warp % 4, column half = (warp - 2) / 4; each CTA drains its own rows.
    // One warp per scheduler cannot hide its own ALU / L1 latencies (r1 ncu: 27 instructions per score at IPC ~0.2
    // left the single fp16 MMA pass waiting for the accumulator), hence two warps per quarter and a lean loop:
    //   survive  <=>  (acc * 2 inv_q) * inv_r - ||r||^2 >= ||q||^2 - tau        (one LDG.128, FMUL, FFMA, FSETP, LOP)
    // Survivor slots are claimed with ONE atomic per (row, 32-column chunk); the atomic's round trip is hidden behind
    // the next chunk's TMEM load and tests, then the chunk's survivors are stored column by column, skipping (warp-
    // uniformly) the columns in which no lane has one.
    constexpr int kParts = kEpi / 4;                 // column parts of a tile (one warp per part and TMEM lane quarter)
    constexpr int kPartCols = kTileN / kParts;       // 128 or 64 columns per warp
    const int quarter = warp & 3, half = (warp - 2) >> 2;    // `half`: this warp's column part (0 .. kParts - 1)
    constexpr int kChunksPerWarp = kPartCols / 32;
    uint32_t it = 0;
    for (int t = worker; t < n_tiles; t += n_workers, ++it) {
      const int ct = t / n_row_tiles, rt = t - ct * n_row_tiles;
      const uint32_t buf = it & 1, use = it >> 1;
      const int row = rt * kPairM + (int)cta_rank * kTileM + quarter * 32 + lane;
      const bool row_ok = row < rows;
      float qnr = 0.f, cq = 0.f, thr = INFINITY;
      if (row_ok) {
        const float4 qm = em.qmeta[q_row0 + row];
        qnr = qm.x;
        cq = 2.f * qm.y;
        // keep approx <= T + 2E (header comment); tau = +inf -> thr = -inf
        thr = qnr - (sel.tau[row] + 2.f * row_err_bound(em, q_row0 + row));
      }
      const int colbase = c0 + ct * kTileN + half * kPartCols;
      if (!first_round) {
        // this tile's column constants -> shared memory, BEFORE waiting for the accumulator (the loads and the barrier of
        // the 8 epilogue warps overlap the tile's MMAs; buffers alternate with the tile parity)
        const int e = (warp - 2) * 32 + lane;
        if (e < kTileN) {
          const int gcol = min(c0 + ct * kTileN + e, c1 - 1);
          const float4 rmc = __ldg(rmeta + gcol);
          cc_inv[(it & 1) * kTileN + e] = rmc.y;
          cc_nrn[(it & 1) * kTileN + e] = -rmc.x;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpi) : "memory");     // the epilogue warps
      }
      mbar_wait(bar_tfull + 8 * buf, use & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * kTileN + half * kPartCols;
      if (first_round) {
        // round 0: tau = +inf, every score is a candidate -> dense store at (col - c0), no counters
#pragma unroll 1
        for (int c = 0; c < kPartCols; c += 32) {
          uint32_t v[32];
          tc_ld32(taddr + c, v);
          if (row_ok) {
            // 16-byte stores: a warp-wide 4-byte store at a 16 KB row stride touches 32 sectors for 128 useful bytes
            // (scores only: candidate i of round 0 IS reference row i, the refine kernel knows)
            float* cd = sel.cand_d2 + (size_t)row * kCandCap + (colbase + c - c0);
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const int col = colbase + c + j;
              float dd[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float4 rm = __ldg(rmeta + min(col + u, c1 - 1));
                dd[u] = make_d2_scaled(qnr, rm.x, __uint_as_float(v[j + u]), cq, rm.y);
              }
              if (col + 3 < c1) {
                *reinterpret_cast<float4*>(cd + j) = make_float4(dd[0], dd[1], dd[2], dd[3]);
              } else {
#pragma unroll
                for (int u = 0; u < 4; ++u)
                  if (col + u < c1) cd[j + u] = dd[u];
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (kCtas == 1) mbar_arrive(bar_tempty + 8 * buf); else mbar_arrive_leader(bar_tempty + 8 * buf); }
      } else {
        float* cd = sel.cand_d2 + (size_t)row * kCandCap;
        int* ci = sel.cand_idx + (size_t)row * kCandCap;
        // test: u = (acc * 2 inv_q) * inv_r - ||r||^2 >= thr; u replaces the accumulator value in v[] because the
        // survivor's approximate d2 is simply ||q||^2 - u (no second look-up of the column constants)
        // The tile's 256 column constants are staged in shared memory once per tile by the 256 epilogue threads (one column
        // each; double-buffered by tile parity, one named barrier per tile) and read back as broadcast 8-byte pairs; the
        // test runs on packed fp32 pairs (mul.f32x2 / fma.f32x2: the same roundings as the scalar form, bit-identical u).
        // r2 ncu at D = 512: with a broadcast LDG.128 per column and scalar FMUL / FFMA the epilogue (~6 us per tile) was
        // twice the tile's MMA time and the tensor pipe 35-38 % active.
        const float2* tinv = reinterpret_cast<const float2*>(cc_inv + (it & 1) * kTileN + half * kPartCols);
        const float2* tnrn = reinterpret_cast<const float2*>(cc_nrn + (it & 1) * kTileN + half * kPartCols);
        const float2 cq2 = make_float2(cq, cq);
        auto test_chunk = [&](uint32_t (&v)[32], int col0) -> uint32_t {
          uint32_t m = 0;
          const int lc = (col0 - colbase) >> 1;              // pair index inside this warp's half of the tile
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float2 a = make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1]));
            const float2 u = __ffma2_rn(__fmul2_rn(a, cq2), tinv[lc + (j >> 1)], tnrn[lc + (j >> 1)]);
            v[j] = __float_as_uint(u.x);
            v[j + 1] = __float_as_uint(u.y);
            if (u.x >= thr) m |= 1u << j;
            if (u.y >= thr) m |= 2u << j;
          }
          const int nv = c1 - col0;                          // columns of this chunk inside the scanned range
          if (nv < 32) m = nv <= 0 ? 0u : (m & ((1u << nv) - 1u));
          return m;
        };
        auto store_chunk = [&](const uint32_t (&v)[32], uint32_t m, int pos, int col0) {
          const uint32_t any = __reduce_or_sync(0xffffffffu, m);
          if (any == 0u) return;
          if (m != 0u && pos + __popc(m) > kCandCap) *sel.overflow = 1;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if ((any >> j) & 1u) {            // warp-uniform skip of columns without survivors
              if ((m >> j) & 1u) {
                if (pos < kCandCap) {
                  cd[pos] = fmaxf(qnr - __uint_as_float(v[j]), 0.f);
                  ci[pos] = col0 + j;
                }
                ++pos;
              }
            }
          }
        };
        uint32_t va[32], vb[32];
        uint32_t ma = 0, mb = 0;
        int pa = 0, pb = 0;
#pragma unroll 1
        for (int cc = 0; cc < kChunksPerWarp; cc += 2) {
          tc_ld32(taddr + cc * 32, va);
          ma = test_chunk(va, colbase + cc * 32);
          if (ma) pa = atomicAdd(sel.cnt + row, __popc(ma));
          if (cc > 0) store_chunk(vb, mb, pb, colbase + (cc - 1) * 32);
          tc_ld32(taddr + (cc + 1) * 32, vb);
          if (cc + 2 >= kChunksPerWarp) {     // last TMEM read of this tile: hand the accumulator back early
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (kCtas == 1) mbar_arrive(bar_tempty + 8 * buf); else mbar_arrive_leader(bar_tempty + 8 * buf); }
          }
          mb = test_chunk(vb, colbase + (cc + 1) * 32);
          if (mb) pb = atomicAdd(sel.cnt + row, __popc(mb));
          store_chunk(va, ma, pa, colbase + cc * 32);
        }
        store_chunk(vb, mb, pb, colbase + (kChunksPerWarp - 1) * 32);
      }
    }
  }
  tc_fence_before();
  if (kCtas == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (kCtas == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ------------------------------------------------------------------------------------------------
// Per-row candidate refinement (one CTA per row).  Radix-select the k-th smallest d2 T (3 passes of 11/11/10 bits over
// the fp32 bit pattern; d2 >= +0 so unsigned order == float order), then by mode:
//   kRefineSelect : approximate scores.  Keep every candidate with d2 <= T + 2E (header comment), tau = T.
//   kRefineFinal  : exact (re-scored) values.  Keep d2 <= T, bitonic sort of the survivors on the 64-bit key
//                   (d2 bits << 32 | idx) => ascending (d2, idx); the first k are written out (global row = offset + idx).
//   kRefineExactK : exact values, conservative schedule.  Same sort, the first k go back to the candidate list
//                   (count = k exactly, ties broken by index like the final order), tau = T.
//   kRefineSelectLoose : kRefineSelect for the rounds BETWEEN scan chunks, where T only has to be an upper bound of the
//                   k-th best: the digit loop stops as soon as the bin holding the k-th value has <= kLooseBin
//                   members and T = that bin's largest possible value (one histogram pass instead of three or four
//                   for ~10 % more survivors; the pass before the exact re-score stays tight).
enum { kRefineSelect = 0, kRefineFinal = 1, kRefineExactK = 2, kRefineSelectLoose = 3 };
constexpr int kLooseBin = 32;
constexpr int kSmallSort = 512;   // up to this many candidates: sort them all, no selection passes

__device__ __forceinline__ void bitonic_sort_keys(unsigned long long* keys, int P, int tid) {
  // every compare-exchange is owned by one thread: t enumerates the P / 2 pairs of a stage, i = t with a zero bit inserted at
  // the stage's distance j (r2: looping over all P indices and skipping the upper partner wasted half the iterations of an
  // issue-bound kernel)
  for (int kk2 = 2; kk2 <= P; kk2 <<= 1) {
    for (int j = kk2 >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (P >> 1); t += 256) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int ixj = i | j;
        const unsigned long long a = keys[i], b = keys[ixj];
        if ((a > b) == ((i & kk2) == 0)) { keys[i] = b; keys[ixj] = a; }
      }
      __syncthreads();
    }
  }
}

// dense_first: the candidates are round 0's dense scores (candidate i = reference row i; the filter kernel does not
// store indices in that round).
__global__ void __launch_bounds__(256)
knn_refine_kernel(SelState sel, ErrModel em, int k, int mode, int dense_first, long long row_offset, int q_row0,
                  float* __restrict__ d2_out, long long* __restrict__ idx_out,
                  unsigned long long* __restrict__ packed_out, int* __restrict__ ref_cnt, int* __restrict__ pair_total) {
  __shared__ __align__(16) unsigned s_v[kCandCap];   // d2 bit patterns   } re-used as 64-bit sort keys
  __shared__ __align__(16) int s_i[kCandCap];        // candidate rows    }
  __shared__ int s_hist[256];
  __shared__ int s_warp[8];
  __shared__ unsigned s_or[8], s_and[8];
  __shared__ int s_bin, s_kk, s_out, s_val;
  static_assert(sizeof(unsigned) * kCandCap * 2 == sizeof(unsigned long long) * kCandCap, "key aliasing");
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(s_v);   // s_v and s_i are contiguous: 32 KB
  const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const bool loose = mode == kRefineSelectLoose;
  if (loose) mode = kRefineSelect;
  int n = sel.cnt[row];
  if (n > kCandCap) n = kCandCap;
  float* cd = sel.cand_d2 + (size_t)row * kCandCap;
  int* ci = sel.cand_idx + (size_t)row * kCandCap;
  // ref_cnt != nullptr (select pass before an inverted re-score): the survivors that still need the exact value are
  // counted per reference row on the way out (first step of the list inversion, see knn_rescore_ref_kernel)
  if (mode != kRefineFinal && n <= k) {
    if (dense_first)
      for (int i = tid; i < n; i += 256) ci[i] = i;
    if (ref_cnt) {
      int mine = 0;
      for (int i = tid; i < n; i += 256) {
        const int col = dense_first ? i : ci[i];
        if (col >= 0) { atomicAdd(ref_cnt + col, 1); ++mine; }
      }
      mine = __reduce_add_sync(0xffffffffu, mine);
      if (lane == 0 && mine) atomicAdd(pair_total, mine);
    }
    return;
  }
  // Select keeps the exact-flag of a candidate; the exact modes order ties by the bare index
  const unsigned idmask = mode == kRefineSelect ? 0xffffffffu : (unsigned)~kExactFlag;
  const float twoE = mode == kRefineSelect ? 2.f * row_err_bound(em, q_row0 + row) : 0.f;
  int c = n;                       // candidates that reach the sort
  int pending = 0;                 // this thread's survivors that still need their exact value (ref_cnt != nullptr)
  unsigned T = 0x7f800000u;        // k-th smallest d2 (bit pattern)
  int P = 1;

  if (n <= kSmallSort) {
    while (P < n) P <<= 1;
    for (int i = tid; i < P; i += 256)
      keys[i] = i < n ? (((unsigned long long)__float_as_uint(cd[i]) << 32) | ((unsigned)(dense_first ? i : ci[i]) & idmask))
                      : ~0ull;
    __syncthreads();
    bitonic_sort_keys(keys, P, tid);
    if (n > k) T = (unsigned)(keys[k - 1] >> 32);
    if (mode == kRefineSelect) {   // n > k here; the survivors are a prefix of the sorted list
      const unsigned keepT = __float_as_uint(__fadd_ru(__uint_as_float(T), twoE));
      int kept = 0;
      for (int i0 = 0; i0 < P; i0 += 256) {
        const int i = i0 + tid;
        const bool ok = i < n && (unsigned)(keys[i] >> 32) <= keepT;
        if (ok) {
          cd[i] = __uint_as_float((unsigned)(keys[i] >> 32));
          ci[i] = (int)(unsigned)keys[i];
          if (ref_cnt && (int)(unsigned)keys[i] >= 0) { atomicAdd(ref_cnt + (int)(unsigned)keys[i], 1); ++pending; }
        }
        kept += __syncthreads_count(ok);
      }
      if (ref_cnt) {
        pending = __reduce_add_sync(0xffffffffu, pending);
        if (lane == 0 && pending) atomicAdd(pair_total, pending);
      }
      if (tid == 0) { sel.cnt[row] = kept; sel.tau[row] = __uint_as_float(T); }
      return;
    }
    if (n > k && tid == 0) sel.tau[row] = __uint_as_float(T);
  } else {
    // ---- radix select of the k-th smallest over the bits in which the candidates differ ----
    unsigned vor = 0u, vand = 0xffffffffu;
    for (int i = tid; i < n; i += 256) {
      const unsigned v = __float_as_uint(cd[i]);
      s_v[i] = v;
      s_i[i] = dense_first ? i : ci[i];
      vor |= v; vand &= v;
    }
    vor = __reduce_or_sync(0xffffffffu, vor);
    vand = __reduce_and_sync(0xffffffffu, vand);
    if (lane == 0) { s_or[w] = vor; s_and[w] = vand; }
    if (tid == 0) s_out = 0;   // (here, not next to its first use: s_bin / s_kk / s_out / s_val share one 16-byte word that the
    __syncthreads();           //  digit loop reads with a vector load -- racecheck flagged the later single-thread write)
    vor = 0u; vand = 0xffffffffu;
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) { vor |= s_or[ww]; vand &= s_and[ww]; }
    // d2 values of one query share their sign / exponent / leading mantissa bits: a digit taken there would send
    // every candidate to the same histogram bin (serialised shared-memory atomics); start below the common prefix
    const bool compact = n > k;                            // (n <= k: final pass of a small bank, sort everything)
    int hi = compact ? 32 - __clz(vor ^ vand) : 0;         // number of low bits that vary (0: all equal)
    unsigned mask = hi >= 32 ? 0u : ~((1u << hi) - 1u);
    unsigned prefix = vand & mask;
    int kk = k;
    while (hi > 0) {
      const int lo = hi > 8 ? hi - 8 : 0;
      const unsigned dm = (1u << (hi - lo)) - 1u;
      s_hist[tid] = 0;
      __syncthreads();
      for (int i = tid; i < n; i += 256) {
        const unsigned v = s_v[i];
        if ((v & mask) == prefix) atomicAdd(&s_hist[(v >> lo) & dm], 1);
      }
      __syncthreads();
      const int val = s_hist[tid];
      int incl = val;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
      if (lane == 31) s_warp[w] = incl;
      __syncthreads();
      int base = 0;
      for (int ww = 0; ww < w; ++ww) base += s_warp[ww];
      const int excl = base + incl - val;
      if (kk > excl && kk <= excl + val) { s_bin = tid; s_kk = kk - excl; s_val = val; }
      __syncthreads();
      prefix |= (unsigned)s_bin << lo;
      mask |= dm << lo;
      kk = s_kk;
      hi = lo;
      if (loose && compact && s_val <= kLooseBin) {   // upper bound of the k-th value: the largest value of its bin
        prefix |= (1u << lo) - 1u;
        break;
      }
    }
    if (compact) T = prefix;
    // approximate scores: everything within 2E of the k-th best may still belong to the exact top k
    const unsigned keepT = mode == kRefineSelect ? __float_as_uint(__fadd_ru(__uint_as_float(T), twoE)) : T;
    for (int i0 = 0; compact && i0 < n; i0 += 256) {
      const int i = i0 + tid;
      const unsigned v = i < n ? s_v[i] : 0xffffffffu;
      const bool ok = i < n && v <= keepT;
      const unsigned bal = __ballot_sync(0xffffffffu, ok);
      int wbase = 0;
      if (lane == 0 && bal) wbase = atomicAdd(&s_out, __popc(bal));
      wbase = __shfl_sync(0xffffffffu, wbase, 0);
      if (ok) {
        const int pos = wbase + __popc(bal & ((1u << lane) - 1u));
        cd[pos] = __uint_as_float(v);
        ci[pos] = s_i[i];
        if (ref_cnt && mode == kRefineSelect && s_i[i] >= 0) { atomicAdd(ref_cnt + s_i[i], 1); ++pending; }
      }
    }
    if (ref_cnt && mode == kRefineSelect) {
      pending = __reduce_add_sync(0xffffffffu, pending);
      if (lane == 0 && pending) atomicAdd(pair_total, pending);
    }
    __syncthreads();
    if (compact) {
      c = s_out;
      if (tid == 0) { sel.cnt[row] = c; sel.tau[row] = __uint_as_float(T); }
    }
    if (mode == kRefineSelect) return;
    // sort the c (~k) survivors by (d2, idx)
    while (P < c) P <<= 1;
    unsigned long long mykeys[kCandCap / 256];
#pragma unroll
    for (int t = 0; t < kCandCap / 256; ++t) {
      const int i = tid + t * 256;
      mykeys[t] = ~0ull;
      if (i < c)
        mykeys[t] = compact ? (((unsigned long long)__float_as_uint(cd[i]) << 32) | ((unsigned)ci[i] & idmask))
                            : (((unsigned long long)s_v[i] << 32) | ((unsigned)s_i[i] & idmask));
    }
    __syncthreads();   // everyone has read its candidates before the smem region is re-purposed
#pragma unroll
    for (int t = 0; t < kCandCap / 256; ++t) {
      const int i = tid + t * 256;
      if (i < P) keys[i] = mykeys[t];
    }
    __syncthreads();
    bitonic_sort_keys(keys, P, tid);
  }
  if (mode == kRefineExactK) {
    const int keep = c < k ? c : k;
    for (int i = tid; i < keep; i += 256) {
      cd[i] = __uint_as_float((unsigned)(keys[i] >> 32));
      ci[i] = (int)(unsigned)keys[i] | kExactFlag;
    }
    if (tid == 0) sel.cnt[row] = keep;
    return;
  }
  // final lists: (d2 fp32, global row int64) like faiss, and / or packed (d2 bits << 32 | int32 global row) -- the payload of
  // the all-gather of a row-sharded search, written straight into the send buffer (padding: +inf | 0xffffffff)
  const size_t o = (size_t)(q_row0 + row) * k;
  for (int i = tid; i < k; i += 256) {
    const bool real = i < c;
    const unsigned d2b = real ? (unsigned)(keys[i] >> 32) : 0x7f800000u;
    const long long gi = real ? row_offset + (long long)(unsigned)keys[i] : -1ll;
    if (d2_out) { d2_out[o + i] = __uint_as_float(d2b); idx_out[o + i] = gi; }
    if (packed_out) packed_out[o + i] = ((unsigned long long)d2b << 32) | (unsigned)(int)gi;
  }
}

// Exact re-scoring of the surviving candidates: the single fp16 pass only has to be good enough to discard (its error
// is bounded, see the header); the <= k + few survivors of each row that are not exact yet are re-evaluated with plain
// fp32 FMAs on the resident fp32 rows (one warp per candidate, lanes stride the channels in float4, shuffle-tree
// reduction), and the final selection + sort uses these values.
// The gather of ~k rows of 4 D bytes per query (13.6 GB for config 2) is what this kernel costs; it runs at ~7.6 TB/s
// effective (HBM + ~25 % L2 hits).  r1 experiments that did NOT help: sweeping the bank in L2-sized waves with
// persistent CTAs (2.1-2.8 ms vs 1.8 ms: the gather is bound by memory-level parallelism, which one CTA per query
// row with every warp on its own candidate maximises), two candidates per warp iteration (register pressure).
__global__ void __launch_bounds__(256)
knn_rescore_kernel(SelState sel, const float* const* __restrict__ q32_slot, const float* const* __restrict__ r32_slot,
                   const float* __restrict__ qn, const float* __restrict__ rn, int q_row0, int D,
                   const int* __restrict__ done_flag) {
  if (done_flag && *done_flag) return;   // the inverted pass (knn_rescore_ref_kernel) re-scored everything
  const float* __restrict__ q32 = *q32_slot;   // fp32 rows: the bank's own copy or the caller's matrix (view banks)
  const float* __restrict__ r32 = *r32_slot;
  const int row = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int n = sel.cnt[row];
  if (n > kCandCap) n = kCandCap;
  const float* qr = q32 + (size_t)(q_row0 + row) * D;
  const float qnr = qn[q_row0 + row];
  float* cd = sel.cand_d2 + (size_t)row * kCandCap;
  int* ci = sel.cand_idx + (size_t)row * kCandCap;
  const bool vec = (D & 3) == 0;
  for (int j = w; j < n; j += 8) {
    const int col = ci[j];
    if (col < 0) continue;   // already exact (kept by an earlier exact selection)
    const float* rr = r32 + (size_t)col * D;
    float acc = 0.f;
    if (vec) {
      const float4* q4 = reinterpret_cast<const float4*>(qr);
      const float4* r4 = reinterpret_cast<const float4*>(rr);
      for (int d = lane; d < (D >> 2); d += 32) {
        const float4 a = __ldg(q4 + d), b = __ldg(r4 + d);
        acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
      }
    } else {
      for (int d = lane; d < D; d += 32) acc = fmaf(__ldg(qr + d), __ldg(rr + d), acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) { cd[j] = make_d2(qnr, rn[col], acc); ci[j] = col | kExactFlag; }
  }
}

// ------------------------------------------------------------------------------------------------
// Re-score BY REFERENCE ROW (the final pass of a resident search).  knn_rescore_kernel above gathers ~k + margin fp32
// reference rows per QUERY: every reference row is fetched ~Nq (k + margin) / Nr times (22 x for config 2) and the
// kernel is bound by DRAM (r1 ncu: 10.8 GB read for a 0.61 GB bank).  Here the candidate lists are inverted first --
// count pairs per reference row, exclusive scan, scatter (query row, slot) -- and one warp per reference row stages that
// row in shared memory ONCE (the bank streams from HBM exactly once) and takes the query rows of its pairs from L2, which
// holds the whole query block (16384 x D fp32 <= 100 MB).  The per-lane summation order is the one of
// knn_rescore_kernel, so both kernels produce bit-identical values.
struct InvState {
  int* ref_off;     // [Nr + 1] pairs per reference row -> exclusive offsets -> (after the scatter) end offsets
  int* blk_sum;     // [ceil(Nr / kScanTile) + 1]
  unsigned* pairs;  // [cap] (query row << 12) | candidate slot
  int* state;       // [0] total pairs, [1] 1: the inverted lists were built (they fit), 0: fall back to the gather kernel
  int cap;
};
constexpr int kScanTile = 4096;       // elements per CTA of the offset scan (1024 threads x 4)
constexpr int kPairsPerRow = 512;     // pair-list capacity per query row of a block (k + margin <= this, else gather kernel)
constexpr int kRefWarps = 8;          // warps (= reference rows in flight) per CTA of knn_rescore_ref_kernel
constexpr int kRefMaxD = 3072;        // 8 rows x 3072 x 4 B = 96 KB of shared memory per CTA

// exclusive scan of ref_off[0 .. n): per-tile scan + tile sums, scan of the tile sums (one CTA), add-back
__global__ void __launch_bounds__(1024)
inv_scan_tiles_kernel(InvState inv, int n) {
  __shared__ int s_w[32];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int base = blockIdx.x * kScanTile + tid * 4;
  int v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = base + i < n ? inv.ref_off[base + i] : 0;
  const int tsum = v[0] + v[1] + v[2] + v[3];
  int incl = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  if (lane == 31) s_w[w] = incl;
  __syncthreads();
  if (w == 0) {
    int x = s_w[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += t; }
    s_w[lane] = x;
  }
  __syncthreads();
  int run = incl - tsum + (w ? s_w[w - 1] : 0);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (base + i < n) inv.ref_off[base + i] = run;
    run += v[i];
  }
  if (tid == 1023) inv.blk_sum[blockIdx.x] = run;
}
__global__ void __launch_bounds__(1024)
inv_scan_sums_kernel(InvState inv, int n_tiles) {
  __shared__ int s_w[32];
  __shared__ int s_carry;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int b0 = 0; b0 < n_tiles; b0 += 1024) {
    const int i = b0 + tid;
    const int v = i < n_tiles ? inv.blk_sum[i] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) s_w[w] = incl;
    __syncthreads();
    if (w == 0) {
      int x = s_w[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += t; }
      s_w[lane] = x;
    }
    __syncthreads();
    const int carry = s_carry;
    if (i < n_tiles) inv.blk_sum[i] = carry + incl - v + (w ? s_w[w - 1] : 0);
    __syncthreads();
    if (tid == 1023) s_carry = carry + s_w[31];
    __syncthreads();
  }
  if (tid == 0) inv.state[1] = (inv.state[0] > 0 && inv.state[0] <= inv.cap) ? 1 : 0;
}
__global__ void __launch_bounds__(1024)
inv_scan_add_kernel(InvState inv, int n) {
  const int add = inv.blk_sum[blockIdx.x];
  const int base = blockIdx.x * kScanTile + threadIdx.x * 4;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (base + i < n) inv.ref_off[base + i] += add;
}

__global__ void __launch_bounds__(256)
inv_scatter_kernel(SelState sel, InvState inv) {
  if (inv.state[1] == 0) return;
  const int row = blockIdx.x;
  int n = sel.cnt[row];
  if (n > kCandCap) n = kCandCap;
  const int* ci = sel.cand_idx + (size_t)row * kCandCap;
  for (int j = threadIdx.x; j < n; j += 256) {
    const int col = ci[j];
    if (col >= 0) inv.pairs[atomicAdd(inv.ref_off + col, 1)] = ((unsigned)row << 12) | (unsigned)j;
  }
}

// grid-stride over reference rows, one warp per row; after the scatter ref_off[r] is the END of row r's pair list and
// ref_off[r - 1] its start.  Shared memory: kRefWarps rows of D floats.
__global__ void __launch_bounds__(kRefWarps * 32)
knn_rescore_ref_kernel(SelState sel, InvState inv, const float* const* __restrict__ q32_slot,
                       const float* const* __restrict__ r32_slot, const float* __restrict__ qn, const float* __restrict__ rn,
                       int q_row0, int Nr, int D) {
  extern __shared__ __align__(16) float s_ref[];
  if (inv.state[1] == 0) return;
  const float* __restrict__ q32 = *q32_slot;
  const float* __restrict__ r32 = *r32_slot;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float* my = s_ref + (size_t)w * D;
  const bool vec = (D & 3) == 0;
  const int nwarps = gridDim.x * kRefWarps;
  for (int r = blockIdx.x * kRefWarps + w; r < Nr; r += nwarps) {
    const int beg = r ? inv.ref_off[r - 1] : 0, end = inv.ref_off[r];
    if (beg == end) continue;
    const float* rr = r32 + (size_t)r * D;
    __syncwarp();
    if (vec) {
      for (int d = lane; d < (D >> 2); d += 32)
        reinterpret_cast<float4*>(my)[d] = __ldcs(reinterpret_cast<const float4*>(rr) + d);   // streamed once
    } else {
      for (int d = lane; d < D; d += 32) my[d] = __ldcs(rr + d);
    }
    __syncwarp();
    const float rnr = rn[r];
    for (int p = beg; p < end; p += 2) {
      const bool two = p + 1 < end;
      const unsigned pa = inv.pairs[p], pb = inv.pairs[two ? p + 1 : p];
      const int rowa = (int)(pa >> 12), rowb = (int)(pb >> 12);
      const float* qa = q32 + (size_t)(q_row0 + rowa) * D;
      const float* qb = q32 + (size_t)(q_row0 + rowb) * D;
      float acca = 0.f, accb = 0.f;
      if (vec) {
        const float4* a4 = reinterpret_cast<const float4*>(qa);
        const float4* b4 = reinterpret_cast<const float4*>(qb);
        const float4* m4 = reinterpret_cast<const float4*>(my);
#pragma unroll 4
        for (int d = lane; d < (D >> 2); d += 32) {
          const float4 a = __ldg(a4 + d), b = __ldg(b4 + d), m = m4[d];
          acca = fmaf(a.x, m.x, acca); acca = fmaf(a.y, m.y, acca); acca = fmaf(a.z, m.z, acca); acca = fmaf(a.w, m.w, acca);
          accb = fmaf(b.x, m.x, accb); accb = fmaf(b.y, m.y, accb); accb = fmaf(b.z, m.z, accb); accb = fmaf(b.w, m.w, accb);
        }
      } else {
        for (int d = lane; d < D; d += 32) {
          const float m = my[d];
          acca = fmaf(__ldg(qa + d), m, acca);
          accb = fmaf(__ldg(qb + d), m, accb);
        }
      }
      acca = warp_sum(acca);
      accb = warp_sum(accb);
      if (lane == 0) {
        const size_t oa = (size_t)rowa * kCandCap + (pa & 4095u);
        sel.cand_d2[oa] = make_d2(qn[q_row0 + rowa], rnr, acca);
        sel.cand_idx[oa] = r | kExactFlag;
        if (two) {
          const size_t ob = (size_t)rowb * kCandCap + (pb & 4095u);
          sel.cand_d2[ob] = make_d2(qn[q_row0 + rowb], rnr, accb);
          sel.cand_idx[ob] = r | kExactFlag;
        }
      }
    }
  }
}

// k-way merge of G per-shard lists: [G][Nq][k] -> [Nq][k]
__global__ void __launch_bounds__(256)
merge_topk_kernel(const float* __restrict__ d2p, const long long* __restrict__ idxp, int G, int Nq, int k,
                  float* __restrict__ d2_out, long long* __restrict__ idx_out) {
  extern __shared__ unsigned long long mkeys[];
  const int row = blockIdx.x, tid = threadIdx.x;
  const int n = G * k;
  int P = 1;
  while (P < n) P <<= 1;
  for (int i = tid; i < P; i += 256) {
    unsigned long long key = ~0ull;
    if (i < n) {
      const int g = i / k, j = i - g * k;
      const size_t o = ((size_t)g * Nq + row) * k + j;
      const long long id = idxp[o];
      float d = d2p[o];
      if (id >= 0) key = ((unsigned long long)__float_as_uint(d > 0.f ? d : 0.f) << 32) | (unsigned)id;
    }
    mkeys[i] = key;
  }
  __syncthreads();
  for (int kk = 2; kk <= P; kk <<= 1) {
    for (int j = kk >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < P; i += 256) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = mkeys[i], b = mkeys[ixj];
          if ((a > b) == ((i & kk) == 0)) { mkeys[i] = b; mkeys[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < k; i += 256) {
    const unsigned long long key = i < P ? mkeys[i] : ~0ull;
    const size_t o = (size_t)row * k + i;
    if (key != ~0ull) {
      d2_out[o] = __uint_as_float((unsigned)(key >> 32));
      idx_out[o] = (long long)(unsigned)key;
    } else {
      d2_out[o] = INFINITY;
      idx_out[o] = -1;
    }
  }
}

// k-way merge of G SORTED per-shard lists in the packed layout of the all-gather: parts[g * part_stride + row * k + j].
// No sort: every element finds its rank in the merged order directly -- its own position j plus, for every other list,
// the number of elements that precede it (binary search; lists with a smaller shard number win ties, which only occur
// between padding entries because global rows are unique) -- and drops itself at that rank if it is < k.  One CTA per
// query row, the G lists staged in shared memory.
__global__ void __launch_bounds__(256)
merge_packed_kernel(const unsigned long long* __restrict__ parts, int G, size_t part_stride, int Nq, int k,
                    float* __restrict__ d2_out, long long* __restrict__ idx_out, unsigned long long* __restrict__ packed_out) {
  extern __shared__ unsigned long long mlists[];   // [G][k]
  const int row = blockIdx.x, tid = threadIdx.x;
  const int n = G * k;
  for (int i = tid; i < n; i += 256) {
    const int g = i / k, j = i - g * k;
    mlists[i] = parts[(size_t)g * part_stride + (size_t)row * k + j];
  }
  __syncthreads();
  // One warp per list, 32 consecutive elements at a time.  Ranks grow along a sorted list, so the first chunk that contains
  // an element of rank >= k ends the list: with balanced shards a list contributes ~k / G elements and the warp stops after
  // one or two chunks instead of ranking all k of them (r2, G = 8: ranking every element cost 0.32 ms per 10 k queries --
  // 9e9 thread instructions of binary search -- more than the all-gather itself).
  const int warp = tid >> 5, lane = tid & 31;
  for (int g = warp; g < G; g += 8) {
    for (int j0 = 0; j0 < k; j0 += 32) {
      const int j = j0 + lane;
      const bool act = j < k;
      const unsigned long long key = act ? mlists[g * k + j] : ~0ull;
      int rank = j;
      for (int h = 0; act && h < G && rank < k; ++h) {
        if (h == g) continue;
        const unsigned long long* L = mlists + h * k;
        // number of elements of list h ordered before (key, g): key' < key, or key' == key and h < g
        int lo = 0, hi = k;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          const unsigned long long v = L[mid];
          if (v < key || (v == key && h < g)) lo = mid + 1; else hi = mid;
        }
        rank += lo;
      }
      if (act && rank < k) {
        const size_t o = (size_t)row * k + rank;
        if (d2_out) {
          d2_out[o] = __uint_as_float((unsigned)(key >> 32));
          idx_out[o] = (long long)(int)(unsigned)key;      // sign-extends the -1 of padding entries
        }
        if (packed_out) packed_out[o] = key;
      }
      if (__any_sync(0xffffffffu, act && rank >= k)) break;   // every later element of this list ranks even higher
    }
  }
}

__global__ void flags_or_kernel(const int* __restrict__ flags, int n, int* __restrict__ out) {
  int v = 0;
  for (int i = threadIdx.x; i < n; i += 32) v |= flags[i];
  v = __reduce_or_sync(0xffffffffu, v);
  if (threadIdx.x == 0) *out = v != 0;
}

// test hook: dense approximate scores of round 0 + the row's error bound
__global__ void knn_debug_copy_kernel(SelState sel, ErrModel em, int q_row0, int rows, int Nr, float* __restrict__ approx_out,
                                      float* __restrict__ bound_out) {
  const int row = blockIdx.x;
  if (row >= rows) return;
  for (int j = threadIdx.x; j < Nr; j += blockDim.x)
    approx_out[(size_t)(q_row0 + row) * Nr + j] = sel.cand_d2[(size_t)row * kCandCap + j];
  if (threadIdx.x == 0) bound_out[q_row0 + row] = row_err_bound(em, q_row0 + row);
}

// ------------------------------------------------------------------------------------------------
// host side
struct KnnLayout {
  SelState sel[2];   // two query blocks can be in flight (one per internal stream)
  InvState inv[2];   // inverted candidate lists of the final re-score (one per block in flight)
  float* qn; float* rn; int* flags; int block_rows; int n_blocks; size_t total;
};
static KnnLayout carve_knn(void* ws, int Nq, int Nr) {
  Carver c(ws);
  KnnLayout L;
  // query blocks: bounded by kQueryBlock rows (workspace) and, above 4096 queries, at least two of them so that one
  // block's refine / re-score / sort (HBM-bound) overlaps the other block's tensor-core scan on a second stream
  // (measured r1: splitting a single block in two to overlap refine/re-score with the other half's tensor-core scan
  //  on a second stream made the step 10 % SLOWER -- the HBM-bound re-score and the TMA-fed scan contend -- so two
  //  blocks are only in flight when the query set exceeds kQueryBlock and SEGVLAD_KNN_DUAL=1)
  int nsplit = (Nq + kQueryBlock - 1) / kQueryBlock;
  if (nsplit < 1) nsplit = 1;
  L.block_rows = (int)align_up((size_t)((Nq + nsplit - 1) / nsplit), 256);
  if (L.block_rows > kQueryBlock) L.block_rows = kQueryBlock;
  if (L.block_rows < 1) L.block_rows = 1;
  L.n_blocks = (Nq + L.block_rows - 1) / L.block_rows;
  for (int i = 0; i < 2; ++i) {
    const int rows = (i == 0 || L.n_blocks > 1) ? L.block_rows : 0;
    L.sel[i].tau = c.take<float>(rows);
    L.sel[i].cnt = c.take<int>(rows);
    L.sel[i].cand_d2 = c.take<float>((size_t)rows * kCandCap);
    L.sel[i].cand_idx = c.take<int>((size_t)rows * kCandCap);
    L.sel[i].overflow = c.take<int>(1);
    L.inv[i].ref_off = c.take<int>(rows ? (size_t)Nr + 1 : 0);
    L.inv[i].blk_sum = c.take<int>(rows ? (size_t)(Nr + kScanTile - 1) / kScanTile + 1 : 0);
    L.inv[i].cap = rows * kPairsPerRow;
    L.inv[i].pairs = c.take<unsigned>((size_t)L.inv[i].cap);
    L.inv[i].state = c.take<int>(4);
  }
  L.qn = c.take<float>(Nq);   // SIMT path only
  L.rn = c.take<float>(Nr);   // SIMT path only
  L.flags = c.take<int>(L.n_blocks + 1);
  L.total = c.total();
  return L;
}

static int make_map(CUtensorMap* m, const __half* base, int n, int Dp, int box_rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return SEGVLAD_ECUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)Dp, (cuuint64_t)n};
  cuuint64_t strides[1] = {(cuuint64_t)Dp * 2};
  cuuint32_t box[2] = {(cuuint32_t)kTileK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return SEGVLAD_ECUDA; }
  return SEGVLAD_OK;
}

// chunk schedule: first chunk fills the buffer, later chunks grow with the number of refs already seen; a chunk is
// sized to leave ~ (C - k) / fill survivors per row (fill = head-room against uneven banks AND the knob that trades
// survivor stores in the filter epilogue against the number of refine rounds)
static double fill_factor() {
  static double f = 0.0;
  if (f == 0.0) {
    const char* e = getenv("SEGVLAD_KNN_FILL");
    f = e ? atof(e) : 2.5;
    if (!(f >= 1.5)) f = 1.5;
  }
  return f;
}
static int next_chunk(int seen, int k, int Nr, bool safe) {
  long long c;
  if (seen == 0) c = kCandCap;
  else if (safe) c = kCandCap - k;
  else c = (long long)((double)seen * (double)(kCandCap - k) / ((double)k * fill_factor()));
  c = c / kTileN * kTileN;
  if (c < kTileN) c = kTileN;
  if (c > Nr - seen) c = Nr - seen;
  return (int)c;
}

// Per-(host thread, device) context of the search: internal streams, a pool of re-usable timing-less events, the SM count
// and the one-time kernel attributes.  Nothing here is created per call (r1: up to 512 events were created and destroyed
// by every host-streamed search, the streams were process-wide statics bound to the first device used).
struct DevCtx {
  int dev = -1, num_sms = 0;
  bool attrs_set = false, merge_attr_set = false;
  cudaStream_t aux[2] = {nullptr, nullptr};
  cudaStream_t copy = nullptr;
  std::vector<cudaEvent_t> events;
  int event(int i, cudaEvent_t* out) {
    while ((int)events.size() <= i) {
      cudaEvent_t e;
      SV_CHECK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      events.push_back(e);
    }
    *out = events[i];
    return SEGVLAD_OK;
  }
};
static int dev_ctx(DevCtx** out) {
  static thread_local std::map<int, DevCtx> ctxs;
  int dev = 0;
  SV_CHECK_CUDA(cudaGetDevice(&dev));
  DevCtx& c = ctxs[dev];
  if (c.dev < 0) {
    SV_CHECK_CUDA(cudaDeviceGetAttribute(&c.num_sms, cudaDevAttrMultiProcessorCount, dev));
    c.dev = dev;
  }
  *out = &c;
  return SEGVLAD_OK;
}
constexpr int kEvFork = 0, kEvJoin0 = 1, kEvJoin1 = 2, kEvHost0 = 3, kEvHostQ = 4, kEvFeed = 8;   // event pool slots

struct TcArgs { BankView q, r; CUtensorMap mq, mr; int num_sms; int ctas; int epi; };
struct KnnOut { float* d2; long long* idx; unsigned long long* packed; };   // final lists: unpacked and / or packed
// asynchronous mode (segvlad_knn_async): no host synchronisation -- the chosen schedule runs, the OR of the query blocks'
// overflow flags goes to overflow_dev and the caller decides (after ITS synchronisation point) whether to repeat the
// call with the conservative schedule
struct KnnAsync { int schedule; int* overflow_dev; };
struct SimtArgs { const float* q; const float* r; };

// Host-streamed reference bank: fp32 rows arrive from pinned host memory on a copy stream, sub-chunk by
// sub-chunk, directly into the bank's fp32 region; the compute stream waits for each sub-chunk's event, splits it
// into the fp16 plane and scans it while the next sub-chunks are still in flight over PCIe.
struct HostFeed {
  const float* r_host;     // [Nr, D] pinned host rows
  cudaStream_t copy;       // copy stream
  int sub_rows;            // rows per copy / scan sub-chunk (multiple of kTileN)
};
struct SubChunk { int c0, c1; int first_round; int last_of_round; int final_pass; };

// one launch of the tcgen05 filter kernel over reference columns [c0, c1) for `rows` queries starting at q_row0
template <int kCtas, int kEpi>
static int launch_filter_t(const TcArgs* ta, const ErrModel& em, int q_row0, int rows, int c0, int c1, int first_round,
                           const SelState& sel, cudaStream_t st) {
  const int chunk = c1 - c0;
  const int n_tiles = ((rows + kCtas * kTileM - 1) / (kCtas * kTileM)) * ((chunk + kTileN - 1) / kTileN);
  int workers = ta->num_sms / kCtas;
  if (n_tiles < workers) workers = n_tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kCtas * workers);
  cfg.blockDim = dim3(64 + 32 * kEpi);
  cfg.dynamicSmemBytes = tc_smem_bytes<kCtas>();
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = kCtas; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  SV_CHECK_CUDA(cudaLaunchKernelEx(&cfg, knn_tc_filter_kernel<kCtas, kEpi>, ta->mq, ta->mr, ta->r.meta, em, q_row0, rows, c0, c1,
                                   ta->q.Dp / kTileK, first_round, sel));
  return SEGVLAD_OK;
}
static int launch_filter(const TcArgs* ta, const ErrModel& em, int q_row0, int rows, int c0, int c1, int first_round,
                         const SelState& sel, cudaStream_t st) {
  if (ta->ctas == 2)
    return ta->epi == 16 ? launch_filter_t<2, 16>(ta, em, q_row0, rows, c0, c1, first_round, sel, st)
                         : launch_filter_t<2, 8>(ta, em, q_row0, rows, c0, c1, first_round, sel, st);
  return ta->epi == 16 ? launch_filter_t<1, 16>(ta, em, q_row0, rows, c0, c1, first_round, sel, st)
                       : launch_filter_t<1, 8>(ta, em, q_row0, rows, c0, c1, first_round, sel, st);
}

struct BlockCtx { SelState sel; InvState inv; const float* qn; const float* rn; };

// 0: always the per-query gather kernel; 1 (default): inverted lists for the final pass of a resident search
static int loose_select() {
  const char* e = getenv("SEGVLAD_KNN_LOOSE");
  return (e && e[0] == '0') ? 0 : 1;
}
static int rescore_by_ref() {
  const char* e = getenv("SEGVLAD_KNN_RESCORE_REF");   // (read per call: the tests flip it to compare both kernels)
  return (e && e[0] == '0') ? 0 : 1;
}
static int run_block(bool tc, const TcArgs* ta, const SimtArgs* sa, const BlockCtx& L, int q_row0, int rows, int Nr,
                     int D, int k, long long row_offset, bool safe, const KnnOut& out, cudaStream_t st, DevCtx* dc,
                     const HostFeed* feed = nullptr) {
  float* d2_out = out.d2;
  long long* idx_out = out.idx;
  unsigned long long* packed_out = out.packed;
  // tensor-core path: filter on the approximate distances with the error-model margin, re-score exactly, keep k
  const ErrModel em = tc ? ErrModel{ta->q.meta, ta->r.stats, acc_err_const(D)} : ErrModel{nullptr, nullptr, 0.f};
  // Host-streamed bank: the GPU waits for PCIe anyway, so every sub-chunk is closed with an exact selection (select ->
  // re-score only the new survivors -> exact k); when the last rows arrive only their few survivors are left to
  // re-score, instead of all k + margin candidates of every query.
  const bool incremental = tc && feed != nullptr && !safe;
  const int first = next_chunk(0, k, Nr, safe);
  sel_init_kernel<<<(rows + 255) / 256, 256, 0, st>>>(L.sel, rows, first);
  SV_CHECK_LAUNCH();
  // schedule: rounds (refine boundaries) split into sub-chunks (one filter launch each)
  std::vector<SubChunk> sched;
  int ns = 0;
  for (int seen = 0; seen < Nr;) {
    const int chunk = next_chunk(seen, k, Nr, safe);
    for (int s0 = seen; s0 < seen + chunk;) {
      // host-streamed: the last rows arrive in small pieces so that little work is left when the copy ends
      int sub = chunk;
      if (feed && seen > 0) sub = (Nr - s0 <= feed->sub_rows) ? (feed->sub_rows / 2 > kTileN ? feed->sub_rows / 2 : kTileN) : feed->sub_rows;
      const int s1 = (s0 + sub < seen + chunk) ? s0 + sub : seen + chunk;
      sched.push_back({s0, s1, seen == 0, incremental || s1 == seen + chunk, s1 == Nr});
      ++ns;
      s0 = s1;
    }
    seen += chunk;
  }
  std::vector<cudaEvent_t> ev(feed ? ns : 0);
  if (feed) {
    for (int i = 0; i < ns; ++i) {
      const int erc = dc->event(kEvFeed + i, &ev[i]);     // pooled: created once per (thread, device), never per call
      if (erc) return erc;
      const size_t off = (size_t)sched[i].c0 * D, cnt = (size_t)(sched[i].c1 - sched[i].c0) * D;
      SV_CHECK_CUDA(cudaMemcpyAsync(const_cast<float*>(ta->r.x32) + off, feed->r_host + off, cnt * sizeof(float),
                                    cudaMemcpyHostToDevice, feed->copy));
      SV_CHECK_CUDA(cudaEventRecord(ev[i], feed->copy));
    }
  }
  for (int i = 0; i < ns; ++i) {
    const int c0 = sched[i].c0, c1 = sched[i].c1, chunk = c1 - c0, first_round = sched[i].first_round;
    if (feed) {
      SV_CHECK_CUDA(cudaStreamWaitEvent(st, ev[i], 0));
      bank_prepare_rows_kernel<<<(chunk + 7) / 8, 256, 0, st>>>(c0, chunk, D, bank_out(ta->r));
      SV_CHECK_LAUNCH();
    }
    const int pslot = prof_begin(SEGVLAD_PROF_KNN_FILTER, st);
    if (tc) {
      const int frc = launch_filter(ta, em, q_row0, rows, c0, c1, first_round, L.sel, st);
      if (frc) return frc;
    } else {
      dim3 grid((chunk + 63) / 64, (rows + 63) / 64);
      knn_simt_filter_kernel<<<grid, 256, 0, st>>>(sa->q + (size_t)q_row0 * D, sa->r, L.qn + q_row0, L.rn, rows, D, c0,
                                                   c1, first_round, L.sel);
    }
    prof_end(pslot, st);
    SV_CHECK_LAUNCH();
    if (!sched[i].last_of_round) continue;
    const bool final_pass = sched[i].final_pass != 0;
    const int dense = (tc && first_round) ? 1 : 0;   // round 0 of the tensor-core filter stores scores only
    if (tc && (safe || incremental || final_pass)) {
      // approximate selection (prunes to ~k + margin), exact re-score of what is not exact yet, exact selection.
      // Final pass of a resident search: every row has ~k + margin candidates to re-score -> by reference row (the bank
      // streams from HBM once; the select pass counts the survivors per reference row on its way out); sub-chunk passes
      // (host-streamed / conservative schedule) re-score few NEW survivors -> gather kernel.  The gather kernel is also
      // launched after the inverted pass and returns at once unless the pair lists did not fit (state[1] == 0, decided
      // on the device).
      const bool by_ref = final_pass && !incremental && !safe && rescore_by_ref() && D <= kRefMaxD && Nr >= 4 * kScanTile;
      int* inv_cnt = by_ref ? L.inv.ref_off : nullptr;
      int* inv_tot = by_ref ? L.inv.state : nullptr;
      if (by_ref) {
        SV_CHECK_CUDA(cudaMemsetAsync(L.inv.ref_off, 0, sizeof(int) * ((size_t)Nr + 1), st));
        SV_CHECK_CUDA(cudaMemsetAsync(L.inv.state, 0, sizeof(int) * 4, st));
      }
      knn_refine_kernel<<<rows, 256, 0, st>>>(L.sel, em, k, kRefineSelect, dense, row_offset, q_row0, d2_out, idx_out,
                                              packed_out, inv_cnt, inv_tot);
      SV_CHECK_LAUNCH();
      const int rslot = prof_begin(SEGVLAD_PROF_KNN_RESCORE, st);
      if (by_ref) {
        const int n_tiles = (Nr + kScanTile - 1) / kScanTile;
        inv_scan_tiles_kernel<<<n_tiles, 1024, 0, st>>>(L.inv, Nr);
        SV_CHECK_LAUNCH();
        inv_scan_sums_kernel<<<1, 1024, 0, st>>>(L.inv, n_tiles);
        SV_CHECK_LAUNCH();
        inv_scan_add_kernel<<<n_tiles, 1024, 0, st>>>(L.inv, Nr);
        SV_CHECK_LAUNCH();
        inv_scatter_kernel<<<rows, 256, 0, st>>>(L.sel, L.inv);
        SV_CHECK_LAUNCH();
        const size_t rsmem = (size_t)kRefWarps * D * sizeof(float);
        int ctas_per_sm = (int)((200 * 1024) / (rsmem + 1024));
        if (ctas_per_sm > 8) ctas_per_sm = 8;
        if (ctas_per_sm < 1) ctas_per_sm = 1;
        int rgrid = ta->num_sms * ctas_per_sm;
        if (rgrid > (Nr + kRefWarps - 1) / kRefWarps) rgrid = (Nr + kRefWarps - 1) / kRefWarps;
        knn_rescore_ref_kernel<<<rgrid, kRefWarps * 32, rsmem, st>>>(L.sel, L.inv, ta->q.x32_slot, ta->r.x32_slot, ta->q.norms,
                                                                    ta->r.norms, q_row0, Nr, D);
        SV_CHECK_LAUNCH();
      }
      knn_rescore_kernel<<<rows, 256, 0, st>>>(L.sel, ta->q.x32_slot, ta->r.x32_slot, ta->q.norms, ta->r.norms, q_row0, D,
                                               by_ref ? L.inv.state + 1 : nullptr);
      prof_end(rslot, st);
      SV_CHECK_LAUNCH();
      knn_refine_kernel<<<rows, 256, 0, st>>>(L.sel, em, k, final_pass ? kRefineFinal : kRefineExactK, 0, row_offset,
                                              q_row0, d2_out, idx_out, packed_out, nullptr, nullptr);
      SV_CHECK_LAUNCH();
    } else {
      const int mode = final_pass ? kRefineFinal : (safe ? kRefineExactK : (loose_select() ? kRefineSelectLoose : kRefineSelect));
      knn_refine_kernel<<<rows, 256, 0, st>>>(L.sel, em, k, mode, dense, row_offset, q_row0, d2_out, idx_out, packed_out,
                                              nullptr, nullptr);
      SV_CHECK_LAUNCH();
    }
  }
  return SEGVLAD_OK;
}

static int knn_driver(bool tc, TcArgs* ta, const SimtArgs* sa, int Nq, int Nr, int D, int k, long long row_offset,
                      const KnnOut& out, void* workspace, size_t workspace_bytes, cudaStream_t st,
                      const HostFeed* feed = nullptr, const KnnAsync* as = nullptr) {
  SV_REQUIRE(Nq >= 0 && Nr >= 0 && D > 0, "knn: bad shape");
  SV_REQUIRE(k > 0 && k <= kMaxK, "knn: k must be in [1, %d] (got %d)", kMaxK, k);
  SV_REQUIRE((out.d2 != nullptr) == (out.idx != nullptr) && (out.d2 || out.packed), "knn: no output buffer");
  SV_REQUIRE(!out.packed || row_offset + (long long)Nr <= 0x7fffffffll, "knn: packed output needs global rows < 2^31");
  if (as && as->overflow_dev) SV_CHECK_CUDA(cudaMemsetAsync(as->overflow_dev, 0, sizeof(int), st));
  if (Nq == 0) return SEGVLAD_OK;
  KnnLayout L = carve_knn(workspace, Nq, Nr);
  if (!workspace || workspace_bytes < L.total) {
    set_error("knn: workspace %zu < required %zu", workspace_bytes, L.total);
    return SEGVLAD_EWORKSPACE;
  }
  const int n_blocks = L.n_blocks, QB = L.block_rows;
  SV_REQUIRE(n_blocks <= 64, "knn: too many query blocks (%d)", n_blocks);
  DevCtx* dc = nullptr;
  { const int rc = dev_ctx(&dc); if (rc) return rc; }
  BlockCtx ctx[2] = {{L.sel[0], L.inv[0], L.qn, L.rn}, {L.sel[1], L.inv[1], L.qn, L.rn}};
  if (Nr == 0) {  // nothing to search: pad like faiss
    for (int b = 0; b < n_blocks; ++b) {
      const int q0 = b * QB, rows = (Nq - q0) < QB ? (Nq - q0) : QB;
      sel_init_kernel<<<(rows + 255) / 256, 256, 0, st>>>(L.sel[0], rows, 0);
      SV_CHECK_LAUNCH();
      knn_refine_kernel<<<rows, 256, 0, st>>>(L.sel[0], ErrModel{nullptr, nullptr, 0.f}, k, kRefineFinal, 0, row_offset, q0,
                                              out.d2, out.idx, out.packed, nullptr, nullptr);
      SV_CHECK_LAUNCH();
    }
    return SEGVLAD_OK;
  }
  if (!tc) {
    row_norms_kernel<<<(Nq + 7) / 8, 256, 0, st>>>(sa->q, Nq, D, L.qn);
    SV_CHECK_LAUNCH();
    row_norms_kernel<<<(Nr + 7) / 8, 256, 0, st>>>(sa->r, Nr, D, L.rn);
    SV_CHECK_LAUNCH();
  }
  // two internal streams when there are >= 2 query blocks and the bank is already resident (development switch)
  const char* denv = getenv("SEGVLAD_KNN_DUAL");
  const bool dual = n_blocks >= 2 && feed == nullptr && denv && denv[0] == '1';
  cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
  if (dual) {
    for (int i = 0; i < 2; ++i)
      if (!dc->aux[i]) SV_CHECK_CUDA(cudaStreamCreateWithFlags(&dc->aux[i], cudaStreamNonBlocking));
    int rc = dc->event(kEvFork, &ev_fork);
    if (!rc) rc = dc->event(kEvJoin0, &ev_join[0]);
    if (!rc) rc = dc->event(kEvJoin1, &ev_join[1]);
    if (rc) return rc;
    SV_CHECK_CUDA(cudaEventRecord(ev_fork, st));
    for (int i = 0; i < 2; ++i) SV_CHECK_CUDA(cudaStreamWaitEvent(dc->aux[i], ev_fork, 0));
  }
  const bool first_safe = as && as->schedule == 1;
  for (int b = 0; b < n_blocks; ++b) {
    const int q0 = b * QB, rows = (Nq - q0) < QB ? (Nq - q0) : QB;
    const int si = dual ? (b & 1) : 0;
    cudaStream_t bs = dual ? dc->aux[si] : st;
    int rc = run_block(tc, ta, sa, ctx[si], q0, rows, Nr, D, k, row_offset, first_safe, out, bs, dc,
                       b == 0 ? feed : nullptr);   // after the first query block the bank is resident
    if (rc) return rc;
    SV_CHECK_CUDA(cudaMemcpyAsync(L.flags + b, ctx[si].sel.overflow, sizeof(int), cudaMemcpyDeviceToDevice, bs));
  }
  if (dual) {
    for (int i = 0; i < 2; ++i) {
      SV_CHECK_CUDA(cudaEventRecord(ev_join[i], dc->aux[i]));
      SV_CHECK_CUDA(cudaStreamWaitEvent(st, ev_join[i], 0));
    }
  }
  if (as) {   // asynchronous contract: publish the flag, never touch the host
    if (as->overflow_dev) {
      flags_or_kernel<<<1, 32, 0, st>>>(L.flags, n_blocks, as->overflow_dev);
      SV_CHECK_LAUNCH();
    }
    return SEGVLAD_OK;
  }
  int hflags[64];
  SV_CHECK_CUDA(cudaMemcpyAsync(hflags, L.flags, sizeof(int) * n_blocks, cudaMemcpyDeviceToHost, st));
  SV_CHECK_CUDA(cudaStreamSynchronize(st));
  for (int b = 0; b < n_blocks; ++b) {
    if (!hflags[b]) continue;
    const int q0 = b * QB, rows = (Nq - q0) < QB ? (Nq - q0) : QB;
    int rc = run_block(tc, ta, sa, ctx[0], q0, rows, Nr, D, k, row_offset, true, out, st, dc);
    if (rc) return rc;
    int f = 0;
    SV_CHECK_CUDA(cudaMemcpyAsync(&f, ctx[0].sel.overflow, sizeof(int), cudaMemcpyDeviceToHost, st));
    SV_CHECK_CUDA(cudaStreamSynchronize(st));
    if (f) { set_error("knn: candidate overflow in the conservative schedule (internal error)"); return SEGVLAD_EOVERFLOW; }
  }
  return SEGVLAD_OK;
}

// where the fp32 rows of a bank live: device word [stats + 8] (written on the stream, after the stats block was cleared)
static cudaError_t set_rows_slot(const BankOut& b, const float* rows, cudaStream_t st) {
  return cudaMemcpyAsync(b.stats + 8, &rows, sizeof(rows), cudaMemcpyHostToDevice, st);   // pageable source: staged at once
}

}  // namespace segvlad

using namespace segvlad;

extern "C" size_t segvlad_bank_bytes(int n, int D) {
  if (n < 0 || D <= 0) return 0;
  Carver c(nullptr);
  const int Dp = padded_dim(D);
  c.take<__half>((size_t)n * Dp);
  c.take<float>(n);
  c.take<float4>(n);
  c.take<unsigned>(64);
  c.take<float>((size_t)n * D);
  return c.total() + 256;
}

extern "C" int segvlad_bank_prepare(const float* x, int n, int D, void* bank, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  SV_REQUIRE(n >= 0 && D > 0 && bank, "bank_prepare: bad arguments");
  SV_REQUIRE((reinterpret_cast<uintptr_t>(bank) & 255) == 0, "bank_prepare: bank must be 256-byte aligned");
  if (n == 0) return SEGVLAD_OK;
  BankOut b = bank_out(bank_view(bank, n, D));
  SV_CHECK_CUDA(cudaMemsetAsync(b.stats, 0, 64 * sizeof(unsigned), st));
  SV_CHECK_CUDA(set_rows_slot(b, b.x32, st));
  bank_prepare_kernel<<<(n + 7) / 8, 256, 0, st>>>(x, n, D, b, 1);
  SV_CHECK_LAUNCH();
  return SEGVLAD_OK;
}

extern "C" int segvlad_bank_prepare_view(const float* x, int n, int D, void* bank, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  SV_REQUIRE(n >= 0 && D > 0 && bank && (x || n == 0), "bank_prepare_view: bad arguments");
  SV_REQUIRE((reinterpret_cast<uintptr_t>(bank) & 255) == 0, "bank_prepare_view: bank must be 256-byte aligned");
  SV_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 || (D & 3) != 0, "bank_prepare_view: x must be 16-byte aligned");
  if (n == 0) return SEGVLAD_OK;
  BankOut b = bank_out(bank_view(bank, n, D));
  SV_CHECK_CUDA(cudaMemsetAsync(b.stats, 0, 64 * sizeof(unsigned), st));
  SV_CHECK_CUDA(set_rows_slot(b, x, st));
  b.x32 = const_cast<float*>(x);          // the split reads the caller's rows; nothing is copied
  bank_prepare_kernel<<<(n + 7) / 8, 256, 0, st>>>(x, n, D, b, 0);
  SV_CHECK_LAUNCH();
  return SEGVLAD_OK;
}

extern "C" int segvlad_bank_prepare_f64(const double* x, int n, int D, int normalize_rows, void* bank, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  SV_REQUIRE(n >= 0 && D > 0 && bank, "bank_prepare_f64: bad arguments");
  SV_REQUIRE((reinterpret_cast<uintptr_t>(bank) & 255) == 0, "bank_prepare_f64: bank must be 256-byte aligned");
  if (n == 0) return SEGVLAD_OK;
  BankOut b = bank_out(bank_view(bank, n, D));
  SV_CHECK_CUDA(cudaMemsetAsync(b.stats, 0, 64 * sizeof(unsigned), st));
  SV_CHECK_CUDA(set_rows_slot(b, b.x32, st));
  bank_prepare_f64_kernel<<<(n + 7) / 8, 256, 0, st>>>(x, n, D, normalize_rows, b);
  SV_CHECK_LAUNCH();
  return SEGVLAD_OK;
}

extern "C" size_t segvlad_knn_workspace_bytes(int Nq, int Nr, int D, int k) {
  (void)D; (void)k;
  if (Nq <= 0 || Nr < 0) return 256;
  return carve_knn(nullptr, Nq, Nr).total;
}

static int tc_setup(TcArgs& ta, const void* qbank, int Nq, const void* rbank, int Nr, int D) {
  ta.q = bank_view(qbank, Nq, D);
  ta.r = bank_view(rbank, Nr, D);
  int rc;
  const char* env = getenv("SEGVLAD_KNN_CTAS");      // 2 (default): CTA pairs / cta_group::2; 1: single-CTA tiles
  ta.ctas = (env && env[0] == '1') ? 1 : 2;
  // epilogue warps: 16 for shallow contractions (D <= 1024: the tile's MMAs are too short to hide 8 warps), else 8
  const char* ee = getenv("SEGVLAD_KNN_EPI");
  ta.epi = ee ? (atoi(ee) == 16 ? 16 : 8) : (padded_dim(D) <= 1024 ? 16 : 8);
  const int r_box = ta.ctas == 2 ? kTileN / 2 : kTileN;
  if ((rc = make_map(&ta.mq, ta.q.h16, Nq, ta.q.Dp, kTileM))) return rc;
  if ((rc = make_map(&ta.mr, ta.r.h16, Nr, ta.r.Dp, r_box))) return rc;
  DevCtx* dc = nullptr;
  if ((rc = dev_ctx(&dc))) return rc;
  ta.num_sms = dc->num_sms;
  if (!dc->attrs_set) {   // once per (thread, device)
    SV_CHECK_CUDA(cudaFuncSetAttribute(knn_tc_filter_kernel<1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)tc_smem_bytes<1>()));
    SV_CHECK_CUDA(cudaFuncSetAttribute(knn_tc_filter_kernel<2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)tc_smem_bytes<2>()));
    SV_CHECK_CUDA(cudaFuncSetAttribute(knn_tc_filter_kernel<1, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)tc_smem_bytes<1>()));
    SV_CHECK_CUDA(cudaFuncSetAttribute(knn_tc_filter_kernel<2, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)tc_smem_bytes<2>()));
    SV_CHECK_CUDA(cudaFuncSetAttribute(knn_rescore_ref_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       kRefWarps * kRefMaxD * (int)sizeof(float)));
    dc->attrs_set = true;
  }
  return SEGVLAD_OK;
}

extern "C" int segvlad_knn(const void* qbank, int Nq, const void* rbank, int Nr, int64_t row_offset, int D, int k,
                           float* d2_out, int64_t* idx_out, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  SV_REQUIRE(D > 0, "knn: bad D");
  TcArgs ta;
  if (Nq > 0 && Nr > 0) {
    int rc = tc_setup(ta, qbank, Nq, rbank, Nr, D);
    if (rc) return rc;
  }
  const KnnOut out{d2_out, reinterpret_cast<long long*>(idx_out), nullptr};
  return knn_driver(true, &ta, nullptr, Nq, Nr, D, k, row_offset, out, workspace, workspace_bytes, st);
}

extern "C" int segvlad_knn_async(const void* qbank, int Nq, const void* rbank, int Nr, int64_t row_offset, int D, int k,
                                 int schedule, float* d2_out, int64_t* idx_out, uint64_t* topk_packed,
                                 int32_t* overflow_dev, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  SV_REQUIRE(D > 0, "knn_async: bad D");
  SV_REQUIRE(schedule == 0 || schedule == 1, "knn_async: schedule must be 0 (fast) or 1 (conservative)");
  SV_REQUIRE(overflow_dev != nullptr, "knn_async: overflow_dev is required (the caller must check it)");
  TcArgs ta;
  if (Nq > 0 && Nr > 0) {
    int rc = tc_setup(ta, qbank, Nq, rbank, Nr, D);
    if (rc) return rc;
  }
  const KnnOut out{d2_out, reinterpret_cast<long long*>(idx_out), reinterpret_cast<unsigned long long*>(topk_packed)};
  const KnnAsync as{schedule, overflow_dev};
  return knn_driver(true, &ta, nullptr, Nq, Nr, D, k, row_offset, out, workspace, workspace_bytes, st, nullptr, &as);
}

extern "C" int segvlad_knn_from_host(const float* q_host, int Nq, const float* r_host, int Nr, int64_t row_offset, int D,
                                     int k, void* qbank, void* rbank, float* d2_out, int64_t* idx_out, void* workspace,
                                     size_t workspace_bytes, void* stream_, void* copy_stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  SV_REQUIRE(D > 0 && Nq > 0 && Nr > 0 && q_host && r_host && qbank && rbank, "knn_from_host: bad arguments");
  SV_REQUIRE((reinterpret_cast<uintptr_t>(qbank) & 255) == 0 && (reinterpret_cast<uintptr_t>(rbank) & 255) == 0,
             "knn_from_host: banks must be 256-byte aligned");
  // everything that can fail on the arguments is checked before any work is enqueued
  SV_REQUIRE(k > 0 && k <= kMaxK, "knn_from_host: k must be in [1, %d] (got %d)", kMaxK, k);
  {
    const KnnLayout Lc = carve_knn(nullptr, Nq, Nr);
    if (!workspace || workspace_bytes < Lc.total) {
      set_error("knn_from_host: workspace %zu < required %zu", workspace_bytes, Lc.total);
      return SEGVLAD_EWORKSPACE;
    }
    SV_REQUIRE(Lc.n_blocks <= 64, "knn_from_host: too many query blocks (%d)", Lc.n_blocks);
  }
  DevCtx* dc = nullptr;
  int rc = dev_ctx(&dc);
  if (rc) return rc;
  cudaStream_t cp = reinterpret_cast<cudaStream_t>(copy_stream_);
  if (!cp) {
    if (!dc->copy) SV_CHECK_CUDA(cudaStreamCreateWithFlags(&dc->copy, cudaStreamNonBlocking));
    cp = dc->copy;
  }
  TcArgs ta;
  rc = tc_setup(ta, qbank, Nq, rbank, Nr, D);
  if (rc) return rc;
  // order the copy stream after everything already queued on the compute stream (buffer re-use across calls)
  cudaEvent_t e0, eq;
  if ((rc = dc->event(kEvHost0, &e0)) || (rc = dc->event(kEvHostQ, &eq))) return rc;
  SV_CHECK_CUDA(cudaEventRecord(e0, st));
  SV_CHECK_CUDA(cudaStreamWaitEvent(cp, e0, 0));
  SV_CHECK_CUDA(cudaMemcpyAsync(const_cast<float*>(ta.q.x32), q_host, sizeof(float) * (size_t)Nq * D,
                                cudaMemcpyHostToDevice, cp));
  SV_CHECK_CUDA(cudaEventRecord(eq, cp));
  SV_CHECK_CUDA(cudaStreamWaitEvent(st, eq, 0));
  SV_CHECK_CUDA(cudaMemsetAsync(const_cast<unsigned*>(ta.q.stats), 0, 64 * sizeof(unsigned), st));
  SV_CHECK_CUDA(cudaMemsetAsync(const_cast<unsigned*>(ta.r.stats), 0, 64 * sizeof(unsigned), st));
  SV_CHECK_CUDA(set_rows_slot(bank_out(ta.q), ta.q.x32, st));
  SV_CHECK_CUDA(set_rows_slot(bank_out(ta.r), ta.r.x32, st));
  bank_prepare_rows_kernel<<<(Nq + 7) / 8, 256, 0, st>>>(0, Nq, D, bank_out(ta.q));
  SV_CHECK_LAUNCH();
  // copy / scan sub-chunks of 8192 rows; banks beyond ~3 M rows use proportionally larger ones (<= ~400 sub-chunks)
  int sub_rows = 8192;
  while ((long long)sub_rows * 384 < Nr) sub_rows *= 2;
  HostFeed feed{r_host, cp, sub_rows};
  const KnnOut out{d2_out, reinterpret_cast<long long*>(idx_out), nullptr};
  return knn_driver(true, &ta, nullptr, Nq, Nr, D, k, row_offset, out, workspace, workspace_bytes, st, &feed);
}

extern "C" int segvlad_knn_debug_approx(const void* qbank, int Nq, const void* rbank, int Nr, int D, float* approx_out,
                                        float* bound_out, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  SV_REQUIRE(D > 0 && Nq > 0 && Nr > 0 && Nr <= kCandCap, "knn_debug_approx: need 0 < Nr <= %d", kCandCap);
  TcArgs ta;
  int rc = tc_setup(ta, qbank, Nq, rbank, Nr, D);
  if (rc) return rc;
  KnnLayout L = carve_knn(workspace, Nq, Nr);
  if (!workspace || workspace_bytes < L.total) {
    set_error("knn_debug_approx: workspace %zu < required %zu", workspace_bytes, L.total);
    return SEGVLAD_EWORKSPACE;
  }
  const ErrModel em{ta.q.meta, ta.r.stats, acc_err_const(D)};
  for (int q0 = 0; q0 < Nq; q0 += L.block_rows) {
    const int rows = (Nq - q0) < L.block_rows ? (Nq - q0) : L.block_rows;
    sel_init_kernel<<<(rows + 255) / 256, 256, 0, st>>>(L.sel[0], rows, Nr);
    SV_CHECK_LAUNCH();
    { const int frc = launch_filter(&ta, em, q0, rows, 0, Nr, 1, L.sel[0], st); if (frc) return frc; }
    SV_CHECK_LAUNCH();
    knn_debug_copy_kernel<<<rows, 256, 0, st>>>(L.sel[0], em, q0, rows, Nr, approx_out, bound_out);
    SV_CHECK_LAUNCH();
  }
  return SEGVLAD_OK;
}

extern "C" int segvlad_knn_simt(const float* q, int Nq, const float* r, int Nr, int64_t row_offset, int D, int k,
                                float* d2_out, int64_t* idx_out, void* workspace, size_t workspace_bytes,
                                void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  SimtArgs sa{q, r};
  const KnnOut out{d2_out, reinterpret_cast<long long*>(idx_out), nullptr};
  return knn_driver(false, nullptr, &sa, Nq, Nr, D, k, row_offset, out, workspace, workspace_bytes, st);
}

extern "C" int segvlad_merge_topk(const float* d2_parts, const int64_t* idx_parts, int G, int Nq, int k, float* d2_out,
                                  int64_t* idx_out, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  SV_REQUIRE(G > 0 && Nq >= 0 && k > 0, "merge_topk: bad shape");
  SV_REQUIRE((long long)G * k <= 16384, "merge_topk: G*k must be <= 16384 (got %lld)", (long long)G * k);
  if (Nq == 0) return SEGVLAD_OK;
  int P = 1;
  while (P < G * k) P <<= 1;
  const size_t smem = (size_t)P * 8;
  SV_CHECK_CUDA(cudaFuncSetAttribute(merge_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8));
  merge_topk_kernel<<<Nq, 256, smem, st>>>(d2_parts, reinterpret_cast<const long long*>(idx_parts), G, Nq, k, d2_out,
                                           reinterpret_cast<long long*>(idx_out));
  SV_CHECK_LAUNCH();
  return SEGVLAD_OK;
}

extern "C" int segvlad_merge_topk_packed(const uint64_t* parts, int G, size_t part_stride, int Nq, int k, float* d2_out,
                                         int64_t* idx_out, uint64_t* packed_out, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  SV_REQUIRE(G > 0 && Nq >= 0 && k > 0 && parts, "merge_topk_packed: bad arguments");
  SV_REQUIRE(part_stride >= (size_t)Nq * k, "merge_topk_packed: part_stride < Nq * k");
  SV_REQUIRE((d2_out != nullptr) == (idx_out != nullptr) && (d2_out || packed_out), "merge_topk_packed: no output buffer");
  SV_REQUIRE((long long)G * k <= 16384, "merge_topk_packed: G*k must be <= 16384 (got %lld)", (long long)G * k);
  if (Nq == 0) return SEGVLAD_OK;
  const size_t smem = (size_t)G * k * 8;
  DevCtx* dc = nullptr;
  { const int rc = dev_ctx(&dc); if (rc) return rc; }
  if (!dc->merge_attr_set) {
    SV_CHECK_CUDA(cudaFuncSetAttribute(merge_packed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8));
    dc->merge_attr_set = true;
  }
  merge_packed_kernel<<<Nq, 256, smem, st>>>(reinterpret_cast<const unsigned long long*>(parts), G, part_stride, Nq, k, d2_out,
                                            reinterpret_cast<long long*>(idx_out),
                                            reinterpret_cast<unsigned long long*>(packed_out));
  SV_CHECK_LAUNCH();
  return SEGVLAD_OK;
}
