// Masked residual aggregation on the 5th-generation tensor cores (sm_100a) -- SURVEY 8 rows a2/a3.
//
// Replaces the per-cluster `M'.double() @ res[inds]` + intra-normalisation + row normalisation of
// vlad_matmuls_per_cluster (func_vpr.py:1191-1205).  For one image and one cluster k the block of all segments is a
// small dense contraction
//       V_k [S x D] = M_k [S x n_k] . R_k [n_k x D]        M = 0/1 SuperSegment membership, R = fp32 residual rows
// with n_k ~ N/K tokens.  The SIMT kernel (aggregate.cu) evaluates it as dense masked FMAs for 8 segments per CTA and
// therefore re-reads every residual row S/8 times and is bound by the FMA pipe and its serial epilogue (r1 profile:
// 32 % of the HBM write roofline).  Here the contraction runs on tcgen05:
//   * R is split once into three bf16 planes (hi + mid + lo = all 24 fp32 mantissa bits; the 0/1 mask is exact in bf16,
//     so every product is exact) stored TRANSPOSED and label-sorted, RT[plane][image][d][i]: the tokens of a cluster
//     are then a contiguous K-major column range that TMA tiles straight into the swizzled operand layout;
//   * one CTA owns (image, tile of 128 segments, cluster): A = mask tile [128 x 64 tokens] written into shared memory
//     by two builder warps from the membership words, B = [128 channels x 64 tokens] x 3 planes by TMA, D = fp32
//     accumulators in TMEM (4 buffers of 128 columns);
//   * the block norm needs all D channels but TMEM holds 512 of them.  Single sweep (r2, default): the channel passes of
//     a block are split over J = ceil(P / 4) sibling CTAs, each keeps its <= 4 accumulators in TMEM, the contraction is
//     issued ONCE, the siblings exchange their partial sums of squares through a small global workspace (a posted value is
//     its own flag) and then scale and write the same accumulators (TMEM -> x scale -> fp64 -> staging box -> bulk tensor
//     store).  Clusters with more than 128 tokens, and SEGVLAD_AGG_RESIDENT=0, take the two-sweep schedule of r1: a norm
//     sweep (TMEM -> sum of squares, no stores) and a write sweep with the contraction issued twice (the operand re-read
//     comes from L2).  The tensor work is ~10 % of the write time either way;
//   * thread = segment row (TMEM lane), so norms and scales never leave the thread (the two column halves of a row meet in
//     shared memory once per item).
// Accumulation is fp32 in TMEM (truncating adds, measured ~4e-8..6e-8 relative per MMA).  The length of one accumulator
// chain is bounded (kTcSubChunks chunks = 128 tokens = 24 MMAs, <= 2.9e-6 relative): a cluster with more tokens -- a
// sky / road dominated image, the whole-image AnyLoc VLAD -- is accumulated as several chains in successive TMEM buffers
// whose partial sums the epilogue adds in registers (fp32 round-to-nearest), so the error does not grow with n_k and
// stays inside the 1e-5 descriptor tolerance for any vocabulary; the planes are accumulated small-to-large.
// Warp roles: warpgroup 0 = warp 0 TMA producer, 1 MMA issuer (+TMEM alloc), 2-3 mask-tile builders (72 registers);
// warpgroups 1-2 = epilogue, warps 4-7 and 8-11 (two warps per TMEM lane quarter, 64 channels of a pass each; 216 registers:
// a spill in the epilogue costs an L2 round trip under the kernel's own store traffic).
#include "aggregate_tc.cuh"

#include <stdlib.h>

#include <mutex>
#include <type_traits>

#include "tc_ptx.cuh"

namespace segvlad {

constexpr int kTcThreadsAgg = 384;   // warps: 0 TMA, 1 MMA, 2-3 mask builders | 4-7 epilogue (column half 0), 8-11 epilogue (half 1)
// Register budget per warpgroup (setmaxnreg): the epilogue warps must not spill -- the kernel's shared memory leaves the L1 a
// few KB, a spilled value comes back from L2, and under the kernel's own store traffic that round trip costs thousands of
// cycles (r2: ~50 reloads per item in ~10 stall points cost more than the second sweep they were meant to save).
constexpr int kTcRegsFront = 72, kTcRegsEpi = 216;   // per scheduler: 72 + 2 x 216 = 504 = 3 x 168, the CTA's own allocation (the pool)
constexpr int kTcStages = 2;
constexpr uint32_t kTcTileBytes = kTcSegTile * kTcTokChunk * 2;          // 16 KB: one [128 x 64] bf16 operand tile
constexpr uint32_t kTcStageBytes = 4 * kTcTileBytes;                     // A + 3 B planes
constexpr int kTcBufs = 512 / kTcPassN;                                  // TMEM accumulator buffers
constexpr double kEpsTc = 1e-12;

constexpr uint32_t kTcBoxBytes = 32 * 128;                               // output staging box: 32 rows x 128 bytes
constexpr uint32_t kTcStageOutBytes = 8 * 2 * kTcBoxBytes;               // 8 epilogue warps x 2 boxes
constexpr int kTcTblTiles = 256;                                         // item tables cached in shared memory when they fit:
constexpr int kTcTblCl = 2048;                                           // [tiles][4] and [B][K + 1] ints (12 KB)
constexpr int kTcEarly = 6;                                              // sibling exchange prefetched into shared memory for J <= 6
__host__ __device__ constexpr size_t agg_tc_smem() {
  return 1024 + (size_t)kTcStages * kTcStageBytes + kTcStageOutBytes + 256 + 2 * 2 * kTcSegTile * 8 +
         4 * (4 * kTcTblTiles + kTcTblCl) + 2 * kTcEarly * kTcSegTile * 8 + 2 * 256 * 4;
}

// ------------------------------------------------------------------------------------------------
// R [B][N][D] fp32 -> RT [3][B][D][Np] bf16 (lo, mid, hi), columns in label-sorted order, zero padded to Np.
__global__ void __launch_bounds__(256)
rt_planes_kernel(const float* __restrict__ R, const int* __restrict__ cl_tok, int B, int N, int D, int Np,
                 __nv_bfloat16* __restrict__ RT) {
  __shared__ float tile[32][65];
  const int i0 = blockIdx.x * 64, d0 = blockIdx.y * 32, b = blockIdx.z;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int t = w * 8 + r, i = i0 + t;
    float v = 0.f;
    if (i < N && d0 + lane < D) {
      const int n = cl_tok[(size_t)b * N + i];
      v = R[((size_t)b * N + n) * D + d0 + lane];
    }
    tile[lane][t] = v;
  }
  __syncthreads();
  const size_t plane = (size_t)B * D * Np;
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) {
    const int dd = w * 4 + rr, d = d0 + dd;
    if (d >= D) continue;
    const float x0 = tile[dd][2 * lane], x1 = tile[dd][2 * lane + 1];
    const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
    const float r0 = x0 - __bfloat162float(h0), r1 = x1 - __bfloat162float(h1);
    const __nv_bfloat16 m0 = __float2bfloat16_rn(r0), m1 = __float2bfloat16_rn(r1);
    const __nv_bfloat16 l0 = __float2bfloat16_rn(r0 - __bfloat162float(m0)), l1 = __float2bfloat16_rn(r1 - __bfloat162float(m1));
    const size_t o = ((size_t)b * D + d) * Np + i0 + 2 * lane;
    *reinterpret_cast<__nv_bfloat162*>(RT + o) = __halves2bfloat162(l0, l1);
    *reinterpret_cast<__nv_bfloat162*>(RT + plane + o) = __halves2bfloat162(m0, m1);
    *reinterpret_cast<__nv_bfloat162*>(RT + 2 * plane + o) = __halves2bfloat162(h0, h1);
  }
}

// ------------------------------------------------------------------------------------------------
// Tokens in the reference's [D][N] layout -> RT directly: r = x / ||x|| - c[label] (the arithmetic of residual_dn_kernel,
// func_vpr.py:1085, 1151), label-sorted and split into the three planes, with no token-major residual matrix in between
// (r1 launch list: residual_dn 151 us + rt_planes 66 us per 16-image batch = two transposes of 150 MB; this kernel reads
// the tokens once, coalesced, and writes RT once).  CTA = (image, CH channels): the channel rows are staged in shared
// memory, gathered by cl_tok, and every plane row is written as contiguous 128-byte segments.
__global__ void __launch_bounds__(256)
rt_from_tokens_kernel(const float* __restrict__ tokens, const float* __restrict__ centers, const int* __restrict__ labels,
                      const float* __restrict__ nrm, const int* __restrict__ cl_tok, int B, int N, int D, int K, int Np,
                      int CH, __nv_bfloat16* __restrict__ RT) {
  extern __shared__ __align__(16) float s_rt[];
  float* s_x = s_rt;                                    // [CH][N] channel rows
  float* s_c = s_rt + (((size_t)CH * N + 3) & ~(size_t)3);   // [K][CH] centre slice
  const int b = blockIdx.y, d0 = blockIdx.x * CH, tid = threadIdx.x;
  const int nch = min(CH, D - d0);
  const float* tok = tokens + ((size_t)b * D + d0) * N;  // rows d0 .. d0+nch-1 are contiguous: nch * N floats
  const int tot = nch * N;
  if ((reinterpret_cast<uintptr_t>(tok) & 15) == 0 && (tot & 3) == 0) {
    for (int i = tid; i < tot / 4; i += 256) reinterpret_cast<float4*>(s_x)[i] = __ldg(reinterpret_cast<const float4*>(tok) + i);
  } else {
    for (int i = tid; i < tot; i += 256) s_x[i] = __ldg(tok + i);
  }
  for (int i = tid; i < K * nch; i += 256) {
    const int k = i / nch, r = i - k * nch;
    s_c[k * CH + r] = __ldg(centers + (size_t)k * D + d0 + r);
  }
  __syncthreads();
  const size_t plane = (size_t)B * D * Np;
  for (int pr = tid; pr < Np / 2; pr += 256) {
    const int i = 2 * pr;
    int n0 = -1, n1 = -1, l0 = 0, l1 = 0;
    float q0 = 1.f, q1 = 1.f;
    if (i < N) { n0 = cl_tok[(size_t)b * N + i]; q0 = nrm[(size_t)b * N + n0]; l0 = labels[(size_t)b * N + n0]; }
    if (i + 1 < N) { n1 = cl_tok[(size_t)b * N + i + 1]; q1 = nrm[(size_t)b * N + n1]; l1 = labels[(size_t)b * N + n1]; }
    for (int r = 0; r < nch; ++r) {
      const float x0 = n0 >= 0 ? s_x[r * N + n0] / q0 - s_c[l0 * CH + r] : 0.f;
      const float x1 = n1 >= 0 ? s_x[r * N + n1] / q1 - s_c[l1 * CH + r] : 0.f;
      const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
      const float r0 = x0 - __bfloat162float(h0), r1 = x1 - __bfloat162float(h1);
      const __nv_bfloat16 m0 = __float2bfloat16_rn(r0), m1 = __float2bfloat16_rn(r1);
      const __nv_bfloat16 e0 = __float2bfloat16_rn(r0 - __bfloat162float(m0)), e1 = __float2bfloat16_rn(r1 - __bfloat162float(m1));
      const size_t o = ((size_t)b * D + d0 + r) * Np + i;
      *reinterpret_cast<__nv_bfloat162*>(RT + o) = __halves2bfloat162(e0, e1);
      *reinterpret_cast<__nv_bfloat162*>(RT + plane + o) = __halves2bfloat162(m0, m1);
      *reinterpret_cast<__nv_bfloat162*>(RT + 2 * plane + o) = __halves2bfloat162(h0, h1);
    }
  }
}

// channels per CTA of rt_from_tokens_kernel (0: the rows do not fit in shared memory -> two-kernel path)
int agg_tc_fused_channels(int N, int K) {
  const char* e = getenv("SEGVLAD_RT_CH");   // development: upper limit of the channels per CTA (measured 2/4/8/16: 125/101/107/135 us)
  const int top = e ? atoi(e) : 4;
  for (int ch = top; ch >= 1; ch >>= 1)
    if (((size_t)ch * N + 4 + (size_t)K * ch) * sizeof(float) <= 100 * 1024) return ch;
  return 0;
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// mbarrier wait that adds the cycles spent to a counter when the development probe is on
// per-pass timeline of CTA 0 (development probe): probe[4096 + 8 * pass + j], j = 0 MMA: buffer free, 1 MMA: first operand
// stage full, 2 MMA: pass committed, 3 producer: first stage of the pass issued, 4 epilogue (warp 2): accumulator full,
// 5 epilogue: buffer handed back; single-sweep items: 4 / 5 = front half (sum of squares), 6 / 7 = write pass begin / end
#define TC_MARK(ti_, j_) do { if (kProbe && probe && blockIdx.x == 0 && (ti_) < 256) probe[4096 + 8 * (ti_) + (j_)] = clock64(); } while (0)
#define TC_TIMED_WAIT(bar, par, acc) do { if (kProbe && probe) { const long long _t = clock64(); mbar_wait(bar, par); acc += clock64() - _t; } else mbar_wait(bar, par); } while (0)

struct TcItem {
  int b, g0, s0, ns, k, p0, p1, rows, x0, nch, pj0, pj1;   // x0: p0 rounded down to 8 tokens (TMA needs 16-byte
  int res;                                                  // aligned global addresses); tokens < p0 are masked out
};                                                          // res: single-sweep item (accumulators stay in TMEM), below
// item id -> (segment tile, cluster, channel split); identical in every warp role
__device__ __forceinline__ TcItem tc_item(int id, const int* __restrict__ tile_tbl, const int* __restrict__ cl_ptr, int K,
                                          int J, int P, int resident) {
  TcItem it;
  const int j = id % J, rest = id / J;
  it.k = rest % K;
  const int tt = rest / K;
  const int4 t = *reinterpret_cast<const int4*>(tile_tbl + 4 * tt);
  it.b = t.x; it.g0 = t.y; it.s0 = t.z; it.ns = t.w;
  it.p0 = cl_ptr[(size_t)it.b * (K + 1) + it.k];
  it.p1 = cl_ptr[(size_t)it.b * (K + 1) + it.k + 1];
  it.rows = it.p1 - it.p0;
  it.x0 = it.p0 & ~7;
  it.nch = (it.p1 - it.x0 + kTcTokChunk - 1) / kTcTokChunk;
  it.pj0 = (int)((long long)j * P / J);
  it.pj1 = (int)((long long)(j + 1) * P / J);
  // Single-sweep ("resident") item: this CTA's channel passes (<= kTcBufs of them) each own a TMEM buffer, the contraction is
  // issued ONCE, the epilogue takes the sums of squares from TMEM, exchanges them with the J - 1 sibling CTAs that hold the
  // other channels of the same block, and then scales and writes the very same accumulators.  Needs one accumulator chain
  // per pass (n_k <= kTcSubChunks chunks); longer clusters take the two-sweep path.  Identical in every role and sibling.
  it.res = resident && it.nch <= kTcSubChunks && (P + J - 1) / J <= kTcBufs;
  return it;
}

// tcgen05.ld split into issue + wait so that the next 32 columns load while the current ones are processed.  The wait
// takes the destination registers as in/out operands: the compiler must not consume them before the wait.
__device__ __forceinline__ void tc_ld32_issue(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_ld_wait(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                 "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                 "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
               :
               : "memory");
}
// partial sum of a further accumulator chain: acc += TMEM piece (fp32 round-to-nearest adds in registers); 8 columns at a
// time -- this is the rare long-cluster path and must not cost the common path registers
__device__ __forceinline__ void tc_acc32(uint32_t taddr, uint32_t (&acc)[32]) {
#pragma unroll
  for (int j0 = 0; j0 < 32; j0 += 8) {
    uint32_t w0, w1, w2, w3, w4, w5, w6, w7;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3), "=r"(w4), "=r"(w5), "=r"(w6), "=r"(w7)
        : "r"(taddr + j0)
        : "memory");
    const uint32_t w[8] = {w0, w1, w2, w3, w4, w5, w6, w7};
#pragma unroll
    for (int j = 0; j < 8; ++j)
      acc[j0 + j] = __float_as_uint(__fadd_rn(__uint_as_float(acc[j0 + j]), __uint_as_float(w[j])));
  }
}
__device__ __forceinline__ float tc_sumsq(const uint32_t (&v)[32], int nb) {
  float f0 = 0.f, f1 = 0.f, f2 = 0.f, f3 = 0.f;
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    if (j < nb) {
      const float a = __uint_as_float(v[j]), b = __uint_as_float(v[j + 1]);
      const float c = __uint_as_float(v[j + 2]), d = __uint_as_float(v[j + 3]);
      f0 = fmaf(a, a, f0); f1 = fmaf(b, b, f1); f2 = fmaf(c, c, f2); f3 = fmaf(d, d, f3);
    }
  }
  return (f0 + f1) + (f2 + f3);
}

// Scaled accumulator piece (32 consecutive channels of one block row) -> global memory with 256-bit stores: every
// lane writes whole 32-byte sectors of its own row, so the row-per-lane pattern costs no write amplification.
template <typename OutT> struct TcOut;
template <> struct TcOut<double> {
  static __device__ __forceinline__ void store(double* p, const uint32_t (&v)[32], double sc, int nb) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      if (j < nb)
        asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p + j), "d"((double)__uint_as_float(v[j]) * sc),
                     "d"((double)__uint_as_float(v[j + 1]) * sc), "d"((double)__uint_as_float(v[j + 2]) * sc),
                     "d"((double)__uint_as_float(v[j + 3]) * sc) : "memory");
    }
  }
  static __device__ __forceinline__ void zero16(double* p) {
#pragma unroll
    for (int j = 0; j < 16; j += 4)
      asm volatile("st.global.v4.f64 [%0], {%1, %1, %1, %1};" ::"l"(p + j), "d"(0.0) : "memory");
  }
};
template <> struct TcOut<float> {
  static __device__ __forceinline__ void store(float* p, const uint32_t (&v)[32], double sc, int nb) {
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      if (j < nb)
        asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p + j),
                     "f"((float)((double)__uint_as_float(v[j]) * sc)), "f"((float)((double)__uint_as_float(v[j + 1]) * sc)),
                     "f"((float)((double)__uint_as_float(v[j + 2]) * sc)), "f"((float)((double)__uint_as_float(v[j + 3]) * sc)),
                     "f"((float)((double)__uint_as_float(v[j + 4]) * sc)), "f"((float)((double)__uint_as_float(v[j + 5]) * sc)),
                     "f"((float)((double)__uint_as_float(v[j + 6]) * sc)), "f"((float)((double)__uint_as_float(v[j + 7]) * sc))
                     : "memory");
    }
  }
  static __device__ __forceinline__ void zero16(float* p) {
#pragma unroll
    for (int j = 0; j < 16; j += 8)
      asm volatile("st.global.v8.f32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"l"(p + j), "f"(0.f) : "memory");
  }
};

// Output through the TMA instead of the LSU: the scaled values of 32 rows x 128 bytes go to a SWIZZLE_128B staging box
// (row = lane, 16-byte chunk c stored at c ^ (row & 7): conflict-free for the row-per-lane pattern) and ONE
// cp.async.bulk.tensor store moves the box.  r1 ncu: with direct 256-bit stores the L1 data pipe (64 bytes per wavefront
// for the row-per-lane pattern) was the busiest unit of the kernel, 100 % busy during the write sweep.
template <typename OutT> struct TcStage;
template <> struct TcStage<double> {
  static constexpr int kCols = 16;
  template <int kOff>
  static __device__ __forceinline__ void put(uint32_t sb, int lane, const uint32_t (&v)[32], double sc) {
#pragma unroll
    for (int c = 0; c < 8; ++c)
      asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(sb + lane * 128 + ((c ^ (lane & 7)) << 4)),
                   "d"((double)__uint_as_float(v[kOff + 2 * c]) * sc), "d"((double)__uint_as_float(v[kOff + 2 * c + 1]) * sc)
                   : "memory");
  }
};
template <> struct TcStage<float> {
  static constexpr int kCols = 32;
  template <int kOff>
  static __device__ __forceinline__ void put(uint32_t sb, int lane, const uint32_t (&v)[32], double sc) {
#pragma unroll
    for (int c = 0; c < 8; ++c)
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sb + lane * 128 + ((c ^ (lane & 7)) << 4)),
                   "f"((float)((double)__uint_as_float(v[4 * c]) * sc)), "f"((float)((double)__uint_as_float(v[4 * c + 1]) * sc)),
                   "f"((float)((double)__uint_as_float(v[4 * c + 2]) * sc)), "f"((float)((double)__uint_as_float(v[4 * c + 3]) * sc))
                   : "memory");
  }
};
// ---- PCA-planes output (row f1): the projection's A operand, written by the aggregation itself -------------------------
// x = v - mean split into three bf16 pieces (hi + mid + lo = all 24 mantissa bits), packed pairs.
// all three planes of a pair in one go
__device__ __forceinline__ void tc_plane_pair3(float x0, float x1, uint32_t& lo, uint32_t& mid, uint32_t& hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
  const float2 hf = __bfloat1622float2(h);
  const float r0 = x0 - hf.x, r1 = x1 - hf.y;
  const __nv_bfloat162 m = __floats2bfloat162_rn(r0, r1);
  const float2 mf = __bfloat1622float2(m);
  const __nv_bfloat162 l = __floats2bfloat162_rn(r0 - mf.x, r1 - mf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  mid = *reinterpret_cast<const uint32_t*>(&m);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
template <int kN>
__device__ __forceinline__ void tc_plane_chunk3(const uint32_t (&v)[kN], int off, uint32_t (&o0)[4], uint32_t (&o1)[4],
                                                uint32_t (&o2)[4]) {
#pragma unroll
  for (int h = 0; h < 4; ++h)
    tc_plane_pair3(__uint_as_float(v[off + 2 * h]), __uint_as_float(v[off + 2 * h + 1]), o0[h], o1[h], o2[h]);
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int x, int y, uint64_t policy) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
               ::"l"(map), "r"(src), "r"(x), "r"(y), "l"(policy) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// kProbe: development build of the kernel with the per-role cycle counters (tools/agg_tc_probe.py); the product
// instantiation carries none of them (they cost the epilogue registers)
// kPlanes (OutT = __nv_bfloat16): instead of the descriptor block the write sweep emits (block - pca_mean) split into three
// bf16 planes, out = [3][S_total][K * D] -- the K-major A operand of the tensor-core PCA projection (project_tc.cu), which
// then needs no converter warps and never reads an fp64 [S, K * D] matrix.
template <typename OutT, bool kProbe, bool kPlanes>
__global__ void __launch_bounds__(kTcThreadsAgg, 1)
aggregate_tc_kernel(const __grid_constant__ CUtensorMap map_rt, const __grid_constant__ CUtensorMap map_out,
                    const int* __restrict__ tile_tbl_g,
                    const int* __restrict__ cl_ptr_g, const uint16_t* __restrict__ memS, const int* __restrict__ cpred,
                    int B, int N, int D, int K, int n_items, int J, OutT* __restrict__ out, double* __restrict__ norms,
                    unsigned long long* __restrict__ probe, int store_hint, const float* __restrict__ pca_mean,
                    int S_total, int resident, double* xnorm) {
  extern __shared__ __align__(1024) uint8_t tc_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* stage_out = smem + kTcStages * kTcStageBytes;   // [8 epilogue warps][2 boxes][4 KB], 1024-byte aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_out + kTcStageOutBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  double* s_ssq = reinterpret_cast<double*>(bars + 32);   // [2 item parities][2 column halves][128 rows] partial sums of squares
  // Item tables in shared memory: every role looks an item up by two dependent loads, and under the kernel's own store
  // traffic a round trip to L2 costs thousands of cycles -- per item, in every role's serial path (r2: ~10 k cycles per item)
  int* s_tbl = reinterpret_cast<int*>(s_ssq + 2 * 2 * kTcSegTile);
  double* s_x = reinterpret_cast<double*>(s_tbl + 4 * kTcTblTiles + kTcTblCl);   // [2 parities][kTcEarly siblings][128 rows]
  int* s_cp = reinterpret_cast<int*>(s_x + 2 * kTcEarly * kTcSegTile);            // [2 parities][256 epilogue threads]
  const int n_tiles = n_items / (K * J);
  const bool tbl_fits = n_tiles <= kTcTblTiles && B * (K + 1) <= kTcTblCl;
  if (tbl_fits) {
    for (int i = threadIdx.x; i < 4 * n_tiles; i += blockDim.x) s_tbl[i] = tile_tbl_g[i];
    for (int i = threadIdx.x; i < B * (K + 1); i += blockDim.x) s_tbl[4 * kTcTblTiles + i] = cl_ptr_g[i];
  }
  const int* tile_tbl = tbl_fits ? s_tbl : tile_tbl_g;
  const int* cl_ptr = tbl_fits ? s_tbl + 4 * kTcTblTiles : cl_ptr_g;
  const uint32_t bar_full = smem_u32(bars + 0);      // [kTcStages] A written (2 builder warps) + B landed (TMA)
  const uint32_t bar_empty = smem_u32(bars + 4);     // [kTcStages] MMAs that read the stage retired
  const uint32_t bar_tfull = smem_u32(bars + 8);     // [kTcBufs]   accumulator pass complete
  const uint32_t bar_tempty = smem_u32(bars + 12);   // [kTcBufs]   accumulator drained by the 4 epilogue warps
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int P = (D + kTcPassN - 1) / kTcPassN;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kTcStages; ++i) { mbar_init(bar_full + 8 * i, 3); mbar_init(bar_empty + 8 * i, 1); }
    for (int i = 0; i < kTcBufs; ++i) { mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kTcRegsFront));
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_rt) : "memory");
      uint32_t stage = 0, phase = 0, tp = 0;
      long long t_wait = 0;
      for (int id = blockIdx.x; id < n_items; id += gridDim.x) {
        const TcItem it = tc_item(id, tile_tbl, cl_ptr, K, J, P, resident);
        if (it.rows == 0) continue;
        const int n_inst = (it.res ? 0 : P) + (it.pj1 - it.pj0);
        for (int inst = 0; inst < n_inst; ++inst, ++tp) {
          const int pass = it.res ? it.pj0 + inst : (inst < P ? inst : it.pj0 + inst - P);
          // (r1: an L2 prefetch of the operand rows 1-3 passes ahead -- the norm sweep is their first touch, ~2.9 k cycles
          // per box from DRAM -- shortens the MMA's operand waits but costs more TMA time than it saves: 0.418 -> 0.434 /
          // 0.442 / 0.447 ms for a look-ahead of 1 / 2 / 3 passes; profiles/r1_agg_experiments.txt)
          for (int c = 0; c < it.nch; ++c) {
            TC_TIMED_WAIT(bar_empty + 8 * stage, phase ^ 1, t_wait);
            if (c == 0) TC_MARK(tp, 3);
            const uint32_t sb = smem_u32(smem + stage * kTcStageBytes) + kTcTileBytes;
            const uint32_t fb = bar_full + 8 * stage;
            mbar_arrive_expect_tx(fb, 3 * kTcTileBytes);
#pragma unroll
            for (int pl = 0; pl < 3; ++pl)
              tma_load_2d(sb + pl * kTcTileBytes, &map_rt, fb, it.x0 + c * kTcTokChunk, (pl * B + it.b) * D + pass * kTcPassN);
            if (++stage == kTcStages) { stage = 0; phase ^= 1; }
          }
        }
      }
      if (kProbe && probe) probe[blockIdx.x * 16 + 9] = t_wait;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, ti = 0;
      long long t_tempty = 0, t_full = 0;
      const long long t_begin = clock64();
      for (int id = blockIdx.x; id < n_items; id += gridDim.x) {
        const TcItem it = tc_item(id, tile_tbl, cl_ptr, K, J, P, resident);
        if (it.rows == 0) continue;
        const int n_inst = (it.res ? 0 : P) + (it.pj1 - it.pj0);
        for (int inst = 0; inst < n_inst; ++inst) {
          const int pass = it.res ? it.pj0 + inst : (inst < P ? inst : it.pj0 + inst - P);
          const int width = min(kTcPassN, D - pass * kTcPassN);
          // kind::f16: D=f32 (bit 4), A=B=bf16 (bits 7, 10), both K-major, N>>3 @17, M>>4 @24
          const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(width >> 3) << 17) |
                                 ((uint32_t)(kTcSegTile >> 4) << 24);
          // one accumulator chain covers at most kTcSubChunks token chunks (header: bounded truncation); a longer cluster
          // continues in the next TMEM buffer and the epilogue adds the partial sums
          uint32_t buf = 0, d_tmem = 0;
          for (int c = 0; c < it.nch; ++c) {
            const int cs = c % kTcSubChunks;
            if (cs == 0) {
              buf = ti % kTcBufs;
              TC_TIMED_WAIT(bar_tempty + 8 * buf, ((ti / kTcBufs) & 1) ^ 1, t_tempty);
              tc_fence_after();
              if (c == 0) TC_MARK(ti, 0);
              d_tmem = tmem_base + buf * kTcPassN;
            }
            TC_TIMED_WAIT(bar_full + 8 * stage, phase, t_full);
            tc_fence_after();
            if (c == 0) TC_MARK(ti, 1);
            const uint32_t sa = smem_u32(smem + stage * kTcStageBytes);
            // descriptors of the stage's tiles differ only in the (address >> 4) field: one encode, constant offsets
            // (r1 timeline: the issuing thread, not the tensor pipe, paced the norm sweep at ~130 cycles per MMA)
            const uint64_t adesc = umma_desc_sw128(sa);
            const uint64_t bdesc = adesc + (kTcTileBytes >> 4);
            const int nk = (min(kTcTokChunk, it.p1 - it.x0 - c * kTcTokChunk) + 15) >> 4;
            for (int kk = 0; kk < nk; ++kk) {
              const uint64_t adv = (uint64_t)(kk * 2);                  // 16 tokens = 32 bytes along K
              // lo, mid, hi: small terms first
              tc_mma_bf16(d_tmem, adesc + adv, bdesc + adv, idesc, (cs | kk) != 0);
              tc_mma_bf16(d_tmem, adesc + adv, bdesc + adv + (kTcTileBytes >> 4), idesc, 1u);
              tc_mma_bf16(d_tmem, adesc + adv, bdesc + adv + 2 * (kTcTileBytes >> 4), idesc, 1u);
            }
            tc_commit(bar_empty + 8 * stage);
            if (++stage == kTcStages) { stage = 0; phase ^= 1; }
            if (cs == kTcSubChunks - 1 || c == it.nch - 1) {   // this chain is complete
              tc_commit(bar_tfull + 8 * buf);
              TC_MARK(ti, 2);
              ++ti;
            }
          }
        }
      }
      if (kProbe && probe) { probe[blockIdx.x * 16 + 6] = t_tempty; probe[blockIdx.x * 16 + 7] = t_full;
                   probe[blockIdx.x * 16 + 8] = clock64() - t_begin; }
    }
  } else {
    // ===================== mask-tile builders (64 threads) =====================
    const int u = threadIdx.x - 64;
    uint32_t stage = 0, phase = 0;
    long long t_bwait = 0;
    const long long t_bbegin = clock64();
    for (int id = blockIdx.x; id < n_items; id += gridDim.x) {
      const TcItem it = tc_item(id, tile_tbl, cl_ptr, K, J, P, resident);
      if (it.rows == 0) continue;
      const int ngrp = (it.ns + 7) >> 3;
      const int n_inst = (it.res ? 0 : P) + (it.pj1 - it.pj0);
      uint32_t wlo[2][4];   // cached membership words of this thread's two (group, 8-token) cells: 8 x u16 each
      int cached_c = -1;
      for (int inst = 0; inst < n_inst; ++inst) {
        for (int c = 0; c < it.nch; ++c) {
          if (c != cached_c) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
              const int q = u + 64 * r, g = q >> 3, c8 = q & 7;
              const int i0 = it.x0 + c * kTcTokChunk + c8 * 8;
              const uint16_t* src = memS + (size_t)(it.g0 + g) * N;
#pragma unroll
              for (int t = 0; t < 8; t += 2) {
                uint32_t a = 0, bb = 0;
                if (g < ngrp) {
                  if (i0 + t >= it.p0 && i0 + t < it.p1) a = src[i0 + t];
                  if (i0 + t + 1 >= it.p0 && i0 + t + 1 < it.p1) bb = src[i0 + t + 1];
                }
                wlo[r][t >> 1] = a | (bb << 16);
              }
            }
            cached_c = c;
          }
          TC_TIMED_WAIT(bar_empty + 8 * stage, phase ^ 1, t_bwait);
          const uint32_t sa = smem_u32(smem + stage * kTcStageBytes);
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const int q = u + 64 * r, g = q >> 3, c8 = q & 7;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              // row = 8 g + j of the SWIZZLE_128B K-major tile: 128-byte row, 16-byte chunk c8 stored at c8 ^ (row & 7)
              // bit j of both 16-bit membership words -> two bf16 1.0 / 0.0 in one shift, mask and multiply
              // (r1 ncu: the select-per-bit form made the two builder warps the critical path of the whole kernel)
              uint32_t o[4];
#pragma unroll
              for (int h = 0; h < 4; ++h) o[h] = ((wlo[r][h] >> j) & 0x00010001u) * 0x3F80u;
              const uint32_t addr = sa + g * 1024 + j * 128 + ((c8 ^ j) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
            }
          }
          tc_fence_async_smem();     // generic-proxy writes -> visible to the tensor core's async-proxy reads
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_full + 8 * stage);
          if (++stage == kTcStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    if (kProbe && probe && threadIdx.x == 64) { probe[blockIdx.x * 16 + 10] = t_bwait; probe[blockIdx.x * 16 + 11] = clock64() - t_bbegin; }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kTcRegsEpi));
    // ===================== epilogue warps: thread = segment row (TMEM lane) =====================
    // Two warps per TMEM lane quarter, each owning 64 of a pass's 128 channels: one warp per scheduler could not hide
    // its own tcgen05.ld -> convert -> store latencies (r1 ncu: 12.5 % warps active, issue slots 19 % busy at 62 % of the
    // HBM write roofline).  The two halves of a row exchange their partial sums of squares through shared memory.
    const int quarter = warp & 3, half = warp >= 8 ? 1 : 0;
    const int row = quarter * 32 + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t my_stage = smem_u32(stage_out) + (uint32_t)(half * 4 + quarter) * 2 * kTcBoxBytes;
    uint64_t pol_out = 0x1000000000000000ull;   // evict_normal
    if (store_hint) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_out));
    const bool tma_dims = (D % 64) == 0;      // every pass piece of this warp is 0 or 32 columns wide
    uint32_t ti = 0, n_done = 0, nbox = 0;
    long long t_nwait = 0, t_bar = 0, t_wwait = 0, t_store = 0, t_xwait = 0;
    const long long t_ebegin = clock64();
    // Look-ahead over the sibling exchange (single-sweep items whose passes fill at most half of the TMEM): the partial sums
    // of item i + 1 are posted BEFORE item i is written, so the wait for the siblings -- and the next item's MMAs -- hide
    // behind a whole write sweep.  A CTA holds at most one posted-but-unwritten ("pending") item.
    const bool la = resident && J > 1 && 2 * ((P + J - 1) / J) <= kTcBufs;
    int pend_id = -1, pend_par = 0;
    uint32_t pend_ti0 = 0;
    TcItem pend_it = {};
    const int e = half * kTcSegTile + row;                           // epilogue thread index
    for (int id = blockIdx.x; id < n_items || pend_id >= 0;) {
      const bool have = id < n_items;
      TcItem cur = {};
      if (have) cur = tc_item(id, tile_tbl, cl_ptr, K, J, P, resident);
      // a two-sweep item cycles through every TMEM buffer: the pending item is written first (front = false: flush only)
      const bool front = have && !(cur.rows != 0 && pend_id >= 0 && !(cur.res && la));
      double ssq = 0.0;
      const uint32_t cur_ti0 = ti;                                   // single-sweep item: first of its TMEM buffers
      // The pending item's exchange slots (posted a write sweep ago) and this item's cluster counts are fetched into shared
      // memory by cp.async while the front half runs: no L2 round trip (~1.5 k cycles under the kernel's own store traffic)
      // sits in the epilogue's serial path, and none of it is held in registers (r2: register prefetch spilled, each load
      // was waited for at once: 6 round trips per item).
      const bool bfront = front && cur.rows != 0;                    // a front half with its barrier
      const int par = n_done & 1;
      bool early = false;
      if (bfront) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(s_cp + par * 256 + e)),
                     "l"(cpred + cur.s0 + (row < cur.ns ? row : 0)) : "memory");
        if (pend_id >= 0 && J <= kTcEarly) {
          early = true;
          const double* xp = xnorm + (size_t)(pend_id - pend_id % J) * kTcSegTile;
          for (int i = e; i < J * (kTcSegTile / 2); i += 256)       // 16-byte pieces, L2 only (the slots change during the kernel)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(s_x + par * kTcEarly * kTcSegTile + 2 * i)),
                         "l"(xp + 2 * i) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
      if (front) {
        if (cur.rows == 0) {
          const bool valid = row < cur.ns;
          const int s = cur.s0 + row;
          OutT* orow = out + ((size_t)(valid ? s : cur.s0) * K + cur.k) * D;
          // empty cluster: the block is zero for every segment
          if (valid) {
            if (cur.pj0 == 0 && half == 0) norms[(size_t)s * K + cur.k] = 0.0;
            const int d_beg = cur.pj0 * kTcPassN, d_end = min(D, cur.pj1 * kTcPassN);
            if constexpr (kPlanes) {
              // zero block: the planes hold -mean
              const size_t KD = (size_t)K * D;
              for (int d = d_beg + 8 * half; d < d_end; d += 16) {
                uint32_t xv[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) xv[j] = __float_as_uint(0.f - __ldg(pca_mean + (size_t)cur.k * D + d + j));
                uint32_t o2[4], o1[4], o0[4];
                tc_plane_chunk3(xv, 0, o0, o1, o2);
                OutT* p0 = out + (size_t)s * KD + (size_t)cur.k * D + d;
                *reinterpret_cast<uint4*>(p0) = make_uint4(o0[0], o0[1], o0[2], o0[3]);
                *reinterpret_cast<uint4*>(p0 + (size_t)S_total * KD) = make_uint4(o1[0], o1[1], o1[2], o1[3]);
                *reinterpret_cast<uint4*>(p0 + 2 * (size_t)S_total * KD) = make_uint4(o2[0], o2[1], o2[2], o2[3]);
              }
            } else {
              for (int d = d_beg + 16 * half; d < d_end; d += 32) TcOut<OutT>::zero16(orow + d);
            }
          }
          id += gridDim.x;
          continue;
        }
        // ---- norm sweep: sum of squares of this warp's half of the block row (fp32 products, fp64 accumulation) ----
        const int nsub = (cur.nch + kTcSubChunks - 1) / kTcSubChunks;   // accumulator chains per pass (1 unless n_k > ~128)
        for (int pass = cur.res ? cur.pj0 : 0; pass < (cur.res ? cur.pj1 : P); ++pass) {
          const int width = min(kTcPassN, D - pass * kTcPassN);
          const int c0 = half * 64;
          const int w0 = min(32, width - c0), w1 = min(32, width - c0 - 32);     // columns in this warp's two pieces
          uint32_t va[32], vb[32];
          for (int sub = 0; sub < nsub; ++sub, ++ti) {
            const uint32_t buf = ti % kTcBufs, use = ti / kTcBufs;
            TC_TIMED_WAIT(bar_tfull + 8 * buf, use & 1, t_nwait);
            tc_fence_after();
            if (warp == 4 && lane == 0) TC_MARK(ti, 4);
            const uint32_t tcol = tlane + buf * kTcPassN + c0;
            if (w0 > 0) {                                  // (warp-uniform)
              if (sub == 0) {
                tc_ld32_issue(tcol, va);
                if (w1 > 0) tc_ld32_issue(tcol + 32, vb);
                tc_ld_wait(va);
                if (w1 > 0) tc_ld_wait(vb);
              } else {
                tc_acc32(tcol, va);
                if (w1 > 0) tc_acc32(tcol + 32, vb);
              }
            }
            if (!cur.res) {                                  // (resident: the buffer keeps the accumulator for the write sweep)
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
            }
            if (warp == 4 && lane == 0) TC_MARK(ti, 5);
          }
          if (w0 > 0) {
            ssq += (double)tc_sumsq(va, w0);
            if (w1 > 0) ssq += (double)tc_sumsq(vb, w1);
          }
        }
        double* xs = s_ssq + (n_done & 1) * 2 * kTcSegTile;
        xs[half * kTcSegTile + row] = ssq;
        asm volatile("cp.async.wait_group 0;" ::: "memory");      // this thread's prefetches: visible to all after the barrier
        { const long long _t = (kProbe && probe) ? clock64() : 0;
          asm volatile("bar.sync 1, 256;" ::: "memory");     // the 8 epilogue warps
          if (kProbe && probe) t_bar += clock64() - _t; }
        ssq = xs[row] + xs[kTcSegTile + row];
        ++n_done;
        if (cur.res && J > 1) {
          // post this CTA's partial sums for the J - 1 sibling CTAs that hold the block's other channels: the 8-byte value is
          // its own flag (the slots are preset to all-ones, which no sum of squares can be), so no fence and no counter --
          // under the kernel's store traffic a fence costs thousands of cycles
          if (half == 0)
            asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(xnorm + (size_t)id * kTcSegTile + row),
                         "l"(__double_as_longlong(ssq)) : "memory");
        }
      }
      // ---- back half: scale and write (w = 0: the pending item, w = 1: the item whose front half just ran) ----
      for (int w = 0; w < 2; ++w) {
      int wid;
      uint32_t ti0;
      if (w == 0) {
        if (pend_id < 0) continue;
        wid = pend_id; ti0 = pend_ti0; pend_id = -1;
      } else {
        if (!front) continue;
        if (cur.res && la) { pend_id = id; pend_ti0 = cur_ti0; pend_it = cur; pend_par = par; continue; }
        wid = id; ti0 = cur_ti0;
      }
      const TcItem it = w == 0 ? pend_it : cur;
      const int cp = s_cp[(w == 0 ? pend_par : par) * 256 + e];
      const bool valid = row < it.ns;
      const bool tma_rows = tma_dims && (quarter * 32 + 32 <= it.ns);
      const int s = it.s0 + row;
      OutT* orow = out + ((size_t)(valid ? s : it.s0) * K + it.k) * D;
      const int nsub = (it.nch + kTcSubChunks - 1) / kTcSubChunks;
      double tot = ssq;
      if (it.res && J > 1) {
        // the siblings' partial sums, added in sibling order: every sibling -- and every run -- gets the same bits.  Siblings
        // are the CTAs of the J consecutive item ids; an item's wait depends only on posts of items with smaller ids + J, so
        // the persistent grid (all CTAs resident: one per SM) cannot deadlock.
        const long long t_x0 = (kProbe && probe) ? clock64() : 0;
        const double* xp = xnorm + (size_t)(wid - wid % J) * kTcSegTile + row;
        unsigned int polls = 0;
        bool got = false;
        tot = 0.0;
        if (w == 0 && early) {
          got = true;
          const double* sx = s_x + par * kTcEarly * kTcSegTile + row;
          for (int q = 0; q < J; ++q) {
            const double v = sx[q * kTcSegTile];
            got = got && __double_as_longlong(v) != -1ll;
            tot += v;
          }
        }
        if (!got) {                                       // immediate exchange, or a sibling more than a write sweep behind
          tot = 0.0;
          for (int j0 = 0; j0 < J; j0 += 4) {
            unsigned long long v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              v[q] = 0ull;
              if (j0 + q < J) asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(v[q]) : "l"(xp + (size_t)(j0 + q) * kTcSegTile) : "memory");
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              while (v[q] == ~0ull) {
                asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(v[q]) : "l"(xp + (size_t)(j0 + q) * kTcSegTile) : "memory");
                if (++polls == (1u << 22)) __trap();      // a sibling never arrived: fail loudly instead of hanging the GPU
              }
              tot += __longlong_as_double((long long)v[q]);
            }
          }
        }
        if (kProbe && probe) t_xwait += clock64() - t_x0;
      }
      const double nrm = sqrt(tot);
      double sc = 0.0;
      if (valid) {
        if (it.pj0 == 0 && half == 0) norms[(size_t)s * K + it.k] = nrm;
        sc = (1.0 / fmax(nrm, kEpsTc)) * (1.0 / fmax(sqrt((double)cp), kEpsTc));
      }
      // ---- write sweep: accumulator x scale -> fp64 -> 32-byte vector stores (each lane fills whole sectors of its row) ----
      uint32_t tw = it.res ? ti0 : ti;                  // accumulator counter of the write sweep
      for (int pass = it.pj0; pass < it.pj1; ++pass) {
        const int width = min(kTcPassN, D - pass * kTcPassN);
        const int c0 = half * 64;
        const int w0 = min(32, width - c0), w1 = min(32, width - c0 - 32);
        OutT* op = orow + (size_t)pass * kTcPassN + c0;
        uint32_t va[32], vb[32];
        long long t_s0 = 0;
        for (int sub = 0; sub < nsub; ++sub, ++tw) {
          const uint32_t buf = tw % kTcBufs, use = tw / kTcBufs;
          if (!it.res) {                                  // (resident: complete since the norm phase)
            TC_TIMED_WAIT(bar_tfull + 8 * buf, use & 1, t_wwait);
            tc_fence_after();
          }
          if (warp == 4 && lane == 0) TC_MARK(tw, it.res ? 6 : 4);
          if (kProbe && sub == 0 && probe) t_s0 = clock64();
          const uint32_t tcol = tlane + buf * kTcPassN + c0;
          if (w0 > 0) {
            if (sub == 0) {
              tc_ld32_issue(tcol, va);
              if (w1 > 0) tc_ld32_issue(tcol + 32, vb);
              tc_ld_wait(va);
              if (w1 > 0) tc_ld_wait(vb);
            } else {
              tc_acc32(tcol, va);
              if (w1 > 0) tc_acc32(tcol + 32, vb);
            }
          }
          // the accumulator is in registers: hand the TMEM buffer back before the (slow) stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
        }
        if constexpr (kPlanes) {
          // x = fl32(block * scale - mean): ONE fused multiply-add per element on the fp32 scale and the fp32 copy of the
          // model mean (both rounded once, ~1e-7 relative of the block value; the projection's operand is fp32-equivalent
          // anyway), kept in place of the accumulator values
          const float scf = (float)sc;
          const int dcol = it.k * D + pass * kTcPassN + c0;   // first of this warp's (up to) 64 columns of the [K * D] row
          if (w0 > 0) {
            // eight columns at a time (the compiler barrier keeps the 64 mean loads from being hoisted into 128 live registers)
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              if (g < 4 || w1 > 0) {
                const float4 m0 = __ldg(reinterpret_cast<const float4*>(pca_mean + dcol + 8 * g));
                const float4 m1 = __ldg(reinterpret_cast<const float4*>(pca_mean + dcol + 8 * g + 4));
                const float mm[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  if (g < 4) va[8 * g + j] = __float_as_uint(fmaf(__uint_as_float(va[8 * g + j]), scf, -mm[j]));
                  else vb[8 * (g - 4) + j] = __float_as_uint(fmaf(__uint_as_float(vb[8 * (g - 4) + j]), scf, -mm[j]));
                }
              }
            }
          }
          const size_t KD = (size_t)K * D;
          // Full 32-row groups go through the TMA: the three planes of a 32-row x 32-column half are staged in three
          // SWIZZLE_64B boxes (2 KB each; row = lane, 16-byte chunk c at c ^ ((row >> 1) & 3): conflict-free) -- a chunk's
          // three planes are computed together (12 registers live) and stored to the three boxes at once -- then three bulk
          // tensor stores.  (r2: per-lane 16-byte global stores ran the planes epilogue at half the fp64 epilogue's rate, the
          // L1 store path being the limit; staging one plane at a time needs the split values of all 64 columns live.)
          if (tma_rows && w0 == 32) {
            const int yrow = it.s0 + quarter * 32;
#pragma unroll
            for (int hseg = 0; hseg < 2; ++hseg) {
              if (hseg == 0 || w1 == 32) {
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // previous boxes read out
                __syncwarp();
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  uint32_t o2[4], o1[4], o0[4];
                  if (hseg == 0) tc_plane_chunk3(va, 8 * c, o0, o1, o2); else tc_plane_chunk3(vb, 8 * c, o0, o1, o2);
                  const uint32_t addr = my_stage + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4);
                  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o0[0]), "r"(o0[1]), "r"(o0[2]), "r"(o0[3]) : "memory");
                  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr + 2048), "r"(o1[0]), "r"(o1[1]), "r"(o1[2]), "r"(o1[3]) : "memory");
                  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr + 4096), "r"(o2[0]), "r"(o2[1]), "r"(o2[2]), "r"(o2[3]) : "memory");
                }
                tc_fence_async_smem();
                __syncwarp();
                if (lane == 0) {
#pragma unroll
                  for (int pl = 0; pl < 3; ++pl)
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
                                 ::"l"(&map_out), "r"(my_stage + pl * 2048), "r"(dcol + 32 * hseg), "r"(pl * S_total + yrow),
                                   "l"(pol_out) : "memory");
                  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
              }
            }
          } else if (valid && w0 > 0) {
            // partial row groups / narrow tails: every lane writes its own row in 16-byte pieces
            OutT* p0 = out + (size_t)s * KD + dcol;
            const int wtot = w0 + (w1 > 0 ? w1 : 0);       // multiple of 16 (D % 16 == 0)
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              if (8 * c < wtot) {
                uint32_t o2[4], o1[4], o0[4];
                if (c < 4) tc_plane_chunk3(va, 8 * c, o0, o1, o2); else tc_plane_chunk3(vb, 8 * (c - 4), o0, o1, o2);
                *reinterpret_cast<uint4*>(p0 + 8 * c) = make_uint4(o0[0], o0[1], o0[2], o0[3]);
                *reinterpret_cast<uint4*>(p0 + 8 * c + (size_t)S_total * KD) = make_uint4(o1[0], o1[1], o1[2], o1[3]);
                *reinterpret_cast<uint4*>(p0 + 8 * c + 2 * (size_t)S_total * KD) = make_uint4(o2[0], o2[1], o2[2], o2[3]);
              }
            }
          }
        } else if (tma_rows && w0 == 32) {                      // (warp-uniform) all 32 rows valid, full 32-column pieces
          const int xcol = it.k * D + pass * kTcPassN + c0;
          const int yrow = it.s0 + quarter * 32;
          auto flush = [&](int x) {                      // box staged by all lanes -> one bulk tensor store
            tc_fence_async_smem();
            __syncwarp();
            if (lane == 0) tma_store_2d(&map_out, my_stage + (nbox & 1) * kTcBoxBytes, x, yrow, pol_out);
            ++nbox;
          };
          auto acquire = [&]() -> uint32_t {             // the box used two stores ago has been read by the TMA
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
            return my_stage + (nbox & 1) * kTcBoxBytes;
          };
          if constexpr (kPlanes) {
          } else if constexpr (sizeof(OutT) == 8) {
            uint32_t sb = acquire(); TcStage<OutT>::template put<0>(sb, lane, va, sc); flush(xcol);
            sb = acquire(); TcStage<OutT>::template put<16>(sb, lane, va, sc); flush(xcol + 16);
            if (w1 == 32) {
              sb = acquire(); TcStage<OutT>::template put<0>(sb, lane, vb, sc); flush(xcol + 32);
              sb = acquire(); TcStage<OutT>::template put<16>(sb, lane, vb, sc); flush(xcol + 48);
            }
          } else {
            uint32_t sb = acquire(); TcStage<OutT>::template put<0>(sb, lane, va, sc); flush(xcol);
            if (w1 == 32) { sb = acquire(); TcStage<OutT>::template put<0>(sb, lane, vb, sc); flush(xcol + 32); }
          }
        } else if (valid && w0 > 0) {
          if constexpr (!kPlanes) {
            TcOut<OutT>::store(op, va, sc, w0);
            if (w1 > 0) TcOut<OutT>::store(op + 32, vb, sc, w1);
          }
        }
        if (kProbe && probe) t_store += clock64() - t_s0;
        if (warp == 4 && lane == 0) TC_MARK(tw - 1, it.res ? 7 : 5);
      }
      if (!it.res) ti = tw;
      }   // w
      if (front) id += gridDim.x;
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // staging boxes drained before exit
    if (kProbe && probe && warp == 4 && lane == 0) {
      unsigned long long* pr = probe + blockIdx.x * 16;
      pr[0] = clock64() - t_ebegin; pr[1] = t_nwait; pr[3] = t_bar; pr[4] = t_wwait; pr[5] = t_store; pr[12] = n_done; pr[13] = t_xwait;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ------------------------------------------------------------------------------------------------
bool agg_tc_supported(int N, int D, int K) {
  const char* e = getenv("SEGVLAD_AGG_TC");   // "0" selects the SIMT kernel of aggregate.cu (kept as a cross-check)
  const bool on = !(e && e[0] == '0');
  return on && N >= 1 && K >= 1 && D >= 16 && D % 16 == 0 && D <= 65536;
}

template <typename OutT, bool kPlanes>
static int launch_tc(const AggTcArgs& a, const CUtensorMap& map, int n_items, int J, int grid, int resident,
                     cudaStream_t st) {
  const size_t smem = agg_tc_smem();
  // L2 evict_first policy on the output stores: the output is a pure stream and should not push the operand planes, which
  // the write sweep re-reads, out of L2 (measured: 0.441 -> 0.427 ms on the bench workload).  "0" switches it off.
  const char* se = getenv("SEGVLAD_AGG_STORE_HINT");
  const int store_hint = se ? atoi(se) : 1;
  // output as a 2-D tensor: the epilogue stores 32-row x 128-byte boxes through the TMA
  //   descriptors [S_total][K*D] fp64 / fp32, or PCA planes [3 * S_total][K*D] bf16 (plane-major rows)
  CUtensorMap map_out;
  {
    PFN_encodeTiled enc = get_encode();
    const cuuint64_t rows = kPlanes ? (cuuint64_t)3 * a.S_total : (cuuint64_t)a.S_total;
    cuuint64_t dims[2] = {(cuuint64_t)a.K * a.D, rows};
    cuuint64_t strides[1] = {(cuuint64_t)a.K * a.D * sizeof(OutT)};
    cuuint32_t box[2] = {(cuuint32_t)(kPlanes ? 32 : 128 / sizeof(OutT)), 32};   // planes: 32 rows x 64 bytes, SWIZZLE_64B
    cuuint32_t estr[2] = {1, 1};
    const CUtensorMapDataType dt = kPlanes ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                           : (sizeof(OutT) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
    CUresult r = enc(&map_out, dt, 2, a.out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     kPlanes ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (output) failed (%d)", (int)r); return SEGVLAD_ECUDA; }
  }
  auto kern = a.probe ? aggregate_tc_kernel<OutT, true, kPlanes> : aggregate_tc_kernel<OutT, false, kPlanes>;
  SV_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int pslot = prof_begin(SEGVLAD_PROF_AGGREGATE, st);
  kern<<<grid, kTcThreadsAgg, smem, st>>>(map, map_out, a.tile_tbl, a.cl_ptr, a.memS, a.cpred, a.B, a.N, a.D, a.K, n_items, J,
                                          reinterpret_cast<OutT*>(a.out), a.norms, a.probe, store_hint, a.pca_mean, a.S_total,
                                          resident, a.xnorm);
  prof_end(pslot, st);
  SV_CHECK_LAUNCH();
  return SEGVLAD_OK;
}

// Single-sweep launches wait for sibling CTAs inside the kernel, which is only safe while every CTA of the grid is resident.
// Two such kernels running side by side on one GPU (two host threads, two streams) could each hold half of the SMs and wait
// for CTAs that cannot start, so launches of one process on one device are chained on the GPU: a launch first waits for the
// event of the previous one (no host blocking).  Other kernels may share the GPU: they finish and free their SMs.
// (Kernels of OTHER processes running concurrently under MPS are not covered: set SEGVLAD_AGG_RESIDENT=0 there.)
struct TcLaunchChain {
  static constexpr int kMaxDev = 64;
  static std::mutex& mu() { static std::mutex m; return m; }
  static cudaEvent_t* events() { static cudaEvent_t ev[kMaxDev] = {}; return ev; }
  bool on;
  cudaStream_t st;
  cudaEvent_t* ev = nullptr;
  TcLaunchChain(bool enable, cudaStream_t s) : on(enable), st(s) {
    if (!on) return;
    mu().lock();
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) { on = false; mu().unlock(); return; }
    ev = events() + dev;
    if (*ev) cudaStreamWaitEvent(st, *ev, 0);
    else if (cudaEventCreateWithFlags(ev, cudaEventDisableTiming) != cudaSuccess) *ev = nullptr;
  }
  ~TcLaunchChain() {
    if (!on) return;
    if (ev && *ev) cudaEventRecord(*ev, st);
    mu().unlock();
  }
};

int agg_tc_run(const AggTcArgs& a, cudaStream_t st) {
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    SV_CHECK_CUDA(cudaGetDevice(&dev));
    SV_CHECK_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int Np = agg_tc_np(a.N);
  // segment tiles: 16 consecutive 8-segment groups of one image (group numbering = aggregate.cu's group table)
  const int max_tiles = agg_tc_max_tiles(a.B, a.S_total);
  int* h = (int*)malloc(sizeof(int) * 4 * (size_t)max_tiles);
  SV_REQUIRE(h, "aggregate: host malloc failed");
  int nt = 0, g = 0;
  for (int b = 0; b < a.B; ++b) {
    const int s0 = a.seg_offsets_host[b], Si = a.seg_offsets_host[b + 1] - s0;
    for (int t0 = 0; t0 < Si; t0 += kTcSegTile) {
      h[4 * nt + 0] = b; h[4 * nt + 1] = g + t0 / 8; h[4 * nt + 2] = s0 + t0; h[4 * nt + 3] = min(kTcSegTile, Si - t0);
      ++nt;
    }
    g += (Si + 7) / 8;
  }
  cudaError_t e = cudaMemcpyAsync(a.tile_tbl, h, sizeof(int) * 4 * (size_t)nt, cudaMemcpyHostToDevice, st);
  free(h);
  SV_CHECK_CUDA(e);
  if (nt == 0) return SEGVLAD_OK;

  if (a.R) {
    rt_planes_kernel<<<dim3(Np / 64, (a.D + 31) / 32, a.B), 256, 0, st>>>(a.R, a.cl_tok, a.B, a.N, a.D, Np, a.RT);
  } else {
    const int ch = agg_tc_fused_channels(a.N, a.K);
    SV_REQUIRE(ch > 0 && a.tokens_dn && a.centers && a.labels && a.nrm, "aggregate: fused residual planes need the tokens");
    const size_t smem = ((((size_t)ch * a.N + 3) & ~(size_t)3) + (size_t)a.K * ch) * sizeof(float);
    SV_CHECK_CUDA(cudaFuncSetAttribute(rt_from_tokens_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rt_from_tokens_kernel<<<dim3((a.D + ch - 1) / ch, a.B), 256, smem, st>>>(a.tokens_dn, a.centers, a.labels, a.nrm, a.cl_tok,
                                                                            a.B, a.N, a.D, a.K, Np, ch, a.RT);
  }
  SV_CHECK_LAUNCH();

  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return SEGVLAD_ECUDA; }
  CUtensorMap map;
  cuuint64_t dims[2] = {(cuuint64_t)Np, (cuuint64_t)3 * a.B * a.D};
  cuuint64_t strides[1] = {(cuuint64_t)Np * 2};
  cuuint32_t box[2] = {(cuuint32_t)kTcTokChunk, (cuuint32_t)kTcPassN};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a.RT, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return SEGVLAD_ECUDA; }

  const int P = (a.D + kTcPassN - 1) / kTcPassN;
  // Single-sweep mode (r2): the channel passes of a block are split over J = ceil(P / 4) sibling CTAs so that each one's
  // accumulators fit its TMEM and the contraction is issued once (tc_item).  "0" selects the two-sweep schedule of r1.
  // With at most kTcBufs / 2 passes per CTA two items fit the TMEM and the sibling exchange of the next item hides behind the
  // write sweep of the current one (look-ahead, "SEGVLAD_AGG_LA=1").  Measured slower than the immediate exchange with
  // kTcBufs passes per CTA (0.398 vs 0.379 ms on the bench workload: twice the items, and each costs a front half), so off
  // by default.
  const char* re = getenv("SEGVLAD_AGG_RESIDENT");
  const char* le = getenv("SEGVLAD_AGG_LA");
  const int per_cta = (le && le[0] == '1') ? kTcBufs / 2 : kTcBufs;
  const int Jres = (P + per_cta - 1) / per_cta;
  const int resident = !(re && re[0] == '0') && Jres <= 32 && 2 * Jres <= num_sms;
  int J = resident ? Jres : 1;
  while ((long long)nt * a.K * J < 2ll * num_sms && J * 2 <= P) J *= 2;   // few items: split the write sweep over channels
  const long long n_items = (long long)nt * a.K * J;
  SV_REQUIRE(n_items < (1ll << 31), "aggregate: too many work items");
  const int grid = (int)(n_items < num_sms ? n_items : num_sms);
  // siblings of an item (J consecutive ids) wait for each other: the grid is never larger than the number of SMs and the
  // kernel's shared memory allows one CTA per SM, so every CTA is resident
  // (their exchange slots are preset to the all-ones pattern = "not posted yet")
  if (resident && J > 1) SV_CHECK_CUDA(cudaMemsetAsync(a.xnorm, 0xFF, sizeof(double) * (size_t)n_items * kTcSegTile, st));
  TcLaunchChain chain(resident && J > 1, st);
  if (a.out_dtype == SEGVLAD_OUT_PCA_PLANES) {
    SV_REQUIRE(a.pca_mean != nullptr && a.D % 64 == 0, "aggregate: the PCA-planes output needs the model mean and D_t %% 64 == 0");
    return launch_tc<__nv_bfloat16, true>(a, map, (int)n_items, J, grid, resident, st);
  }
  return a.out_dtype == SEGVLAD_OUT_F64 ? launch_tc<double, false>(a, map, (int)n_items, J, grid, resident, st)
                                        : launch_tc<float, false>(a, map, (int)n_items, J, grid, resident, st);
}

}  // namespace segvlad
