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
//     are then a contiguous K-major column range that TMA tiles straight into the SWIZZLE_128B operand layout;
//   * one CTA owns (image, tile of 128 segments, cluster): A = mask tile [128 x 64 tokens] written into shared memory
//     by two builder warps from the membership words, B = [128 channels x 64 tokens] x 3 planes by TMA, D = fp32
//     accumulators in TMEM (4 buffers of 128 columns);
//   * the block norm needs all D channels but TMEM holds 512 of them, so the contraction is issued twice -- a norm
//     sweep (TMEM -> sum of squares, no stores) and a write sweep (TMEM -> x scale -> fp64 -> swizzled staging box ->
//     bulk tensor store).  The two sweeps are drained by DIFFERENT warps from different TMEM buffers and the norm sweep
//     of item i+1 is interleaved pass by pass with the write sweep of item i (tc_schedule), so the output stream to
//     HBM never pauses.  The tensor work is ~10 % of the write time, the operand re-read comes from L2;
//   * thread = segment row (TMEM lane), so norms and scales never leave the thread; the norm warps hand the row sums
//     of squares to the write warps through a two-slot shared-memory mailbox (mbarrier full/empty pair).
// Accumulation is fp32 in TMEM (truncating adds, measured ~4e-8 relative per MMA): <= 3 * n_k / 16 MMAs per element,
// i.e. ~1e-6 relative for the largest clusters -- inside the 1e-5 descriptor tolerance; the planes are accumulated
// small-to-large.  Warp roles: 0 TMA producer, 1 MMA issuer (+TMEM alloc), 2-3 mask-tile builders, 4-7 norm sweep,
// 8-15 write sweep (two warps per TMEM lane quarter, 64 channels of a pass each).
#include "aggregate_tc.cuh"

#include <stdlib.h>

#include "tc_ptx.cuh"

namespace segvlad {

constexpr int kTcThreadsAgg = 512;   // warps: 0 TMA, 1 MMA, 2-3 mask builders, 4-7 norm sweep, 8-15 write sweep (two per lane quarter)
constexpr int kTcStages = 4;
constexpr uint32_t kTcTileBytes = kTcSegTile * kTcTokChunk * 2;          // 8 KB: one [128 x 32] bf16 operand tile
constexpr uint32_t kTcStageBytes = 4 * kTcTileBytes;                     // A + 3 B planes
static_assert(512 / kTcPassN == 4, "TMEM: two accumulator buffers per sweep");
constexpr double kEpsTc = 1e-12;

constexpr uint32_t kTcBoxBytes = 32 * 128;                               // output staging box: 32 rows x 128 bytes
constexpr uint32_t kTcStageOutBytes = 8 * 2 * kTcBoxBytes;               // 8 epilogue warps x 2 boxes
__host__ __device__ constexpr size_t agg_tc_smem() {
  return 1024 + (size_t)kTcStages * kTcStageBytes + kTcStageOutBytes + 256 + 2 * kTcSegTile * 8;
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
__device__ __forceinline__ void tc_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// mbarrier wait that adds the cycles spent to a counter when the development probe is on
#define TC_TIMED_WAIT(bar, par, acc) do { if (probe) { const long long _t = clock64(); mbar_wait(bar, par); acc += clock64() - _t; } else mbar_wait(bar, par); } while (0)

struct TcItem {
  int id, b, g0, s0, ns, k, p0, p1, rows, x0, nch, pj0, pj1;   // x0: p0 rounded down to 8 tokens (TMA needs 16-byte
};                                                          // aligned global addresses); tokens < p0 are masked out
// item id -> (segment tile, cluster, channel split); identical in every warp role
__device__ __forceinline__ TcItem tc_item(int id, const int* __restrict__ tile_tbl, const int* __restrict__ cl_ptr, int K,
                                          int J, int P) {
  TcItem it;
  it.id = id;
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
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(src), "r"(x), "r"(y) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// Issue order of the accumulator passes of one CTA, identical in every warp role.  The norm sweep of the NEXT non-empty
// item is interleaved pass by pass with the write sweep of the current one: the two sweeps are drained by different warps
// from different TMEM buffers, so the output stream to HBM never pauses for a norm sweep (r1 probes: with the sweeps back
// to back the write sweep ran at the HBM rate for 66 % of the kernel and the memory system idled for the rest).
template <typename FN, typename FW>
__device__ __forceinline__ void tc_schedule(int first, int stride, int n_items, const int* __restrict__ tile_tbl,
                                            const int* __restrict__ cl_ptr, int K, int J, int P, FN&& norm_pass,
                                            FW&& write_pass) {
  auto next_nonempty = [&](int from, TcItem& o) -> int {
    for (int i = from; i < n_items; i += stride) {
      o = tc_item(i, tile_tbl, cl_ptr, K, J, P);
      if (o.rows) return i;
    }
    return -1;
  };
  TcItem cur, nxt;
  int cid = next_nonempty(first, cur);
  if (cid < 0) return;
  for (int p = 0; p < P; ++p) norm_pass(cur, p);
  for (;;) {
    const int nid = next_nonempty(cid + stride, nxt);
    const int nw = cur.pj1 - cur.pj0;
    const int steps = max(nw, nid >= 0 ? P : 0);
    for (int s = 0; s < steps; ++s) {
      if (nid >= 0 && s < P) norm_pass(nxt, s);
      if (s < nw) write_pass(cur, cur.pj0 + s);
    }
    if (nid < 0) break;
    cur = nxt;
    cid = nid;
  }
}

template <typename OutT>
__global__ void __launch_bounds__(kTcThreadsAgg, 1)
aggregate_tc_kernel(const __grid_constant__ CUtensorMap map_rt, const __grid_constant__ CUtensorMap map_out,
                    const __grid_constant__ CUtensorMap map_lin, const int* __restrict__ tile_tbl,
                    const int* __restrict__ cl_ptr, const uint16_t* __restrict__ memS, const int* __restrict__ cpred,
                    int B, int N, int D, int K, int n_items, int J, OutT* __restrict__ out, double* __restrict__ norms,
                    unsigned long long* __restrict__ probe, int exp) {
  extern __shared__ __align__(1024) uint8_t tc_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* stage_out = smem + kTcStages * kTcStageBytes;   // [8 write warps][2 boxes][4 KB], 1024-byte aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_out + kTcStageOutBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
  double* s_ssq = reinterpret_cast<double*>(bars + 32);   // [2 slots][128 rows] block sums of squares, norm -> write warps
  const uint32_t bar_full = smem_u32(bars + 0);      // [kTcStages] A written (2 builder warps) + B landed (TMA)
  const uint32_t bar_empty = smem_u32(bars + 4);     // [kTcStages] MMAs that read the stage retired
  const uint32_t bar_tfull = smem_u32(bars + 8);     // [4] accumulator pass complete   (buffers 0-1 norm, 2-3 write)
  const uint32_t bar_tempty = smem_u32(bars + 12);   // [4] accumulator drained
  const uint32_t bar_sfull = smem_u32(bars + 16);    // [2] sums of squares of an item published by the 4 norm warps
  const uint32_t bar_sempty = smem_u32(bars + 18);   // [2] ... consumed by the 8 write warps
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int P = (D + kTcPassN - 1) / kTcPassN;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kTcStages; ++i) { mbar_init(bar_full + 8 * i, 3); mbar_init(bar_empty + 8 * i, 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, i < 2 ? 4 : 8); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_sfull + 8 * i, 4); mbar_init(bar_sempty + 8 * i, 8); }
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

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_rt) : "memory");
      uint32_t stage = 0, phase = 0;
      long long t_wait = 0;
      auto load_pass = [&](const TcItem& it, int pass) {
        for (int c = 0; c < it.nch; ++c) {
          TC_TIMED_WAIT(bar_empty + 8 * stage, phase ^ 1, t_wait);
          const uint32_t sb = smem_u32(smem + stage * kTcStageBytes) + kTcTileBytes;
          const uint32_t fb = bar_full + 8 * stage;
          if (exp & 1) mbar_arrive(fb);
          else {
          mbar_arrive_expect_tx(fb, 3 * kTcTileBytes);
#pragma unroll
          for (int pl = 0; pl < 3; ++pl)
            tma_load_2d(sb + pl * kTcTileBytes, &map_rt, fb, it.x0 + c * kTcTokChunk, (pl * B + it.b) * D + pass * kTcPassN);
          }
          if (++stage == kTcStages) { stage = 0; phase ^= 1; }
        }
      };
      tc_schedule(blockIdx.x, gridDim.x, n_items, tile_tbl, cl_ptr, K, J, P, load_pass, load_pass);
      if (probe) probe[blockIdx.x * 16 + 9] = t_wait;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, tn = 0, tw = 0;
      long long t_tempty_n = 0, t_tempty_w = 0, t_full = 0;
      const long long t_begin = clock64();
      auto mma_pass = [&](const TcItem& it, int pass, uint32_t buf) {
        const int width = min(kTcPassN, D - pass * kTcPassN);
        // kind::f16: D=f32 (bit 4), A=B=bf16 (bits 7, 10), both K-major, N>>3 @17, M>>4 @24
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(width >> 3) << 17) |
                               ((uint32_t)(kTcSegTile >> 4) << 24);
        const uint32_t d_tmem = tmem_base + buf * kTcPassN;
        for (int c = 0; c < it.nch; ++c) {
          TC_TIMED_WAIT(bar_full + 8 * stage, phase, t_full);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kTcStageBytes);
          const uint64_t adesc = umma_desc_sw64(sa);
          const int nk = (min(kTcTokChunk, it.p1 - it.x0 - c * kTcTokChunk) + 15) >> 4;
          for (int kk = 0; kk < ((exp & 4) ? 0 : nk); ++kk) {
            const uint64_t adv = (uint64_t)(kk * 32 >> 4);
#pragma unroll
            for (int pl = 0; pl < 3; ++pl)   // lo, mid, hi: small terms first
              tc_mma_bf16(d_tmem, adesc + adv, umma_desc_sw64(sa + (1 + pl) * kTcTileBytes) + adv, idesc,
                          (c | kk | pl) != 0);
          }
          tc_commit(bar_empty + 8 * stage);
          if (++stage == kTcStages) { stage = 0; phase ^= 1; }
        }
        tc_commit(bar_tfull + 8 * buf);
      };
      auto norm_pass = [&](const TcItem& it, int pass) {
        const uint32_t buf = tn & 1, use = tn >> 1;
        ++tn;
        TC_TIMED_WAIT(bar_tempty + 8 * buf, (use & 1) ^ 1, t_tempty_n);
        tc_fence_after();
        mma_pass(it, pass, buf);
      };
      auto write_pass = [&](const TcItem& it, int pass) {
        const uint32_t buf = 2 + (tw & 1), use = tw >> 1;
        ++tw;
        TC_TIMED_WAIT(bar_tempty + 8 * buf, (use & 1) ^ 1, t_tempty_w);
        tc_fence_after();
        mma_pass(it, pass, buf);
      };
      tc_schedule(blockIdx.x, gridDim.x, n_items, tile_tbl, cl_ptr, K, J, P, norm_pass, write_pass);
      if (probe) { unsigned long long* pr = probe + blockIdx.x * 16;
                   pr[6] = t_tempty_w; pr[7] = t_full; pr[8] = clock64() - t_begin; pr[13] = t_tempty_n; }
    }
  } else if (warp == 2 || warp == 3) {
    // ===================== mask-tile builders (64 threads) =====================
    const int u = threadIdx.x - 64;
    uint32_t stage = 0, phase = 0;
    long long t_bwait = 0;
    const long long t_bbegin = clock64();
    // cached membership words of this thread's (group, 8-token) cell (8 x u16), one cache per sweep: the two
    // sweeps in flight belong to different items
    struct Cache { uint32_t w[4]; int s0, k, c; };
    Cache cn, cw;
    cn.s0 = cw.s0 = -1; cn.k = cw.k = -1; cn.c = cw.c = -1;
    // this thread's cell of the [128 segments x 32 tokens] tile: 8-segment group g x 8-token column block c8
    const int g = u >> 2, c8 = u & 3;
    auto build_pass = [&](const TcItem& it, Cache& cc) {
      const int ngrp = (it.ns + 7) >> 3;
      for (int c = 0; c < it.nch; ++c) {
        if (cc.c != c || cc.s0 != it.s0 || cc.k != it.k) {
          const int i0 = it.x0 + c * kTcTokChunk + c8 * 8;
          const uint16_t* src = memS + (size_t)(it.g0 + g) * N;
#pragma unroll
          for (int t = 0; t < 8; t += 2) {
            uint32_t a = 0, bb = 0;
            if (g < ngrp) {
              if (i0 + t >= it.p0 && i0 + t < it.p1) a = src[i0 + t];
              if (i0 + t + 1 >= it.p0 && i0 + t + 1 < it.p1) bb = src[i0 + t + 1];
            }
            cc.w[t >> 1] = a | (bb << 16);
          }
          cc.c = c; cc.s0 = it.s0; cc.k = it.k;
        }
        TC_TIMED_WAIT(bar_empty + 8 * stage, phase ^ 1, t_bwait);
        const uint32_t sa = smem_u32(smem + stage * kTcStageBytes);
        if (!(exp & 2))
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          // row = 8 g + j of the SWIZZLE_64B K-major tile: 64-byte row, 16-byte chunk c8 stored at c8 ^ ((row >> 1) & 3).
          // Odd groups take their rows in the order j ^ 1, so that eight consecutive lanes (two groups x 4 chunks) fill
          // one 128-byte bank window: conflict-free.
          // bit j of both 16-bit membership words -> two bf16 1.0 / 0.0 in one shift, mask and multiply
          // (r1 ncu: the select-per-bit form made the two builder warps the critical path of the whole kernel)
          const int j = jj ^ (g & 1);
          uint32_t o[4];
#pragma unroll
          for (int h = 0; h < 4; ++h) o[h] = ((cc.w[h] >> j) & 0x00010001u) * 0x3F80u;
          const uint32_t addr = sa + g * 512 + j * 64 + ((c8 ^ ((j >> 1) & 3)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
        }
        tc_fence_async_smem();     // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_full + 8 * stage);
        if (++stage == kTcStages) { stage = 0; phase ^= 1; }
      }
    };
    auto norm_pass = [&](const TcItem& it, int) { build_pass(it, cn); };
    auto write_pass = [&](const TcItem& it, int) { build_pass(it, cw); };
    tc_schedule(blockIdx.x, gridDim.x, n_items, tile_tbl, cl_ptr, K, J, P, norm_pass, write_pass);
    if (probe && threadIdx.x == 64) { probe[blockIdx.x * 16 + 10] = t_bwait; probe[blockIdx.x * 16 + 11] = clock64() - t_bbegin; }
  } else if (warp < 8) {
    // ===================== norm warps (4): thread = segment row (TMEM lane), whole passes =====================
    // Sum of squares of the block row over all D channels (fp32 products, fp64 accumulation per 32 columns), published
    // to the write warps through a two-slot shared-memory mailbox.
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    uint32_t tn = 0, n_done = 0;
    long long t_nwait = 0, t_swait = 0;
    const long long t_nbegin = clock64();
    for (int id = blockIdx.x; id < n_items; id += gridDim.x) {
      const TcItem it = tc_item(id, tile_tbl, cl_ptr, K, J, P);
      if (it.rows == 0) continue;
      double ssq = 0.0;
      for (int pass = 0; pass < P; ++pass, ++tn) {
        const int width = min(kTcPassN, D - pass * kTcPassN);
        const int np = (width + 31) >> 5;
        const uint32_t buf = tn & 1, use = tn >> 1;
        TC_TIMED_WAIT(bar_tfull + 8 * buf, use & 1, t_nwait);
        tc_fence_after();
        const uint32_t tcol = tlane + buf * kTcPassN;
        uint32_t va[32], vb[32];
        tc_ld32_issue(tcol, va);
        tc_ld_wait(va);
#pragma unroll 1
        for (int cc = 0; cc < np; cc += 2) {
          if (cc + 1 < np) tc_ld32_issue(tcol + (cc + 1) * 32, vb);      // next piece loads while this one is reduced
          ssq += (double)tc_sumsq(va, min(32, width - cc * 32));
          if (cc + 1 < np) {
            tc_ld_wait(vb);
            if (cc + 2 < np) tc_ld32_issue(tcol + (cc + 2) * 32, va);
            ssq += (double)tc_sumsq(vb, min(32, width - (cc + 1) * 32));
            if (cc + 2 < np) tc_ld_wait(va);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
      }
      const uint32_t slot = n_done & 1, suse = n_done >> 1;
      ++n_done;
      TC_TIMED_WAIT(bar_sempty + 8 * slot, (suse & 1) ^ 1, t_swait);
      s_ssq[slot * kTcSegTile + row] = ssq;
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_sfull + 8 * slot);     // (release: the stores above are visible to the waiters)
    }
    if (probe && warp == 4 && lane == 0) {
      unsigned long long* pr = probe + blockIdx.x * 16;
      pr[1] = t_nwait; pr[2] = t_swait; pr[3] = clock64() - t_nbegin;
    }
  } else {
    // ===================== write warps (8): thread = segment row (TMEM lane), 64 channels of a pass =====================
    // Two warps per TMEM lane quarter: one warp per scheduler could not hide its own tcgen05.ld -> convert -> store
    // latencies (r1 ncu: 12.5 % warps active, issue slots 19 % busy at 62 % of the HBM write roofline).
    const int quarter = warp & 3, half = warp >= 12 ? 1 : 0;
    const int row = quarter * 32 + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t my_stage = smem_u32(stage_out) + (uint32_t)(half * 4 + quarter) * 2 * kTcBoxBytes;
    const bool tma_dims = (D % 64) == 0;      // every pass piece of this warp is 0 or 32 columns wide
    uint32_t tw = 0, n_done = 0, nbox = 0;
    long long t_swait = 0, t_wwait = 0, t_store = 0;
    const long long t_ebegin = clock64();
    for (int id = blockIdx.x; id < n_items; id += gridDim.x) {
      const TcItem it = tc_item(id, tile_tbl, cl_ptr, K, J, P);
      const bool valid = row < it.ns;
      const bool tma_rows = tma_dims && (quarter * 32 + 32 <= it.ns);
      const int s = it.s0 + row;
      OutT* orow = out + ((size_t)(valid ? s : it.s0) * K + it.k) * D;
      if (it.rows == 0) {
        // empty cluster: the block is zero for every segment
        if (valid) {
          if (it.pj0 == 0 && half == 0) norms[(size_t)s * K + it.k] = 0.0;
          const int d_beg = it.pj0 * kTcPassN, d_end = min(D, it.pj1 * kTcPassN);
          for (int d = d_beg + 16 * half; d < d_end; d += 32) TcOut<OutT>::zero16(orow + d);
        }
        continue;
      }
      const uint32_t slot = n_done & 1, suse = n_done >> 1;
      ++n_done;
      TC_TIMED_WAIT(bar_sfull + 8 * slot, suse & 1, t_swait);
      const double nrm = sqrt(s_ssq[slot * kTcSegTile + row]);
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_sempty + 8 * slot);
      double sc = 0.0;
      if (valid) {
        if (it.pj0 == 0 && half == 0) norms[(size_t)s * K + it.k] = nrm;
        sc = (1.0 / fmax(nrm, kEpsTc)) * (1.0 / fmax(sqrt((double)cpred[s]), kEpsTc));
      }
      // ---- write sweep: accumulator x scale -> fp64 -> staging box -> bulk tensor store ----
      for (int pass = it.pj0; pass < it.pj1; ++pass, ++tw) {
        const int width = min(kTcPassN, D - pass * kTcPassN);
        const int c0 = half * 64;
        const int w0 = min(32, width - c0), w1 = min(32, width - c0 - 32);     // columns in this warp's two pieces
        const uint32_t buf = 2 + (tw & 1), use = tw >> 1;
        TC_TIMED_WAIT(bar_tfull + 8 * buf, use & 1, t_wwait);
        tc_fence_after();
        const long long t_s0 = probe ? clock64() : 0;
        const uint32_t tcol = tlane + buf * kTcPassN + c0;
        OutT* op = orow + (size_t)pass * kTcPassN + c0;
        uint32_t va[32], vb[32];
        if (w0 > 0) {                                  // (warp-uniform)
          tc_ld32_issue(tcol, va);
          if (w1 > 0) tc_ld32_issue(tcol + 32, vb);
          tc_ld_wait(va);
          if (w1 > 0) tc_ld_wait(vb);
        }
        // the accumulator is in registers: hand the TMEM buffer back before the (slow) stores
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
        if (exp & 8) { }
        else if (tma_rows && w0 == 32) {                 // (warp-uniform) all 32 rows valid, full 32-column pieces
          const int xcol = it.k * D + pass * kTcPassN + c0;
          const int yrow = it.s0 + quarter * 32;
          auto flush = [&](int x) {                      // box staged by all lanes -> one bulk tensor store
            tc_fence_async_smem();
            __syncwarp();
            if (exp & 16) {
              if (lane == 0) tma_store_2d(&map_lin, my_stage + (nbox & 1) * kTcBoxBytes, 0, ((blockIdx.x * 8 + (warp - 8)) * 336 + (nbox % 336)) * 32);
            } else
            if (lane == 0) tma_store_2d(&map_out, my_stage + (nbox & 1) * kTcBoxBytes, x, yrow);
            ++nbox;
          };
          auto acquire = [&]() -> uint32_t {             // the box used two stores ago has been read by the TMA
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
            return my_stage + (nbox & 1) * kTcBoxBytes;
          };
          if constexpr (sizeof(OutT) == 8) {
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
          TcOut<OutT>::store(op, va, sc, w0);
          if (w1 > 0) TcOut<OutT>::store(op + 32, vb, sc, w1);
        }
        if (probe) t_store += clock64() - t_s0;
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // staging boxes drained before exit
    if (probe && warp == 8 && lane == 0) {
      unsigned long long* pr = probe + blockIdx.x * 16;
      pr[0] = clock64() - t_ebegin; pr[4] = t_wwait; pr[5] = t_store; pr[12] = n_done; pr[14] = t_swait;
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

template <typename OutT>
static int launch_tc(const AggTcArgs& a, const CUtensorMap& map, int n_items, int J, int grid, cudaStream_t st) {
  const size_t smem = agg_tc_smem();
  // output [S_total][K*D] as a 2-D tensor: the epilogue stores 32-row x 128-byte boxes through the TMA
  CUtensorMap map_out;
  {
    PFN_encodeTiled enc = get_encode();
    cuuint64_t dims[2] = {(cuuint64_t)a.K * a.D, (cuuint64_t)a.S_total};
    cuuint64_t strides[1] = {(cuuint64_t)a.K * a.D * sizeof(OutT)};
    cuuint32_t box[2] = {(cuuint32_t)(128 / sizeof(OutT)), 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&map_out, sizeof(OutT) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, a.out,
                     dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (output) failed (%d)", (int)r); return SEGVLAD_ECUDA; }
  }
  CUtensorMap map_lin;
  {
    PFN_encodeTiled enc = get_encode();
    cuuint64_t dims[2] = {(cuuint64_t)(128 / sizeof(OutT)), (cuuint64_t)a.K * a.D * a.S_total * sizeof(OutT) / 128};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {(cuuint32_t)(128 / sizeof(OutT)), 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&map_lin, sizeof(OutT) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, a.out,
                     dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (lin) failed (%d)", (int)r); return SEGVLAD_ECUDA; }
  }
  const char* ee = getenv("SEGVLAD_AGG_EXP");   // development: timing experiments with parts of the kernel switched off
  const int exp = ee ? atoi(ee) : 0;
  SV_CHECK_CUDA(cudaFuncSetAttribute(aggregate_tc_kernel<OutT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int pslot = prof_begin(SEGVLAD_PROF_AGGREGATE, st);
  aggregate_tc_kernel<OutT><<<grid, kTcThreadsAgg, smem, st>>>(map, map_out, map_lin, a.tile_tbl, a.cl_ptr, a.memS, a.cpred, a.B, a.N, a.D, a.K,
                                                             n_items, J, reinterpret_cast<OutT*>(a.out), a.norms, a.probe, exp);
  prof_end(pslot, st);
  SV_CHECK_LAUNCH();
  return SEGVLAD_OK;
}

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

  static_assert(kTcNpAlign == 64, "rt_planes_kernel tiles 64 tokens");
  rt_planes_kernel<<<dim3(Np / 64, (a.D + 31) / 32, a.B), 256, 0, st>>>(a.R, a.cl_tok, a.B, a.N, a.D, Np, a.RT);
  SV_CHECK_LAUNCH();

  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return SEGVLAD_ECUDA; }
  CUtensorMap map;
  cuuint64_t dims[2] = {(cuuint64_t)Np, (cuuint64_t)3 * a.B * a.D};
  cuuint64_t strides[1] = {(cuuint64_t)Np * 2};
  cuuint32_t box[2] = {(cuuint32_t)kTcTokChunk, (cuuint32_t)kTcPassN};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a.RT, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return SEGVLAD_ECUDA; }

  const int P = (a.D + kTcPassN - 1) / kTcPassN;
  int J = 1;
  while ((long long)nt * a.K * J < 2ll * num_sms && J * 2 <= P) J *= 2;   // few items: split the write sweep over channels
  const long long n_items = (long long)nt * a.K * J;
  SV_REQUIRE(n_items < (1ll << 31), "aggregate: too many work items");
  const int grid = (int)(n_items < num_sms ? n_items : num_sms);
  return a.out_dtype == SEGVLAD_OUT_F64 ? launch_tc<double>(a, map, (int)n_items, J, grid, st)
                                        : launch_tc<float>(a, map, (int)n_items, J, grid, st);
}

}  // namespace segvlad
