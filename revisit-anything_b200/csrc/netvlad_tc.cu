// NetVLAD + anti-burst aggregation on the 5th-generation tensor cores (sm_100a) -- SURVEY 8 row a9, BASELINE config 5.
//
// Replaces VLAD-BuFF/models/aggregators/aggregation.py:266-361 NetVLAD.forward (antiburst=True; getWeights :148-162).
// Per image the forward is three dense contractions around two row-wise non-linearities:
//     S   = x_hat^T x_hat          [N x N]   self-similarity  (:298)   -> w_p = (sum_q sigmoid(ab_w (2 S_pq - 2) + ab_b))^ab_p
//     L   = x_hat^T W^T            [N x K]   1x1-conv logits  (:289)   -> a_kp = softmax_k(L_pk) / w_p              (:290, :337)
//     V   = a x_hat                [K x D]   weighted sum     (:346-358), V_kd -= c_kd sum_p a_kp; intra-norm; L2   (:359-361)
// netvlad.cu runs them as fp32 FFMA tiles (16 TFLOP/s, correctness-first).  Here they run on tcgen05 with fp32-EQUIVALENT
// operands: every operand is split into two fp16 pieces (hi + lo = 22 mantissa bits; |x_hat|, a <= 1, so fp16's range is
// enough) and three products (hh | hl + lh; the dropped ll term is 2^-22 relative) are accumulated in fp32 TMEM, the main
// product and the two small ones in separate accumulators (TMEM adds truncate: csrc/project_tc.cu).
//   nv_prep_kernel      x [B][D][N] fp32 -> x_hat planes, token-major XH [2][B][Np][Dp] and channel-major XT [2][B][Dp][Np]
//   nv_assign_tc_kernel CTA = (image, 128 tokens): Np/128 self-similarity tiles + one logits tile through one TMA / MMA
//                       pipeline (double-buffered 128-column accumulators); epilogue thread = token row: sigmoid row sum
//                       (the N x N matrix never leaves TMEM), softmax over the K logits, a = softmax / w written as fp16
//                       planes AP [2][B][128][Np] + per-warp partial sums of a over the tokens
//   nv_vlad_tc_kernel   CTA = (image, 128 channels): V tile [128 clusters x 128 channels] over the Np tokens; epilogue
//                       thread = cluster row: V - c * sum_p a  -> V [B][K][D] fp32
//   nv_finalize_kernel  (netvlad.cu) intra-norm + L2.
// Everything is deterministic (no float atomics).  Supported: K <= 128 and K % 16 == 0 (else the FFMA path).
#include <cuda_fp16.h>
#include <stdlib.h>

#include "tc_ptx.cuh"

namespace segvlad {

constexpr int kNvTile = 128;                  // tokens per tile (M) / tile width
constexpr int kNvCh = 64;                     // contraction elements per stage (one 128-byte swizzle row of fp16)
constexpr int kNvStages = 3;
constexpr uint32_t kNvTileB = kNvTile * kNvCh * 2;   // 16 KB: one [128 x 64] fp16 operand tile
constexpr uint32_t kNvStageB = 4 * kNvTileB;         // A hi, A lo, B hi, B lo
constexpr int kNvThreads = 192;               // warp 0 TMA producer, warp 1 MMA issuer (+ TMEM), warps 2-5 epilogue
__host__ __device__ constexpr size_t nv_tc_smem() { return 1024 + (size_t)kNvStages * kNvStageB + 256; }

__device__ __forceinline__ void nv_split_store(float v, __half* hi, __half* lo) {
  const __half h = __float2half_rn(v);
  *hi = h;
  *lo = __float2half_rn(v - __half2float(h));
}

// tokens [B][D][N] -> unit-norm x_hat split into fp16 (hi, lo) planes in both operand layouts, zero padded.
//   XH [2][B][Np][Dp]  row = token (K-major for contractions over the channels)
//   XT [2][B][Dp][Np]  row = channel (K-major for the contraction over the tokens)
// CTA = (image, 32 tokens): the whole [D][32] fp32 slab is staged in shared memory (D x 33 floats, 101 KB at D = 768), so x
// is read from HBM ONCE for both the norms and the planes (r2 launch lists: 2-byte stores 1.06 ms per 512 images; 64-token
// CTAs with 4-byte stores but a second pass over x 1.24 ms -- with ~1200 CTAs' slabs in flight the second read misses L2),
// and written out in both orientations with 4-byte (half2) stores.
__device__ __forceinline__ void nv_split2(float v0, float v1, __half2& hi, __half2& lo) {
  hi = __floats2half2_rn(v0, v1);
  const float2 hf = __half22float2(hi);
  lo = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
}
constexpr int kNvPrepTok = 32;
__global__ void __launch_bounds__(256)
nv_prep_kernel(const float* __restrict__ x, int B, int N, int D, int Np, int Dp, __half* __restrict__ XH,
               __half* __restrict__ XT) {
  extern __shared__ float s_slab[];             // [Dp][33]: channel-major slab of this CTA's 32 tokens (zero padded)
  __shared__ float s_part[8][32];
  __shared__ float s_nrm[32];
  const int b = blockIdx.y, p0 = blockIdx.x * kNvPrepTok;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const float* xb = x + (size_t)b * D * N;
  const int p = p0 + lane;
  const bool ok = p < N;
  // warp w: channels w, w + 8, ...; lane = token: 128-byte row pieces.  All of a thread's 4-byte copies are issued as
  // cp.async before anything waits (r2: the first version's register loads kept ~8 KB in flight per SM and ran the kernel at
  // 2.5 TB/s); padding tokens / channels are zero-filled by a copy of source size 0.
  for (int d = w; d < Dp; d += 8) {
    const bool in = d < D && ok;
    const float* src = in ? xb + (size_t)d * N + p : xb;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(s_slab + d * 33 + lane)), "l"(src),
                 "r"(in ? 4 : 0) : "memory");
  }
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
  float ss = 0.f;
  for (int d = w; d < Dp; d += 8) {             // (each thread reads back exactly what it copied: no barrier needed yet)
    const float v = s_slab[d * 33 + lane];
    ss = fmaf(v, v, ss);
  }
  s_part[w][lane] = ss;
  __syncthreads();
  if (w == 0) {
    float t = 0.f;
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) t += s_part[ww][lane];
    s_nrm[lane] = fmaxf(sqrtf(t), 1e-12f);      // max(||x_p||, eps) like F.normalize
  }
  __syncthreads();
  // normalise the slab in place: ONE fp32 division per element (x / max(||x||, eps) like F.normalize), by the thread that
  // copied it (r2 ncu: with the division inside both write-out loops the kernel was issue-bound at 58 %)
  {
    const float nl = s_nrm[lane];
    for (int d = w; d < Dp; d += 8) s_slab[d * 33 + lane] = s_slab[d * 33 + lane] / nl;
  }
  __syncthreads();
  const size_t plane_h = (size_t)B * Np * Dp, plane_t = (size_t)B * Dp * Np;
  // channel-major planes: a warp writes two channel rows per step (half-warp = one row of 16 token pairs = 64 bytes)
  {
    const int hw = lane >> 4, l = lane & 15;
    for (int d = 2 * w + hw; d < Dp; d += 16) {
      __half2 hi, lo;
      nv_split2(s_slab[d * 33 + 2 * l], s_slab[d * 33 + 2 * l + 1], hi, lo);
      const size_t o = ((size_t)b * Dp + d) * Np + p0 + 2 * l;
      *reinterpret_cast<__half2*>(XT + o) = hi;
      *reinterpret_cast<__half2*>(XT + plane_t + o) = lo;
    }
  }
  // token-major planes: warp w writes tokens w, w + 8, ...; lane = channel pair, 128 contiguous bytes per store
  for (int t = w; t < kNvPrepTok; t += 8) {
    const size_t row = ((size_t)b * Np + p0 + t) * Dp;
    for (int d = 2 * lane; d < Dp; d += 64) {
      __half2 hi, lo;
      nv_split2(s_slab[d * 33 + t], s_slab[(d + 1) * 33 + t], hi, lo);
      *reinterpret_cast<__half2*>(XH + row + d) = hi;
      *reinterpret_cast<__half2*>(XH + plane_h + row + d) = lo;
    }
  }
}

// W [K][D] fp32 -> planes [2][Kp][Dp] fp16 (hi, lo), zero padded
__global__ void __launch_bounds__(256)
nv_wplanes_kernel(const float* __restrict__ W, int K, int D, int Kp, int Dp, __half* __restrict__ WP) {
  const size_t n = (size_t)Kp * Dp;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    const int k = (int)(i / Dp), d = (int)(i % Dp);
    const float v = (k < K && d < D) ? W[(size_t)k * D + d] : 0.f;
    nv_split_store(v, WP + i, WP + n + i);
  }
}

__device__ __forceinline__ void nv_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void nv_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// kind::f16 instruction descriptor: D = f32 (bit 4), A = B = f16 (format 0), both K-major, N >> 3 @17, M >> 4 @24
__device__ __forceinline__ uint32_t nv_idesc(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kNvTile >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------
// Self-similarity row sums + soft assignment.  Tiles t = 0 .. n_qt-1: S[p-tile, q-tile t]; tile n_qt: logits (B = W planes).
__global__ void __launch_bounds__(kNvThreads, 1)
nv_assign_tc_kernel(const __grid_constant__ CUtensorMap map_xh, const __grid_constant__ CUtensorMap map_w, int B, int N, int Np,
                    int Dp, int K, int Kp, float ab_w, float ab_b, float ab_p, __half* __restrict__ AP,
                    float* __restrict__ asum_part) {
  extern __shared__ __align__(1024) uint8_t nv_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(nv_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kNvStages * kNvStageB);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  const uint32_t bar_full = smem_u32(bars + 0);      // [stages]
  const uint32_t bar_empty = smem_u32(bars + 4);     // [stages]
  const uint32_t bar_tfull = smem_u32(bars + 8);     // [2] accumulator pair complete
  const uint32_t bar_tempty = smem_u32(bars + 10);   // [2] accumulator pair drained (4 epilogue warps)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, pt = blockIdx.x, p0 = pt * kNvTile;
  const int n_qt = Np / kNvTile, n_kb = Dp / kNvCh;
  const int n_tiles = n_qt + 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kNvStages; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, 4); }
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
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_xh) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
      uint32_t stage = 0, phase = 0;
      for (int t = 0; t < n_tiles; ++t) {
        const bool logits = t == n_qt;
        const uint32_t bbytes = logits ? (uint32_t)Kp * kNvCh * 2 : kNvTileB;
        for (int kb = 0; kb < n_kb; ++kb) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t sa = smem_u32(smem + stage * kNvStageB);
          const uint32_t fb = bar_full + 8 * stage;
          mbar_arrive_expect_tx(fb, 2 * kNvTileB + 2 * bbytes);
#pragma unroll
          for (int pl = 0; pl < 2; ++pl) {
            tma_load_2d(sa + pl * kNvTileB, &map_xh, fb, kb * kNvCh, (pl * B + b) * Np + p0);
            if (logits) tma_load_2d(sa + (2 + pl) * kNvTileB, &map_w, fb, kb * kNvCh, pl * Kp);
            else tma_load_2d(sa + (2 + pl) * kNvTileB, &map_xh, fb, kb * kNvCh, (pl * B + b) * Np + t * kNvTile);
          }
          if (++stage == kNvStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int t = 0; t < n_tiles; ++t) {
        const uint32_t buf = t & 1, use = t >> 1;
        const uint32_t idesc = nv_idesc(t == n_qt ? Kp : kNvTile);
        mbar_wait(bar_tempty + 8 * buf, (use & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_main = tmem_base + buf * 256, d_small = d_main + 128;
        for (int kb = 0; kb < n_kb; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint64_t a_hi = umma_desc_sw128(smem_u32(smem + stage * kNvStageB));
          const uint64_t a_lo = a_hi + (kNvTileB >> 4), b_hi = a_hi + 2 * (kNvTileB >> 4), b_lo = a_hi + 3 * (kNvTileB >> 4);
#pragma unroll
          for (int kk = 0; kk < kNvCh / 16; ++kk) {
            const uint64_t adv = (uint64_t)(kk * 2);              // 16 fp16 = 32 bytes along K
            const uint32_t acc = (uint32_t)((kb | kk) != 0);
            tc_mma_bf16(d_small, a_hi + adv, b_lo + adv, idesc, acc);    // (kind::f16 wrapper; operand type from idesc)
            tc_mma_bf16(d_small, a_lo + adv, b_hi + adv, idesc, 1u);
            tc_mma_bf16(d_main, a_hi + adv, b_hi + adv, idesc, acc);
          }
          tc_commit(bar_empty + 8 * stage);
          if (++stage == kNvStages) { stage = 0; phase ^= 1; }
        }
        tc_commit(bar_tfull + 8 * buf);
      }
    }
  } else {
    // ===================== epilogue: thread = token row = TMEM lane =====================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int p = p0 + row;
    const bool valid = p < N;
    const uint32_t tl = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const float za = 2.f * ab_w, zb = ab_b - 2.f * ab_w;       // z = ab_w (2 S - 2) + ab_b
    float wsum = 0.f;
    for (int t = 0; t < n_qt; ++t) {
      const uint32_t buf = t & 1, use = t >> 1;
      mbar_wait(bar_tfull + 8 * buf, use & 1);
      tc_fence_after();
      const int qn = min(kNvTile, N - t * kNvTile);             // valid columns of this tile (> 0: Np - N < 128)
      for (int c = 0; c < kNvTile; c += 16) {
        uint32_t vm[16], vs[16];
        nv_ld16(tl + buf * 256 + c, vm);
        nv_ld16(tl + buf * 256 + 128 + c, vs);
        nv_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float s = __uint_as_float(vm[j]) + __uint_as_float(vs[j]);
          const float z = fmaf(s, za, zb);
          const float sg = 1.f / (1.f + __expf(-z));
          if (c + j < qn) wsum += sg;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
    }
    const float w = (ab_p == 1.f) ? wsum : powf(wsum, ab_p);
    // ---- logits tile: softmax over the K clusters, a = softmax / w ----
    {
      const uint32_t buf = n_qt & 1, use = n_qt >> 1;
      mbar_wait(bar_tfull + 8 * buf, use & 1);
      tc_fence_after();
      const uint32_t tm = tl + buf * 256;
      float mx = -INFINITY;
      for (int c = 0; c < Kp; c += 16) {
        uint32_t vm[16], vs[16];
        nv_ld16(tm + c, vm);
        nv_ld16(tm + 128 + c, vs);
        nv_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (c + j < K) mx = fmaxf(mx, __uint_as_float(vm[j]) + __uint_as_float(vs[j]));
      }
      float den = 0.f;
      for (int c = 0; c < Kp; c += 16) {
        uint32_t vm[16], vs[16];
        nv_ld16(tm + c, vm);
        nv_ld16(tm + 128 + c, vs);
        nv_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (c + j < K) den += expf(__uint_as_float(vm[j]) + __uint_as_float(vs[j]) - mx);
      }
      const float scale = valid ? 1.f / (den * w) : 0.f;        // padding tokens get a = 0
      const size_t plane = (size_t)B * kNvTile * Np;
      __half* ap = AP + (size_t)b * kNvTile * Np + p;             // AP[pl][b][k][p]
      float* part = asum_part + ((size_t)b * (Np / 32) + (size_t)(pt * 4 + quarter)) * kNvTile;
      for (int c = 0; c < Kp; c += 16) {
        uint32_t vm[16], vs[16];
        nv_ld16(tm + c, vm);
        nv_ld16(tm + 128 + c, vs);
        nv_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int k = c + j;
          float a = 0.f;
          if (k < K) a = expf(__uint_as_float(vm[j]) + __uint_as_float(vs[j]) - mx) * scale;
          nv_split_store(a, ap + (size_t)k * Np, ap + plane + (size_t)k * Np);
          const float tot = warp_sum(a);                         // partial sum of a_k over this warp's 32 tokens
          if (lane == 0) part[k] = tot;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
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
// V[b][k][d0 .. d0+127] = sum_p a[k][p] x_hat[p][d] - c[k][d] sum_p a[k][p]
__global__ void __launch_bounds__(kNvThreads, 1)
nv_vlad_tc_kernel(const __grid_constant__ CUtensorMap map_ap, const __grid_constant__ CUtensorMap map_xt, int B, int Np, int Dp,
                  int D, int K, const float* __restrict__ cent, const float* __restrict__ asum_part, float* __restrict__ V) {
  extern __shared__ __align__(1024) uint8_t nv_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(nv_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kNvStages * kNvStageB);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  const uint32_t bar_full = smem_u32(bars + 0), bar_empty = smem_u32(bars + 4), bar_done = smem_u32(bars + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, d0 = blockIdx.x * kNvTile;
  const int n_kb = Np / kNvCh;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kNvStages; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_ap) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_xt) : "memory");
      for (int kb = 0; kb < n_kb; ++kb) {
        const int stage = kb % kNvStages, use = kb / kNvStages;
        mbar_wait(bar_empty + 8 * stage, (use & 1) ^ 1);
        const uint32_t sa = smem_u32(smem + stage * kNvStageB);
        const uint32_t fb = bar_full + 8 * stage;
        mbar_arrive_expect_tx(fb, 4 * kNvTileB);
#pragma unroll
        for (int pl = 0; pl < 2; ++pl) {
          tma_load_2d(sa + pl * kNvTileB, &map_ap, fb, kb * kNvCh, (pl * B + b) * kNvTile);
          tma_load_2d(sa + (2 + pl) * kNvTileB, &map_xt, fb, kb * kNvCh, (pl * B + b) * Dp + d0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = nv_idesc(kNvTile);
      const uint32_t d_main = tmem_base, d_small = tmem_base + 128;
      for (int kb = 0; kb < n_kb; ++kb) {
        const int stage = kb % kNvStages, use = kb / kNvStages;
        mbar_wait(bar_full + 8 * stage, use & 1);
        tc_fence_after();
        const uint64_t a_hi = umma_desc_sw128(smem_u32(smem + stage * kNvStageB));
        const uint64_t a_lo = a_hi + (kNvTileB >> 4), b_hi = a_hi + 2 * (kNvTileB >> 4), b_lo = a_hi + 3 * (kNvTileB >> 4);
#pragma unroll
        for (int kk = 0; kk < kNvCh / 16; ++kk) {
          const uint64_t adv = (uint64_t)(kk * 2);
          const uint32_t acc = (uint32_t)((kb | kk) != 0);
          tc_mma_bf16(d_small, a_hi + adv, b_lo + adv, idesc, acc);
          tc_mma_bf16(d_small, a_lo + adv, b_hi + adv, idesc, 1u);
          tc_mma_bf16(d_main, a_hi + adv, b_hi + adv, idesc, acc);
        }
        tc_commit(bar_empty + 8 * stage);
      }
      tc_commit(bar_done);
    }
  } else {
    // epilogue: thread = cluster row
    const int quarter = warp & 3;
    const int k = quarter * 32 + lane;
    const bool vec4 = (D & 3) == 0 && ((reinterpret_cast<uintptr_t>(V) | reinterpret_cast<uintptr_t>(cent)) & 15) == 0;
    float asum = 0.f;
    if (k < K)
      for (int i = 0; i < Np / 32; ++i) asum += asum_part[((size_t)b * (Np / 32) + i) * kNvTile + k];   // fixed order
    mbar_wait(bar_done, 0);
    tc_fence_after();
    const uint32_t tl = tmem_base + ((uint32_t)(quarter * 32) << 16);
    for (int c = 0; c < kNvTile; c += 16) {
      uint32_t vm[16], vs[16];
      nv_ld16(tl + c, vm);
      nv_ld16(tl + 128 + c, vs);
      nv_ld_wait();
      if (k < K) {
        float* o = V + ((size_t)b * K + k) * D + d0 + c;
        const float* cr = cent + (size_t)k * D + d0 + c;
        if (vec4 && d0 + c + 16 <= D) {          // 64 contiguous bytes per thread: whole sectors, 16-byte accesses
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 cv = __ldg(reinterpret_cast<const float4*>(cr + j));
            float4 r;
            r.x = (__uint_as_float(vm[j]) + __uint_as_float(vs[j])) - cv.x * asum;
            r.y = (__uint_as_float(vm[j + 1]) + __uint_as_float(vs[j + 1])) - cv.y * asum;
            r.z = (__uint_as_float(vm[j + 2]) + __uint_as_float(vs[j + 2])) - cv.z * asum;
            r.w = (__uint_as_float(vm[j + 3]) + __uint_as_float(vs[j + 3])) - cv.w * asum;
            *reinterpret_cast<float4*>(o + j) = r;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (d0 + c + j < D) o[j] = (__uint_as_float(vm[j]) + __uint_as_float(vs[j])) - cr[j] * asum;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
  }
}

// ------------------------------------------------------------------------------------------------
bool nv_tc_supported(int N, int D, int K) {
  const char* e = getenv("SEGVLAD_NETVLAD_TC");   // "0" selects the fp32 FFMA kernels of netvlad.cu (kept as the cross-check)
  const bool on = !(e && e[0] == '0');
  return on && N >= 1 && D >= 8 && D <= 1536 && K >= 16 && K <= 128 && K % 16 == 0;   // D: shared-memory slab of nv_prep_kernel
}
static inline int nv_np(int N) { return (int)align_up((size_t)N, kNvTile); }
static inline int nv_dp(int D) { return (int)align_up((size_t)D, kNvTile); }   // 128: also the channel tile of the V kernel

size_t nv_tc_workspace_bytes(int B, int N, int D, int K) {
  Carver c(nullptr);
  const int Np = nv_np(N), Dp = nv_dp(D), Kp = (int)align_up((size_t)K, 16);
  c.take<__half>((size_t)2 * B * Np * Dp);        // XH
  c.take<__half>((size_t)2 * B * Dp * Np);        // XT
  c.take<__half>((size_t)2 * Kp * Dp);            // W planes
  c.take<__half>((size_t)2 * B * kNvTile * Np);   // AP
  c.take<float>((size_t)B * (Np / 32) * kNvTile); // partial sums of a
  c.take<float>((size_t)B * K * D);               // V
  return c.total();
}

static int nv_map(CUtensorMap* m, const void* base, uint64_t cols, uint64_t rows, uint32_t box_rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return SEGVLAD_ECUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)kNvCh, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (netvlad) failed (%d)", (int)r); return SEGVLAD_ECUDA; }
  return SEGVLAD_OK;
}

// defined in netvlad.cu
void nv_finalize_launch(const float* V, int B, int D, int K, float* out, cudaStream_t st);

int nv_tc_run(const float* x, int B, int N, int D, const float* centroids, const float* conv_weight, int K, float ab_w,
              float ab_b, float ab_p, float* out, void* workspace, cudaStream_t st) {
  const int Np = nv_np(N), Dp = nv_dp(D), Kp = (int)align_up((size_t)K, 16);
  Carver c(workspace);
  __half* XH = c.take<__half>((size_t)2 * B * Np * Dp);
  __half* XT = c.take<__half>((size_t)2 * B * Dp * Np);
  __half* WP = c.take<__half>((size_t)2 * Kp * Dp);
  __half* AP = c.take<__half>((size_t)2 * B * kNvTile * Np);
  float* asum_part = c.take<float>((size_t)B * (Np / 32) * kNvTile);
  float* V = c.take<float>((size_t)B * K * D);
  // cluster rows K .. 127 of AP are never written by the assignment kernel: the V kernel's A tile must read zeros there
  if (K < kNvTile) SV_CHECK_CUDA(cudaMemsetAsync(AP, 0, sizeof(__half) * (size_t)2 * B * kNvTile * Np, st));
  const size_t psmem = (size_t)Dp * 33 * sizeof(float);
  SV_REQUIRE(psmem <= 200 * 1024, "netvlad: D too large for the tensor-core path");
  SV_CHECK_CUDA(cudaFuncSetAttribute(nv_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
  nv_prep_kernel<<<dim3(Np / kNvPrepTok, B), 256, psmem, st>>>(x, B, N, D, Np, Dp, XH, XT);
  SV_CHECK_LAUNCH();
  nv_wplanes_kernel<<<64, 256, 0, st>>>(conv_weight, K, D, Kp, Dp, WP);
  SV_CHECK_LAUNCH();
  CUtensorMap map_xh, map_w, map_ap, map_xt;
  int rc;
  if ((rc = nv_map(&map_xh, XH, Dp, (uint64_t)2 * B * Np, kNvTile))) return rc;
  if ((rc = nv_map(&map_w, WP, Dp, (uint64_t)2 * Kp, (uint32_t)Kp))) return rc;
  if ((rc = nv_map(&map_ap, AP, Np, (uint64_t)2 * B * kNvTile, kNvTile))) return rc;
  if ((rc = nv_map(&map_xt, XT, Np, (uint64_t)2 * B * Dp, kNvTile))) return rc;
  const size_t smem = nv_tc_smem();
  SV_CHECK_CUDA(cudaFuncSetAttribute(nv_assign_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  SV_CHECK_CUDA(cudaFuncSetAttribute(nv_vlad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  nv_assign_tc_kernel<<<dim3(Np / kNvTile, B), kNvThreads, smem, st>>>(map_xh, map_w, B, N, Np, Dp, K, Kp, ab_w, ab_b, ab_p, AP,
                                                                      asum_part);
  SV_CHECK_LAUNCH();
  nv_vlad_tc_kernel<<<dim3(Dp / kNvTile, B), kNvThreads, smem, st>>>(map_ap, map_xt, B, Np, Dp, D, K, centroids, asum_part, V);
  SV_CHECK_LAUNCH();
  nv_finalize_launch(V, B, D, K, out, st);
  SV_CHECK_LAUNCH();
  return SEGVLAD_OK;
}

}  // namespace segvlad
