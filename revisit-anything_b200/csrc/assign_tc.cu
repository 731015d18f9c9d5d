// Cosine cluster assignment on the 5th-generation tensor cores (sm_100a) -- SURVEY 8 row a2,
// `labels = argmax_k( x . c_hat_k )` of vlad_single (func_vpr.py:1145-1146), tokens in the reference's [D][N] layout.
//
// The product is a small dense GEMM  S [N x K] = X^T [N x D] . C_hat^T [D x K]  (4.8 GFLOP for a 16-image batch).  The SIMT
// kernel (aggregate.cu: split-K register tiles + finalize) runs it at ~20 TFLOP/s (r1 launch list: 248 + 74 us, a third of
// the batch).  Here it runs on tcgen05 with fp32-equivalent operands: x and c_hat are split into three bf16 pieces
// (hi + mid + lo = all 24 mantissa bits) and the six products down to 2^-16 (hh, hm, mh, hl, lh, mm) are accumulated in
// fp32 TMEM (hh and the five small ones in separate accumulators) -- the dropped terms are <= 2^-23 relative, the rounding
// level of the reference's SGEMM.
//   * CTA = (image, 128 tokens).  The tokens arrive channel-major, so the A operand (tokens x channels, K-major) is
//     written by 8 converter warps: coalesced LDG along the tokens, split, 16-byte st.shared into the SWIZZLE_128B
//     layout (thread = token row: conflict-free), ||x||^2 accumulated on the way.
//   * B = c_hat planes [3][Kp][D] bf16 (prepared once per call), one TMA box per plane and 64-channel stage.
//   * warp 0 issues the MMAs (M=128, N=Kp, K=16), 2-stage ring; the epilogue reads the accumulator row of each token
//     (thread = TMEM lane) and takes the argmax in registers: first index on ties, as torch.argmax.
#include <stdlib.h>

#include "aggregate_tc.cuh"
#include "tc_ptx.cuh"

namespace segvlad {

constexpr int kAtTok = 128;                 // tokens per CTA (M)
constexpr int kAtCh = 64;                   // channels per stage (one 128-byte swizzle row of bf16)
constexpr int kAtStages = 2;
constexpr int kAtThreads = 320;             // warp 0 MMA, warps 1-8 converters (1-4 also epilogue), warp 9 TMA producer
constexpr uint32_t kAtATile = kAtTok * kAtCh * 2;   // 16 KB per plane
constexpr float kEpsAt = 1e-12f;

__host__ __device__ inline size_t assign_tc_stage_bytes(int Kp) { return 3 * (size_t)kAtATile + 3 * (size_t)Kp * kAtCh * 2; }
__host__ __device__ inline size_t assign_tc_smem(int Kp) {
  // stage size is a multiple of 1024 (Kp % 16 == 0 -> Kp * 128 B % 2048 == 0)
  return 1024 + kAtStages * assign_tc_stage_bytes(Kp) + 128 + 2 * kAtTok * sizeof(float);
}

// c_hat = c / max(||c||, eps) split into bf16 planes [3][Kp][D] (lo, mid, hi); rows k >= K are zero.
__global__ void __launch_bounds__(256)
chat_planes_kernel(const float* __restrict__ c, int K, int Kp, int D, __nv_bfloat16* __restrict__ planes) {
  const int k = blockIdx.x;
  __shared__ float s_w[8];
  __shared__ float s_n;
  float ss = 0.f;
  if (k < K)
    for (int d = threadIdx.x; d < D; d += 256) { const float v = c[(size_t)k * D + d]; ss += v * v; }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += s_w[i];
    s_n = fmaxf(sqrtf(t), kEpsAt);
  }
  __syncthreads();
  const float nrm = s_n;
  const size_t plane = (size_t)Kp * D;
  for (int d = threadIdx.x; d < D; d += 256) {
    const float x = k < K ? c[(size_t)k * D + d] / nrm : 0.f;
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    const float r = x - __bfloat162float(h);
    const __nv_bfloat16 m = __float2bfloat16_rn(r);
    const __nv_bfloat16 l = __float2bfloat16_rn(r - __bfloat162float(m));
    const size_t o = (size_t)k * D + d;
    planes[o] = l; planes[plane + o] = m; planes[2 * plane + o] = h;
  }
}

// two fp32 -> packed bf16x2 pieces (lo, mid, hi); element 0 in the low half
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& lo, uint32_t& mid, uint32_t& hi) {
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

__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(kAtThreads, 1)
assign_tc_kernel(const __grid_constant__ CUtensorMap map_c, const float* __restrict__ tokens, int N, int D, int K, int Kp,
                 int tmem_cols, int prenorm, int* __restrict__ labels, float* __restrict__ nrm_out) {
  extern __shared__ __align__(1024) uint8_t at_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at_smem_raw) + 1023) & ~(uintptr_t)1023);
  const size_t stage_bytes = assign_tc_stage_bytes(Kp);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kAtStages * stage_bytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  float* s_ssq = reinterpret_cast<float*>(bars + 16);          // [2 channel halves][128 tokens]
  const uint32_t bar_full = smem_u32(bars + 0);     // [stages] A written (8 converter warps) + B landed (TMA)
  const uint32_t bar_empty = smem_u32(bars + 2);    // [stages] MMAs that read the stage retired
  const uint32_t bar_done = smem_u32(bars + 4);     // accumulator complete
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, p0 = blockIdx.x * kAtTok;
  const int n_st = (D + kAtCh - 1) / kAtCh;
  const bool stacked = 6 * Kp <= 512;          // every product in its own TMEM columns (three wide MMAs per k-step)

  if (threadIdx.x == 0) {
    for (int i = 0; i < kAtStages; ++i) { mbar_init(bar_full + 8 * i, 9); mbar_init(bar_empty + 8 * i, 1); }
    mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 9) {
    // ===================== TMA producer: c_hat planes =====================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c) : "memory");
      for (int s = 0; s < n_st; ++s) {
        const int stage = s % kAtStages, use = s / kAtStages;
        mbar_wait(bar_empty + 8 * stage, (use & 1) ^ 1);
        const uint32_t sb = smem_u32(smem + stage * stage_bytes) + 3 * kAtATile;
        const uint32_t fb = bar_full + 8 * stage;
        const uint32_t btile = (uint32_t)Kp * kAtCh * 2;
        mbar_arrive_expect_tx(fb, 3 * btile);
#pragma unroll
        for (int pl = 0; pl < 3; ++pl) tma_load_2d(sb + pl * btile, &map_c, fb, s * kAtCh, pl * Kp);
      }
    }
  } else if (warp == 0) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // kind::f16: D=f32 (bit 4), A=B=bf16 (bits 7, 10), both K-major, N>>3 @17, M>>4 @24
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(Kp >> 3) << 17) | ((uint32_t)(kAtTok >> 4) << 24);
      const uint32_t btile = (uint32_t)Kp * kAtCh * 2;
      for (int s = 0; s < n_st; ++s) {
        const int stage = s % kAtStages, use = s / kAtStages;
        mbar_wait(bar_full + 8 * stage, use & 1);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * stage_bytes);
        const uint32_t sb = sa + 3 * kAtATile;
        // planes: 0 = lo, 1 = mid, 2 = hi.  TMEM adds truncate (every accumulating MMA costs up to an ulp of the running
        // sum, csrc/project_tc.cu), so the main product (h,h) never shares an accumulator with the small ones.
        if (stacked) {
          // The three c_hat planes of a stage are consecutive K-major row blocks: ONE MMA per token piece covers several
          // of them (x_hi . [lo | mid | hi] with N = 3 Kp, x_mid . [mid | hi] with N = 2 Kp, x_lo . [hi] with N = Kp) --
          // three MMAs per k-step instead of six (the small N = Kp MMAs are latency-bound, ~125 cycles each), every
          // product in its own accumulator columns: [0,3Kp) = hl hm hh, [3Kp,5Kp) = mm mh, [5Kp,6Kp) = lh.
          const uint32_t ibase = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kAtTok >> 4) << 24);
#pragma unroll
          for (int kk = 0; kk < kAtCh / 16; ++kk) {
            const uint64_t adv = (uint64_t)(kk * 32 >> 4);
            const uint32_t acc = (s | kk) != 0;
            tc_mma_bf16(tmem_base + 5 * Kp, umma_desc_sw128(sa) + adv, umma_desc_sw128(sb + 2 * btile) + adv,
                        ibase | ((uint32_t)(Kp >> 3) << 17), acc);
            tc_mma_bf16(tmem_base + 3 * Kp, umma_desc_sw128(sa + kAtATile) + adv, umma_desc_sw128(sb + btile) + adv,
                        ibase | ((uint32_t)(2 * Kp >> 3) << 17), acc);
            tc_mma_bf16(tmem_base, umma_desc_sw128(sa + 2 * kAtATile) + adv, umma_desc_sw128(sb) + adv,
                        ibase | ((uint32_t)(3 * Kp >> 3) << 17), acc);
          }
        } else {
          // many clusters (6 Kp > 512 TMEM columns): (h,h) in [0,Kp), the five small products chained in [Kp,2Kp)
          const int pa[5] = {0, 2, 1, 1, 2}, pb[5] = {2, 0, 1, 2, 1};
#pragma unroll
          for (int kk = 0; kk < kAtCh / 16; ++kk) {
            const uint64_t adv = (uint64_t)(kk * 32 >> 4);
#pragma unroll
            for (int q = 0; q < 5; ++q)
              tc_mma_bf16(tmem_base + Kp, umma_desc_sw128(sa + pa[q] * kAtATile) + adv, umma_desc_sw128(sb + pb[q] * btile) + adv,
                          idesc, (s | kk | q) != 0);
            tc_mma_bf16(tmem_base, umma_desc_sw128(sa + 2 * kAtATile) + adv, umma_desc_sw128(sb + 2 * btile) + adv, idesc,
                        (s | kk) != 0);
          }
        }
        tc_commit(bar_empty + 8 * stage);
      }
      tc_commit(bar_done);
    }
  } else {
    // ===================== converters (warps 1-8): thread = token row x half of the stage's channels =====================
    const int u = threadIdx.x - 32;              // 0 .. 255
    const int t = u & (kAtTok - 1), h = u >> 7;  // token row of the tile, channel half (32 channels)
    const int p = p0 + t;
    const bool valid = p < N;
    const float* tok = tokens + (size_t)b * D * N + p;
    float ssq = 0.f;
    float x[32], xn[32];   // this stage's channels and the next stage's (loads in flight while this one is converted)
#pragma unroll
    for (int j = 0; j < 32; ++j) xn[j] = (valid && h * 32 + j < D) ? __ldg(tok + (size_t)(h * 32 + j) * N) : 0.f;
    for (int s = 0; s < n_st; ++s) {
      const int stage = s % kAtStages, use = s / kAtStages;
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = xn[j];
      if (s + 1 < n_st) {
        const int d1 = (s + 1) * kAtCh + h * 32;
#pragma unroll
        for (int j = 0; j < 32; ++j) xn[j] = (valid && d1 + j < D) ? __ldg(tok + (size_t)(d1 + j) * N) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) ssq = fmaf(x[j], x[j], ssq);
      mbar_wait(bar_empty + 8 * stage, (use & 1) ^ 1);
      const uint32_t sa = smem_u32(smem + stage * stage_bytes);
#pragma unroll
      for (int c = 0; c < 4; ++c) {              // 16-byte chunk = 8 channels; chunk index h*4+c stored at (..) ^ (row & 7)
        uint32_t lo[4], mid[4], hi[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) split2(x[8 * c + 2 * e], x[8 * c + 2 * e + 1], lo[e], mid[e], hi[e]);
        const uint32_t addr = sa + (uint32_t)(t >> 3) * 1024 + (uint32_t)(t & 7) * 128 + (uint32_t)(((h * 4 + c) ^ (t & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr + kAtATile), "r"(mid[0]), "r"(mid[1]), "r"(mid[2]), "r"(mid[3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr + 2 * kAtATile), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_full + 8 * stage);
    }
    s_ssq[h * kAtTok + t] = ssq;
    asm volatile("bar.sync 1, 256;" ::: "memory");    // the 8 converter warps
    if (warp <= 4) {
      // ===================== epilogue (warps 1-4): thread = token = TMEM lane =====================
      const int quarter = warp & 3;
      const int row = quarter * 32 + lane;
      const int pr = p0 + row;
      mbar_wait(bar_done, 0);
      tc_fence_after();
      float bestv = -INFINITY;
      int besti = 0;
      for (int c0 = 0; c0 < Kp; c0 += 16) {
        const uint32_t tl = tmem_base + ((uint32_t)(quarter * 32) << 16) + c0;
        uint32_t v[16], w[16];
        float sm[16];
        if (stacked) {            // small products first: (hl + lh + mm) + (hm + mh), then hh
          tc_ld16(tl, v);                      // hl
          tc_ld16(tl + 5 * Kp, w);             // lh
#pragma unroll
          for (int j = 0; j < 16; ++j) sm[j] = __uint_as_float(v[j]) + __uint_as_float(w[j]);
          tc_ld16(tl + 3 * Kp, v);             // mm
          tc_ld16(tl + Kp, w);                 // hm
#pragma unroll
          for (int j = 0; j < 16; ++j) sm[j] += __uint_as_float(v[j]);
          tc_ld16(tl + 4 * Kp, v);             // mh
#pragma unroll
          for (int j = 0; j < 16; ++j) sm[j] += __uint_as_float(w[j]) + __uint_as_float(v[j]);
          tc_ld16(tl + 2 * Kp, v);             // hh
        } else {
          tc_ld16(tl, v);
          tc_ld16(tl + Kp, w);
#pragma unroll
          for (int j = 0; j < 16; ++j) sm[j] = __uint_as_float(w[j]);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float f = __uint_as_float(v[j]) + sm[j];
          if (c0 + j < K && f > bestv) { bestv = f; besti = c0 + j; }
        }
      }
      if (pr < N) {
        labels[(size_t)b * N + pr] = besti;
        const float tq = s_ssq[row] + s_ssq[kAtTok + row];
        nrm_out[(size_t)b * N + pr] = prenorm ? 1.0f : fmaxf(sqrtf(tq), kEpsAt);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
  }
}

// ------------------------------------------------------------------------------------------------
bool assign_tc_supported(int N, int D, int K) {
  const char* e = getenv("SEGVLAD_ASSIGN_TC");   // "0" selects the SIMT kernels of aggregate.cu (kept as a cross-check)
  const bool on = !(e && e[0] == '0');
  return on && N >= 1 && K >= 1 && K <= 128 && D >= kAtCh && D % 8 == 0;   // TMA: 16-byte row pitch, box inside the tensor
}
size_t assign_tc_workspace_elems(int D, int K) { return (size_t)3 * align_up((size_t)K, 16) * D; }   // bf16 elements

int assign_tc_run(const float* tokens, int B, int N, int D, const float* centers, int K, int prenorm,
                  __nv_bfloat16* chat_planes, int* labels, float* nrm, cudaStream_t st) {
  const int Kp = (int)align_up((size_t)K, 16);
  chat_planes_kernel<<<Kp, 256, 0, st>>>(centers, K, Kp, D, chat_planes);
  SV_CHECK_LAUNCH();
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return SEGVLAD_ECUDA; }
  CUtensorMap map;
  cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)3 * Kp};
  cuuint64_t strides[1] = {(cuuint64_t)D * 2};
  cuuint32_t box[2] = {(cuuint32_t)kAtCh, (cuuint32_t)Kp};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, chat_planes, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (centres) failed (%d)", (int)r); return SEGVLAD_ECUDA; }
  int tmem_cols = 32;
  while (tmem_cols < (6 * Kp <= 512 ? 6 * Kp : 2 * Kp)) tmem_cols *= 2;   // six product accumulators, or main + small
  const size_t smem = assign_tc_smem(Kp);
  SV_CHECK_CUDA(cudaFuncSetAttribute(assign_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  assign_tc_kernel<<<dim3((N + kAtTok - 1) / kAtTok, B), kAtThreads, smem, st>>>(map, tokens, N, D, K, Kp, tmem_cols, prenorm,
                                                                              labels, nrm);
  SV_CHECK_LAUNCH();
  return SEGVLAD_OK;
}

}  // namespace segvlad
