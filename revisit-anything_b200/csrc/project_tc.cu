// PCA-whitening projection on the 5th-generation tensor cores (sm_100a) -- SURVEY 8f row f1,
//     Y = ((X - mean) @ components^T) / sqrt(explained_variance)      (func_vpr.py:1419-1443, sklearn PCA.transform, whiten)
// X [S, D_in] fp64 (VLAD descriptors), components [D_out, D_in] fp32.  For the published configuration (49152 -> 1024) this
// is 100 GFLOP per 1000 segments: the fp64 CUDA-core kernel (project.cu, kept as the cross-check) makes it the slowest
// stage of the PCA pipeline by far.  Here it runs on tcgen05 with fp32-equivalent operands and a two-level accumulation:
//   * (x - mean) is formed in fp64, rounded to fp32 and split into three bf16 pieces (all 24 mantissa bits); the
//     components are split the same way once per model (segvlad_pca_prepare_planes); the six products down to 2^-16
//     (hh, hm, mh, hl, lh, mm) are accumulated in fp32 TMEM -- the main product hh in one accumulator, the five small
//     ones in a second (TMEM adds truncate: every accumulating MMA costs an ulp of the running sum);
//   * fp32 accumulation only runs over CHUNKS of 512 channels (32 main MMAs per element); the chunk sums are added in
//     fp64 to accumulators that live in the other half of TMEM as (lo, hi) word pairs (tcgen05.ld / add / tcgen05.st by
//     the thread that owns the row), so the rounding of a 49152-term sum stays at ~2e-6 relative -- inside the 1e-5
//     descriptor tolerance of north_star (tests/test_gpu_pca.py);
//   * CTA = 128 rows x 128 components x one K split.  X is K-major already (row = segment), so the 16 converter warps read
//     it coalesced (512 contiguous bytes per warp and row), subtract the mean, split and write 4-byte pieces into the
//     SWIZZLE_128B A tile (one 128-byte row per warp store: conflict-free); the component planes come by TMA; warp 0 issues
//     24 MMAs (M=128, N=128, K=16) per 64-channel stage; after every chunk the converter warps (four per TMEM lane
//     quarter) fold the two fp32 chunk sums into the fp64 accumulators and hand them back to the MMA warp;
//   * split K over blockIdx.z for small S; the fp64 partials are summed, scaled by 1/sqrt(ev) and optionally
//     row-normalised by pca_finalize_kernel (project.cu) in a fixed order: deterministic.
#include <stdlib.h>

#include "tc_ptx.cuh"

namespace segvlad {

constexpr int kPtRows = 128;                 // rows per CTA (M)
constexpr int kPtCols = 128;                 // components per CTA (N)
constexpr int kPtCh = 64;                    // channels per stage (one 128-byte swizzle row of bf16)
constexpr int kPtChunk = 8;                  // stages per fp32 accumulation chunk (512 channels)
constexpr int kPtStages = 2;
constexpr int kPtConv = 16;                  // converter warps
constexpr int kPtThreads = 64 + 32 * kPtConv; // warp 0 MMA, warp 1 TMA producer, warps 2-17 converters + accumulator drain
constexpr uint32_t kPtTile = kPtRows * kPtCh * 2;          // 16 KB: one [128 x 64] bf16 operand tile
constexpr uint32_t kPtStageBytes = 6 * kPtTile;            // 3 A planes + 3 B planes

__host__ __device__ constexpr size_t pca_tc_smem() { return 1024 + (size_t)kPtStages * kPtStageBytes + 256; }

// components [Dout][Din] fp32 -> planes [3][Dout_p][Din] bf16 (lo, mid, hi), rows >= Dout zero
__global__ void __launch_bounds__(256)
pca_planes_kernel(const float* __restrict__ W, int Dout, int Dout_p, int Din, __nv_bfloat16* __restrict__ planes) {
  const size_t n = (size_t)Dout_p * Din, plane = n;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    const size_t o = i / Din;
    const float x = o < (size_t)Dout ? W[i] : 0.f;
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    const float r = x - __bfloat162float(h);
    const __nv_bfloat16 m = __float2bfloat16_rn(r);
    const __nv_bfloat16 l = __float2bfloat16_rn(r - __bfloat162float(m));
    planes[i] = l; planes[plane + i] = m; planes[2 * plane + i] = h;
  }
}

__device__ __forceinline__ void pt_split2(float x0, float x1, uint32_t& lo, uint32_t& mid, uint32_t& hi) {
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

__device__ __forceinline__ void pt_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void pt_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
        "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
        "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}

// part[z][S][Dout] (fp64) = sum over this split's channels of (X - mean) . W^T
// kPlanesA: the A operand arrives as ready-made bf16 planes of (X - mean), [3][S][Din] (written by the aggregation kernel's
// SEGVLAD_OUT_PCA_PLANES epilogue), through TMA like the component planes -- no converter work, no fp64 X in HBM (row f1);
// the 16 former converter warps only fold the chunk sums into the fp64 accumulators.
template <bool kPlanesA>
__global__ void __launch_bounds__(kPtThreads, 1)
pca_tc_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x,
              const double* __restrict__ X, const double* __restrict__ mean,
              int S, int Din, int Dout, int Dout_p, int stages_per_split, double* __restrict__ part) {
  extern __shared__ __align__(1024) uint8_t pt_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(pt_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kPtStages * kPtStageBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  const uint32_t bar_full = smem_u32(bars + 0);      // [stages] A written (8 converter warps) + B landed (TMA)
  const uint32_t bar_empty = smem_u32(bars + 2);     // [stages] MMAs that read the stage retired
  const uint32_t bar_tfull = smem_u32(bars + 4);     // chunk accumulators (main, small) complete
  const uint32_t bar_tempty = smem_u32(bars + 6);    // chunk accumulators drained by the converter warps
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * kPtCols, row0 = blockIdx.y * kPtRows;
  const int n_st_total = (Din + kPtCh - 1) / kPtCh;
  const int s_begin = blockIdx.z * stages_per_split;
  const int s_end = min(n_st_total, s_begin + stages_per_split);
  const int n_st = s_end - s_begin;                  // > 0 by construction of the grid
  const int n_chunks = (n_st + kPtChunk - 1) / kPtChunk;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kPtStages; ++i) { mbar_init(bar_full + 8 * i, kPlanesA ? 1 : kPtConv + 1); mbar_init(bar_empty + 8 * i, 1); }
    mbar_init(bar_tfull, 1);
    mbar_init(bar_tempty, kPtConv);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 1) {
    // ===================== TMA producer: component planes =====================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
      for (int s = 0; s < n_st; ++s) {
        const int stage = s % kPtStages, use = s / kPtStages;
        mbar_wait(bar_empty + 8 * stage, (use & 1) ^ 1);
        const uint32_t sb = smem_u32(smem + stage * kPtStageBytes) + 3 * kPtTile;
        const uint32_t fb = bar_full + 8 * stage;
        mbar_arrive_expect_tx(fb, (kPlanesA ? 6 : 3) * kPtTile);
#pragma unroll
        for (int pl = 0; pl < 3; ++pl) tma_load_2d(sb + pl * kPtTile, &map_w, fb, (s_begin + s) * kPtCh, pl * Dout_p + n0);
        if (kPlanesA) {
          // rows beyond S of the last row tile read the next plane's rows (or zero fill past the end): finite garbage in
          // accumulator rows that are never stored
#pragma unroll
          for (int pl = 0; pl < 3; ++pl)
            tma_load_2d(sb - 3 * kPtTile + pl * kPtTile, &map_x, fb, (s_begin + s) * kPtCh, pl * S + row0);
        }
      }
    }
  } else if (warp == 0) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // kind::f16: D=f32 (bit 4), A=B=bf16 (bits 7, 10), both K-major, N>>3 @17, M>>4 @24
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kPtCols >> 3) << 17) | ((uint32_t)(kPtRows >> 4) << 24);
      // planes: 0 = lo, 1 = mid, 2 = hi.  The TMEM accumulation truncates, i.e. every accumulating MMA costs up to one
      // ulp of the RUNNING SUM whatever the size of its addend (measured: 5e-6 relative with all six products chained in
      // one accumulator over a 512-channel chunk).  The main product (h,h) therefore has its own accumulator (32 MMAs per
      // chunk) and the five small ones -- (l,h) (h,l) (m,m) (m,h) (h,m), 2^-8 of the sum -- share a second one, whose
      // truncation errors are 2^-8 smaller; both are added in fp64 when the chunk is drained.
      const int pa[5] = {0, 2, 1, 1, 2}, pb[5] = {2, 0, 1, 2, 1};
      for (int c = 0; c < n_chunks; ++c) {
        mbar_wait(bar_tempty, (c & 1) ^ 1);                 // the previous chunk has been added to the fp64 accumulators
        tc_fence_after();
        const uint32_t d_main = tmem_base, d_small = tmem_base + kPtCols;
        const int cs0 = c * kPtChunk, cs1 = min(n_st, cs0 + kPtChunk);
        for (int s = cs0; s < cs1; ++s) {
          const int stage = s % kPtStages, use = s / kPtStages;
          mbar_wait(bar_full + 8 * stage, use & 1);
          tc_fence_after();
          const uint64_t adesc = umma_desc_sw128(smem_u32(smem + stage * kPtStageBytes));
#pragma unroll
          for (int kk = 0; kk < kPtCh / 16; ++kk) {
            const uint32_t first = (uint32_t)((s > cs0) | kk);
#pragma unroll
            for (int q = 0; q < 5; ++q)
              tc_mma_bf16(d_small, adesc + (uint64_t)(pa[q] * (kPtTile >> 4) + 2 * kk),
                          adesc + (uint64_t)((3 + pb[q]) * (kPtTile >> 4) + 2 * kk), idesc, first | (uint32_t)q);
            tc_mma_bf16(d_main, adesc + (uint64_t)(2 * (kPtTile >> 4) + 2 * kk),
                        adesc + (uint64_t)(5 * (kPtTile >> 4) + 2 * kk), idesc, first);
          }
          tc_commit(bar_empty + 8 * stage);
        }
        tc_commit(bar_tfull);
      }
    }
  } else {
    // ===================== converters + accumulator drain (warps 2-9) =====================
    const int cw = warp - 2;                               // 0 .. 15: rows cw, cw + 16, ... of the tile
    const int quarter = warp & 3, cpart = cw >> 2;         // TMEM lane quarter of this warp, 32-column part
    const int orow = quarter * 32 + lane;                  // output row of this thread within the tile
    // TMEM columns: [0,128) main fp32 chunk sum, [128,256) small-product fp32 chunk sum, [256,512) the fp64 accumulators of
    // the tile as (lo, hi) word pairs -- they live in tensor memory because 32 doubles per thread next to the converter
    // state spilled (ncu: 12 GB of local-memory traffic per call with register accumulators)
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t t_acc = t_lane + 2 * kPtCols + cpart * 64;
    {
      uint32_t z[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) z[j] = 0u;
      pt_st32(t_acc, z);
      pt_st32(t_acc + 32, z);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    auto drain = [&](int c) {                              // chunk c: fp64 accumulators += main + small (fp32)
      mbar_wait(bar_tfull, c & 1);
      tc_fence_after();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t vm[16], vs[16], aw[32];
        pt_ld16(t_lane + cpart * 32 + 16 * h, vm);
        pt_ld16(t_lane + kPtCols + cpart * 32 + 16 * h, vs);
        tc_ld32(t_acc + 32 * h, aw);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          double a = __hiloint2double((int)aw[2 * j + 1], (int)aw[2 * j]);
          a += (double)__uint_as_float(vs[j]);             // small products first
          a += (double)__uint_as_float(vm[j]);
          aw[2 * j] = (uint32_t)__double2loint(a);
          aw[2 * j + 1] = (uint32_t)__double2hiint(a);
        }
        pt_st32(t_acc + 32 * h, aw);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty);
    };
    if (kPlanesA) {
      for (int c = 0; c < n_chunks; ++c) drain(c);
    }
    for (int s = 0; !kPlanesA && s < n_st; ++s) {
      const int stage = s % kPtStages, use = s / kPtStages;
      const int d = (s_begin + s) * kPtCh + 2 * lane;      // this lane's two channels
      const bool dok = d < Din;                            // Din is even (multiple of 8)
      double2 mu = make_double2(0.0, 0.0);
      if (dok) mu = *reinterpret_cast<const double2*>(mean + d);
      // When the slot frees, the MMAs of stage s - 2 have retired: if that was the last stage of a chunk, its sums are
      // complete and are folded in first (before this stage's loads, so that their registers are not live across it)
      const bool fold = s >= 2 && (s - 1) % kPtChunk == 0;
      if (fold) {
        mbar_wait(bar_empty + 8 * stage, (use & 1) ^ 1);
        drain((s - 2) / kPtChunk);
      }
      double2 x[8];                                        // issued before the slot wait: in flight while the MMAs retire
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = row0 + cw + kPtConv * i;
        x[i] = (dok && r < S) ? __ldg(reinterpret_cast<const double2*>(X + (size_t)r * Din + d)) : mu;
      }
      if (!fold) mbar_wait(bar_empty + 8 * stage, (use & 1) ^ 1);
      const uint32_t sa = smem_u32(smem + stage * kPtStageBytes);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int t = cw + kPtConv * i;                    // row of the tile
        uint32_t lo, mid, hi;
        pt_split2((float)(x[i].x - mu.x), (float)(x[i].y - mu.y), lo, mid, hi);
        // SWIZZLE_128B K-major tile: row t at (t >> 3) * 1024 + (t & 7) * 128, 16-byte chunk c stored at c ^ (t & 7);
        // this lane's 4 bytes are word (lane & 3) of chunk (lane >> 2): the warp fills one 128-byte row
        const uint32_t addr = sa + (uint32_t)(t >> 3) * 1024 + (uint32_t)(t & 7) * 128 +
                              (uint32_t)((((lane >> 2) ^ (t & 7)) << 4) + ((lane & 3) << 2));
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(lo) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr + kPtTile), "r"(mid) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr + 2 * kPtTile), "r"(hi) : "memory");
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_full + 8 * stage);
    }
    // chunks not drained inside the loop: chunk c was drained at stage 8 (c + 1) + 1 iff that stage exists
    for (int c = 0; !kPlanesA && c < n_chunks; ++c)
      if ((c + 1) * kPtChunk + 1 > n_st - 1) drain(c);
    const int r = row0 + orow;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint32_t aw[32];
      tc_ld32(t_acc + 32 * h, aw);                         // (warp-collective: executed by every lane)
      if (r < S) {
        double* o = part + ((size_t)blockIdx.z * S + r) * Dout + n0 + cpart * 32 + 16 * h;
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (n0 + cpart * 32 + 16 * h + j < Dout) o[j] = __hiloint2double((int)aw[2 * j + 1], (int)aw[2 * j]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// defined in project.cu
void pca_finalize_launch(const double* part, int n_split, const float* ev, int S, int Dout, int normalize_rows, double* Y,
                         cudaStream_t st);

static int pca_tc_splits(int S, int Din, int Dout, int* stages_per_split) {
  const int n_st = (Din + kPtCh - 1) / kPtCh;
  const long long tiles = (long long)((S + kPtRows - 1) / kPtRows) * ((Dout + kPtCols - 1) / kPtCols);
  int zmax = n_st / (2 * kPtChunk);                         // >= two accumulation chunks per split
  if (zmax < 1) zmax = 1;
  if (zmax > 16) zmax = 16;                                 // bounds the fp64 partials (z x S x Dout x 8 bytes)
  // one CTA per SM (192 KB of shared memory each): take the smallest K split whose last wave is at least 95 % full,
  // else the best one (2048 x 1024: 128 tiles -> z = 8, 1024 CTAs = 6.9 waves instead of 1.7)
  int best_z = 1;
  double best_eff = 0.0;
  for (int z = 1; z <= zmax; ++z) {
    const long long ctas = tiles * z, waves = (ctas + 147) / 148;
    const double eff = (double)ctas / (double)(waves * 148);
    if (eff > best_eff + 1e-9) { best_eff = eff; best_z = z; }
    if (eff >= 0.95) { best_z = z; break; }
  }
  int sps = (n_st + best_z - 1) / best_z;
  sps = (sps + kPtChunk - 1) / kPtChunk * kPtChunk;
  *stages_per_split = sps;
  return (n_st + sps - 1) / sps;
}

}  // namespace segvlad

using namespace segvlad;

extern "C" size_t segvlad_pca_planes_bytes(int D_in, int D_out) {
  if (D_in <= 0 || D_out <= 0) return 256;
  return align_up((size_t)3 * align_up((size_t)D_out, kPtCols) * D_in * sizeof(__nv_bfloat16), 256);
}

extern "C" int segvlad_pca_prepare_planes(const float* components, int D_out, int D_in, void* planes, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  SV_REQUIRE(components && planes && D_out > 0 && D_in > 0, "pca_prepare_planes: bad arguments");
  const int Dout_p = (int)align_up((size_t)D_out, kPtCols);
  pca_planes_kernel<<<148 * 8, 256, 0, st>>>(components, D_out, Dout_p, D_in, reinterpret_cast<__nv_bfloat16*>(planes));
  SV_CHECK_LAUNCH();
  return SEGVLAD_OK;
}

extern "C" int segvlad_pca_tc_supported(int D_in, int D_out) {
  const char* e = getenv("SEGVLAD_PCA_TC");   // "0" selects the fp64 CUDA-core kernel of project.cu (kept as a cross-check)
  const bool on = !(e && e[0] == '0');
  return (on && D_in >= kPtCh && D_in % 8 == 0 && D_out >= 1) ? 1 : 0;
}

extern "C" size_t segvlad_pca_tc_workspace_bytes(int S, int D_in, int D_out) {
  if (S <= 0 || D_in <= 0 || D_out <= 0) return 256;
  int sps = 0;
  const int z = pca_tc_splits(S, D_in, D_out, &sps);
  return align_up((size_t)z * S * D_out * sizeof(double), 256) + 256;
}

extern "C" int segvlad_pca_project_tc(const double* X, int S, int D_in, const void* planes, const double* mean,
                                      const float* explained_variance, int D_out, int normalize_rows, double* Y,
                                      void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  SV_REQUIRE(X && planes && mean && explained_variance && Y, "pca_project_tc: null pointer");
  SV_REQUIRE(S >= 0 && D_in >= kPtCh && D_in % 8 == 0 && D_out > 0, "pca_project_tc: unsupported shape (D_in %d, D_out %d)", D_in, D_out);
  SV_REQUIRE((reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(mean) & 15) == 0,
             "pca_project_tc: X and mean must be 16-byte aligned");
  if (S == 0) return SEGVLAD_OK;
  const size_t need = segvlad_pca_tc_workspace_bytes(S, D_in, D_out);
  if (!workspace || workspace_bytes < need) {
    set_error("pca_project_tc: workspace %zu < required %zu", workspace_bytes, need);
    return SEGVLAD_EWORKSPACE;
  }
  int sps = 0;
  const int z = pca_tc_splits(S, D_in, D_out, &sps);
  const int Dout_p = (int)align_up((size_t)D_out, kPtCols);
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return SEGVLAD_ECUDA; }
  CUtensorMap map;
  cuuint64_t dims[2] = {(cuuint64_t)D_in, (cuuint64_t)3 * Dout_p};
  cuuint64_t strides[1] = {(cuuint64_t)D_in * 2};
  cuuint32_t box[2] = {(cuuint32_t)kPtCh, (cuuint32_t)kPtCols};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(planes), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (components) failed (%d)", (int)r); return SEGVLAD_ECUDA; }
  double* part = reinterpret_cast<double*>(workspace);
  const size_t smem = pca_tc_smem();
  SV_CHECK_CUDA(cudaFuncSetAttribute(pca_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(Dout_p / kPtCols, (S + kPtRows - 1) / kPtRows, z);
  const int pslot = prof_begin(SEGVLAD_PROF_PCA, st);
  pca_tc_kernel<false><<<grid, kPtThreads, smem, st>>>(map, map, X, mean, S, D_in, D_out, Dout_p, sps, part);
  prof_end(pslot, st);
  SV_CHECK_LAUNCH();
  pca_finalize_launch(part, z, explained_variance, S, D_out, normalize_rows, Y, st);
  SV_CHECK_LAUNCH();
  return SEGVLAD_OK;
}

extern "C" int segvlad_pca_project_planes(const void* x_planes, int S, int D_in, const void* planes, const float* explained_variance,
                                          int D_out, int normalize_rows, double* Y, void* workspace, size_t workspace_bytes,
                                          void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  SV_REQUIRE(x_planes && planes && explained_variance && Y, "pca_project_planes: null pointer");
  SV_REQUIRE(S >= 0 && D_in >= kPtCh && D_in % 8 == 0 && D_out > 0, "pca_project_planes: unsupported shape (D_in %d, D_out %d)", D_in,
             D_out);
  SV_REQUIRE((reinterpret_cast<uintptr_t>(x_planes) & 15) == 0, "pca_project_planes: x_planes must be 16-byte aligned");
  if (S == 0) return SEGVLAD_OK;
  const size_t need = segvlad_pca_tc_workspace_bytes(S, D_in, D_out);
  if (!workspace || workspace_bytes < need) {
    set_error("pca_project_planes: workspace %zu < required %zu", workspace_bytes, need);
    return SEGVLAD_EWORKSPACE;
  }
  int sps = 0;
  const int z = pca_tc_splits(S, D_in, D_out, &sps);
  const int Dout_p = (int)align_up((size_t)D_out, kPtCols);
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return SEGVLAD_ECUDA; }
  CUtensorMap map_w, map_x;
  cuuint32_t estr[2] = {1, 1};
  {
    cuuint64_t dims[2] = {(cuuint64_t)D_in, (cuuint64_t)3 * Dout_p};
    cuuint64_t strides[1] = {(cuuint64_t)D_in * 2};
    cuuint32_t box[2] = {(cuuint32_t)kPtCh, (cuuint32_t)kPtCols};
    CUresult r = enc(&map_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(planes), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (components) failed (%d)", (int)r); return SEGVLAD_ECUDA; }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)D_in, (cuuint64_t)3 * S};
    cuuint64_t strides[1] = {(cuuint64_t)D_in * 2};
    cuuint32_t box[2] = {(cuuint32_t)kPtCh, (cuuint32_t)kPtRows};
    CUresult r = enc(&map_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(x_planes), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (x planes) failed (%d)", (int)r); return SEGVLAD_ECUDA; }
  }
  double* part = reinterpret_cast<double*>(workspace);
  const size_t smem = pca_tc_smem();
  SV_CHECK_CUDA(cudaFuncSetAttribute(pca_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(Dout_p / kPtCols, (S + kPtRows - 1) / kPtRows, z);
  const int pslot = prof_begin(SEGVLAD_PROF_PCA, st);
  pca_tc_kernel<true><<<grid, kPtThreads, smem, st>>>(map_w, map_x, nullptr, nullptr, S, D_in, D_out, Dout_p, sps, part);
  prof_end(pslot, st);
  SV_CHECK_LAUNCH();
  pca_finalize_launch(part, z, explained_variance, S, D_out, normalize_rows, Y, st);
  SV_CHECK_LAUNCH();
  return SEGVLAD_OK;
}
