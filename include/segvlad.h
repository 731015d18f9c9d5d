/* segvlad.h -- C ABI of libsegvlad.so: the B200-native SegVLAD hot path (aggregate -> match -> vote).
 *
 * The reference (AnyLoc/Revisit-Anything) is pure Python and has no FFI; its "API" for this path is
 * a handful of module-level functions.  Each entry point below names the reference
 * function(s) it replaces (paths relative to the reference root).  The Python host layer
 * (revisit-anything_b200/func_vpr.py, place_rec_main.py) binds these with ctypes and keeps the
 * reference signatures.
 *
 * Conventions
 *  - every function returns 0 on success, a negative SEGVLAD_E* code otherwise;
 *    segvlad_last_error() gives the message of the last failure on the calling thread.
 *  - all data pointers are DEVICE pointers unless the parameter name ends in `_host`.
 *  - `stream` is a cudaStream_t passed as void*; every launch goes on it; no internal
 *    synchronisation except where a function is documented to read a device flag.
 *  - the library never allocates device memory: callers pass a workspace sized by the matching
 *    *_workspace_bytes() query (plain device memory, 256-byte aligned, contents undefined).
 *  - no torch types, no C++ types.
 */
#ifndef SEGVLAD_H_
#define SEGVLAD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SEGVLAD_OK 0
#define SEGVLAD_EINVAL (-1)      /* bad shape / unsupported parameter            */
#define SEGVLAD_ECUDA (-2)       /* CUDA runtime / driver error                  */
#define SEGVLAD_EWORKSPACE (-3)  /* workspace too small                          */
#define SEGVLAD_EOVERFLOW (-4)   /* candidate buffers overflowed in every schedule */

#define SEGVLAD_OUT_F64 0
#define SEGVLAD_OUT_F32 1
#define SEGVLAD_OUT_PCA_PLANES 2 /* segvlad_aggregate_batch_pca only: bf16 planes of (descriptor - pca mean) */

#define SEGVLAD_TOKENS_DN 0 /* per image [D_t, N]  (reference h5 layout [1,D_t,dh,dw]) */
#define SEGVLAD_TOKENS_ND 1 /* per image [N, D_t]                                      */
#define SEGVLAD_TOKENS_PRENORMALIZED 2 /* OR-flag: tokens are already unit-norm, skip the L2-normalise
                                          (vlad_single receives normalised descriptors, func_vpr.py:1096) */

int segvlad_version(void);
const char* segvlad_last_error(void);
/* cumulative number of CUDA kernels this library has launched in the process (bench.py's gpu_launches) */
uint64_t segvlad_launch_count(void);
/* Optional kernel timing for measurement (bench.py roofline): when enabled, the dominant kernels are bracketed
 * by CUDA events on their launch stream.  Tags: */
#define SEGVLAD_PROF_KNN_FILTER 1 /* tcgen05 (or SIMT) all-pairs + filter kernel */
#define SEGVLAD_PROF_AGGREGATE 2  /* masked residual aggregation kernel          */
#define SEGVLAD_PROF_KNN_RESCORE 3
#define SEGVLAD_PROF_PCA 4         /* tensor-core PCA projection kernel            */
void segvlad_profile_enable(int on);
int segvlad_profile_read(int tag, double* total_ms, int* launches); /* synchronises the recorded events */
void segvlad_profile_reset(void);

/* ---------------------------------------------------------------------------------------------
 * Aggregation: per-(Super)Segment masked hard-assignment VLAD.
 * Replaces  func_vpr.py:1140-1179 vlad_single  +  func_vpr.py:1181-1210 vlad_matmuls_per_cluster
 * (called from func_vpr.py:1065-1101 seg_vlad_gpu_single / :1103-1138 seg_vlad_gpu_single_img),
 * batched over images.
 *
 *  tokens        n_images * N * D_t fp32, layout per `token_layout` (NOT yet normalised: the kernel does
 *                the channel L2-normalise of func_vpr.py:1085)
 *  centers       [K, D_t] fp32, un-normalised (c_centers.pt)
 *  member_bits   [S_total, ceil(N/32)] uint32: base-segment patch membership bitmask
 *                (bit p%32 of word p/32 = mask_idx[s,p] of func_vpr.py:1090-1092)
 *  seg_offsets_host [n_images+1] int32 prefix sum of segments per image (HOST memory)
 *  adj           concatenated per-image [S_i, S_i] uint8 (0/1) neighbourhood matrices
 *                (func_vpr.py:1309-1347 output), or NULL for identity (order 0)
 *  out           [S_total, K*D_t] fp64 (SEGVLAD_OUT_F64, the reference dtype) or fp32
 *  labels_out    optional [n_images*N] int32 cluster label per token (NULL to skip)
 */
size_t segvlad_aggregate_workspace_bytes(int n_images, int N, int D_t, int K, int S_total);
int segvlad_aggregate_batch(const float* tokens, int n_images, int N, int D_t, int token_layout,
                            const float* centers, int K, const uint32_t* member_bits,
                            const int32_t* seg_offsets_host, const uint8_t* adj, void* out,
                            int out_dtype, int32_t* labels_out, void* workspace,
                            size_t workspace_bytes, void* stream);

/* Aggregation fused with the front half of the PCA-whitening projection (place_rec_main.py:261-272 projects
 * every aggregated batch with func_vpr.py:1419-1443): instead of the [S_total, K*D_t] fp64 descriptor matrix the kernel's
 * epilogue writes (descriptor - pca_mean) directly as the three bf16 planes the tensor-core projection consumes,
 * x_planes = [3][S_total][K*D_t] bf16 (plane 0 = lo, 1 = mid, 2 = hi; 6 bytes per element instead of 8, and no fp64 matrix
 * in HBM at all).  pca_mean_f32: [K*D_t] fp32 copy of pca.mean_.  Needs D_t % 64 == 0.  Feed x_planes to
 * segvlad_pca_project_planes.  Same workspace as segvlad_aggregate_batch. */
int segvlad_aggregate_batch_pca(const float* tokens, int n_images, int N, int D_t, int token_layout,
                                const float* centers, int K, const uint32_t* member_bits,
                                const int32_t* seg_offsets_host, const uint8_t* adj, const float* pca_mean_f32,
                                void* x_planes, int32_t* labels_out, void* workspace, size_t workspace_bytes,
                                void* stream);

/* Inner stage alone: aggregate caller-provided residual rows [n_images*N, D_t] fp32 with caller-provided
 * labels [n_images*N] int32.  Replaces func_vpr.py:1181-1210 vlad_matmuls_per_cluster(num_c, masks, res,
 * clus_labels, adjMat).  Same workspace size as segvlad_aggregate_batch. */
int segvlad_aggregate_residuals(const float* residuals, const int32_t* labels, int n_images, int N, int D_t,
                                int K, const uint32_t* member_bits, const int32_t* seg_offsets_host,
                                const uint8_t* adj, void* out, int out_dtype, void* workspace,
                                size_t workspace_bytes, void* stream);

/* Pixel masks -> patch membership bitmask (func_vpr.py:1088-1092 + lookup table place_rec_main.py:187-194).
 *  masks  [S, Hm, Wm] uint8 (0/1) at SAM resolution; nearest-neighbour upsampled to HxW, then a patch
 *  is a member if any pixel of its (clipped) patch x patch cell is set.  member_bits [S, ceil(N/32)]. */
int segvlad_mask_to_membership(const uint8_t* masks, int S, int Hm, int Wm, int H, int W, int patch,
                               uint32_t* member_bits, void* stream);
/* Mask centroids (x = mean column, y = mean row, fp64; NaN for an empty mask): the per-mask reduction of
 * func_vpr.py:1314 `np.array(np.nonzero(m)).mean(1)[::-1]` that feeds scipy's Delaunay in nbrMasksAGGFastSingle
 * (func_vpr.py:1309-1347; the triangulation itself stays on the host as in the reference).  centroids_xy [S, 2]. */
int segvlad_mask_centroids(const uint8_t* masks, int S, int Hm, int Wm, double* centroids_xy, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Matching: exhaustive squared-L2 kNN of query-segment descriptors against a (shard of the)
 * reference bank.  Replaces faiss.IndexFlatL2.add/search at place_rec_main.py:53-61.
 *
 * A "bank" is the resident, kernel-ready form of an [n, D] fp32 descriptor matrix: one fp16 plane (each row scaled by
 * a power of two, padded to a multiple of 64 columns) for the tensor-core scan, per-row fp32 squared norm / scale /
 * rounding residual (the scan's error bound), and the fp32 rows for the exact re-score of the surviving candidates.
 */
size_t segvlad_bank_bytes(int n, int D);
int segvlad_bank_prepare(const float* x, int n, int D, void* bank, void* stream);
/* Same, but the bank REFERENCES the caller's fp32 rows instead of copying them (the exact re-score reads x directly): x must
 * stay valid and unchanged for as long as the bank is searched.  Saves the 4 n D byte copy of every index.add
 * (place_rec_main.py:55, 59 re-adds the whole reference set on every recall_segloc call).  x 16-byte aligned. */
int segvlad_bank_prepare_view(const float* x, int n, int D, void* bank, void* stream);
/* fp64 input (the reference keeps segFtVLAD1/2 in fp64): optional row L2-normalise in fp64 WITHOUT eps
 * (func_vpr.py:1673-1676 normalizeFeat, place_rec_main.py:55-56), then the fp32 cast faiss's Python
 * wrapper performs, then the split. */
int segvlad_bank_prepare_f64(const double* x, int n, int D, int normalize_rows, void* bank, void* stream);

size_t segvlad_knn_workspace_bytes(int Nq, int Nr, int D, int k);
/* d2_out [Nq,k] fp32 ascending (ties: smaller index first), idx_out [Nq,k] int64 = row_offset + local
 * row; rows with fewer than k refs are padded with (+inf, -1) like faiss.
 * Synchronises `stream` once at the end to read the overflow flag (and re-runs with the conservative
 * chunk schedule if any candidate buffer overflowed). */
int segvlad_knn(const void* qbank, int Nq, const void* rbank, int Nr, int64_t row_offset, int D,
                int k, float* d2_out, int64_t* idx_out, void* workspace, size_t workspace_bytes,
                void* stream);
/* Asynchronous form of segvlad_knn for pipelines (search -> [all-gather -> merge] -> vote) that must not stall the host
 * between stages: NO host synchronisation.  `schedule` 0 = fast chunk schedule (candidate buffers can overflow on
 * adversarially ordered / massively duplicated banks), 1 = conservative schedule (cannot overflow, slower).  The OR of the
 * query blocks' overflow flags is written to overflow_dev (device int32, required): the caller reads it at ITS next
 * synchronisation point (after the vote) and, if it is set, repeats the call with schedule = 1.
 * Outputs (at least one): d2_out / idx_out as segvlad_knn, and / or topk_packed [Nq, k] uint64 =
 * (fp32 bits of d2) << 32 | (uint32)(int32 global row) -- ascending as unsigned integers, padding = +inf | 0xffffffff --
 * the payload of the ONE all-gather of a row-sharded search, written by the final selection kernel
 * straight into the send buffer.  Packed output requires row_offset + Nr < 2^31 (checked on the host). */
int segvlad_knn_async(const void* qbank, int Nq, const void* rbank, int Nr, int64_t row_offset, int D, int k,
                      int schedule, float* d2_out, int64_t* idx_out, uint64_t* topk_packed, int32_t* overflow_dev,
                      void* workspace, size_t workspace_bytes, void* stream);
/* Same search with both descriptor matrices in (pinned) HOST memory -- what the reference hands to faiss
 * (segFtVLAD1/2 are CPU tensors, place_rec_main.py:53-61).  The fp32 rows are copied sub-chunk by sub-chunk on a copy
 * stream straight into the banks' fp32 regions while earlier sub-chunks are split and scanned on `stream`, so the PCIe
 * transfer overlaps the tensor-core scan.  qbank / rbank: device buffers of segvlad_bank_bytes() (outputs; the banks are
 * resident and re-usable with segvlad_knn afterwards).  copy_stream may be NULL (library-owned stream). */
int segvlad_knn_from_host(const float* q_host, int Nq, const float* r_host, int Nr, int64_t row_offset, int D,
                          int k, void* qbank, void* rbank, float* d2_out, int64_t* idx_out, void* workspace,
                          size_t workspace_bytes, void* stream, void* copy_stream);
/* Same contract on the raw fp32 matrices with plain fp32 FFMA inner products (no tensor cores):
 * the on-device cross-check for the tcgen05 path (tests / debugging); same workspace size. */
int segvlad_knn_simt(const float* q, int Nq, const float* r, int Nr, int64_t row_offset, int D, int k,
                     float* d2_out, int64_t* idx_out, void* workspace, size_t workspace_bytes,
                     void* stream);

/* Test hook for the scan's error model (Nr <= 4096): the approximate d2 the single fp16 tensor-core pass assigns to
 * every (query, reference) pair, approx_out [Nq, Nr], and the per-query bound E the filter assumes on
 * |approx - fp32 value|, bound_out [Nq].  Workspace as for segvlad_knn. */
int segvlad_knn_debug_approx(const void* qbank, int Nq, const void* rbank, int Nr, int D, float* approx_out,
                             float* bound_out, void* workspace, size_t workspace_bytes, void* stream);

/* k-way merge of per-shard results after the all-gather: parts are [G, Nq, k]. */
int segvlad_merge_topk(const float* d2_parts, const int64_t* idx_parts, int G, int Nq, int k,
                       float* d2_out, int64_t* idx_out, void* stream);
/* The same merge on the packed all-gather payload of segvlad_knn_async: shard g's sorted lists start at
 * parts + g * part_stride (part_stride >= Nq * k elements; the gather buffer may carry per-shard trailer words).
 * A true k-way merge (rank by binary search, no sort, no unpacking pass); outputs d2_out / idx_out [Nq, k] (faiss
 * layout, int64 only here) and / or packed_out [Nq, k].  G * k <= 16384. */
int segvlad_merge_topk_packed(const uint64_t* parts, int G, size_t part_stride, int Nq, int k, float* d2_out,
                              int64_t* idx_out, uint64_t* packed_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Vote: segment hits -> ranked reference images.
 * Replaces func_vpr.py:80-243 get_matches, branch :207-224 ("max_seg_topk_wt_borda_Im") with
 * func_vpr.py:61-77 weighted_borda_count, and the integer bincount of :118-125 ("max_seg_topk").
 *
 *  matches [Nq, ld] int64 (first k_vote columns used), sims [Nq, ld] fp32 (or d2 if sims_is_d2 != 0:
 *  sims = 2 - d2 in fp32, place_rec_main.py:78-81)
 *  qimg_offsets [n_qimg+1] int32: query image i owns rows [off[i], off[i+1])  (segRangeQuery)
 *  qrow_index   optional [off[n_qimg]] int32: position p of the concatenated segRangeQuery lists reads row
 *               qrow_index[p] of matches / sims (NULL: identity, the contiguous ranges place_rec_main.py:354-355
 *               builds).  The min / max normalisation is always taken over ALL Nq rows, like np.min / np.max over the
 *               whole sims array at func_vpr.py:211-212; padding entries (match < 0) are excluded from it.
 *  rseg_to_rimg [Nr] int32 (imIndsRef)
 *  preds        [n_qimg, n_pred] int32 ref-image ids, -1 padded;  pred_scores [n_qimg, n_pred] fp64
 *  scores_dense / counts_dense: optional [n_qimg, n_rimg] fp64 / int32 (NULL to skip)
 *  minmax_out   optional [2] fp32 (global min, max of sims)
 */
size_t segvlad_vote_workspace_bytes(int Nq, int k_vote, int n_qimg, int max_segs_per_qimg);
int segvlad_vote(const int64_t* matches, const float* sims, int ld, int sims_is_d2, int k_vote, int Nq,
                 const int32_t* qimg_offsets, const int32_t* qrow_index, int n_qimg, int max_segs_per_qimg,
                 const int32_t* rseg_to_rimg, int Nr, int n_rimg, int n_pred, int32_t* preds,
                 double* pred_scores, double* scores_dense, int32_t* counts_dense, float* minmax_out,
                 void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * PCA-whitening projection of segment descriptors (SURVEY 8f "next" row f1).
 * Replaces func_vpr.py:1419-1443 apply_pca_transform_from_pkl (sklearn PCA.transform, whiten=True; model fitted at
 * place_rec_pca.py:339-342) and, with normalize_rows != 0, the following func_vpr.py:1673-1676 normalizeFeat.
 *  X [S, D_in] fp64 (aggregation output), components [D_out, D_in] fp32 (pca.components_), mean [D_in] fp64
 *  (pca.mean_), explained_variance [D_out] fp32;  Y [S, D_out] fp64 = ((X - mean) @ components^T) / sqrt(ev).
 */
size_t segvlad_pca_workspace_bytes(int S, int D_in, int D_out);
int segvlad_pca_project(const double* X, int S, int D_in, const float* components, const double* mean,
                        const float* explained_variance, int D_out, int normalize_rows, double* Y,
                        void* workspace, size_t workspace_bytes, void* stream);
/* Tensor-core version of the same projection (csrc/project_tc.cu): the components are split once per model into bf16
 * planes (segvlad_pca_prepare_planes, planes = segvlad_pca_planes_bytes(D_in, D_out) bytes of device memory, kept by
 * the caller next to the model -- the reference re-unpickles the model on every batch, func_vpr.py:1431-1436), the
 * projection accumulates 512-channel chunks in fp32 on tcgen05 and the chunk sums in fp64 (~1e-6 relative, inside the
 * 1e-5 descriptor tolerance).  segvlad_pca_tc_supported: D_in >= 64 and a multiple of 8, and SEGVLAD_PCA_TC != "0";
 * otherwise use segvlad_pca_project (fp64 CUDA cores, also the cross-check).  X and mean must be 16-byte aligned. */
int segvlad_pca_tc_supported(int D_in, int D_out);
size_t segvlad_pca_planes_bytes(int D_in, int D_out);
int segvlad_pca_prepare_planes(const float* components, int D_out, int D_in, void* planes, void* stream);
size_t segvlad_pca_tc_workspace_bytes(int S, int D_in, int D_out);
int segvlad_pca_project_tc(const double* X, int S, int D_in, const void* planes, const double* mean,
                           const float* explained_variance, int D_out, int normalize_rows, double* Y,
                           void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * NetVLAD + anti-burst aggregation (BASELINE config 5; data parallel over images, no collective).
 * Replaces VLAD-BuFF/models/aggregators/aggregation.py:266-361 NetVLAD.forward (antiburst=True, getWeights
 * :148-162, defaults ab_relu/ab_inv/ab_soft False).
 *  x [B, D, N] fp32 backbone tokens (reference [B,D,H,W]); centroids [K,D]; conv_weight [K,D] (= alpha * c_hat,
 *  aggregation.py:245-256); (ab_w, ab_b, ab_p) = ab_params (8,7,1 by default); out [B, K*D] fp32.
 */
size_t segvlad_netvlad_workspace_bytes(int B, int N, int D, int K);
int segvlad_netvlad_antiburst(const float* x, int B, int N, int D, const float* centroids,
                              const float* conv_weight, int K, float ab_w, float ab_b, float ab_p, float* out,
                              void* workspace, size_t workspace_bytes, void* stream);

/* Back half of the fused path: Y = (x_planes . components^T) / sqrt(ev) with the A operand already split (see
 * segvlad_aggregate_batch_pca); both operands arrive by TMA, fp32 chunk sums are folded into fp64 accumulators in TMEM as
 * in segvlad_pca_project_tc.  Workspace: segvlad_pca_tc_workspace_bytes(S, D_in, D_out). */
int segvlad_pca_project_planes(const void* x_planes, int S, int D_in, const void* planes, const float* explained_variance,
                               int D_out, int normalize_rows, double* Y, void* workspace, size_t workspace_bytes,
                               void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SEGVLAD_H_ */
