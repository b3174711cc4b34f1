/*
 * picopose_b200 -- C ABI of the B200-native PicoPose correspondence hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes,
 * no torch types.  Every pointer is a DEVICE pointer owned by the caller
 * (PyTorch allocates, the library never allocates, frees or retains memory);
 * `stream` is a cudaStream_t passed as void*; all work is enqueued
 * asynchronously on it.  Return value: 0 on success, negative pp_status on
 * error with a message in pp_last_error().  There is no CPU fallback: calls
 * fail with PP_ERR_DEVICE on anything but an sm_100 device.
 *
 * Each entry point names the reference interface it replaces
 * (paths relative to the PicoPose repository).
 */
#ifndef PICOPOSE_B200_H
#define PICOPOSE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PP_VERSION 100 /* 0.1.0 */

#if defined(__GNUC__)
#define PP_API __attribute__((visibility("default")))
#else
#define PP_API
#endif

typedef enum pp_status {
    PP_OK = 0,
    PP_ERR_ARG = -1,       /* bad shape / null pointer / unsupported option (mirrors the reference's asserts) */
    PP_ERR_ALIGN = -2,     /* pointer or stride not aligned as required */
    PP_ERR_WORKSPACE = -3, /* caller-provided workspace too small */
    PP_ERR_DEVICE = -4,    /* not an sm_100 device / driver entry point missing */
    PP_ERR_LAUNCH = -5,    /* CUDA launch or API failure */
    PP_ERR_KERNEL = -6     /* a kernel reported an internal fault (pipeline timeout) */
} pp_status;

/* OR into the `cluster` argument of pp_match_scores: cheaper reduction keys in the contraction's epilogue (similarities
 * resolved to 7.6e-6 absolute instead of 2^-19 relative).  Meant for PP_MODE_BF16 operands, whose own error is ~3e-4; the
 * one-call entry points (pp_match_templates, pp_match_templates_dense) set it themselves for that mode.  Query masks must
 * lie in [0, 1] (the reference's are 0/1). */
#define PP_MATCH_FAST_KEYS 0x100

/* arithmetic mode of the stage-1 contraction */
typedef enum pp_mode {
    PP_MODE_BF16 = 0,   /* bf16 operands, fp32 accumulate (tcgen05 kind::f16)                 */
    PP_MODE_FP32 = 1,   /* fp32-accurate: 3-way bf16 split, 6 cross terms on the same tensor path */
    PP_MODE_BF16X3 = 2  /* 2-way bf16 split, 3 cross terms (~1e-6 abs error)                   */
} pp_mode;

PP_API int pp_version(void);
PP_API const char* pp_last_error(void);
/* Reads and clears the device-side fault record (synchronises the device). 0 = no fault; PP_ERR_KERNEL with a message
 * for a pipeline timeout (kernel trapped), a peer that never arrived at an exchange, or a bank index out of range. */
PP_API int pp_check_device_faults(void);
/* Number of kernels this library has launched in this process (measurement aid). */
PP_API long long pp_launch_count(void);
/* Measurement aid: the next tensor-core GEMM launched from this thread records `start_event` right
 * before and `stop_event` right after itself on its stream (both cudaEvent_t; pass NULLs to cancel). */
PP_API void pp_profile_gemm_events(void* start_event, void* stop_event);

/* ------------------------------------------------------------------------------------------
 * Stage 3 -- correlation-window lookup.
 * Replaces CorrLookup.forward, utils/corr_lookup.py:100-134 (called from
 * model/stage3/flow_decoder.py:61).
 *   pyr_ptrs[l] : level-l volume, (B*H*W, 1, pyr_h[l], pyr_w[l]) fp32 contiguous   (HOST array of DEVICE ptrs)
 *   flow        : (B, 2, H, W) fp32, channel 0 = x, 1 = y
 *   out         : (B, L*(2r+1)^2, H, W) fp32; channel l*D*D + a*D + b samples (x/2^l + a-r, y/2^l + b-r),
 *                 bilinear, zero padding, align_corners=True arithmetic of F.grid_sample.
 * ------------------------------------------------------------------------------------------ */
PP_API int pp_corr_lookup(const void* const* pyr_ptrs, const int* pyr_h, const int* pyr_w, int L,
                   const float* flow, int B, int H, int W, int radius,
                   float* out, void* stream);

/* Same lookup on volumes in the TILED layout pp_correlation_pyramid_tiled writes: every (h x w) slice stored as 4-row x
 * 8-column tiles of 32 floats (one 128-byte line per tile), tiles row-major inside the slice; element (y, x) lives at
 * ((y/4) * (w/8) + x/8) * 32 + (y%4) * 8 + x%8.  DRAM moves whole 128-byte lines, and a lookup window crosses ~40 % fewer
 * tiles than row segments.  Every level needs h % 4 == 0 and w % 8 == 0; radius 1..8.  Same output as pp_corr_lookup. */
PP_API int pp_corr_lookup_tiled(const void* const* pyr_ptrs, const int* pyr_h, const int* pyr_w, int L,
                         const float* flow, int B, int H, int W, int radius,
                         float* out, void* stream);
/* Layout change of `slices` (h x w) fp32 slices: to_tiled != 0 reference row-major -> tiled, else the way back. */
PP_API int pp_volume_retile(const float* in, float* out, int64_t slices, int h, int w, int to_tiled, void* stream);

/* Replaces bilinear_sample, utils/corr_lookup.py:29-65 (mode='bilinear', padding_mode='zeros');
 * used by FlowDecoder.feature_sample, model/stage3/flow_decoder.py:49-56.
 *   feat (N,C,Hf,Wf) fp32; grid (N,Ho,Wo,2) if grid_chw == 0 else (N,2,Ho,Wo); out (N,C,Ho,Wo).
 *   scale != 0: grid holds pixel coordinates and is normalised as the reference does
 *   (the caller's grid is NOT modified, unlike the reference's in-place scaling). */
PP_API int pp_bilinear_sample(const float* feat, const float* grid, int N, int C, int Hf, int Wf,
                       int Ho, int Wo, int grid_chw, int align_corners, int scale,
                       float* out, void* stream);

/* The remaining argument combinations of bilinear_sample / CorrLookup (utils/corr_lookup.py:29-65,88-98 hand `mode` and
 * `padding_mode` straight to F.grid_sample): nearest and bicubic interpolation, border and reflection padding, after
 * ATen's GridSampler rules.  PicoPose itself only uses bilinear + zeros (the tuned kernels above); this is the
 * compatibility path.  Same tensors as pp_bilinear_sample; (PP_SAMPLE_BILINEAR, PP_PAD_ZEROS) forwards to it. */
enum pp_sample_mode { PP_SAMPLE_BILINEAR = 0, PP_SAMPLE_NEAREST = 1, PP_SAMPLE_BICUBIC = 2 };
enum pp_pad_mode { PP_PAD_ZEROS = 0, PP_PAD_BORDER = 1, PP_PAD_REFLECTION = 2 };
PP_API int pp_grid_sample(const float* feat, const float* grid, int N, int C, int Hf, int Wf,
                   int Ho, int Wo, int grid_chw, int align_corners, int scale, int mode, int padding_mode,
                   float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Stage 1 -- query-vs-template matching.
 *
 * Operand preparation ("prologue"): cast to bf16 (split into terms for the fp32 modes), lay out
 * K-major for the tensor cores and compute the inverse L2 norm over C per patch
 * (1 / max(||x||, 1e-12), F.normalize's rule); the contraction's epilogue applies the two inverse norms.
 *   feats : (G, C, P) fp32, P = H*W contiguous      -> prepared : (G, P, Kp) bf16, rnorm : (G, P) fp32
 *   Kp = pp_match_kp(C, mode).  is_query selects which side of the split-term pairing is written.
 * Replaces the F.normalize + rearrange lines utils/matching.py:13-14,18-19,40-41,43-44.
 * ------------------------------------------------------------------------------------------ */
PP_API int pp_match_kp(int C, int mode);
PP_API int pp_match_prepare(const float* feats, int64_t G, int C, int P, int mode, int is_query,
                     void* prepared, float* rnorm, void* stream);

/* Query side of pp_match_scores: resizes the query masks (nearest, as F.interpolate at utils/matching.py:38-39),
 * drops masked patches (they are rows of zeros in the reference and never need the tensor cores), and prepares the
 * remaining patches NORMALISED and multiplied by their mask value (x * m / max(||x||, 1e-12), utils/matching.py:40-41,48)
 * before the bf16 split, so the contraction's epilogue has no per-row factor left to apply (q_rnorm is 1 for live rows);
 * also records the bookkeeping pp_match_scores needs.
 *   tar_feat (B,C,H,W) fp32, tar_mask (B,Hm,Wm) fp32  ->  q_prep (B, H*W, Kp) bf16 (unmasked patches first),
 *   q_rnorm (B, H*W) fp32, q_meta: pp_match_query_meta_bytes(B, H*W) bytes (opaque). */
PP_API size_t pp_match_query_meta_bytes(int B, int T);
PP_API int pp_match_prepare_query(const float* tar_feat, const float* tar_mask, int B, int C, int H, int W, int Hm, int Wm,
                           int mode, void* q_prep, float* q_rnorm, void* q_meta, void* stream);

/* Bytes of scratch pp_match_scores needs for (B detections, N views, T = H*W patches). */
PP_API size_t pp_match_scores_workspace(int B, int N, int T);

/* Fused similarity GEMM + bidirectional max/argmax + validity-masked mean.
 * Replaces utils/matching.py:38-39,47-67 (the sim tensor never reaches HBM).
 *   q_prep, q_rnorm, q_meta : outputs of pp_match_prepare_query for the same B, H, W
 *   bank_prep, bank_rnorm : (n_banks, N, T, Kp) bf16 / (n_banks, N, T) fp32   prepared template banks
 *   bank_of_det : (B,) int32 device array, bank used by detection b; NULL = identity (n_banks == B)
 *   sim_avg   : (B, N) fp32 out
 *   optional outs (NULL to skip): score_t2s (B,N,T) fp32, idx_t2s (B,N,T) int32, idx_s2t (B,N,T) int32,
 *   mutual_nn (B,N,T) uint8 = 1 where query patch t and template patch idx_t2s[t] are mutual nearest neighbours
 *   cluster   : 0 = default, 1 = one CTA per tile, 2 = CTA pairs (cta_group::2)
 */
PP_API int pp_match_scores(const void* q_prep, const float* q_rnorm, const void* q_meta, const void* bank_prep,
                    const float* bank_rnorm, int64_t n_banks, const int32_t* bank_of_det, int B, int N, int H, int W, int Kp,
                    float* sim_avg, float* score_t2s, int32_t* idx_t2s, int32_t* idx_s2t, uint8_t* mutual_nn,
                    void* workspace, size_t workspace_bytes, int cluster, void* stream);

/* Row-wise top-k (descending, lowest index first on ties).  Replaces torch.topk, utils/matching.py:68.
 *   scores (B,N) fp32 -> out_score (B,k) fp32, out_idx (B,k) int64 (+ idx_offset, for sharded banks). */
PP_API int pp_topk(const float* scores, int B, int N, int k, int64_t idx_offset,
            float* out_score, int64_t* out_idx, void* stream);

/* One-call form of the stage-1 ranking against prepared banks: pp_match_prepare_query + pp_match_scores + pp_topk
 * with a single caller-provided workspace.  Replaces matching_templates, utils/matching.py:29-69, for banks that
 * were prepared once with pp_match_prepare (the reference re-normalises them on every call, :43).
 *   out_score (B,k) fp32, out_idx (B,k) int64; sim_avg_out (B,N) fp32 optional dense scores (NULL to skip). */
PP_API size_t pp_match_templates_workspace(int B, int N, int C, int H, int W, int mode);
PP_API int pp_match_templates(const float* tar_feat, const float* tar_mask, const void* bank_prep, const float* bank_rnorm,
                       int64_t n_banks, const int32_t* bank_of_det, int B, int N, int C, int H, int W, int Hm, int Wm,
                       int mode, int k, float* out_score, int64_t* out_idx, float* sim_avg_out,
                       void* workspace, size_t workspace_bytes, int cluster, void* stream);

/* matching_templates exactly as the reference calls it (utils/matching.py:29-69, from model/picopose.py:102): dense
 * fp32 template features in, prepared on the way.  src_feats (G, N, C, H, W) with G == B (one bank per detection)
 * or any G with bank_of_det (B,) (NULL: detection b uses bank b).  The bank prologue (pp_match_prepare) and the
 * query prologue (pp_match_prepare_query) are independent, so they run concurrently on a forked internal stream
 * that joins `stream` before the contraction.  bank_prep (G, N, H*W, Kp) bf16 and bank_rnorm (G, N, H*W) fp32 are
 * caller-provided outputs (reusable with pp_match_templates afterwards); workspace as pp_match_templates. */
PP_API int pp_match_templates_dense(const float* src_feats, int64_t G, const float* tar_feat, const float* tar_mask,
                             const int32_t* bank_of_det, int B, int N, int C, int H, int W, int Hm, int Wm, int mode,
                             int k, void* bank_prep, float* bank_rnorm, float* out_score, int64_t* out_idx,
                             float* sim_avg_out, void* workspace, size_t workspace_bytes, int cluster, void* stream);

/* Multi-GPU merge of sharded template banks: pp_topk_pairs writes each rank's local top-k as (score, global index)
 * pairs of doubles, (B, k, 2), padded with (-inf, -1) when the shard holds fewer than k views -- one tensor to
 * all-gather; pp_topk_merge reduces the gathered (R, B, k_in, 2) lists to the global top-k of every row
 * (ties: lowest rank, then lowest slot).  Together they replace torch.topk over the full view axis, utils/matching.py:68. */
PP_API int pp_topk_pairs(const float* scores, int B, int N, int k, int64_t idx_offset, double* out_pairs, void* stream);
PP_API int pp_topk_merge(const double* pairs, int R, int B, int k_in, int k, float* out_score, int64_t* out_idx,
                  void* stream);

/* The same exchange over NVLink peer memory in ONE kernel (one node, one process per GPU): block b ranks detection b's
 * local scores, stores its k (score, global index) pairs into every peer's exchange buffer through the peer mapping,
 * releases a per-(rank, detection) flag on every peer, waits for the peers' flags of that detection and merges.
 *   pp_xchg_bytes(world, max_b, k_max)   size of one rank's exchange buffer
 *   pp_xchg_create                        cudaMalloc + zero + cudaIpcGetMemHandle (64-byte handle for the peers)
 *   pp_xchg_open / pp_xchg_close          cudaIpcOpenMemHandle / cudaIpcCloseMemHandle on a peer's handle
 *   pp_topk_exchange                      peers_dev: device array of `world` buffer pointers (entry `rank` = own buffer);
 *                                         epoch: non-zero, incremented by 1 per call, the same on every rank; every rank
 *                                         must make the call (it is a collective).  Waiting for a peer is bounded by
 *                                         PICOPOSE_B200_XCHG_TIMEOUT_S (default 120 s, 0 = for ever): a rank that gives up
 *                                         records a fault (pp_check_device_faults -> PP_ERR_KERNEL), writes NaN / -1 into
 *                                         its outputs and returns -- no trap, the context survives.  B may exceed the
 *                                         number of co-resident blocks: the call is cut into launches that fit. */
PP_API size_t pp_xchg_bytes(int world, int max_b, int k_max);
PP_API int pp_xchg_create(size_t bytes, void** buf, void* handle64);
PP_API int pp_xchg_open(const void* handle64, void** peer_buf);
PP_API int pp_xchg_close(void* peer_buf);
PP_API int pp_xchg_destroy(void* buf);
/* Bulk pushes through the same kind of buffer (used for the query gather): pp_xchg_push copies `bytes` from `src` to
 * offset `dst_offset` of EVERY peer's buffer with copy-engine peer copies (peers_host: HOST array of `world` buffer
 * pointers), pp_xchg_signal then writes `epoch` into flag [rank] (u32 array at `flag_offset`) of every peer's buffer and
 * pp_xchg_wait spins on the `world` flags of the own buffer; all three are stream-ordered. */
PP_API int pp_xchg_push(const void* src, size_t bytes, const void* const* peers_host, int world, size_t dst_offset,
                 void* stream);
/* One call for the push side of a gather: `bytes` from `src` to offset `dst_offset` and the small `payload` (e.g. the
 * query masks) to `payload_offset` of every peer's buffer with the copy engines (on `stream`; PICOPOSE_B200_PUSH_STREAMS
 * = n > 1 deals the bulk copies over n internal streams that are joined back into `stream`, which pays only while the
 * copy engines are contended), then flag [rank] = `epoch` on every peer as pp_xchg_signal does.  pp_xchg_wait is the
 * receiving side. */
PP_API int pp_xchg_push_signal(const void* src, size_t bytes, size_t dst_offset, const void* payload, size_t payload_bytes,
                        size_t payload_offset, const void* const* peers_host, const void* const* peers_dev,
                        size_t flag_offset, int rank, int world, uint32_t epoch, void* stream);
PP_API int pp_xchg_signal(const void* const* peers_dev, size_t flag_offset, int rank, int world, uint32_t epoch,
                   void* stream);
PP_API int pp_xchg_wait(const void* own_buf, size_t flag_offset, int world, uint32_t epoch, void* stream);
PP_API int pp_topk_exchange(const float* scores, int B, int N, int k, int64_t idx_offset, const void* const* peers_dev,
                     int rank, int world, int max_b, int k_max, uint32_t epoch, float* out_score, int64_t* out_idx,
                     void* stream);

/* Stage-2 input volume.  Replaces matching_features_similarity, utils/matching.py:6-26.
 *   q_prep/q_rnorm, s_prep/s_rnorm : (B, T, Kp) / (B, T) prepared query / template features (pp_match_prepare, is_query = 1 / 0);
 *   src_mask (B,Hm,Wm) fp32
 *   out : (B, S, H, W) fp32 with out[b,s,h,w] = max(0, sim[b, t = w*H+h, s] * mask_s)
 * Norms, template mask (nearest-resized on the fly), clamp and the "(w h)" layout are applied by the contraction's
 * epilogue: one launch, no staging (workspace may be NULL; pp_match_similarity_workspace returns 0). */
PP_API size_t pp_match_similarity_workspace(int B, int T);
PP_API int pp_match_similarity(const void* q_prep, const float* q_rnorm, const void* s_prep, const float* s_rnorm,
                        const float* src_mask,
                        int B, int H, int W, int Kp, int Hm, int Wm, float* out,
                        void* workspace, size_t workspace_bytes, int cluster, void* stream);
/* The same from fp32 features, as the reference calls it (model/picopose.py:81): src_feat, tar_feat (B, C, H, W).
 * Two launches: one prologue for both operands, one contraction.  workspace >= pp_match_similarity_dense_workspace bytes. */
PP_API size_t pp_match_similarity_dense_workspace(int B, int C, int H, int W, int mode);
PP_API int pp_match_similarity_dense(const float* src_feat, const float* tar_feat, const float* src_mask, int B, int C,
                              int H, int W, int Hm, int Wm, int mode, float* out, void* workspace, size_t workspace_bytes,
                              int cluster, void* stream);

/* All-pairs correlation pyramid.  Replaces CorrelationPyramid.forward, model/stage3/raft_decoder.py:30-53
 * (torch.matmul / sqrt(C) + AvgPool2d(2,2) levels), the producer of the volumes pp_corr_lookup reads.
 *   f1_prep, f2_prep : (N, H*W, Kp) prepared features (pp_match_prepare; norms unused)
 *   level_ptrs[l]    : (N*H*W, 1, H>>l, W>>l) fp32 outputs (HOST array of DEVICE pointers), scale = 1/sqrt(C) */
PP_API int pp_correlation_pyramid(const void* f1_prep, const void* f2_prep, int N, int H, int W, int Kp, float scale,
                           int num_levels, void* const* level_ptrs, int cluster, void* stream);
/* same volumes in the tiled layout of pp_corr_lookup_tiled (written that way by the contraction's epilogue and the
 * pooling kernel; no extra pass) */
PP_API int pp_correlation_pyramid_tiled(const void* f1_prep, const void* f2_prep, int N, int H, int W, int Kp, float scale,
                                 int num_levels, void* const* level_ptrs, int cluster, void* stream);

/* Fused CorrelationPyramid + CorrLookup (model/stage3/raft_decoder.py:30-53 followed by
 * utils/corr_lookup.py:100-134, as called back to back in model/stage3/flow_decoder.py:59-61) that never builds the
 * all-pairs volume: the window samples are blended from correlations with the (2r+3)^2 integer neighbours in the
 * average-pooled FEATURE maps.
 *   pp_windowed_correlation_prepare : feat (N,C,H,W) fp32 -> out (N, (H>>level)*(W>>level), C) fp32, 2^level x 2^level
 *                                     average pooled, position-major.  Call with level 0 for feat1 and levels 0..L-1 for feat2.
 *   pp_windowed_correlation         : f1t (N,H*W,C); f2t_levels[l] (HOST array of DEVICE pointers) from the call above;
 *                                     flow (N,2,H,W) -> out (N, L*(2r+1)^2, H, W), same channel order as pp_corr_lookup. */
PP_API int pp_windowed_correlation_prepare(const float* feat, int N, int C, int H, int W, int level, float* out, void* stream);
/* same for feat1 (level 0 -> f1t) and feat2 (levels 0..L-1 -> f2t_levels[l]) in one launch */
PP_API int pp_windowed_correlation_prepare_all(const float* feat1, const float* feat2, int N, int C, int H, int W, int L,
                                        float* f1t, void* const* f2t_levels, void* stream);
PP_API int pp_windowed_correlation(const float* f1t, const void* const* f2t_levels, int L, const float* flow, int N, int C,
                            int H, int W, int radius, float* out, void* stream);

/* The same with the first layer of the reference's MotionEncoder fused in: the 1x1 convolution
 * corr_inch = L*(2r+1)^2 -> cout of MotionEncoder.corr_net[0] (model/stage3/raft_decoder.py:113-116,127-129, applied at
 * :157 to the lookup of flow_decoder.py:61) runs on the block's lookup tile while it is still in shared memory; the
 * (N, L*D*D, H, W) lookup tensor is never written.
 *   weight (cout, L*D*D) fp32 = Conv2d.weight[:, :, 0, 0]; bias (cout) or NULL; relu != 0 applies max(0, .)
 *   out (N, cout, H, W) fp32.  Covered shapes: radius <= 2, C % 32 == 0, L <= 4, cout % 8 == 0, cout * L*D*D <= 40960;
 *   anything else returns PP_ERR_ARG and is the caller's to run unfused. */
PP_API int pp_windowed_correlation_conv1x1(const float* f1t, const void* const* f2t_levels, int L, const float* flow, int N,
                                    int C, int H, int W, int radius, const float* weight, const float* bias, int cout,
                                    int relu, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Correspondence glue.
 * pp_init_correspondences replaces compute_init_correspondences, utils/correspondence.py:10-26:
 *   Ms (B,3,3) fp32, tem_mask (B,Hm,Wm) fp32 -> flow (B,2,h,w), certainty (B,1,h,w).
 * pp_stage3_correspondences replaces compute_stage3_correspondences, utils/correspondence.py:28-59
 *   (no host synchronisation, unlike the reference's torch.nonzero):
 *   flow (B,2,H,W), certainty (B,1,H,W) -> tar_pts, src_pts (B, H*W, 2) int64, flat index w*H+h.
 * ------------------------------------------------------------------------------------------ */
PP_API int pp_init_correspondences(const float* Ms, const float* tem_mask, int B, int Hm, int Wm,
                            int h, int w, float* flow, float* certainty, void* stream);
PP_API int pp_stage3_correspondences(const float* flow, const float* certainty, int B, int H, int W,
                              float threshold, int64_t* tar_pts, int64_t* src_pts, void* stream);

/* ------------------------------------------------------------------------------------------
 * Hypothesis selection between stage 1 and stages 2/3.
 * Replaces Net.select_template_data, model/picopose.py:52-70 (six torch.gather + repeat per hypothesis) and, with
 * hyp_sel < 0, the per-hypothesis loop around it (:107-110): ONE launch gathers the selected template view of every
 * detection out of every per-view tensor.
 *   tensor i: src_ptrs[i] (B, N, view_bytes[i]) -> dst_ptrs[i] (rows, view_bytes[i])   (HOST arrays, DEVICE pointers, <= 16)
 *   pred_id (B, K) int64 (pp_match_templates' out_idx)
 *   hyp_sel = k >= 0 : rows = B,   row b       <- view pred_id[b, k]
 *   hyp_sel < 0      : rows = B*K, row k*B + b <- view pred_id[b, k]   (hypothesis-major: chunk k is one stage-2/3 batch)
 * ------------------------------------------------------------------------------------------ */
PP_API int pp_select_templates(const void* const* src_ptrs, void* const* dst_ptrs, const int64_t* view_bytes, int n_tensors,
                        int B, int N, const int64_t* pred_id, int K, int hyp_sel, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PICOPOSE_B200_H */
