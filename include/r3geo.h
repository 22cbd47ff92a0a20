/*
 * r3geo.h — C ABI of libr3geo.so: the B200 (sm_100a) rotated-geometry hot path of R3Det.
 *
 * Drop-in boundary: each entry point replaces one pybind11 function of the reference's native
 * extensions (r3det/ops/<op>/src, cited per function).  Plain pointers and sizes only — no torch types.
 * Conventions
 *   - all pointers are DEVICE pointers on the current CUDA device unless the name ends in `_host`;
 *   - the caller owns every buffer (inputs, outputs, workspace); the library never allocates, frees or
 *     synchronises the device, and enqueues all work on `stream` (a cudaStream_t passed as void*);
 *   - return value 0 = success, < 0 = error; r3g_last_error() returns a thread-local message;
 *   - boxes are float32 rows <cx, cy, w, h, angle[rad]> with a row stride given in floats (5 or 6);
 *   - variant: 1/2/3 = the reference's angle/geometry conventions v1/v2/v3;
 *   - there is no CPU fallback: without a CUDA device every compute entry returns R3G_ERR_CUDA.
 */
#ifndef R3GEO_H_
#define R3GEO_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define R3G_OK 0
#define R3G_ERR_ARG (-1)       /* bad argument (null pointer, bad variant/mode, negative size ...) */
#define R3G_ERR_WORKSPACE (-2) /* workspace too small */
#define R3G_ERR_CUDA (-3)      /* CUDA runtime error (message has the cudaError string) */

#define R3G_V1 1
#define R3G_V2 2
#define R3G_V3 3

#define R3G_MODE_IOU 0
#define R3G_MODE_IOF 1

/* flags for the IoU entry points */
#define R3G_FLAG_STRICT 1       /* re-evaluate degenerate pairs with the reference's own point-set algorithm */
#define R3G_FLAG_SMALL_MASK 2   /* v3 wrapper: rows/cols with min(w,h) < 1e-3 are 0 (box_iou_rotated_wrapper.py:54-60) */
#define R3G_FLAG_EMULATE_ALL 4  /* diagnostic: every overlapping pair goes through the reference's own algorithm */

/* flags for r3g_nms_f32 */
#define R3G_NMS_INCLUSIVE 1     /* suppress when IoU >= thr (reference CPU rule); default IoU > thr (reference GPU rule) */
#define R3G_NMS_ORDER_INDEX 2   /* keep list in ascending original index (rnms_kernel.cu:331-334); default descending score */
#define R3G_NMS_DROP_SMALL 4    /* boxes with min(w,h) < 1e-3 take no part (nms_rotated_wrapper.py:40-46) */
#define R3G_NMS_STRICT 8        /* decide near-threshold / degenerate pairs with the reference's own algorithm */
#define R3G_NMS_SORT_PATH 16    /* diagnostic: take the radix-sort path that serves K > 16384 whatever K is */
#define R3G_NMS_LABEL_BITS(n) (((n) & 63) << 8) /* optional promise: every label is < 2^n (saves radix passes); 0 = unknown */

const char* r3g_last_error(void);
int r3g_version(void);

/* ---- pairwise rotated IoU / IoF -------------------------------------------------------------------
 * replaces  rbbox_geo_cuda.mat_iou_iof(rb1, rb2, iof)        r3det/ops/rbbox_geo/src/rbbox_geo_cuda.cpp:25-30   (variant 1)
 *           box_iou_rotated_ext.overlaps(b1, b2, iou_or_iof)  r3det/ops/box_iou_rotated/src/box_iou_rotated_ext.cpp:36-38 (variant 3)
 *           mmcv.ops.box_iou_rotated(b1, b2, mode)            call site rotate_iou2d_calculator.py:156          (variant 2)
 * out is row-major (m, n).  Workspace holds the prepared boxes: r3g_iou_workspace_bytes(m, n). */
int r3g_iou_workspace_bytes(int64_t m, int64_t n, size_t* bytes);
int r3g_iou_matrix_f32(const float* boxes1, int64_t m, int64_t stride1,
                       const float* boxes2, int64_t n, int64_t stride2,
                       int variant, int mode, int flags, float* out,
                       void* workspace, size_t workspace_bytes, void* stream);

/* The two halves of r3g_iou_matrix_f32, exposed so that callers with static boxes (anchors) can prepare once
 * and so that the matrix kernel can be timed alone: prepare = per-box trig/extents into the workspace
 * (O(m+n)); matrix_prepared = the pair kernel over a prepared workspace. */
int r3g_iou_prepare_f32(const float* boxes1, int64_t m, int64_t stride1,
                        const float* boxes2, int64_t n, int64_t stride2, int variant,
                        void* workspace, size_t workspace_bytes, void* stream);
int r3g_iou_matrix_prepared_f32(const float* boxes1, int64_t m, int64_t stride1,
                                const float* boxes2, int64_t n, int64_t stride2,
                                int variant, int mode, int flags, float* out,
                                void* workspace, size_t workspace_bytes, void* stream);

/* replaces rbbox_geo_cuda.vec_iou_iof (rbbox_geo_cuda.cpp:25-30): out has max(n1, n2) elements,
 * element i pairs boxes1[i % n1] with boxes2[i % n2] (rbbox_geo_kernel.cu:278-281). */
int r3g_iou_aligned_f32(const float* boxes1, int64_t n1, int64_t stride1,
                        const float* boxes2, int64_t n2, int64_t stride2,
                        int variant, int mode, int flags, float* out, void* stream);

/* ---- fused max-IoU assignment ----------------------------------------------------------
 * mmdet-2.19 MaxIoUAssigner.assign_wrt_overlaps over overlaps = calculator(gt, anchors) WITHOUT materialising the
 * (G, A) matrix (the reference path stores it: rotate_anchor_head.py:220-228 -> rotate_iou2d_calculator.py:42-43).
 * Outputs (device): assigned_gt_inds (A) int64 [-1 ignore, 0 background, i+1 = GT i]; optional (may be NULL)
 * max_overlaps (A), argmax_overlaps (A), gt_max_overlaps (G), gt_argmax_overlaps (G).  Ties resolve to the lowest
 * index like torch.max; with gt_max_assign_all every anchor whose overlap equals a GT's maximum is assigned to it. */
int r3g_assign_workspace_bytes(int64_t G, int64_t A, size_t* bytes);
int r3g_max_iou_assign_f32(const float* gt, int64_t G, int64_t gt_stride,
                           const float* anchors, int64_t A, int64_t anchor_stride,
                           int variant, int flags, float pos_iou_thr, float neg_iou_thr, float min_pos_iou,
                           int match_low_quality, int gt_max_assign_all,
                           int64_t* assigned_gt_inds, float* max_overlaps, int64_t* argmax_overlaps,
                           float* gt_max_overlaps, int64_t* gt_argmax_overlaps,
                           void* workspace, size_t workspace_bytes, void* stream);

/* counters of the last r3g_iou_matrix_f32 launch on this workspace (device-side, 4 x uint64 at the
 * start of the workspace): pairs passing the circumradius test, passing the separating-axis test,
 * re-evaluated by the strict path, total pairs.  For roofline accounting (bench.py). */
/* A batch of images in ONE launch sequence (B <= 64): gt = the images' GT boxes concatenated (gt_counts[b] rows each,
 * HOST array), anchors = (A, stride) shared by the batch (anchors_shared = 1, the anchor-head stage) or (B, A, stride)
 * contiguous (anchors_shared = 0, the refine stage's per-image boxes).  Per-anchor outputs are (B, A); per-GT outputs are
 * concatenated like gt, with indices relative to the image. */
int r3g_assign_batched_workspace_bytes(int64_t B, const int64_t* gt_counts, int64_t A, int anchors_shared, size_t* bytes);
int r3g_max_iou_assign_batched_f32(int64_t B, const float* gt, const int64_t* gt_counts, int64_t gt_stride,
                                   const float* anchors, int64_t A, int64_t anchor_stride, int anchors_shared,
                                   int variant, int flags, float pos_iou_thr, float neg_iou_thr, float min_pos_iou,
                                   int match_low_quality, int gt_max_assign_all,
                                   int64_t* assigned_gt_inds, float* max_overlaps, int64_t* argmax_overlaps,
                                   float* gt_max_overlaps, int64_t* gt_argmax_overlaps,
                                   void* workspace, size_t workspace_bytes, void* stream);
#define R3G_IOU_STATS_U64 4

/* ---- rotated NMS ------------------------------------------------------------------------------------
 * replaces  rnms_ext.rnms(dets[K,6], thr)                      r3det/ops/rnms/src/rnms_ext.cpp:11-20          (variant 1, ORDER_INDEX)
 *           nms_rotated_ext.nms_rotated(dets, scores, thr)     r3det/ops/nms_rotated/src/nms_rotated_ext.cpp:53-56 (variant 3, ORDER_SCORE)
 *           ml_nms_rotated(dets, scores, labels, thr)          r3det/ops/ml_nms_rotated/src/nms_rotated.h:23-38   (variant 2, labels)
 * boxes: (K, stride) floats; scores: (K); labels: (K) int64 in [0, 2^31) or NULL (single class).  A box
 * suppresses a lower-scored box of the same label when IoU > thr (>= with R3G_NMS_INCLUSIVE).  Candidates whose score is
 * -inf or NaN are padding: they take no part and are never kept.
 * keep_out: (K) int64 original indices, first *num_keep_out valid (both device memory).
 * With labels, `class_offset` != NULL points to ONE device float: the reference's per-class coordinate
 * offset scale (rnms_wrapper.py:61-64 / nms_rotated_wrapper.py:84-90); boxes are then evaluated at
 * x + label*scale, y + label*scale in FP32 exactly as the reference's batched wrappers do.
 * The greedy selection runs in rounds inside one persistent kernel (only pairs of kept rows are evaluated); the
 * workspace is linear in K (~0.5 KB per candidate).  K <= 2^26 per call. */
int r3g_nms_workspace_bytes(int64_t K, size_t* bytes);
int r3g_nms_f32(const float* boxes, int64_t stride, const float* scores, const int64_t* labels,
                int64_t K, float thr, int variant, int flags, const float* class_offset,
                int64_t* keep_out, int64_t* num_keep_out,
                void* workspace, size_t workspace_bytes, void* stream);

/* Multi-image form of r3g_nms_f32: one launch sequence for a whole batch of images (BASELINE configs[3]: images are
 * independent, so (image, class) pairs are simply more segments).  batch_ids: (K) int64 in [0, n_batches) or NULL
 * (n_batches = 1); labels in [0, 65536) when batch_ids is given (the segment key packs image and label into 32 bits) —
 * candidates with an image id outside [0, n_batches) or a label outside [0, 65536) take no part and are never kept;
 * class_offset: n_batches device floats (per-image scale) or
 * NULL.  keep_out: kept original indices grouped by image — ascending index (R3G_NMS_ORDER_INDEX; candidates are
 * expected to be concatenated image by image) or descending score within each image; num_keep_out: n_batches int64. */
int r3g_nms_batched_f32(const float* boxes, int64_t stride, const float* scores, const int64_t* labels,
                        const int64_t* batch_ids, int n_batches,
                        int64_t K, float thr, int variant, int flags, const float* class_offset,
                        int64_t* keep_out, int64_t* num_keep_out,
                        void* workspace, size_t workspace_bytes, void* stream);

/* The same with a device-side candidate count: the K arrays are a CAPACITY, only *count_dev candidates are real and the rest is
 * padding that must carry score -inf, image id 65535 and label 65535 (what r3g_mc_candidates_batched_f32 writes), so that it
 * sorts behind every candidate; the selection kernel then never visits it.  count_dev NULL = all K are real. */
int r3g_nms_batched_counted_f32(const float* boxes, int64_t stride, const float* scores, const int64_t* labels,
                                const int64_t* batch_ids, int n_batches,
                                int64_t K, const int64_t* count_dev, float thr, int variant, int flags, const float* class_offset,
                                int64_t* keep_out, int64_t* num_keep_out,
                                void* workspace, size_t workspace_bytes, void* stream);

/* ---- polygon NMS ----------------------------------------------------------------------
 * replaces nms_rotated_ext.nms_poly   r3det/ops/nms_rotated/src/poly_nms_cuda.cu:122-262 (mask kernel + host scan)
 * polys: K rows of `stride` >= 8 floats [x0, y0, ..., x3, y3] (arbitrary quadrilaterals), scores (K).  Greedy in
 * descending score (ties: lower index first), suppress when IoU > thr; keep_out[0 .. *num_keep_out) = kept original
 * indices in descending-score order (device memory, like r3g_nms_f32).  labels (K, or NULL): class-wise NMS in one call
 * (polygons with different labels never suppress each other) — the per-class loop of dota1.py:646-657. */
int r3g_poly_nms_workspace_bytes(int64_t K, size_t* bytes);
int r3g_poly_nms_f32(const float* polys, int64_t stride, const float* scores, const int64_t* labels, int64_t K, float thr,
                     int64_t* keep_out, int64_t* num_keep_out, void* workspace, size_t workspace_bytes, void* stream);

/* Batched, synchronisation-free form of r3g_mc_candidates_f32 for B <= 64 images of n_per_image boxes each (multi_bboxes /
 * multi_scores are the (B * n_per_image) concatenated rows).  Outputs have CAPACITY B * n_per_image * C: the candidates are
 * compacted to the front, the tail is padding with score -inf, image id and label 65535 (r3g_nms_batched_f32 gives such candidates
 * no part; r3g_nms_batched_counted_f32 with count_out does not even visit them), out_batch holds the image of every candidate and out_src its flat (row * C + class) index (-1 in the padding).  scale_out (B floats) is
 * the per-image class-offset scale of the reference's batched wrappers — offset_rule 1: max over the candidate boxes + 1
 * (rnms_wrapper.py:61-64), 2: span of their horizontal bounding boxes + 1 (nms_rotated_wrapper.py:84-90), 0: 1 — so the whole
 * multiclass NMS of a batch runs without reading anything back to the host.  Workspace: r3g_mc_candidates_workspace_bytes(B *
 * n_per_image, C) + 512 bytes. */
int r3g_mc_candidates_batched_f32(const float* multi_bboxes, int box_cols, const float* multi_scores, int64_t score_stride,
                                  int64_t n_per_image, int B, int C, float score_thr, int offset_rule,
                                  float* out_boxes, float* out_scores, int64_t* out_labels, int64_t* out_batch,
                                  int64_t* out_src, int64_t* count_out, float* scale_out,
                                  void* workspace, size_t workspace_bytes, void* stream);

/* ---- padded keep records ------------------------------------------------------------------------------------
 * replaces the per-image `dets[keep][:max_num]` slicing after NMS (r3det/core/post_processing/bbox_nms_rotated.py:127-131,
 * rotate_anchor_head.py:626-673), which needs a host read of the keep count per image.  keep / num_keep are the outputs of
 * r3g_nms_f32 / r3g_nms_batched_f32 (keep grouped by image).  Image b's first min(num_keep[b] - drop_last, max_per_img)
 * kept candidates are written to out_dets[b, o, 0..5] = <x, y, w, h, a, score> and out_labels[b, o]; the remaining rows are
 * zero; out_counts[b] = rows filled.  drop_last = 1 reproduces the `v2` branch's `inds[:-1]` for max_num = -1 (:119-121).
 * n_batches <= 8192. */
int r3g_nms_pack_f32(const float* boxes, int64_t stride, const float* scores, const int64_t* labels,
                     const int64_t* keep, const int64_t* num_keep, const int64_t* batch_ids, int n_batches, int64_t K,
                     int max_per_img, int drop_last, float* out_dets, int64_t* out_labels, int64_t* out_counts, void* stream);

/* ---- multiclass candidate extraction -----------------------------------------------------------------------
 * replaces the torch prologue of multiclass_nms_rotated (r3det/core/post_processing/bbox_nms_rotated.py:34-41,
 * 98-103): candidates = (box, class) pairs with multi_scores[i, c] > score_thr for c < C (the last, background,
 * column is never read), enumerated row-major.  multi_bboxes is (n, 5) or (n, 5*C) (box_cols); score rows have
 * `score_stride` floats.  Outputs are sized for n*C candidates; *count_out (device) receives K.
 * out_src[k] = i*C + c.  score_factors (n) or NULL multiplies the emitted score (:100-101). */
int r3g_mc_candidates_workspace_bytes(int64_t n, int C, size_t* bytes);
int r3g_mc_candidates_f32(const float* multi_bboxes, int box_cols, const float* multi_scores, int64_t score_stride,
                          const float* score_factors, int64_t n, int C, float score_thr,
                          float* out_boxes, float* out_scores, int64_t* out_labels, int64_t* out_src,
                          int64_t* count_out, void* workspace, size_t workspace_bytes, void* stream);

/* ---- FRM feature refinement ---------------------------------------------------------------------------
 * replaces feature_refine_cuda.forward / .backward   r3det/ops/fr/src/feature_refine_cuda.cpp:24-67
 * feat/out/grad: (N, C, H, W) float32 contiguous; boxes: (N*H*W, 5); points in {1, 5}.
 * forward:  out = feat + sum_p bilinear(feat, y_p, x_p)   (out need not be pre-zeroed)
 * backward: grad_in = grad_out + scatter of bilinear weights, atomic-free (sorted taps in workspace). */
int r3g_frm_forward_f32(const float* feat, const float* boxes, int N, int C, int H, int W,
                        float spatial_scale, int points, float* out, void* stream);
int r3g_frm_backward_workspace_bytes(int N, int H, int W, int points, size_t* bytes);
int r3g_frm_backward_f32(const float* grad_out, const float* boxes, int N, int C, int H, int W,
                         float spatial_scale, int points, float* grad_in,
                         void* workspace, size_t workspace_bytes, void* stream);

/* All FPN levels of a batch in ONE launch sequence (FeatureRefineModule.forward loops over the levels,
 * r3det/ops/fr/feature_refine_module.py:112-127).  Pointer arrays, level_hw {H_0, W_0, H_1, ...} and spatial_scales
 * are HOST arrays of L <= 8 entries; level l holds (N, C, H_l, W_l) features and (N*H_l*W_l, 5) boxes.
 * residuals (array or NULL; entries may be NULL): out_l = residuals_l + FRM(feats_l) — the module's final
 * `x_scale + feat_refined_scale` (:126) folded into the epilogue.  The backward shares one tap sort / CSR build. */
int r3g_frm_forward_multi_f32(int L, const float* const* feats, const float* const* boxes, const float* const* residuals,
                              int N, int C, const int* level_hw, const float* spatial_scales, int points,
                              float* const* outs, void* stream);
int r3g_frm_backward_multi_workspace_bytes(int L, int N, const int* level_hw, int points, size_t* bytes);
int r3g_frm_backward_multi_f32(int L, const float* const* grad_outs, const float* const* boxes, int N, int C,
                               const int* level_hw, const float* spatial_scales, int points, float* const* grad_ins,
                               void* workspace, size_t workspace_bytes, void* stream);

/* The same backward in two halves: PLAN depends on the boxes only (tap sort + per-target CSR into the workspace) and can
 * run on a side stream while the forward pass executes; APPLY is the gather.  plan + apply == r3g_frm_backward_multi_f32
 * bit for bit.  The workspace (r3g_frm_backward_multi_workspace_bytes) must stay untouched between the two calls, and the
 * level list (empty levels included) must be the same. */
int r3g_frm_backward_plan_multi_f32(int L, const float* const* boxes, int N, const int* level_hw, const float* spatial_scales,
                                    int points, void* workspace, size_t workspace_bytes, void* stream);
int r3g_frm_backward_apply_multi_f32(int L, const float* const* grad_outs, const float* const* boxes, int N, int C,
                                     const int* level_hw, const float* spatial_scales, int points, float* const* grad_ins,
                                     const void* workspace, size_t workspace_bytes, void* stream);

/* ---- box transforms (r3det/core/bbox/rtransforms.py) ---------------------------------------------------
 * obb2poly :367-440, poly2obb :190-277, obb2hbb :443-537, hbb2obb :540-592, obb2xyxy :595-651.
 * n boxes; version 1/2/3. */
int r3g_obb2poly_f32(const float* obb, int64_t n, int version, float* poly, void* stream);
int r3g_poly2obb_f32(const float* poly, int64_t n, int version, float* obb, void* stream);
int r3g_obb2hbb_f32(const float* obb, int64_t n, int version, float* hbb, void* stream);
int r3g_hbb2obb_f32(const float* hbb, int64_t n, int version, float* obb, void* stream);
int r3g_obb2xyxy_f32(const float* obb, int64_t n, int version, float* xyxy, void* stream);

/* ---- rotated box coder and the steps around NMS / FRM that use it -------------------
 * replaces DeltaXYWHAOBBoxCoder.encode / .decode   r3det/core/bbox/coder/delta_xywha_rbbox_coder.py:46-101
 *   (bbox2delta_v1/v2/v3 :104-139 / :214-252 / :314-360, delta2bbox_v1/v2/v3 :142-211 / :255-311 / :363-423)
 * means / stds: 5 floats each in HOST memory.  deltas: (n, 5*groups), rois: n rows of `roi_stride` floats, out like
 * deltas.  max_shape_hw: HOST {H, W} or NULL; honoured by v1 only, exactly as in the reference (:97-99). */
int r3g_delta2bbox_f32(const float* rois, int64_t n, int64_t roi_stride, const float* deltas, int64_t groups,
                       const float* means, const float* stds, int variant, const int* max_shape_hw,
                       double wh_ratio_clip, int add_ctr_clamp, float ctr_clamp, float* out, void* stream);
int r3g_bbox2delta_f32(const float* proposals, int64_t proposal_stride, const float* gt, int64_t gt_stride, int64_t n,
                       const float* means, const float* stds, int variant, float* out, void* stream);
/* replaces RRetinaHead.filter_bboxes   r3det/models/dense_heads/rotate_retina_head.py:117-179   for one level:
 * cls_score (B, A*C, H, W), bbox_pred (B, A*5, H, W), anchors (H*W*A, 5) -> out (B, H*W, 5): per location the anchor
 * with the largest class logit (first maximum), decoded.  Network outputs are read in NCHW, no permuted copies. */
int r3g_filter_bboxes_f32(const float* cls_score, const float* bbox_pred, const float* anchors, int64_t B, int64_t A,
                          int64_t C, int64_t H, int64_t W, const float* means, const float* stds, int variant,
                          double wh_ratio_clip, int add_ctr_clamp, float ctr_clamp, float* out, void* stream);
/* replaces RRetinaRefineHead.refine_bboxes   r3det/models/dense_heads/rotate_retina_refine_head.py:56-97   for one
 * level: bbox_pred (B, 5, H, W), rois (B, H*W, 5) -> out (B, H*W, 5) — the (N*H*W, 5) layout r3g_frm_* consume. */
int r3g_refine_bboxes_f32(const float* bbox_pred, const float* rois, int64_t B, int64_t H, int64_t W, const float* means,
                          const float* stds, int variant, double wh_ratio_clip, int add_ctr_clamp, float ctr_clamp,
                          float* out, void* stream);
/* replaces RAnchorHead._get_bboxes_single up to the NMS call, for a whole batch
 *   r3det/models/dense_heads/rotate_anchor_head.py:626-664 (use_sigmoid_cls=True):
 * per level l < L: cls_scores[l] (B, A*C, H_l, W_l), bbox_preds[l] (B, A*5, H_l, W_l), anchors[l] (H_l*W_l*A, 5) shared
 * by the batch (anchor_batch_strides[l] == 0 or NULL array) or per image (stride in floats).  The pointer arrays,
 * level_hw {H_0, W_0, H_1, ...}, max_shapes_hw {H, W} per image (or NULL; v1 only) and scale_factors (4 per image, or
 * NULL = no rescale) are HOST arrays.  Levels with more than nms_pre rows keep their nms_pre best rows by
 * max-class sigmoid score in descending order (ties: lower row index); smaller levels are kept whole in row order.
 * boxes_out (B, rows_per_image, 5), scores_out (B, rows_per_image, C + 1) with a zero last column.
 * At most 8 levels and 64 images per call. */
int r3g_select_decode_sizes(int64_t L, int64_t B, int64_t A, const int64_t* level_hw, int64_t nms_pre,
                            int64_t* rows_per_image, size_t* workspace_bytes);
int r3g_select_decode_f32(int64_t L, const float* const* cls_scores, const float* const* bbox_preds,
                          const float* const* anchors, const int64_t* anchor_batch_strides, const int64_t* level_hw,
                          int64_t B, int64_t A, int64_t C, int64_t nms_pre,
                          const float* means, const float* stds, int variant, double wh_ratio_clip, int add_ctr_clamp,
                          float ctr_clamp, const int* max_shapes_hw, const float* scale_factors,
                          float* boxes_out, float* scores_out, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* R3GEO_H_ */
