/*
 * b2d_glue.h - C ABI of the B200 (sm_100a) two-stage detection glue.
 *
 * Drop-in boundary for the hot path of mathild7/faster_rcnn_pytorch_multimodal
 * (SURVEY.md §8).  The reference has no native layer of its own: its hot path calls
 * torchvision's `_C.so` (nms, roi_align) and ATen.  Each entry point below names the
 * reference Python interface (file:line under /root/reference/lib) it replaces.
 *
 * Conventions
 *  - Every pointer is a DEVICE pointer unless the name ends in `_host`.
 *  - All tensors are dense fp32 / int32 in the layouts stated per function.
 *  - `stream` is a cudaStream_t passed as void*.  Calls are stream-ordered and never
 *    synchronise the device, except the `_host` entry points, which block until
 *    their outputs are in host memory.
 *  - The library allocates nothing and keeps no global state: the caller owns every
 *    buffer, including `workspace` (size from the matching *_workspace_bytes()).
 *  - Return value: 0 on success, negative b2d_status otherwise.  Nothing throws.
 *  - Re-entrant; concurrent calls are safe on distinct streams + workspaces.
 *  - The batched stages (proposal, NMS, RoIAlign, anchor targets, final detections, the host
 *    pipeline) take a leading `num_frames`; frames are independent.  The element-wise codecs,
 *    the RoI sampler, the MC reductions and the BEV rasteriser take a flat row count: stack the
 *    rows of several frames to batch them.
 */
#ifndef B2D_GLUE_H_
#define B2D_GLUE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum b2d_status {
  B2D_OK = 0,
  B2D_ERR_INVALID_ARG = -1,
  B2D_ERR_WORKSPACE = -2,   /* workspace NULL or too small */
  B2D_ERR_CUDA = -3,        /* launch / runtime error (see b2d_last_cuda_error) */
  B2D_ERR_UNSUPPORTED = -4  /* shape outside the compiled limits */
} b2d_status;

int b2d_abi_version(void);
const char* b2d_status_string(int status);
/* cudaError_t of the last failing runtime call made by this thread (0 if none). */
int b2d_last_cuda_error(void);
/* Number of kernels this library launched since load (all threads). */
uint64_t b2d_launch_count(void);
/* Capacity of the in-CTA sort: up to this many boxes (pre_nms_topN, argsort n) sort in shared memory; more take
 * the chunk-sort + merge path (no upper limit). */
int b2d_max_pre_nms(void);

/* ------------------------------------------------------------------------------------
 * Proposal stage: fg-score gather + pre-NMS top-k (radix select + in-CTA sort) +
 * bbox_transform_inv + clip_boxes + greedy NMS + post-NMS gather.
 * Replaces proposal_layer()        layer_utils/proposal_layer.py:18-57
 *          bbox_transform_inv()     model/bbox_transform.py:75-105   (fused, K=1)
 *          clip_boxes()             model/bbox_transform.py:235-257  (fused)
 *          torchvision.ops.nms      call site proposal_layer.py:46
 *
 * Inputs (per frame f of num_frames):
 *   cls_prob   [F, n_loc, 2A]   fg score of anchor a at location l = cls_prob[f][l][A + a]
 *   bbox_pred  [F, n_loc, 4A]   deltas (dx,dy,dw,dh) of flat anchor n = l*A + a at [f][n*4..]
 *   info       [F, 7]           x_min,x_max,y_min,y_max,z_min,z_max,scale (only 0..3 used)
 *   anchors    [N, 4]           N = n_loc*A, shared by all frames
 *   anchors_3d [N, 7] or NULL   gathered alongside (lidar); shared by all frames
 * Parameters: pre_nms (<=0: all N), post_nms (<=0: no cap), nms_thresh (double, compared
 * like torchvision's CPU kernel: (double)iou > nms_thresh).
 * Outputs (padded to `max_out` rows per frame, rows >= num_out[f] are zero):
 *   rois        [F, max_out, 5]  col0 = f * batch_index_stride (0 reproduces the reference)
 *   roi_scores  [F, max_out]
 *   roi_a3d     [F, max_out, 7]  or NULL
 *   roi_anchor  [F, max_out]     int32 flat anchor index of each kept proposal, or NULL
 *   num_out     [F]              int32
 * max_out = post_nms if post_nms > 0 else min(pre_nms, N).
 * Tie order of equal scores: lower flat index first.
 * ---------------------------------------------------------------------------------- */
size_t b2d_proposal_workspace_bytes(int num_frames, int n_loc, int num_anchors, int pre_nms, int post_nms);
int b2d_proposal(int num_frames, int n_loc, int num_anchors,
                 const float* cls_prob, const float* bbox_pred, const float* info,
                 const float* anchors, const float* anchors_3d,
                 int pre_nms, int post_nms, double nms_thresh, int batch_index_stride,
                 float* rois, float* roi_scores, float* roi_a3d, int32_t* roi_anchor, int32_t* num_out,
                 void* workspace, size_t workspace_bytes, void* stream);

/* proposal_top_layer()  layer_utils/proposal_top_layer.py:18-59 (n >= top_n branch):
 * top_n by score (same tie rule), decode + clip AFTER selection, no NMS.
 *   rois [F, top_n, 5], roi_scores [F, top_n], roi_anchors [F, top_n, 4] (selected anchors). */
int b2d_proposal_top(int num_frames, int n_loc, int num_anchors,
                     const float* cls_prob, const float* bbox_pred, const float* info, const float* anchors,
                     int top_n, int batch_index_stride,
                     float* rois, float* roi_scores, float* roi_anchors,
                     void* workspace, size_t workspace_bytes, void* stream);

/* Intermediate products of the proposal stage, for parity tests ("keep-indices bit-exact
 * given identical decoded boxes"): after b2d_proposal() returns, the workspace holds the
 * pre-NMS sorted list.  These copy it out (device to device, stream-ordered).
 *   sorted_boxes [F, k, 4], sorted_scores [F, k], sorted_index [F, k] (int32), k = min(pre_nms, N). */
int b2d_proposal_debug_sorted(int num_frames, int n_loc, int num_anchors, int pre_nms, int post_nms,
                              const void* workspace, float* sorted_boxes, float* sorted_scores,
                              int32_t* sorted_index, void* stream);

/* ------------------------------------------------------------------------------------
 * Greedy NMS on boxes already sorted by descending score (torchvision semantics:
 * area=(x2-x1)*(y2-y1), suppress iff (double)iou > thresh, NaN never suppresses).
 * Replaces torchvision.ops.nms at proposal_layer.py:46, filter_predictions.py:67,69.
 *   boxes [F, n, 4]; n_valid [F] int32 or NULL (all n valid);
 *   keep [F, max_keep] int32 positions (ascending), num_keep [F] int32.
 * The sweep stops once max_keep boxes are kept (== keep[:max_keep] of the full result).
 * ---------------------------------------------------------------------------------- */
/* workspace: optional scratch of b2d_nms_workspace_bytes() bytes.  With it, calls of a few frames run one
 * 16-CTA thread-block cluster per frame (chunked bitmask + one-warp sweep) instead of one CTA per frame. */
size_t b2d_nms_workspace_bytes(int num_frames, int max_keep);
int b2d_nms_sorted(int num_frames, int n, const float* boxes, const int32_t* n_valid, double thresh,
                   int max_keep, int32_t* keep, int32_t* num_keep, void* workspace, size_t workspace_bytes,
                   void* stream);

/* Stable descending sort of scores (ties: lower index first), order [F, n] int32.  Used by the nms() wrapper for
 * unsorted input.  n <= b2d_max_pre_nms() sorts in one CTA per frame and needs no workspace; longer lists are
 * chunk-sorted and merged in b2d_argsort_workspace_bytes() bytes of scratch. */
size_t b2d_argsort_workspace_bytes(int num_frames, int n);
int b2d_argsort_desc(int num_frames, int n, const float* scores, int32_t* order, void* workspace,
                     size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * RoIAlign forward / backward (torchvision.ops.roi_align semantics, NCHW fp32).
 * Replaces Network._crop_pool_layer (missing lib/nets/network.py) and
 * torchvision.ops.roi_align at utils/torchpoolers.py:165-170,194-197.
 *   feat [F, C, H, W]; rois [R, 5] (col0 = frame index in [0,F));
 *   roi_ids [n_roi_ids] int32 or NULL: the RoI *list* is rois[roi_ids[e]] (FPN level lists),
 *     else the list is rows 0..R-1;
 *   seg_count [F] int32 or NULL with seg_stride: frame f owns list entries
 *     [f*seg_stride, f*seg_stride + seg_count[f]) (the padded layout b2d_proposal() emits);
 *     entries past seg_count[f] produce zero rows.  When NULL every frame filters the list by col0
 *     (rows whose col0 is outside [0,F) come back as zeros).
 *   out  [R, C, PH, PW]; grad_out same shape; grad_feat [F, C, H, W] (fully overwritten
 *   unless accumulate != 0).
 * Backward is deterministic and uses no atomics.
 * ---------------------------------------------------------------------------------- */
/* per_frame: most list entries one frame can own (seg_stride; or num_rois when frames filter by col0). */
size_t b2d_roi_align_workspace_bytes(int num_frames, int channels, int height, int width, int num_rois,
                                     int per_frame);
int b2d_roi_align_forward(int num_frames, int channels, int height, int width,
                          const float* feat, const float* rois, int num_rois,
                          const int32_t* roi_ids, int n_roi_ids, const int32_t* seg_count, int seg_stride,
                          int pooled_h, int pooled_w, float spatial_scale, int sampling_ratio, int aligned,
                          float* out, void* workspace, size_t workspace_bytes, void* stream);
/* Same call with the kernel family pinned (parity tests cover every family on small shapes that the
 * automatic dispatch would send elsewhere): AUTO = what b2d_roi_align_forward does; ROWS = streaming
 * kernels, never the small-call gather; ROWS_COOP = rows kernel with its cp.async fill; PLANES = plane
 * resident kernel; GATHER = one thread per output. */
typedef enum b2d_roi_route {
  B2D_ROI_ROUTE_AUTO = 0,
  B2D_ROI_ROUTE_ROWS = 1,
  B2D_ROI_ROUTE_ROWS_COOP = 2,
  B2D_ROI_ROUTE_PLANES = 3,
  B2D_ROI_ROUTE_GATHER = 4
} b2d_roi_route;
int b2d_roi_align_forward_route(int num_frames, int channels, int height, int width,
                                const float* feat, const float* rois, int num_rois,
                                const int32_t* roi_ids, int n_roi_ids, const int32_t* seg_count, int seg_stride,
                                int pooled_h, int pooled_w, float spatial_scale, int sampling_ratio, int aligned,
                                int route, float* out, void* workspace, size_t workspace_bytes, void* stream);
int b2d_roi_align_backward(int num_frames, int channels, int height, int width,
                           const float* grad_out, const float* rois, int num_rois,
                           const int32_t* roi_ids, int n_roi_ids, const int32_t* seg_count, int seg_stride,
                           int pooled_h, int pooled_w, float spatial_scale, int sampling_ratio, int aligned,
                           int accumulate, float* grad_feat, void* workspace, size_t workspace_bytes,
                           void* stream);

/* MultiScaleRoIAlign.forward  utils/torchpoolers.py:137-200 for small calls (one frame's RoIs) in ONE launch:
 * replaces the per-level nonzero / index_select / roi_align / scatter loop :187-199.
 *   feats[num_levels] device pointers to [F, C, heights[l], widths[l]] fp32 maps (HOST arrays of num_levels
 *   entries, read during the call), scales[l] = spatial scale of level l, rois [R,5], levels [R] int32
 *   (b2d_fpn_level_map) -> out [R, C, pooled_h, pooled_w], every row written.  One thread per output element
 *   straight from L2: meant for R * C * pooled_h * pooled_w up to a few million; larger batches run
 *   b2d_roi_align_forward once per level with an index list. */
int b2d_roi_align_forward_levels(int num_levels, int num_frames, int channels, const float* const* feats,
                                 const int32_t* heights, const int32_t* widths, const float* scales,
                                 const float* rois, const int32_t* levels, int num_rois, int pooled_h, int pooled_w,
                                 int sampling_ratio, int aligned, float* out, void* stream);

/* FPN level assignment, LevelMapper.__call__  utils/torchpoolers.py:39-51.
 *   boxes [R,4] -> levels [R] int32 in [0, k_max-k_min]. */
int b2d_fpn_level_map(int num_rois, const float* boxes, int k_min, int k_max, float canonical_scale,
                      int canonical_level, float eps, int32_t* levels, void* stream);

/* ------------------------------------------------------------------------------------
 * Box codecs and IoU.
 * ---------------------------------------------------------------------------------- */
/* bbox_overlaps()  utils/bbox.py:5-33  -> overlaps [n, k] (+1 pixel convention). */
int b2d_bbox_overlaps(int n, int k, const float* boxes, int box_stride, const float* query, int query_stride,
                      float* overlaps, void* stream);
/* bbox_transform()  model/bbox_transform.py:52-70 -> targets [n,4]. */
int b2d_bbox_transform(int n, const float* ex_rois, int ex_stride, const float* gt_rois, int gt_stride,
                       float* targets, void* stream);
/* bbox_transform_inv() model/bbox_transform.py:75-105; deltas/out [n, 4*k]; inv_scale
 * multiplies boxes first when use_scale (boxes/scales).  clip != 0 also applies clip_boxes
 * with info[0..3]. */
int b2d_bbox_transform_inv(int n, int k, const float* boxes, int box_stride, const float* deltas,
                           int use_scale, float scale, int clip, const float* info, float* out, void* stream);
/* clip_boxes()  model/bbox_transform.py:235-257; boxes [n, 4*k] -> out. */
int b2d_clip_boxes(int n, int k, const float* boxes, const float* info, float* out, void* stream);
/* lidar_3d_bbox_transform()  model/bbox_transform.py:16-49 -> targets [n,7]. */
int b2d_lidar_bbox_transform(int n, const float* ex_rois, int roi_stride, const float* ex_anchors,
                             const float* gt_rois, int gt_stride, float* targets, void* stream);
/* lidar_3d_bbox_transform_inv() :174-233 and lidar_3d_uncertainty_transform_inv() :132-169.
 * rois [n,4] (stride), boxes [n,7], deltas [n,7k]; mode 0 = decode, 1 = uncertainty (deltas =
 * uncertainty).  The caller applies the reference's in-place scale mutation. */
int b2d_lidar_bbox_transform_inv(int n, int k, const float* rois, int roi_stride, const float* boxes,
                                 const float* deltas, int mode, float* out, void* stream);
/* bbaa_graphics_gems_torch()  utils/bbox.py:296-336; boxes [n,7] -> aabb [n,4]. */
int b2d_bbaa(int n, const float* boxes7, int clip, float width, float height, float* aabb, void* stream);
/* generate_anchors_pre()  layer_utils/snippets.py:13-40: base [A,4] fp64 (host) tiled over H x W. */
int b2d_generate_anchors(int height, int width, int feat_stride, int num_base, const double* base_host,
                         float* anchors, void* stream);

/* ------------------------------------------------------------------------------------
 * Training targets.
 * ---------------------------------------------------------------------------------- */
/* anchor_target_layer_torch()  layer_utils/anchor_target_layer.py:22-165, split at its two
 * randperm() draws so that the host can draw them from torch's generator exactly as the
 * reference does:
 *   phase 1: inside filter, IoU, argmax both ways, labels before subsampling; writes ordered
 *            fg / bg index lists (positions among inside anchors, ascending) and counts[F,4] =
 *            {n_inside, n_fg, n_bg, 0}.
 *   phase 2: applies the disable lists (fg_disable / bg_disable: positions into the fg / bg
 *            lists, i.e. perm[num_fg:] and perm[num_bg:]), computes targets and weights and
 *            writes the four outputs in the reference's layouts:
 *            labels [F, A, H, W], bbox_targets / inside_w / outside_w [F, H, W, 4A].
 * gt_boxes [F, G, 5] (x1,y1,x2,y2,cls) with num_gt[F] valid rows each.
 */
size_t b2d_anchor_target_workspace_bytes(int num_frames, int n_anchors_total, int max_gt);
int b2d_anchor_target_phase1(int num_frames, int n_total, int max_gt, const float* anchors,
                             const float* gt_boxes, const int32_t* num_gt, const float* info,
                             float neg_overlap, float pos_overlap, int clobber_positives,
                             int32_t* counts, void* workspace, size_t workspace_bytes, void* stream);
int b2d_anchor_target_phase2(int num_frames, int n_total, int max_gt, int num_anchors, int height, int width,
                             const float* anchors, const float* gt_boxes, const int32_t* counts,
                             const int64_t* fg_disable, const int32_t* n_fg_disable,
                             const int64_t* bg_disable, const int32_t* n_bg_disable, int disable_stride,
                             const float* inside_weights4, float positive_weight,
                             float* labels, float* bbox_targets, float* inside_w, float* outside_w,
                             void* workspace, size_t workspace_bytes, void* stream);

/* proposal_target_layer()  layer_utils/proposal_target_layer.py:22-262, split the same way:
 *   phase 1: IoU(rois, gt) -> max_overlap[R], assignment[R] (int32), ordered fg / bg lists
 *            (bg_mode 0 = reference-as-run: never any bg; 1 = [lo,hi)), counts {n_fg, n_bg}.
 *   phase 2: gathers keep_inds [S] (int64, from the host-side sampler) and writes
 *            labels [S,1], rois [S,5], anchors_3d [S,7], scores [S], bbox_targets / inside_w /
 *            outside_w [S, K*E] (E = 4 image codec, 7 lidar codec), normalised by means/stds[E].
 */
int b2d_proposal_target_phase1(int num_rois, int num_gt, const float* rois, const float* gt_boxes,
                               float fg_thresh, float bg_hi, float bg_lo, int bg_mode,
                               float* max_overlap, int32_t* assignment, int32_t* fg_list, int32_t* bg_list,
                               int32_t* counts, void* stream);
int b2d_proposal_target_phase2(int num_keep, int fg_count, const int64_t* keep_inds,
                               const float* rois, const float* scores, const float* anchors_3d,
                               const float* gt_boxes, const float* true_gt_boxes, const int32_t* assignment,
                               int num_classes, int num_elem, int normalize, const float* means,
                               const float* stds, float* labels, float* out_rois, float* out_a3d,
                               float* out_scores, float* bbox_targets, float* inside_w, float* outside_w,
                               void* stream);

/* ------------------------------------------------------------------------------------
 * MC-dropout reductions.
 * compute_bbox_var()  utils/loss_utils.py:114-120: samples [T, m] -> var [m]
 *   (single-pass (sum x^2 - (sum x)^2/T)/(T-1), clamp_min 0, as the reference).
 * mode 1 = compute_bbox_cov() diagonal, :103-112: mean(x^2) - mean(x)^2, clamp_min 0.
 * categorical_mutual_information() :132-141: logits [T, n, K] -> mi [n]; entropy of the mean
 * softmax -> ent [n] (either output may be NULL).
 * var_sort: argsort of mean_j var[r, j] (datasets/db.py:264-303), ties by lower index.
 * ---------------------------------------------------------------------------------- */
int b2d_mc_variance(int num_samples, int m, const float* samples, int mode, float* var, void* stream);
int b2d_mc_class_uncertainty(int num_samples, int n, int num_classes, const float* logits, float* mutual_info,
                             float* entropy, void* stream);
int b2d_var_sort(int n, int cols, const float* var, int descending, float* key, int32_t* order, void* stream);

/* ------------------------------------------------------------------------------------
 * Tail of the detection head over the MC-dropout stack, ONE launch for all frames.
 * Restates, element by element: per-class de-normalisation `bbox_pred * stds + means`  model/config.py:219-223;
 * the mean over the T samples and compute_bbox_var  utils/loss_utils.py:114-120;
 * lidar_3d_bbox_transform_inv  model/bbox_transform.py:174-233 (mode 1) or bbox_transform_inv + clip_boxes
 * :75-105,235-257 (mode 0); lidar_3d_uncertainty_transform_inv  :132-169 on the epistemic variance and on an
 * optional aleatoric variance input; softmax + mean class probability, categorical_entropy and
 * categorical_mutual_information  utils/loss_utils.py:122-141.  (Network.test_frame, which strings these
 * together, is in the missing lib/nets/network.py: the composition is [INFERRED])
 *   bbox_pred [T, F, R, K*E], cls_score [T, F, R, K]: T head passes over the same F*R RoIs, stacked;
 *   rois [F, R, 5]; anchors_3d [F, R, 7] (mode 1); info [F, 7]; a_bbox_var_in [F, R, K*E] or NULL;
 *   means_host / stds_host [E]: HOST arrays (copied into the launch);
 *   use_scale: divide the RoIs by info[f][6] first; clip (mode 0): clip_boxes to info.
 * Outputs, in b2d_final_detections' input layout: boxes [F, R, K*E]; probs [F, R, K];
 *   e_bbox_var [F, R, K*E] (0 when T == 1) or NULL; a_bbox_var [F, R, K*E] or NULL; entropy [F, R] or NULL;
 *   mutual_info [F, R] or NULL.  mode 0 returns raw variances (the reference's image flavour of the
 *   uncertainty transform is unusable).  K <= 16.
 * ---------------------------------------------------------------------------------- */
int b2d_head_tail_decode(int num_frames, int num_samples, int num_rois, int num_classes, int num_elem,
                         const float* bbox_pred, const float* cls_score, const float* rois, const float* anchors_3d,
                         const float* info, const float* a_bbox_var_in, const float* means_host,
                         const float* stds_host, int mode, int use_scale, int clip, float* boxes, float* probs,
                         float* e_bbox_var, float* a_bbox_var, float* entropy, float* mutual_info, void* stream);

/* ------------------------------------------------------------------------------------
 * LiDAR BEV rasterisation for one frame of points.
 * Replaces the CPU path of _get_lidar_blob()  roi_data_layer/minibatch.py:428-512:
 *   filter_points() :232-235, z shift :454, spconv.utils.VoxelGeneratorV2.generate() :445-456
 *   (spconv==1.0, req.txt:261: voxels in order of first appearance, capped at max_voxels; the
 *   first max_pts_per_voxel points of a voxel in input order), the per-voxel max-height slices
 *   :463-479, the density / tanh(mean intensity) / tanh(mean elongation) channels :481-509 (last
 *   voxel of an (x, y) column wins, as the reference's fancy-index assignment does) and the
 *   transpose :512.
 *   points [num_points, num_feat] fp32 on the device (x, y, z, intensity[, elongation]);
 *   ranges = cfg.LIDAR.{X,Y,Z}_RANGE, voxel_len = cfg.LIDAR.VOXEL_LEN / scale;
 *   bev_map [ny, nx, nz + num_meta] fp32 (fully written, zeros included);
 *   num_voxels (optional, device) = voxels kept.
 * ---------------------------------------------------------------------------------- */
size_t b2d_bev_workspace_bytes(int max_points, int nx, int ny, int nz);
int b2d_bev_rasterize(int num_points, int num_feat, const float* points, float x_lo, float x_hi, float y_lo,
                      float y_hi, float z_lo, float z_hi, float voxel_len, float voxel_height, int nx, int ny,
                      int nz, int max_pts_per_voxel, int max_voxels, int num_meta, int elongation,
                      float* bev_map, int32_t* num_voxels, void* workspace, size_t workspace_bytes,
                      void* stream);

/* ------------------------------------------------------------------------------------
 * Final per-class detection filter, batched over frames and classes.
 * Replaces nms_hstack_torch() / filter_and_draw_prep()  utils/filter_predictions.py:45-130
 *          and the max-dets filter of the test loop       model/test.py:213-221.
 *   cls_score [F,R,K]; pred_boxes [F,R,K*E] (already decoded); num_rois [F] or NULL; info [F,7].
 *   lidar = 0: boxes clamped to [0, w/scale-1] x [0, h/scale-1] (:82-91), NMS on the box;
 *   lidar = 1: NMS on the un-rotated AABB of (xc, yc, l, w) = elements 0,1,3,4 (:58-62);
 *   lidar = 2: image boxes taken as they are (nms_hstack_torch() called on its own, :45-72).
 *   Per class c >= 1: rois with score > score_thresh, descending score (ties: lower roi first),
 *   greedy NMS (torchvision semantics, (double)iou > nms_thresh), then max_dets (> 0): keep
 *   scores >= the max_dets-th best (ties kept).
 *   uc_row [F,R,n_uc_row] per-roi uncertainty columns (entropy, mutual information, ...) and
 *   uc_cls [F,R,n_uc_cls,K*E] per-class-box columns (bbox variances) are gathered alongside.
 * Outputs, padded to max_out rows per (frame, class) with zeros / -1:
 *   dets [F,K,max_out,E+1] = box, score; det_roi [F,K,max_out] source roi; counts [F,K];
 *   out_uc_row [F,K,max_out,n_uc_row]; out_uc_cls [F,K,max_out,n_uc_cls*E].  R <= 4096 (shared-memory staging of one frame x class).
 * ---------------------------------------------------------------------------------- */
int b2d_final_detections(int num_frames, int num_rois_max, int num_classes, int num_elem,
                         const float* cls_score, const float* pred_boxes, const int32_t* num_rois,
                         const float* info, int lidar, float score_thresh, double nms_thresh, int max_dets,
                         int max_out, const float* uc_row, int n_uc_row, const float* uc_cls, int n_uc_cls,
                         float* dets, int32_t* det_roi, float* out_uc_row, float* out_uc_cls, int32_t* counts,
                         void* stream);

/* ------------------------------------------------------------------------------------
 * End-to-end entry with HOST buffers (bench.py `e2e`): H2D of a frame batch, proposal
 * stage, RoIAlign forward, D2H of rois / scores / counts / pooled features.  Host buffers
 * should be pinned.  device_ws must hold b2d_pipeline_device_bytes().  Blocks until done.
 * ---------------------------------------------------------------------------------- */
size_t b2d_pipeline_device_bytes(int num_frames, int n_loc, int num_anchors, int channels, int height,
                                 int width, int pre_nms, int post_nms, int pooled);
int b2d_proposal_crop_host(int num_frames, int n_loc, int num_anchors, int channels, int height, int width,
                           const float* cls_prob_host, const float* bbox_pred_host, const float* info_host,
                           const float* anchors_dev, const float* feat_host,
                           int pre_nms, int post_nms, double nms_thresh,
                           int pooled, float spatial_scale, int sampling_ratio,
                           float* rois_host, float* scores_host, int32_t* num_out_host, float* pooled_host,
                           void* device_ws, size_t device_ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Result post-processing and evaluation.
 *
 * b2d_bbox_voxel_grid_to_pc: utils/bbox.py:140-162, applied by model/test.py:224 to a lidar frame's detections
 *   before they are stacked into all_boxes.  boxes [n, row_stride] fp32, transformed in place:
 *   x = x*fx + x0, y = y*fy + y0, then (aabb) x2, y2 likewise or (7-DoF) l *= fx, w *= fy.  The caller computes
 *   fx = (ext[3]-ext[0])/(s_info[1]-s_info[0]), fy = (ext[4]-ext[1])/(s_info[3]-s_info[2]), x0 = ext[0], y0 = ext[1]
 *   in fp32 with s_info = info[0:6]/info[6].
 *
 * b2d_eval_match: the confidence-ordered greedy matching loop of datasets/waymo_eval.py:120-215 (kitti_eval.py and
 *   cadc_eval.py share it).  Detections of one class in DESCENDING confidence order, float64 as the reference
 *   parses them from the result file; frames ("recs") are independent:
 *     det_boxes [n_det, box_elem]; rec_det_offset [n_rec+1] + rec_det_index: for every frame the positions of its
 *     detections in that order (ascending); gt_offset [n_rec+1], gt_boxes [sum G, box_elem], gt_flags [sum G]
 *     (bit 0: ignore, bits 8..: difficulty); dc_offset [n_rec+1], dc_boxes [sum D, box_elem] (don't-care boxes,
 *     used when ignore_dc != 0, cfg.TEST.IGNORE_DC).  mode 0 = '2d' boxes [x1,y1,x2,y2] (+1 pixel convention),
 *     mode 1 = 'bev_aa' axis-aligned footprint of [xc,yc,zc,l,w,h,ry].
 *   Outputs per detection: code (0 nothing recorded, 1 true positive, 2 false positive on an already matched box,
 *   3 false positive below the overlap threshold; detections of frames that are not evaluated are simply not
 *   listed and keep whatever the caller initialised), ovmax (fp64, -inf without ground truth), jmax (first argmax),
 *   difficulty of the matched box (-1 unless code 1 or 2).  hit_scratch: sum G bytes.
 * ---------------------------------------------------------------------------------- */
int b2d_bbox_voxel_grid_to_pc(int n, int row_stride, float fx, float fy, float x0, float y0, int aabb, float* boxes,
                              void* stream);
int b2d_eval_match(int n_det, int n_rec, int box_elem, int mode, const double* det_boxes,
                   const int32_t* rec_det_offset, const int32_t* rec_det_index, const int32_t* gt_offset,
                   const double* gt_boxes, const int32_t* gt_flags, const int32_t* dc_offset, const double* dc_boxes,
                   double ovthresh, double ovthresh_dc, int ignore_dc, int32_t* code, double* ovmax, int32_t* jmax,
                   int32_t* difficulty, unsigned char* hit_scratch, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B2D_GLUE_H_ */
