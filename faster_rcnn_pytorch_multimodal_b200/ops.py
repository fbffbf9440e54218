"""Tensor-level operators over the C ABI: the replacements for the two torchvision calls on
the reference's hot path (``torchvision.ops.nms`` at layer_utils/proposal_layer.py:46 and
utils/filter_predictions.py:67,69; ``torchvision.ops.roi_align`` at utils/torchpoolers.py:165,194)
plus the batched proposal stage.  Everything here is stream-ordered on torch's current stream.
"""
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import check, f32c, lib, ptr, require_cuda, stream_ptr, workspaces


# ------------------------------------------------------------------------------------------
# proposal stage (batched over frames; no host sync)
# ------------------------------------------------------------------------------------------
def proposal_batched(cls_prob: torch.Tensor, bbox_pred: torch.Tensor, info: torch.Tensor, anchors: torch.Tensor,
                     anchors_3d: Optional[torch.Tensor], num_anchors: int, pre_nms: int, post_nms: int,
                     nms_thresh: float, batch_index_stride: int = 1, want_anchor_index: bool = False):
    """cls_prob [F,H,W,2A], bbox_pred [F,H,W,4A], info [F,7] -> padded per-frame proposals.

    Returns (rois [F,M,5], scores [F,M], a3d [F,M,7] | None, anchor_index [F,M] | None, num_out [F] int32),
    all on the device; rows >= num_out[f] are zero.  M = post_nms (or min(pre_nms, N) if post_nms <= 0).
    """
    require_cuda(cls_prob, bbox_pred, info, anchors, anchors_3d)
    dev = cls_prob.device
    F = cls_prob.shape[0]
    A = int(num_anchors)
    n_loc = cls_prob[0].numel() // (2 * A)
    N = n_loc * A
    if bbox_pred.numel() != F * N * 4 or anchors.numel() != N * 4:
        raise _lib.B2DError("proposal: cls_prob / bbox_pred / anchors shapes disagree")
    cls_prob, bbox_pred, info, anchors = f32c(cls_prob), f32c(bbox_pred), f32c(info.reshape(F, -1)), f32c(anchors)
    if info.shape[1] < 4:
        raise _lib.B2DError("info needs at least [x_min, x_max, y_min, y_max]")
    if info.shape[1] != 7:
        pad = torch.zeros(F, 7, device=dev)
        pad[:, :info.shape[1]] = info
        info = pad
    k = pre_nms if 0 < pre_nms < N else N
    M = post_nms if 0 < post_nms < k else k
    rois = torch.empty(F, M, 5, device=dev)
    scores = torch.empty(F, M, device=dev)
    a3d_out = None
    if anchors_3d is not None:
        anchors_3d = f32c(anchors_3d)
        a3d_out = torch.empty(F, M, 7, device=dev)
    aidx = torch.empty(F, M, dtype=torch.int32, device=dev) if want_anchor_index else None
    num_out = torch.empty(F, dtype=torch.int32, device=dev)
    L = lib(dev)
    nbytes = L.b2d_proposal_workspace_bytes(F, n_loc, A, pre_nms, post_nms)
    ws = workspaces.get(dev, "proposal", nbytes)
    check(L.b2d_proposal(F, n_loc, A, ptr(cls_prob), ptr(bbox_pred), ptr(info), ptr(anchors), ptr(anchors_3d),
                         int(pre_nms), int(post_nms), float(nms_thresh), int(batch_index_stride), ptr(rois),
                         ptr(scores), ptr(a3d_out), ptr(aidx), ptr(num_out), ptr(ws), ws.numel(), stream_ptr(dev)),
          "b2d_proposal")
    return rois, scores, a3d_out, aidx, num_out


def proposal_sorted_debug(F, n_loc, A, pre_nms, post_nms, device):
    """Pre-NMS sorted list left in the workspace by the last proposal_batched() call."""
    N = n_loc * A
    k = pre_nms if 0 < pre_nms < N else N
    boxes = torch.empty(F, k, 4, device=device)
    scores = torch.empty(F, k, device=device)
    index = torch.empty(F, k, dtype=torch.int32, device=device)
    L = lib(device)
    ws = workspaces.get(device, "proposal", L.b2d_proposal_workspace_bytes(F, n_loc, A, pre_nms, post_nms))
    check(L.b2d_proposal_debug_sorted(F, n_loc, A, pre_nms, post_nms, ptr(ws), ptr(boxes), ptr(scores), ptr(index),
                                      stream_ptr(device)), "b2d_proposal_debug_sorted")
    return boxes, scores, index


def proposal_top_batched(cls_prob, bbox_pred, info, anchors, num_anchors, top_n, batch_index_stride=1):
    require_cuda(cls_prob, bbox_pred, info, anchors)
    dev = cls_prob.device
    F = cls_prob.shape[0]
    A = int(num_anchors)
    n_loc = cls_prob[0].numel() // (2 * A)
    cls_prob, bbox_pred, anchors = f32c(cls_prob), f32c(bbox_pred), f32c(anchors)
    info7 = torch.zeros(F, 7, device=dev)
    info = f32c(info.reshape(F, -1))
    info7[:, :info.shape[1]] = info
    rois = torch.empty(F, top_n, 5, device=dev)
    scores = torch.empty(F, top_n, device=dev)
    anc = torch.empty(F, top_n, 4, device=dev)
    L = lib(dev)
    nbytes = L.b2d_proposal_workspace_bytes(F, n_loc, A, top_n, top_n)
    ws = workspaces.get(dev, "proposal", nbytes)
    check(L.b2d_proposal_top(F, n_loc, A, ptr(cls_prob), ptr(bbox_pred), ptr(info7), ptr(anchors), int(top_n),
                             int(batch_index_stride), ptr(rois), ptr(scores), ptr(anc), ptr(ws), ws.numel(),
                             stream_ptr(dev)), "b2d_proposal_top")
    return rois, scores, anc


# ------------------------------------------------------------------------------------------
# NMS (torchvision.ops.nms drop-in)
# ------------------------------------------------------------------------------------------
def nms_sorted(boxes: torch.Tensor, thresh: float, max_keep: int = -1,
               n_valid: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """boxes [F,n,4] already in descending score order -> (keep [F,max_keep] int32, num_keep [F] int32)."""
    require_cuda(boxes, n_valid)
    boxes, n_valid = f32c(boxes), _lib.i32c(n_valid, "n_valid")
    F, n = boxes.shape[0], boxes.shape[1]
    mk = n if max_keep <= 0 else min(max_keep, n)
    keep = torch.empty(F, max(mk, 1), dtype=torch.int32, device=boxes.device)
    num = torch.zeros(F, dtype=torch.int32, device=boxes.device)
    if n > 0:
        L = lib(boxes.device)
        nbytes = L.b2d_nms_workspace_bytes(F, mk)
        ws = workspaces.get(boxes.device, "nms", nbytes) if nbytes else None
        check(L.b2d_nms_sorted(F, n, ptr(boxes), ptr(n_valid), float(thresh), mk, ptr(keep), ptr(num), ptr(ws),
                               int(nbytes), stream_ptr(boxes.device)), "b2d_nms_sorted")
    return keep, num


def argsort_desc(scores: torch.Tensor) -> torch.Tensor:
    """Stable descending argsort of [F,n] scores (ties: lower index first), int32."""
    require_cuda(scores)
    scores = f32c(scores)
    F, n = scores.shape
    order = torch.empty(F, n, dtype=torch.int32, device=scores.device)
    if n > 0:
        L = lib(scores.device)
        nbytes = L.b2d_argsort_workspace_bytes(F, n)
        ws = workspaces.get(scores.device, "argsort", nbytes) if nbytes else None
        check(L.b2d_argsort_desc(F, n, ptr(scores), ptr(order), ptr(ws), int(nbytes), stream_ptr(scores.device)),
              "b2d_argsort_desc")
    return order


def nms(boxes: torch.Tensor, scores: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    """Drop-in for ``torchvision.ops.nms``: int64 indices of kept boxes, by decreasing score.

    One host sync (the number of kept boxes), as in torchvision's own CUDA path.
    """
    require_cuda(boxes, scores)
    n = boxes.shape[0]
    if n == 0:
        return torch.empty(0, dtype=torch.int64, device=boxes.device)
    order = argsort_desc(scores.reshape(1, n))[0].long()
    keep, num = nms_sorted(f32c(boxes)[order].unsqueeze(0), iou_threshold)
    return order[keep[0, :int(num.item())].long()]


# ------------------------------------------------------------------------------------------
# RoIAlign (torchvision.ops.roi_align drop-in, with autograd)
# ------------------------------------------------------------------------------------------
# kernel family of RoIAlign forward (include/b2d_glue.h b2d_roi_route); parity tests pin it, production leaves AUTO
ROI_ROUTES = {"auto": 0, "rows": 1, "rows_coop": 2, "planes": 3, "gather": 4}
ROI_ROUTE = "auto"


def _roi_align_forward(feat, rois, out_hw, scale, sampling_ratio, aligned, roi_ids=None, seg_count=None,
                       seg_stride=0, out=None, route=None):
    roi_ids, seg_count = _lib.i32c(roi_ids, "roi_ids"), _lib.i32c(seg_count, "seg_count")
    Fr, Cc, H, W = feat.shape
    R = rois.shape[0]
    if out is None:
        out = torch.empty(R, Cc, out_hw[0], out_hw[1], device=feat.device)
        if roi_ids is not None:
            out.zero_()
    if R == 0 or (roi_ids is not None and roi_ids.numel() == 0):
        return out
    n_ids = 0 if roi_ids is None else roi_ids.numel()
    L = lib(feat.device)
    n_list = n_ids if roi_ids is not None else R
    per_frame = int(seg_stride) if seg_count is not None else n_list
    ws = workspaces.get(feat.device, "roi_align", L.b2d_roi_align_workspace_bytes(Fr, Cc, H, W, n_list, per_frame))
    check(L.b2d_roi_align_forward_route(Fr, Cc, H, W, ptr(feat), ptr(rois), R, ptr(roi_ids), n_ids, ptr(seg_count),
                                        int(seg_stride), out_hw[0], out_hw[1], float(scale), int(sampling_ratio),
                                        int(bool(aligned)), ROI_ROUTES[route or ROI_ROUTE], ptr(out), ptr(ws),
                                        ws.numel(), stream_ptr(feat.device)), "b2d_roi_align_forward")
    return out


def _roi_align_backward(grad_out, rois, feat_shape, out_hw, scale, sampling_ratio, aligned, roi_ids=None,
                        seg_count=None, seg_stride=0, grad_feat=None):
    roi_ids, seg_count = _lib.i32c(roi_ids, "roi_ids"), _lib.i32c(seg_count, "seg_count")
    Fr, Cc, H, W = feat_shape
    accumulate = grad_feat is not None
    if grad_feat is None:
        grad_feat = torch.empty(feat_shape, device=grad_out.device)
    if rois.shape[0] == 0 or (roi_ids is not None and roi_ids.numel() == 0):
        return grad_feat if accumulate else grad_feat.zero_()
    n_ids = 0 if roi_ids is None else roi_ids.numel()
    L = lib(grad_out.device)
    n_list = n_ids if roi_ids is not None else rois.shape[0]
    per_frame = int(seg_stride) if seg_count is not None else n_list
    ws = workspaces.get(grad_out.device, "roi_align", L.b2d_roi_align_workspace_bytes(Fr, Cc, H, W, n_list, per_frame))
    check(L.b2d_roi_align_backward(Fr, Cc, H, W, ptr(grad_out), ptr(rois), rois.shape[0], ptr(roi_ids), n_ids,
                                   ptr(seg_count), int(seg_stride), out_hw[0], out_hw[1], float(scale),
                                   int(sampling_ratio), int(bool(aligned)), int(accumulate), ptr(grad_feat),
                                   ptr(ws), ws.numel(), stream_ptr(grad_out.device)), "b2d_roi_align_backward")
    return grad_feat


class _RoIAlignFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, rois, out_h, out_w, scale, sampling_ratio, aligned, seg_count, seg_stride):
        require_cuda(feat, rois, seg_count)
        feat_c, rois_c = f32c(feat), f32c(rois)
        ctx.save_for_backward(rois_c, seg_count)
        ctx.meta = (tuple(feat.shape), (out_h, out_w), scale, sampling_ratio, aligned, seg_stride)
        return _roi_align_forward(feat_c, rois_c, (out_h, out_w), scale, sampling_ratio, aligned,
                                  seg_count=seg_count, seg_stride=seg_stride)

    @staticmethod
    def backward(ctx, grad_out):
        rois, seg_count = ctx.saved_tensors
        shape, out_hw, scale, sr, aligned, seg_stride = ctx.meta
        g = _roi_align_backward(f32c(grad_out), rois, shape, out_hw, scale, sr, aligned, seg_count=seg_count,
                                seg_stride=seg_stride)
        return g, None, None, None, None, None, None, None, None


def roi_align(input: torch.Tensor, boxes, output_size, spatial_scale: float = 1.0, sampling_ratio: int = -1,
              aligned: bool = False, seg_count: Optional[torch.Tensor] = None, seg_stride: int = 0) -> torch.Tensor:
    """Drop-in for ``torchvision.ops.roi_align`` (NCHW fp32).  ``boxes`` is a Tensor[K,5] or a
    list of Tensor[L,4] (one per frame).  ``seg_count``/``seg_stride`` select the padded per-frame
    layout that ``proposal_batched`` emits."""
    if isinstance(output_size, int):
        output_size = (output_size, output_size)
    if isinstance(boxes, (list, tuple)):
        boxes = torch.cat([torch.cat((b.new_full((b.shape[0], 1), float(i)), b), dim=1)
                           for i, b in enumerate(boxes)], dim=0)
    out = _RoIAlignFn.apply(input, boxes, int(output_size[0]), int(output_size[1]), float(spatial_scale),
                            int(sampling_ratio), bool(aligned), seg_count, int(seg_stride))
    # the kernels compute in fp32; like torchvision the result comes back in the input's dtype (half / AMP features)
    return out if out.dtype == input.dtype else out.to(input.dtype)


class RoIAlign(torch.nn.Module):
    """Module form (``torchvision.ops.RoIAlign``; imported at nets/imagenet.py:15, lidarnet.py:16)."""

    def __init__(self, output_size, spatial_scale: float, sampling_ratio: int, aligned: bool = False):
        super().__init__()
        self.output_size = output_size
        self.spatial_scale = spatial_scale
        self.sampling_ratio = sampling_ratio
        self.aligned = aligned

    def forward(self, input, rois):
        return roi_align(input, rois, self.output_size, self.spatial_scale, self.sampling_ratio, self.aligned)


def roi_align_forward_levels(feats, scales, rois, levels_i32, out_hw, sampling_ratio, aligned=False):
    """One-launch multi-level RoIAlign forward (small calls): feats = per-level [F,C,H,W] maps, levels_i32 [R]."""
    import ctypes as C
    require_cuda(rois, levels_i32, *feats)
    feats = [f32c(f) for f in feats]
    rois, levels_i32 = f32c(rois), _lib.i32c(levels_i32, "levels")
    n = len(feats)
    Fr, Cc = feats[0].shape[:2]
    R = rois.shape[0]
    out = torch.empty(R, Cc, out_hw[0], out_hw[1], device=rois.device)
    fp = (C.c_void_p * n)(*[f.data_ptr() for f in feats])
    hh = (C.c_int32 * n)(*[f.shape[2] for f in feats])
    ww = (C.c_int32 * n)(*[f.shape[3] for f in feats])
    sc = (C.c_float * n)(*[float(s) for s in scales])
    check(lib(rois.device).b2d_roi_align_forward_levels(n, Fr, Cc, fp, hh, ww, sc, ptr(rois), ptr(levels_i32), R, out_hw[0], out_hw[1],
                                             int(sampling_ratio), int(bool(aligned)), ptr(out), stream_ptr(rois.device)),
          "b2d_roi_align_forward_levels")
    return out


def fpn_level_map(boxes: torch.Tensor, k_min: int, k_max: int, canonical_scale: float = 224,
                  canonical_level: int = 4, eps: float = 1e-6, as_int32: bool = False) -> torch.Tensor:
    """LevelMapper.__call__ (utils/torchpoolers.py:39-51) on one concatenated [R,4] tensor -> int64 [R]."""
    require_cuda(boxes)
    boxes = f32c(boxes)
    out = torch.empty(boxes.shape[0], dtype=torch.int32, device=boxes.device)
    check(lib(boxes.device).b2d_fpn_level_map(boxes.shape[0], ptr(boxes), int(k_min), int(k_max), float(canonical_scale),
                                  int(canonical_level), float(eps), ptr(out), stream_ptr(boxes.device)),
          "b2d_fpn_level_map")
    return out if as_int32 else out.long()


# ------------------------------------------------------------------------------------------
# Final per-class detection filter (utils/filter_predictions.py:45-130, model/test.py:213-221)
# ------------------------------------------------------------------------------------------
def final_detections(cls_score, pred_boxes, info, num_elem, db_type, score_thresh, nms_thresh, max_dets=0,
                     num_rois=None, uc_row=None, uc_cls=None, max_out=None):
    """Batched on-device replacement of the per-class threshold + NMS + gather loop.

    cls_score [F,R,K], pred_boxes [F,R,K*E] (decoded), info [F,7]; db_type 'image' (boxes clamped to the
    frame) | 'lidar' | 'image_noclamp'.
    uc_row [F,R,U] per-roi and uc_cls [F,R,U2,K*E] per-class-box uncertainty columns are optional.
    -> dets [F,K,max_out,E+1] (box, score; descending score; zero padded), det_roi [F,K,max_out] int32
       (source roi, -1 padded), counts [F,K] int32, out_uc_row [F,K,max_out,U] | None,
       out_uc_cls [F,K,max_out,U2*E] | None.  One launch, no host sync.
    """
    require_cuda(cls_score, pred_boxes, info, num_rois, uc_row, uc_cls)
    modes = {"image": 0, "lidar": 1, "image_noclamp": 2}
    if db_type not in modes:
        raise _lib.B2DError(f"final_detections: db_type must be one of {sorted(modes)}, got {db_type!r}")
    sc, pb, inf = f32c(cls_score), f32c(pred_boxes), f32c(info)
    F, R, K = sc.shape
    E = int(num_elem)
    if pb.shape != (F, R, K * E) or inf.shape != (F, 7):
        raise _lib.B2DError("final_detections: shape mismatch")
    mo = int(max_out) if max_out else max(R, 1)
    dev = sc.device
    dets = torch.empty(F, K, mo, E + 1, device=dev)
    det_roi = torch.empty(F, K, mo, dtype=torch.int32, device=dev)
    counts = torch.empty(F, K, dtype=torch.int32, device=dev)
    if K > 0:
        dets[:, 0].zero_()
        det_roi[:, 0].fill_(-1)
    ur = f32c(uc_row) if uc_row is not None else None
    ucl = f32c(uc_cls) if uc_cls is not None else None
    n_ur = ur.shape[2] if ur is not None else 0
    n_uc = ucl.shape[2] if ucl is not None else 0
    o_ur = torch.zeros(F, K, mo, n_ur, device=dev) if n_ur else None
    o_uc = torch.zeros(F, K, mo, n_uc * E, device=dev) if n_uc else None
    nr = num_rois.to(torch.int32).contiguous() if num_rois is not None else None
    check(lib(dev).b2d_final_detections(F, R, K, E, ptr(sc), ptr(pb), ptr(nr), ptr(inf), modes[db_type],
                                     float(score_thresh), float(nms_thresh), int(max_dets), mo, ptr(ur), n_ur,
                                     ptr(ucl), n_uc, ptr(dets), ptr(det_roi), ptr(o_ur), ptr(o_uc), ptr(counts),
                                     stream_ptr(dev)), "b2d_final_detections")
    return dets, det_roi, counts, o_ur, o_uc


# ------------------------------------------------------------------------------------------
# Tail of the detection head over the MC-dropout stack, one launch for all frames
# ------------------------------------------------------------------------------------------
def head_tail_decode(bbox_pred, cls_score, rois, anchors_3d, info, net_type, a_bbox_var=None, use_scale=False,
                     clip=True, means=None, stds=None):
    """bbox_pred [T,F,R,K*E] (normalised deltas of T stacked head passes), cls_score [T,F,R,K], rois [F,R,5],
    anchors_3d [F,R,7] (lidar), info [F,7] -> dict, every tensor in ``final_detections``' input layout:

      boxes [F,R,K*E]        de-normalised (model/config.py:219-223) mean prediction, decoded with
                             lidar_3d_bbox_transform_inv / bbox_transform_inv + clip_boxes
      probs [F,R,K]          mean softmax
      e_bbox_var [F,R,K*E]   compute_bbox_var of the de-normalised samples through lidar_3d_uncertainty_transform_inv
      a_bbox_var [F,R,K*E]   the same transform of the aleatoric head output (only when ``a_bbox_var`` is given)
      e_entropy, e_mutual_info [F,R]   utils/loss_utils.py:122-141

    T = 1 (no MC-dropout) gives the plain decode with zero epistemic variance.  ``means`` / ``stds`` default to
    ``cfg.TRAIN.{LIDAR,IMAGE}.BBOX_NORMALIZE_{MEANS,STDS}``."""
    import ctypes as C
    from .model.config import cfg
    require_cuda(bbox_pred, cls_score, rois, anchors_3d, info, a_bbox_var)
    lidar = net_type == "lidar"
    E = 7 if lidar else 4
    bp, cs, ro, inf = f32c(bbox_pred), f32c(cls_score), f32c(rois), f32c(info)
    if bp.dim() == 3:                      # [F,R,K*E]: a single head pass
        bp, cs = bp.unsqueeze(0), cs.unsqueeze(0)
    T, F, R, KE = bp.shape
    K = cs.shape[3]
    if KE != K * E or cs.shape[:3] != (T, F, R) or ro.shape != (F, R, 5) or inf.shape != (F, 7):
        raise _lib.B2DError("head_tail_decode: shape mismatch")
    if lidar and (anchors_3d is None or tuple(anchors_3d.shape) != (F, R, 7)):
        raise _lib.B2DError("head_tail_decode: lidar needs anchors_3d [F,R,7]")
    a3 = f32c(anchors_3d) if anchors_3d is not None else None
    av = f32c(a_bbox_var) if a_bbox_var is not None else None
    tr = cfg.TRAIN.LIDAR if lidar else cfg.TRAIN.IMAGE
    means = list(tr.BBOX_NORMALIZE_MEANS if means is None else means)
    stds = list(tr.BBOX_NORMALIZE_STDS if stds is None else stds)
    if len(means) != E or len(stds) != E:
        raise _lib.B2DError("head_tail_decode: means / stds need one value per box element")
    dev = bp.device
    out = {"boxes": torch.empty(F, R, KE, device=dev), "probs": torch.empty(F, R, K, device=dev),
           "e_bbox_var": torch.empty(F, R, KE, device=dev), "e_entropy": torch.empty(F, R, device=dev),
           "e_mutual_info": torch.empty(F, R, device=dev)}
    if av is not None:
        out["a_bbox_var"] = torch.empty(F, R, KE, device=dev)
    check(lib(dev).b2d_head_tail_decode(F, T, R, K, E, ptr(bp), ptr(cs), ptr(ro), ptr(a3), ptr(inf), ptr(av),
                                        (C.c_float * E)(*means), (C.c_float * E)(*stds), 1 if lidar else 0,
                                        int(bool(use_scale)), int(bool(clip)), ptr(out["boxes"]), ptr(out["probs"]),
                                        ptr(out["e_bbox_var"]), ptr(out.get("a_bbox_var")), ptr(out["e_entropy"]),
                                        ptr(out["e_mutual_info"]), stream_ptr(dev)), "b2d_head_tail_decode")
    return out
