"""proposal_target_layer (lib/layer_utils/proposal_target_layer.py:22-262).

Phase 1 (device): RoI x GT IoU, max/argmax, ordered fg/bg lists.  Host: the reference's sampling
branches with the same torch RNG calls (torch_choice, :265-284).  Phase 2 (device): gather,
regression targets (image or lidar codec), normalisation, per-class expansion, weights.
"""
import math

import torch

from .._lib import B2DError, check, f32c, lib, ptr, require_cuda, stream_ptr
from ..model.config import cfg
from ..utils.bbox import bbox_overlaps


# Device whose torch generator feeds the sampler.  None = the tensors' device (the reference's
# behaviour, proposal_target_layer.py:210,217).  Parity tests set 'cpu' to share the CPU oracle's draws.
RNG_DEVICE = None


def torch_choice(max_idx, num_elem, dev, to_replace=False):
    if to_replace:
        return torch.randint(max_idx, (num_elem,), device=dev)                   # :275-277
    if num_elem > max_idx:                                                       # :279-285
        factor = math.ceil(num_elem / max_idx)
        idx = torch.arange(max_idx).repeat(factor)
        perm = torch.randperm(idx.shape[0], device=dev)
        return idx.to(dev)[perm][:num_elem]
    return torch.randperm(max_idx, device=dev)[:num_elem]


def proposal_target_layer(rpn_rois, rpn_scores, anchors_3d, gt_boxes, true_gt_boxes, gt_boxes_dc, _num_classes,
                          num_bbox_elem):
    """-> labels [S,1], rois [S,5], anchors_3d [S,7], roi_scores [S], bbox_targets / inside_w / outside_w [S,K*E]."""
    require_cuda(rpn_rois, rpn_scores, gt_boxes)
    dev = rpn_rois.device
    all_rois, all_scores, all_a3d = rpn_rois, rpn_scores, anchors_3d
    if cfg.TRAIN.USE_GT:                                                          # :35-41
        zeros = rpn_rois.new_zeros(gt_boxes.shape[0], 1)
        all_rois = torch.cat((all_rois, torch.cat((zeros, gt_boxes[:, :-1]), 1)), 0)
        all_scores = torch.cat((all_scores, zeros), 0)
        all_a3d = torch.cat((all_a3d, true_gt_boxes[:, :-1]), 0)
    rois_per_frame = cfg.TRAIN.ROI_BATCH_SIZE / 1
    fg_rois_per_frame = int(round(cfg.TRAIN.FG_FRACTION * rois_per_frame))
    if cfg.TRAIN.IGNORE_DC and list(gt_boxes_dc.size())[0] > 0:                   # :184-190
        mx_dc, _ = bbox_overlaps(all_rois[:, 1:5].contiguous(), gt_boxes_dc[:, :4].contiguous()).max(1)
        dc_inds = (mx_dc < cfg.TRAIN.DC_THRESH).nonzero().view(-1)
        all_rois, all_scores, all_a3d = all_rois[dc_inds, :], all_scores[dc_inds, :], all_a3d[dc_inds, :]
    rois = f32c(all_rois)
    scores = f32c(all_scores).view(-1)
    gt = f32c(gt_boxes)
    R, G = rois.shape[0], gt.shape[0]
    if G == 0:
        raise B2DError("proposal_target_layer: no ground-truth boxes")
    E = int(num_bbox_elem)
    lidar = cfg.NET_TYPE == 'lidar'
    if lidar and E != 7 or (not lidar and E != 4):
        raise B2DError("num_bbox_elem must be 7 for NET_TYPE 'lidar' and 4 for 'image'")
    max_ov = torch.empty(R, device=dev)
    assign = torch.empty(R, dtype=torch.int32, device=dev)
    fg_list = torch.empty(max(R, 1), dtype=torch.int32, device=dev)
    bg_list = torch.empty(max(R, 1), dtype=torch.int32, device=dev)
    counts = torch.empty(2, dtype=torch.int32, device=dev)
    L, st = lib(dev), stream_ptr(dev)
    bg_mode = 0 if cfg.TRAIN.get('BG_MODE', 'strict') == 'strict' else 1
    check(L.b2d_proposal_target_phase1(R, G, ptr(rois), ptr(gt), float(cfg.TRAIN.FG_THRESH),
                                       float(cfg.TRAIN.BG_THRESH_HI), float(cfg.TRAIN.BG_THRESH_LO), bg_mode,
                                       ptr(max_ov), ptr(assign), ptr(fg_list), ptr(bg_list), ptr(counts), st),
          "b2d_proposal_target_phase1")
    n_fg, n_bg = counts.tolist()                                                  # the reference syncs on numel()
    fg_inds, bg_inds = fg_list[:n_fg].long(), bg_list[:n_bg].long()
    tdev = dev
    if RNG_DEVICE is not None:
        dev = torch.device(RNG_DEVICE)
        fg_inds, bg_inds = fg_inds.to(dev), bg_inds.to(dev)
    if n_fg > 0 and n_bg > 0:                                                     # :206-217
        fg_rois_per_frame = min(fg_rois_per_frame, n_fg)
        fg_inds = fg_inds[torch_choice(n_fg, int(fg_rois_per_frame), dev, to_replace=False)]
        bg_rois_per_frame = rois_per_frame - fg_rois_per_frame
        bg_inds = bg_inds[torch_choice(n_bg, int(bg_rois_per_frame), dev, to_replace=n_bg < bg_rois_per_frame)]
    elif n_fg > 0:                                                                # :218-224
        fg_inds = fg_inds[torch_choice(n_fg, int(rois_per_frame), dev, to_replace=n_fg < rois_per_frame)]
        fg_rois_per_frame = rois_per_frame
    elif n_bg > 0:                                                                # :225-231
        bg_inds = bg_inds[torch_choice(n_bg, int(rois_per_frame), dev, to_replace=n_bg < rois_per_frame)]
        fg_rois_per_frame = 0
    else:
        raise B2DError("proposal_target_layer: no foreground and no background RoIs "
                       "(the reference drops into pdb here, proposal_target_layer.py:232-235)")
    dev = tdev
    keep = torch.cat([fg_inds, bg_inds], 0).to(dev).contiguous()
    S = keep.numel()
    K = int(_num_classes)
    labels = torch.empty(S, 1, device=dev)
    out_rois = torch.empty(S, 5, device=dev)
    out_a3d = torch.empty(S, 7, device=dev) if all_a3d is not None else None
    out_scores = torch.empty(S, device=dev)
    targets = torch.empty(S, K * E, device=dev)
    inside_w = torch.empty_like(targets)
    outside_w = torch.empty_like(targets)
    norm = bool(cfg.TRAIN.BBOX_NORMALIZE_TARGETS_PRECOMPUTED)
    stat = cfg.TRAIN.LIDAR if lidar else cfg.TRAIN.IMAGE
    means = torch.tensor(stat.BBOX_NORMALIZE_MEANS, dtype=torch.float32, device=dev)
    stds = torch.tensor(stat.BBOX_NORMALIZE_STDS, dtype=torch.float32, device=dev)
    a3d_c = f32c(all_a3d) if all_a3d is not None else None
    gt8 = f32c(true_gt_boxes) if (lidar and true_gt_boxes is not None) else None
    check(L.b2d_proposal_target_phase2(S, int(fg_rois_per_frame), ptr(keep), ptr(rois), ptr(scores), ptr(a3d_c),
                                       ptr(gt), ptr(gt8), ptr(assign), K, E, int(norm), ptr(means), ptr(stds),
                                       ptr(labels), ptr(out_rois), ptr(out_a3d), ptr(out_scores), ptr(targets),
                                       ptr(inside_w), ptr(outside_w), st), "b2d_proposal_target_phase2")
    return labels, out_rois, out_a3d, out_scores, targets, inside_w, outside_w
