"""Train-mode targets of SEVERAL frames at once: ``anchor_target_layer_torch`` + ``proposal_target_layer``
(lib/layer_utils/anchor_target_layer.py:22-165, lib/layer_utils/proposal_target_layer.py:22-262) for a batch.

The reference runs one frame per call and stops the device three times per frame (``len(fg_inds)``,
``len(bg_inds)``, the sampler's ``numel()``).  Here the device phases of all frames are issued back to back, the
host reads every count in ONE transfer, makes the ``randperm`` / ``randint`` draws in exactly the order a
frame-by-frame run makes them (frame 0: anchor fg, anchor bg, RoI fg, RoI bg; frame 1: ...) and the second device
phases follow - so a seeded run returns, frame for frame, what the per-frame functions return.
"""
from typing import List, Optional, Sequence

import numpy as np
import torch

from .._lib import B2DError, check, f32c, lib, ptr, require_cuda, stream_ptr, workspaces
from ..model.config import cfg
from . import proposal_target_layer as ptl


def _pad_gt(gts: Sequence[torch.Tensor], device):
    F = len(gts)
    G = max(int(g.shape[0]) for g in gts)
    if G == 0 or min(int(g.shape[0]) for g in gts) == 0:
        raise B2DError("batched targets: every frame needs at least one ground-truth box")
    pad = torch.zeros(F, G, 5, device=device)
    for f, g in enumerate(gts):
        pad[f, :g.shape[0], :min(5, g.shape[1])] = f32c(g)[:, :5]
    num = torch.tensor([int(g.shape[0]) for g in gts], dtype=torch.int32, device=device)
    return pad, num, G


def train_targets_batched(gt_boxes: Sequence[torch.Tensor], info: torch.Tensor, all_anchors: torch.Tensor,
                          num_anchors: int, height: int, width: int, rois: torch.Tensor, roi_scores: torch.Tensor,
                          anchors_3d: Optional[torch.Tensor], num_rois: torch.Tensor,
                          true_gt_boxes: Optional[Sequence[torch.Tensor]], num_classes: int, num_bbox_elem: int,
                          dev=None):
    """gt_boxes: F tensors [G_f,5]; info [F,7]; rois [F,M,5], roi_scores [F,M], anchors_3d [F,M,7] | None and
    num_rois [F] as ``ops.proposal_batched`` returns them; true_gt_boxes: F tensors [G_f,8] (lidar) or None.

    -> (rpn_labels [F,A,H,W], rpn_bbox_targets / inside_w / outside_w [F,H,W,4A],
        roi_outputs: list over frames of the 7-tuple ``proposal_target_layer`` returns).
    ``dev`` is the device whose torch generator feeds the draws (None: the tensors' device, as the reference)."""
    require_cuda(all_anchors, rois, roi_scores, num_rois, info)
    device = all_anchors.device
    rng = torch.device(dev) if dev is not None else device
    F = len(gt_boxes)
    A, H, W = int(num_anchors), int(height), int(width)
    anchors = f32c(all_anchors)
    N = anchors.shape[0]
    gt, num_gt, G = _pad_gt(gt_boxes, device)
    info7 = torch.zeros(F, 7, device=device)
    info_c = f32c(info.reshape(F, -1))
    info7[:, :info_c.shape[1]] = info_c
    L, st = lib(device), stream_ptr(device)
    E, K = int(num_bbox_elem), int(num_classes)
    lidar = cfg.NET_TYPE == 'lidar'
    if lidar and E != 7 or (not lidar and E != 4):
        raise B2DError("num_bbox_elem must be 7 for NET_TYPE 'lidar' and 4 for 'image'")
    if cfg.TRAIN.USE_GT or cfg.TRAIN.IGNORE_DC:
        raise B2DError("train_targets_batched: cfg.TRAIN.USE_GT / IGNORE_DC change the candidate set per frame; "
                       "use the per-frame proposal_target_layer for those configurations")

    # ---- phase 1 of both layers, every frame, no host sync in between
    ws = workspaces.get(device, "anchor_target", L.b2d_anchor_target_workspace_bytes(F, N, G))
    at_counts = torch.empty(F, 4, dtype=torch.int32, device=device)
    check(L.b2d_anchor_target_phase1(F, N, G, ptr(anchors), ptr(gt), ptr(num_gt), ptr(info7),
                                     float(cfg.TRAIN.RPN_NEGATIVE_OVERLAP), float(cfg.TRAIN.RPN_POSITIVE_OVERLAP),
                                     int(bool(cfg.TRAIN.RPN_CLOBBER_POSITIVES)), ptr(at_counts), ptr(ws), ws.numel(), st),
          "b2d_anchor_target_phase1")
    nr = num_rois.tolist()                                   # the reference's own sync (keep.numel(), proposal_layer.py:48)
    rois_c, scores_c = f32c(rois), f32c(roi_scores)
    a3d_c = f32c(anchors_3d) if anchors_3d is not None else None
    M = rois_c.shape[1]
    max_ov = torch.empty(F, M, device=device)
    assign = torch.empty(F, M, dtype=torch.int32, device=device)
    fg_list = torch.empty(F, max(M, 1), dtype=torch.int32, device=device)
    bg_list = torch.empty(F, max(M, 1), dtype=torch.int32, device=device)
    pt_counts = torch.empty(F, 2, dtype=torch.int32, device=device)
    bg_mode = 0 if cfg.TRAIN.get('BG_MODE', 'strict') == 'strict' else 1
    gt_rows = [f32c(g) for g in gt_boxes]
    for f in range(F):
        check(L.b2d_proposal_target_phase1(int(nr[f]), int(gt_rows[f].shape[0]), ptr(rois_c[f]), ptr(gt_rows[f]),
                                           float(cfg.TRAIN.FG_THRESH), float(cfg.TRAIN.BG_THRESH_HI),
                                           float(cfg.TRAIN.BG_THRESH_LO), bg_mode, ptr(max_ov[f]), ptr(assign[f]),
                                           ptr(fg_list[f]), ptr(bg_list[f]), ptr(pt_counts[f]), st),
              "b2d_proposal_target_phase1")
    # ---- ONE transfer with every count the samplers need
    counts = torch.cat((at_counts, pt_counts), dim=1).cpu().tolist()

    # ---- the draws, in the order a frame-by-frame run makes them
    num_fg_cap = int(cfg.TRAIN.RPN_FG_FRACTION * cfg.TRAIN.RPN_BATCHSIZE)
    rois_per_frame = cfg.TRAIN.ROI_BATCH_SIZE / 1
    fg_dis, bg_dis, keeps, fg_counts = [], [], [], []
    for f in range(F):
        _, n_fg, n_bg, _, p_fg, p_bg = counts[f]
        fd = bd = None
        if n_fg > num_fg_cap:                                             # anchor_target_layer.py:95-98
            fd = torch.randperm(n_fg, device=rng)[num_fg_cap:]
        fg_left = n_fg - (0 if fd is None else fd.numel())
        num_bg = cfg.TRAIN.RPN_BATCHSIZE - fg_left                        # :101-102
        if n_bg > num_bg:                                                 # :104-107
            bd = torch.randperm(n_bg, device=rng)[num_bg:]
        fg_dis.append(fd)
        bg_dis.append(bd)
        fg_inds, bg_inds = fg_list[f, :p_fg].long().to(rng), bg_list[f, :p_bg].long().to(rng)
        fg_per = int(round(cfg.TRAIN.FG_FRACTION * rois_per_frame))
        if p_fg > 0 and p_bg > 0:                                         # proposal_target_layer.py:206-217
            fg_per = min(fg_per, p_fg)
            fg_inds = fg_inds[ptl.torch_choice(p_fg, int(fg_per), rng, to_replace=False)]
            bg_per = rois_per_frame - fg_per
            bg_inds = bg_inds[ptl.torch_choice(p_bg, int(bg_per), rng, to_replace=p_bg < bg_per)]
        elif p_fg > 0:                                                    # :218-224
            fg_inds = fg_inds[ptl.torch_choice(p_fg, int(rois_per_frame), rng, to_replace=p_fg < rois_per_frame)]
            fg_per = rois_per_frame
        elif p_bg > 0:                                                    # :225-231
            bg_inds = bg_inds[ptl.torch_choice(p_bg, int(rois_per_frame), rng, to_replace=p_bg < rois_per_frame)]
            fg_per = 0
        else:
            raise B2DError(f"train_targets_batched: frame {f} has no foreground and no background RoIs "
                           "(the reference drops into pdb here, proposal_target_layer.py:232-235)")
        keeps.append(torch.cat([fg_inds, bg_inds], 0).to(device).contiguous())
        fg_counts.append(int(fg_per))

    # ---- phase 2: anchors (one launch for all frames), then the RoI gathers
    stride = max([1] + [0 if d is None else d.numel() for d in fg_dis + bg_dis])
    fd_t = torch.zeros(F, stride, dtype=torch.int64, device=device)
    bd_t = torch.zeros(F, stride, dtype=torch.int64, device=device)
    for f in range(F):
        if fg_dis[f] is not None:
            fd_t[f, :fg_dis[f].numel()] = fg_dis[f].to(device)
        if bg_dis[f] is not None:
            bd_t[f, :bg_dis[f].numel()] = bg_dis[f].to(device)
    nfd = torch.tensor([0 if d is None else d.numel() for d in fg_dis], dtype=torch.int32, device=device)
    nbd = torch.tensor([0 if d is None else d.numel() for d in bg_dis], dtype=torch.int32, device=device)
    labels = torch.empty(F, A, H, W, device=device)
    targets = torch.empty(F, H, W, 4 * A, device=device)
    inside_w, outside_w = torch.empty_like(targets), torch.empty_like(targets)
    iw4 = torch.tensor(np.array(cfg.TRAIN.RPN_BBOX_INSIDE_WEIGHTS, dtype=np.float32), device=device)
    pw = float(cfg.TRAIN.RPN_POSITIVE_WEIGHT)
    if pw >= 0:
        assert 0 < pw < 1                                                 # :127-128
    check(L.b2d_anchor_target_phase2(F, N, G, A, H, W, ptr(anchors), ptr(gt), ptr(at_counts), ptr(fd_t), ptr(nfd),
                                     ptr(bd_t), ptr(nbd), stride, ptr(iw4), pw, ptr(labels), ptr(targets),
                                     ptr(inside_w), ptr(outside_w), ptr(ws), ws.numel(), st), "b2d_anchor_target_phase2")
    norm = bool(cfg.TRAIN.BBOX_NORMALIZE_TARGETS_PRECOMPUTED)
    stat = cfg.TRAIN.LIDAR if lidar else cfg.TRAIN.IMAGE
    means = torch.tensor(stat.BBOX_NORMALIZE_MEANS, dtype=torch.float32, device=device)
    stds = torch.tensor(stat.BBOX_NORMALIZE_STDS, dtype=torch.float32, device=device)
    roi_out: List[tuple] = []
    for f in range(F):
        keep = keeps[f]
        S = keep.numel()
        o_lab = torch.empty(S, 1, device=device)
        o_rois = torch.empty(S, 5, device=device)
        o_a3d = torch.empty(S, 7, device=device) if a3d_c is not None else None
        o_sc = torch.empty(S, device=device)
        o_t = torch.empty(S, K * E, device=device)
        o_iw, o_ow = torch.empty_like(o_t), torch.empty_like(o_t)
        gt8 = f32c(true_gt_boxes[f]) if (lidar and true_gt_boxes is not None) else None
        check(L.b2d_proposal_target_phase2(S, fg_counts[f], ptr(keep), ptr(rois_c[f]), ptr(scores_c[f]),
                                           ptr(a3d_c[f]) if a3d_c is not None else None, ptr(gt_rows[f]), ptr(gt8),
                                           ptr(assign[f]), K, E, int(norm), ptr(means), ptr(stds), ptr(o_lab),
                                           ptr(o_rois), ptr(o_a3d), ptr(o_sc), ptr(o_t), ptr(o_iw), ptr(o_ow), st),
              "b2d_proposal_target_phase2")
        roi_out.append((o_lab, o_rois, o_a3d, o_sc, o_t, o_iw, o_ow))
    return labels, targets, inside_w, outside_w, roi_out
