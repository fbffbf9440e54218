"""anchor_target_layer_torch (lib/layer_utils/anchor_target_layer.py:22-165).

Phase 1 (device): inside filter, IoU, both argmaxes, labels before subsampling, ordered fg/bg
lists.  Host: the two data-dependent ``torch.randperm`` draws, made exactly as the reference
makes them (same generator, same sizes, same order) so sampled indices are bit-identical.
Phase 2 (device): disable, regression targets, weights, unmap, output layouts.
"""
import numpy as np
import torch

from .._lib import B2DError, check, f32c, lib, ptr, require_cuda, stream_ptr, workspaces
from ..model.config import cfg
from .proposal_layer import _info_tensor


def anchor_target_layer_torch(gt_boxes, gt_boxes_dc, info, all_anchors, num_anchors, height, width, dev):
    """-> labels [1,A,H,W], bbox_targets / inside_w / outside_w [1,H,W,4A] (fp32, on ``dev``).

    ``cfg.TRAIN.IGNORE_DC`` has no effect on the reference's result here (it writes -1 into
    labels that are still all -1, anchor_target_layer.py:46,57-62), so gt_boxes_dc is unused.
    """
    require_cuda(gt_boxes, all_anchors)
    device = all_anchors.device
    A = int(num_anchors)
    anchors = f32c(all_anchors)
    N = anchors.shape[0]
    gt = f32c(gt_boxes)
    G = gt.shape[0]
    if G == 0:
        raise B2DError("anchor_target_layer: no ground-truth boxes (the reference fails here too)")
    if gt.shape[1] != 5:
        gt = f32c(torch.cat((gt[:, :4], gt.new_zeros(G, 5 - 4)), 1)) if gt.shape[1] == 4 else f32c(gt[:, :5])
    info_t = _info_tensor(info, device)
    if info_t.shape[1] != 7:
        pad = torch.zeros(1, 7, device=device)
        pad[:, :info_t.shape[1]] = info_t
        info_t = pad
    L = lib(device)
    ws = workspaces.get(device, "anchor_target", L.b2d_anchor_target_workspace_bytes(1, N, G))
    num_gt = torch.tensor([G], dtype=torch.int32, device=device)
    counts = torch.empty(1, 4, dtype=torch.int32, device=device)
    st = stream_ptr(device)
    check(L.b2d_anchor_target_phase1(1, N, G, ptr(anchors), ptr(gt), ptr(num_gt), ptr(info_t),
                                     float(cfg.TRAIN.RPN_NEGATIVE_OVERLAP), float(cfg.TRAIN.RPN_POSITIVE_OVERLAP),
                                     int(bool(cfg.TRAIN.RPN_CLOBBER_POSITIVES)), ptr(counts), ptr(ws), ws.numel(), st),
          "b2d_anchor_target_phase1")
    _, n_fg, n_bg, _ = counts[0].tolist()                         # the reference syncs at len(fg_inds) too
    num_fg = int(cfg.TRAIN.RPN_FG_FRACTION * cfg.TRAIN.RPN_BATCHSIZE)
    fg_dis = bg_dis = None
    if n_fg > num_fg:                                             # :95-98
        fg_dis = torch.randperm(n_fg, device=dev)[num_fg:].contiguous()
    fg_left = n_fg - (0 if fg_dis is None else fg_dis.numel())
    num_bg = cfg.TRAIN.RPN_BATCHSIZE - fg_left                    # :101-102
    if n_bg > num_bg:                                             # :104-107
        bg_dis = torch.randperm(n_bg, device=dev)[num_bg:].contiguous()
    n_fd = 0 if fg_dis is None else fg_dis.numel()
    n_bd = 0 if bg_dis is None else bg_dis.numel()
    stride = max(n_fd, n_bd, 1)
    nfd_t = torch.tensor([n_fd], dtype=torch.int32, device=device)
    nbd_t = torch.tensor([n_bd], dtype=torch.int32, device=device)
    if fg_dis is None:
        fg_dis = torch.zeros(1, dtype=torch.int64, device=device)
    if bg_dis is None:
        bg_dis = torch.zeros(1, dtype=torch.int64, device=device)
    fg_dis, bg_dis = fg_dis.to(device), bg_dis.to(device)
    labels = torch.empty(1, A, int(height), int(width), device=device)
    targets = torch.empty(1, int(height), int(width), 4 * A, device=device)
    inside_w = torch.empty_like(targets)
    outside_w = torch.empty_like(targets)
    iw4 = torch.tensor(np.array(cfg.TRAIN.RPN_BBOX_INSIDE_WEIGHTS, dtype=np.float32), device=device)
    pw = float(cfg.TRAIN.RPN_POSITIVE_WEIGHT)
    if pw >= 0:
        assert 0 < pw < 1                                          # :127-128
    check(L.b2d_anchor_target_phase2(1, N, G, A, int(height), int(width), ptr(anchors), ptr(gt), ptr(counts),
                                     ptr(fg_dis), ptr(nfd_t), ptr(bg_dis), ptr(nbd_t), stride, ptr(iw4), pw,
                                     ptr(labels), ptr(targets), ptr(inside_w), ptr(outside_w), ptr(ws), ws.numel(),
                                     st), "b2d_anchor_target_phase2")
    return labels, targets, inside_w, outside_w


def anchor_target_layer(gt_boxes, gt_boxes_dc, info, _feat_stride, all_anchors, num_anchors, height, width):
    """numpy-typed variant (anchor_target_layer.py:171-332; broken upstream on numpy>=1.24,
    SURVEY.md F6).  Same computation on the current CUDA device, numpy in / numpy out."""
    dev = torch.device("cuda", torch.cuda.current_device())
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device=dev)
    out = anchor_target_layer_torch(t(gt_boxes), t(gt_boxes_dc), info, t(all_anchors), num_anchors, height, width, dev)
    return tuple(o.cpu().numpy() for o in out)
