"""CPU restatement of the reference's detection-glue arithmetic (oracle; tests only).

Every function names the reference ``file:line`` it restates (paths relative to
``/root/reference/lib``).  The restatement keeps the reference's *operation order*
in fp32 torch so that integer-valued results (sort order, NMS keep, labels,
sampled indices) are bit-identical, but it is written against explicit
parameters (``GlueCfg``) instead of the reference's module-global ``cfg``.

Third-party arithmetic: ``torchvision.ops.nms`` / ``roi_align`` (reference pins
torchvision==0.4.0, ``req.txt:282``; the operative copy is the container's
0.26.0).  ``nms_greedy_np`` / ``roi_align_np`` / ``roi_align_backward_np`` below
restate their published algorithm in numpy and are pinned against that copy in
``tests/test_oracle.py``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional, Sequence, Tuple

import numpy as np
import torch


# --------------------------------------------------------------------------------------
# configuration (values: model/config.py, lines cited per field)
# --------------------------------------------------------------------------------------
@dataclass
class GlueCfg:
    net_type: str = "image"                      # config.py:54 (default 'lidar'; CLIs set it)
    # RPN proposal selection, keyed by cfg_key
    train_pre_nms: int = 12000                   # config.py:192
    train_post_nms: int = 2000                   # config.py:195
    train_nms_thresh: float = 0.7                # config.py:189
    test_pre_nms: int = 6000                     # config.py:253
    test_post_nms: int = 300                     # config.py:256
    test_nms_thresh: float = 0.7                 # config.py:250
    test_rpn_top_n: int = 5000                   # config.py:266
    # anchor targets
    rpn_positive_overlap: float = 0.7            # config.py:174
    rpn_negative_overlap: float = 0.3            # config.py:177
    rpn_clobber_positives: bool = False          # config.py:180
    rpn_fg_fraction: float = 0.5                 # config.py:183
    rpn_batchsize: int = 256                     # config.py:186
    rpn_bbox_inside_weights: Tuple[float, ...] = (1.0, 1.0, 1.0, 1.0)   # config.py:198
    rpn_positive_weight: float = -1.0            # config.py:203
    ignore_dc: bool = False                      # config.py:210
    dc_thresh: float = 0.5                       # config.py:130
    # proposal targets
    use_gt: bool = False                         # config.py:108
    roi_batch_size: int = 256                    # config.py:123
    fg_fraction: float = 0.25                    # config.py:126
    fg_thresh: float = 0.6                       # config.py:129
    bg_thresh_hi: float = 0.5                    # config.py:133
    bg_thresh_lo: float = 0.0                    # config.py:134
    normalize_targets_precomputed: bool = True   # config.py:161
    image_means: Tuple[float, ...] = (0.0, 0.0, 0.0, 0.0)               # config.py:222
    image_stds: Tuple[float, ...] = (0.1, 0.1, 0.2, 0.2)                # config.py:223
    lidar_means: Tuple[float, ...] = (0.0,) * 7                         # config.py:219
    lidar_stds: Tuple[float, ...] = (0.1, 0.1, 0.1, 0.2, 0.2, 0.2, 1.0)  # config.py:220
    # lidar geometry
    lidar_z_range: Tuple[float, float] = (-3.0, 3.0)                    # config.py:399
    lidar_voxel_len: float = 0.1                 # config.py:400
    lidar_voxel_height: float = 0.5              # config.py:401
    lidar_anchors: Tuple[Tuple[float, float, float], ...] = ((4.73, 2.08, 1.77),)  # config.py:421
    test_nms_final: float = 0.6                  # config.py:234

    def rpn(self, cfg_key: str) -> Tuple[int, int, float]:
        if isinstance(cfg_key, bytes):           # proposal_layer.py:25-26
            cfg_key = cfg_key.decode("utf-8")
        if cfg_key == "TRAIN":
            return self.train_pre_nms, self.train_post_nms, self.train_nms_thresh
        if cfg_key == "TEST":
            return self.test_pre_nms, self.test_post_nms, self.test_nms_thresh
        raise KeyError(cfg_key)


DEFAULT_CFG = GlueCfg()


# --------------------------------------------------------------------------------------
# anchors  (layer_utils/generate_anchors.py:41-105, layer_utils/snippets.py:13-40)
# --------------------------------------------------------------------------------------
def _box_whc(box):
    """width, height, centre of an inclusive-pixel box (generate_anchors.py:57-66)."""
    w = box[2] - box[0] + 1
    h = box[3] - box[1] + 1
    return w, h, box[0] + 0.5 * (w - 1), box[1] + 0.5 * (h - 1)


def _boxes_from_whc(ws, hs, cx, cy):
    """generate_anchors.py:69-79."""
    ws = np.asarray(ws, dtype=np.float64).reshape(-1, 1)
    hs = np.asarray(hs, dtype=np.float64).reshape(-1, 1)
    return np.hstack((cx - 0.5 * (ws - 1), cy - 0.5 * (hs - 1),
                      cx + 0.5 * (ws - 1), cy + 0.5 * (hs - 1)))


def generate_anchors(base_size=16, ratios=(0.5, 1, 2), scales=(8, 16, 32)) -> np.ndarray:
    """A = len(ratios)*len(scales) base windows, ratio-major / scale-minor, fp64.

    generate_anchors.py:41-54 with _ratio_enum :82-93 (np.round = half-to-even) and
    _scale_enum :96-105.
    """
    ratios = np.asarray(ratios, dtype=np.float64)
    scales = np.asarray(scales, dtype=np.float64)
    base = np.array([0, 0, base_size - 1, base_size - 1], dtype=np.float64)
    w, h, cx, cy = _box_whc(base)
    ws = np.round(np.sqrt((w * h) / ratios))
    hs = np.round(ws * ratios)
    per_ratio = _boxes_from_whc(ws, hs, cx, cy)
    out = []
    for row in per_ratio:
        rw, rh, rcx, rcy = _box_whc(row)
        out.append(_boxes_from_whc(rw * scales, rh * scales, rcx, rcy))
    return np.vstack(out)


def generate_anchors_pre(height, width, feat_stride, anchor_scales=(8, 16, 32),
                         anchor_ratios=(0.5, 1, 2), frame_scale=1.0):
    """All N=height*width*A anchors in (h, w, a) order, fp32; snippets.py:13-40."""
    base = generate_anchors(ratios=np.asarray(anchor_ratios, dtype=np.float64),
                            scales=np.asarray(anchor_scales) * frame_scale)
    sx = np.arange(0, width) * feat_stride
    sy = np.arange(0, height) * feat_stride
    gx, gy = np.meshgrid(sx, sy)                       # [H, W]
    shifts = np.stack((gx.ravel(), gy.ravel(), gx.ravel(), gy.ravel()), axis=1)
    allb = (base[None, :, :] + shifts[:, None, :]).reshape(-1, 4)
    allb = allb.astype(np.float32, copy=False)
    return allb, np.int32(allb.shape[0])


def generate_3d_anchors(height, width, feature_stride, anchor_scales, anchor_rotations,
                        frame_scale, cfg: GlueCfg = DEFAULT_CFG):
    """3-D BEV anchors [x,y,z,l,w,h,ry] in (y, x, size, rot) order, fp32.

    generate_3d_anchors.py:15-44 (extent/size set-up) and :47-118 (tiling).
    """
    scales = np.asarray(anchor_scales).reshape(-1)
    assert scales.size == 1                                            # :31
    x_max = width * feature_stride - 1
    y_max = height * feature_stride - 1
    voxel_len = cfg.lidar_voxel_len / frame_scale
    sizes = np.asarray(cfg.lidar_anchors, dtype=np.float64) / np.array([voxel_len, voxel_len, 1.0]) * scales[0]
    rots = np.asarray(anchor_rotations, dtype=np.float64)
    xc = np.array(np.arange(0, x_max, step=feature_stride), dtype=np.float32)   # :72-73
    yc = np.array(np.arange(0, y_max, step=feature_stride), dtype=np.float32)   # :77-78
    grid = np.meshgrid(xc, yc, np.arange(len(sizes)), np.arange(len(rots)))     # 'xy' -> [Y, X, S, R]
    flat = np.stack(grid, axis=4).reshape(-1, 4)
    n = flat.shape[0]
    out = np.zeros((n, 7), dtype=np.float32)
    out[:, 0] = flat[:, 0]
    out[:, 1] = flat[:, 1]
    out[:, 2] = np.zeros_like(flat[:, 0]) + sizes[0][2] / 2.0                   # :99
    out[:, 3:6] = sizes[flat[:, 2].astype(np.int32)]
    out[:, 6] = rots[flat[:, 3].astype(np.int32)]
    return n, out


def bbaa_graphics_gems(bboxes: np.ndarray, width, height, clip=True) -> np.ndarray:
    """Rotated BEV box -> enclosing axis-aligned box (Arvo); utils/bbox.py:256-293.

    numpy flavour: cos/sin and the min/max products run in the input dtype promoted
    with fp64 zeros (so fp64), the per-axis sums are cast to fp32 (:279-280), then the
    centre is added (:282-283).
    """
    b = np.asarray(bboxes)
    rot = b[:, 6]                       # cos/sin stay in the input dtype (fp32 anchors -> fp32)
    c, s = np.cos(rot), np.sin(rot)
    # M[i] = [[c, s], [-s, c]]; a = M * Amin (per column k), b = M * Amax
    half = np.zeros((b.shape[0], 2))
    half[:, 0] = b[:, 3] / 2.0
    half[:, 1] = b[:, 4] / 2.0
    M = np.stack((np.stack((c, s), axis=1), np.stack((-s, c), axis=1)), axis=1)  # [n,2(j),2(k)]
    lo = M * (-half)[:, None, :]
    hi = M * (half)[:, None, :]
    bmin = np.minimum(lo, hi).sum(axis=2).astype(np.float32)
    bmax = np.maximum(lo, hi).sum(axis=2).astype(np.float32)
    bmin = bmin + b[:, 0:2]
    bmax = bmax + b[:, 0:2]
    out = np.concatenate((bmin[:, 0:1], bmin[:, 1:2], bmax[:, 0:1], bmax[:, 1:2]), axis=1)
    if clip:                                              # _bbox_clip, utils/bbox.py:93-98
        out[:, 0] = np.clip(out[:, 0], 0, width - 1)
        out[:, 2] = np.clip(out[:, 2], 0, width - 1)
        out[:, 1] = np.clip(out[:, 1], 0, height - 1)
        out[:, 3] = np.clip(out[:, 3], 0, height - 1)
    return out


def bbaa_graphics_gems_torch(bboxes: torch.Tensor, width, height, clip=True) -> torch.Tensor:
    """fp32 torch flavour; utils/bbox.py:296-336."""
    rot = bboxes[:, 6]
    c, s = torch.cos(rot), torch.sin(rot)
    half = torch.stack((bboxes[:, 3] / 2.0, bboxes[:, 4] / 2.0), dim=1)
    M = torch.stack((torch.stack((c, s), dim=1), torch.stack((-s, c), dim=1)), dim=1)
    lo = M * (-half)[:, None, :]
    hi = M * half[:, None, :]
    bmin = torch.min(lo, hi).sum(dim=2).float() + bboxes[:, 0:2]
    bmax = torch.max(lo, hi).sum(dim=2).float() + bboxes[:, 0:2]
    out = torch.cat((bmin[:, 0:1], bmin[:, 1:2], bmax[:, 0:1], bmax[:, 1:2]), dim=1)
    if clip:
        out[:, 0] = out[:, 0].clamp(0, width - 1)
        out[:, 2] = out[:, 2].clamp(0, width - 1)
        out[:, 1] = out[:, 1].clamp(0, height - 1)
        out[:, 3] = out[:, 3].clamp(0, height - 1)
    return out


# --------------------------------------------------------------------------------------
# box codecs  (model/bbox_transform.py)
# --------------------------------------------------------------------------------------
def bbox_transform(ex_rois: torch.Tensor, gt_rois: torch.Tensor) -> torch.Tensor:
    """Encoder; centre deltas normalised by the box diagonal. bbox_transform.py:52-70."""
    ew = ex_rois[:, 2] - ex_rois[:, 0] + 1.0
    eh = ex_rois[:, 3] - ex_rois[:, 1] + 1.0
    diag = torch.sqrt(ew * ew + eh * eh)
    ecx = ex_rois[:, 0] + 0.5 * ew
    ecy = ex_rois[:, 1] + 0.5 * eh
    gw = gt_rois[:, 2] - gt_rois[:, 0] + 1.0
    gh = gt_rois[:, 3] - gt_rois[:, 1] + 1.0
    gcx = gt_rois[:, 0] + 0.5 * gw
    gcy = gt_rois[:, 1] + 0.5 * gh
    return torch.stack(((gcx - ecx) / diag, (gcy - ecy) / diag,
                        torch.log(gw / ew), torch.log(gh / eh)), dim=1)


def bbox_transform_inv(boxes: torch.Tensor, deltas: torch.Tensor, scales=None) -> torch.Tensor:
    """Decoder for [N,4K] deltas. bbox_transform.py:75-105."""
    if scales is not None:
        boxes = boxes / scales
    if len(boxes) == 0:
        return deltas.detach() * 0
    w = boxes[:, 2] - boxes[:, 0] + 1.0
    h = boxes[:, 3] - boxes[:, 1] + 1.0
    diag = torch.sqrt(w * w + h * h)
    cx = boxes[:, 0] + 0.5 * w
    cy = boxes[:, 1] + 0.5 * h
    pcx = deltas[:, 0::4] * diag[:, None] + cx[:, None]
    pcy = deltas[:, 1::4] * diag[:, None] + cy[:, None]
    pw = torch.exp(deltas[:, 2::4]) * w[:, None]
    ph = torch.exp(deltas[:, 3::4]) * h[:, None]
    out = torch.stack((pcx - 0.5 * pw, pcy - 0.5 * ph, pcx + 0.5 * pw, pcy + 0.5 * ph), dim=2)
    return out.reshape(len(boxes), -1)


def clip_boxes(boxes: torch.Tensor, info) -> torch.Tensor:
    """x in [info0, info1-1], y in [info2, info3-1]. bbox_transform.py:235-257."""
    b = boxes.reshape(boxes.shape[0], -1, 4)
    out = torch.stack((b[:, :, 0].clamp(info[0], info[1] - 1),
                       b[:, :, 1].clamp(info[2], info[3] - 1),
                       b[:, :, 2].clamp(info[0], info[1] - 1),
                       b[:, :, 3].clamp(info[2], info[3] - 1)), dim=2)
    return out.reshape(boxes.shape[0], -1)


def lidar_3d_bbox_transform(ex_rois, ex_anchors, gt_rois) -> torch.Tensor:
    """7-DoF encoder on the RoI's AABB extent. bbox_transform.py:16-49."""
    rl = ex_rois[:, 2] - ex_rois[:, 0] + 1
    rw = ex_rois[:, 3] - ex_rois[:, 1] + 1
    eh = ex_anchors[:, 5]
    cx = ex_rois[:, 0] + rl / 2.0
    cy = ex_rois[:, 1] + rw / 2.0
    cz = ex_anchors[:, 2]
    diag = torch.sqrt(rl * rl + rw * rw)
    return torch.stack(((gt_rois[:, 0] - cx) / diag,
                        (gt_rois[:, 1] - cy) / diag,
                        (gt_rois[:, 2] - cz) / eh,
                        torch.log(gt_rois[:, 3] / rl),
                        torch.log(gt_rois[:, 4] / rw),
                        torch.log(gt_rois[:, 5] / eh),
                        gt_rois[:, 6]), dim=1)


def lidar_3d_bbox_transform_inv(rois, boxes, deltas, scales=None) -> torch.Tensor:
    """7-DoF decoder; NOTE mutates ``boxes`` in place when scales is given (:178-180).
    bbox_transform.py:174-233."""
    if scales is not None:
        boxes[:, 0:2] = boxes[:, 0:2] / scales
        boxes[:, 3:5] = boxes[:, 3:5] / scales
        rois = rois / scales
    if len(boxes) == 0:
        return deltas.detach() * 0
    rl = rois[:, 2] - rois[:, 0] + 1
    rw = rois[:, 3] - rois[:, 1] + 1
    hh = boxes[:, 5]
    cx = rois[:, 0] + rl / 2.0
    cy = rois[:, 1] + rw / 2.0
    cz = boxes[:, 2]
    diag = torch.sqrt(rl * rl + rw * rw)
    cols = (deltas[:, 0::7] * diag[:, None] + cx[:, None],
            deltas[:, 1::7] * diag[:, None] + cy[:, None],
            deltas[:, 2::7] * hh[:, None] + cz[:, None],
            torch.exp(deltas[:, 3::7]) * rl[:, None],
            torch.exp(deltas[:, 4::7]) * rw[:, None],
            torch.exp(deltas[:, 5::7]) * hh[:, None],
            deltas[:, 6::7])
    return torch.stack(cols, dim=2).reshape(len(boxes), -1)


def lidar_3d_uncertainty_transform_inv(rois, boxes, deltas, uncertainty, scales=None) -> torch.Tensor:
    """bbox_transform.py:132-169 (same in-place scale mutation as the decoder)."""
    if scales is not None:
        boxes[:, 0:2] = boxes[:, 0:2] / scales
        boxes[:, 3:5] = boxes[:, 3:5] / scales
        rois = rois / scales
    rl = rois[:, 2] - rois[:, 0] + 1
    rw = rois[:, 3] - rois[:, 1] + 1
    hh = boxes[:, 5]
    cols = (uncertainty[:, 0::7] * rl[:, None],
            uncertainty[:, 1::7] * rw[:, None],
            uncertainty[:, 2::7] * hh[:, None],
            torch.exp(uncertainty[:, 3::7]) - 1,
            torch.exp(uncertainty[:, 4::7]) - 1,
            torch.exp(uncertainty[:, 5::7]) - 1,
            uncertainty[:, 6::7])
    inv = torch.stack(cols, dim=2).reshape(len(boxes), -1)
    return torch.pow(inv, 2)


# --------------------------------------------------------------------------------------
# IoU  (utils/bbox.py:5-33)
# --------------------------------------------------------------------------------------
def bbox_overlaps(boxes, query_boxes):
    """Dense [N,K] IoU with the legacy +1 pixel convention. utils/bbox.py:5-33."""
    as_np = isinstance(boxes, np.ndarray)
    if as_np:
        boxes = torch.from_numpy(boxes)
        query_boxes = torch.from_numpy(query_boxes)
    ba = (boxes[:, 2] - boxes[:, 0] + 1) * (boxes[:, 3] - boxes[:, 1] + 1)
    qa = (query_boxes[:, 2] - query_boxes[:, 0] + 1) * (query_boxes[:, 3] - query_boxes[:, 1] + 1)
    iw = (torch.min(boxes[:, 2:3], query_boxes[:, 2:3].t())
          - torch.max(boxes[:, 0:1], query_boxes[:, 0:1].t()) + 1).clamp(min=0)
    ih = (torch.min(boxes[:, 3:4], query_boxes[:, 3:4].t())
          - torch.max(boxes[:, 1:2], query_boxes[:, 1:2].t()) + 1).clamp(min=0)
    ua = ba.view(-1, 1) + qa.view(1, -1) - iw * ih
    ov = iw * ih / ua
    return ov.numpy() if as_np else ov


# --------------------------------------------------------------------------------------
# NMS  (torchvision.ops.nms; call sites proposal_layer.py:46, filter_predictions.py:67,69)
# --------------------------------------------------------------------------------------
def nms(boxes: torch.Tensor, scores: torch.Tensor, thresh: float) -> torch.Tensor:
    """The reference's third-party NMS (torchvision CPU kernel)."""
    from torchvision.ops import nms as tv_nms
    return tv_nms(boxes.float().cpu(), scores.float().cpu(), float(thresh))


def nms_greedy_np(boxes: np.ndarray, scores: np.ndarray, thresh: float) -> np.ndarray:
    """numpy restatement of torchvision's greedy NMS (csrc/ops/cpu/nms_kernel.cpp).

    Semantics: stable descending score order; area=(x2-x1)*(y2-y1) (no +1); box j is
    suppressed by an earlier kept box i iff  inter/(area_i+area_j-inter) > thresh,
    evaluated in fp32 and compared against the *double* threshold; 0/0=NaN never
    suppresses.  Returns kept original indices in score order (int64).
    """
    b = np.asarray(boxes, dtype=np.float32)
    n = b.shape[0]
    if n == 0:
        return np.zeros((0,), dtype=np.int64)
    order = np.argsort(-np.asarray(scores, dtype=np.float32), kind="stable")
    x1, y1, x2, y2 = (b[order, k] for k in range(4))
    area = (x2 - x1) * (y2 - y1)
    dead = np.zeros(n, dtype=bool)
    keep = []
    zero = np.float32(0)
    with np.errstate(invalid="ignore", divide="ignore"):
        for i in range(n):
            if dead[i]:
                continue
            keep.append(order[i])
            if i + 1 == n:
                break
            w = np.maximum(zero, np.minimum(x2[i], x2[i + 1:]) - np.maximum(x1[i], x1[i + 1:]))
            h = np.maximum(zero, np.minimum(y2[i], y2[i + 1:]) - np.maximum(y1[i], y1[i + 1:]))
            inter = w * h
            ovr = inter / (area[i] + area[i + 1:] - inter)
            dead[i + 1:] |= ovr.astype(np.float64) > float(thresh)
    return np.asarray(keep, dtype=np.int64)


# --------------------------------------------------------------------------------------
# proposal layers  (layer_utils/proposal_layer.py:18-57, proposal_top_layer.py:18-59)
# --------------------------------------------------------------------------------------
def proposal_layer(rpn_cls_prob, rpn_bbox_pred, info, cfg_key, anchors, anchors_3d, num_anchors,
                   cfg: GlueCfg = DEFAULT_CFG, stable_sort: bool = True):
    """RPN outputs -> (blob[R,5], scores[R,1], anchors_3d[R,7]).

    ``stable_sort=True`` pins tie order to lower-flat-index-first (the
    reference's ``sort(descending=True)`` at proposal_layer.py:39 is not stable).
    """
    pre, post, thr = cfg.rpn(cfg_key)
    scores = rpn_cls_prob[:, :, :, num_anchors:].contiguous().view(-1)          # :32,34
    deltas = rpn_bbox_pred.view(-1, 4)                                           # :33
    props = clip_boxes(bbox_transform_inv(anchors, deltas), info)                # :35-36
    scores, order = scores.sort(descending=True, stable=stable_sort)             # :39
    if pre > 0:                                                                  # :40-42
        order = order[:pre]
        scores = scores[:pre]
    props = props[order]
    a3d = anchors_3d[order]
    keep = nms(props, scores, thr)                                               # :46
    if post > 0:
        keep = keep[:post]
    props, scores, a3d = props[keep], scores[keep].view(-1, 1), a3d[keep]
    blob = torch.cat((props.new_zeros(props.shape[0], 1), props), dim=1)         # :54-55
    return blob, scores, a3d


def proposal_top_layer(rpn_cls_prob, rpn_bbox_pred, info, anchors, num_anchors,
                       cfg: GlueCfg = DEFAULT_CFG, stable_sort: bool = True, rng=None):
    """TEST.MODE=='top': top-N by score, no NMS. proposal_top_layer.py:18-59."""
    top_n = cfg.test_rpn_top_n
    scores = rpn_cls_prob[:, :, :, num_anchors:].contiguous().view(-1, 1)
    deltas = rpn_bbox_pred.view(-1, 4)
    n = scores.shape[0]
    if n < top_n:                                                                # :33-38
        rng = np.random if rng is None else rng
        top = torch.from_numpy(rng.choice(n, size=top_n, replace=True)).long()
    else:
        top = scores.view(-1).sort(descending=True, stable=stable_sort)[1][:top_n]
    anc = anchors[top].contiguous()
    props = clip_boxes(bbox_transform_inv(anc, deltas[top].contiguous()), info)
    blob = torch.cat((props.new_zeros(props.shape[0], 1), props), dim=1)
    return blob, scores[top].contiguous(), anc


# --------------------------------------------------------------------------------------
# anchor targets  (layer_utils/anchor_target_layer.py:22-165)
# --------------------------------------------------------------------------------------
def anchor_target_layer(gt_boxes, gt_boxes_dc, info, all_anchors, num_anchors, height, width,
                        cfg: GlueCfg = DEFAULT_CFG, generator: Optional[torch.Generator] = None,
                        return_debug: bool = False):
    """RPN labels / regression targets / weights.

    Restates ``anchor_target_layer_torch``.  The two ``torch.randperm`` draws
    (:96, :105) consume ``generator`` (or torch's global CPU generator) exactly
    as the reference does, so seeded runs are bit-identical.
    """
    A = num_anchors
    total = all_anchors.shape[0]
    inside = torch.where((all_anchors[:, 0] >= info[0]) & (all_anchors[:, 1] >= info[2]) &
                         (all_anchors[:, 2] < info[1]) & (all_anchors[:, 3] < info[3]))[0]   # :37-42
    anc = all_anchors[inside]
    n_in = inside.numel()
    labels = torch.full((n_in,), -1, dtype=torch.int64)
    ov = bbox_overlaps(anc.contiguous(), gt_boxes.contiguous())           # uses all 5 cols' first 4
    if cfg.ignore_dc:                                                     # :57-62
        ov_dc = bbox_overlaps(anc.contiguous(), gt_boxes_dc.contiguous())
        labels[torch.argwhere(ov_dc > cfg.dc_thresh)[:, 0]] = -1
    arg = ov.argmax(dim=1)                                                # :64
    mx = ov[torch.arange(n_in), arg]
    gt_arg = ov.argmax(dim=0)
    gt_mx = ov[gt_arg, torch.arange(ov.shape[1])]
    gt_mx = torch.clamp(gt_mx, torch.finfo(torch.float32).eps, float("inf"))   # :71
    gt_pos = torch.where(ov == gt_mx)[0]                                  # :72
    if not cfg.rpn_clobber_positives:
        labels[mx < cfg.rpn_negative_overlap] = 0                         # :77
    labels[gt_pos] = 1                                                    # :81
    labels[mx >= cfg.rpn_positive_overlap] = 1                            # :86
    if cfg.rpn_clobber_positives:
        labels[mx < cfg.rpn_negative_overlap] = 0
    num_fg = int(cfg.rpn_fg_fraction * cfg.rpn_batchsize)
    fg = torch.where(labels == 1)[0]
    if len(fg) > num_fg:                                                  # :95-98
        perm = torch.randperm(fg.numel(), generator=generator)[num_fg:]
        labels[fg[perm]] = -1
    num_bg = cfg.rpn_batchsize - torch.sum(labels == 1)
    bg = torch.where(labels == 0)[0]
    if len(bg) > num_bg:                                                  # :104-107
        perm = torch.randperm(bg.numel(), generator=generator)[num_bg:]
        labels[bg[perm]] = -1
    targets = bbox_transform(anc, gt_boxes[arg, :][:, :4]).float()        # :110, _compute_targets :361-370
    inside_w = torch.zeros((n_in, 4), dtype=torch.float32)
    inside_w[labels == 1, :] = torch.tensor(cfg.rpn_bbox_inside_weights, dtype=torch.float32)
    outside_w = torch.zeros((n_in, 4), dtype=torch.float32)
    if cfg.rpn_positive_weight < 0:                                       # :118-125
        n_ex = torch.sum(labels >= 0)
        pos_w = neg_w = 1.0 / float(n_ex)
    else:
        pos_w = cfg.rpn_positive_weight / torch.sum(labels == 1)
        neg_w = (1.0 - cfg.rpn_positive_weight) / torch.sum(labels == 0)
    outside_w[labels == 1, :] = pos_w
    outside_w[labels == 0, :] = neg_w

    def unmap(data, fill):                                                # _unmap :335-358
        shape = (total,) + tuple(data.shape[1:])
        ret = torch.full(shape, fill, dtype=torch.float32)
        ret[inside] = data
        return ret

    lab_full = unmap(labels.float(), -1)
    out = (lab_full.reshape(1, height, width, A).permute(0, 3, 1, 2),     # :145
           unmap(targets, 0).reshape(1, height, width, A * 4),
           unmap(inside_w, 0).reshape(1, height, width, A * 4),
           unmap(outside_w, 0).reshape(1, height, width, A * 4))
    if return_debug:
        return out + ({"inside": inside, "max_overlaps": mx, "argmax": arg, "gt_max": gt_mx},)
    return out


# --------------------------------------------------------------------------------------
# proposal targets  (layer_utils/proposal_target_layer.py:22-284)
# --------------------------------------------------------------------------------------
def _choice(n, k, replace, generator):
    """torch_choice :265-284."""
    if replace:
        return torch.randint(n, (k,), generator=generator)
    if k > n:
        factor = math.ceil(k / n)
        idx = torch.arange(n).repeat(factor)
        return idx[torch.randperm(idx.shape[0], generator=generator)][:k]
    return torch.randperm(n, generator=generator)[:k]


def proposal_target_layer(rpn_rois, rpn_scores, anchors_3d, gt_boxes, true_gt_boxes, gt_boxes_dc,
                          num_classes, num_bbox_elem, cfg: GlueCfg = DEFAULT_CFG,
                          generator: Optional[torch.Generator] = None, bg_mode: str = "strict",
                          return_debug: bool = False):
    """RoI-head sampling + per-class regression targets.

    ``bg_mode='strict'`` reproduces the reference as it runs on torch>=1.2 bool
    semantics: ``(a<HI)+(b>=LO)==2`` is all-False, so there is never a
    background candidate (:203-204).  ``bg_mode='intended'`` uses ``&``.
    Raises RuntimeError where the reference would drop into pdb (:232-235).
    """
    rois, scores, a3d = rpn_rois, rpn_scores, anchors_3d
    if cfg.use_gt:                                                        # :35-41
        z = rpn_rois.new_zeros(gt_boxes.shape[0], 1)
        rois = torch.cat((rois, torch.cat((z, gt_boxes[:, :-1]), 1)), 0)
        scores = torch.cat((scores, z), 0)
        a3d = torch.cat((a3d, true_gt_boxes[:, :-1]), 0)
    per_frame = cfg.roi_batch_size / 1
    fg_per_frame = int(round(cfg.fg_fraction * per_frame))
    if cfg.ignore_dc and gt_boxes_dc.shape[0] > 0:                        # :184-190
        mx_dc = bbox_overlaps(rois[:, 1:5], gt_boxes_dc[:, :4]).max(1)[0]
        sel = (mx_dc < cfg.dc_thresh).nonzero().view(-1)
        rois, scores, a3d = rois[sel], scores[sel], a3d[sel]
    ov = bbox_overlaps(rois[:, 1:5], gt_boxes[:, :4])                     # :195
    mx, assign = ov.max(1)
    labels = gt_boxes[assign, [4]]                                        # :198
    fg = (mx >= cfg.fg_thresh).nonzero().view(-1)                         # :200
    if bg_mode == "strict":
        bg = (((mx < cfg.bg_thresh_hi) + (mx >= cfg.bg_thresh_lo)) == 2).nonzero().view(-1)
    else:
        bg = ((mx < cfg.bg_thresh_hi) & (mx >= cfg.bg_thresh_lo)).nonzero().view(-1)
    if fg.numel() > 0 and bg.numel() > 0:                                 # :206-217
        fg_n = min(fg_per_frame, fg.numel())
        fg = fg[_choice(fg.numel(), int(fg_n), False, generator)]
        bg_n = per_frame - fg_n
        bg = bg[_choice(bg.numel(), int(bg_n), bg.numel() < bg_n, generator)]
    elif fg.numel() > 0:                                                  # :218-224
        fg = fg[_choice(fg.numel(), int(per_frame), fg.numel() < per_frame, generator)]
        fg_n = per_frame
    elif bg.numel() > 0:                                                  # :225-231
        bg = bg[_choice(bg.numel(), int(per_frame), bg.numel() < per_frame, generator)]
        fg_n = 0
    else:
        raise RuntimeError("no fg and no bg RoIs (reference enters pdb here)")
    keep = torch.cat([fg, bg], 0)
    labels = labels[keep].contiguous()
    labels[int(fg_n):] = 0                                                # :242
    rois_k, scores_k, a3d_k = rois[keep].contiguous(), scores[keep].contiguous(), a3d[keep].contiguous()
    if cfg.net_type == "lidar":                                           # :248-252, :134-149
        t = lidar_3d_bbox_transform(rois_k[:, 1:5], a3d_k, true_gt_boxes[assign[keep]][:, :-1])
        means, stds = cfg.lidar_means, cfg.lidar_stds
    else:                                                                 # :254-257, :151-164
        t = bbox_transform(rois_k[:, 1:5], gt_boxes[assign[keep]][:, :4])
        means, stds = cfg.image_means, cfg.image_stds
    if cfg.normalize_targets_precomputed:
        t = (t - t.new_tensor(means)) / t.new_tensor(stds)
    E = num_bbox_elem
    tgt = labels.new_zeros(labels.numel(), E * num_classes)               # :64-103
    inw = labels.new_zeros(tgt.shape)
    pos = (labels > 0).nonzero().view(-1)
    for r in pos.tolist():
        c = int(labels[r].item())
        tgt[r, E * c:E * (c + 1)] = t[r]
        inw[r, E * c:E * (c + 1)] = 1.0
    out = (labels.view(-1, 1), rois_k.view(-1, 5), a3d_k, scores_k.view(-1), tgt, inw, (inw > 0).float())
    if return_debug:
        return out + ({"keep": keep, "max_overlaps": mx, "assign": assign, "fg_n": int(fg_n)},)
    return out


# --------------------------------------------------------------------------------------
# RoI align  (torchvision.ops.roi_align; utils/torchpoolers.py:20-51,137-200)
# --------------------------------------------------------------------------------------
def roi_align(feat, rois, output_size=(7, 7), spatial_scale=1.0 / 16, sampling_ratio=2, aligned=False):
    """The reference's third-party RoIAlign (torchvision CPU kernel)."""
    from torchvision.ops import roi_align as tv_roi_align
    return tv_roi_align(feat, rois, output_size, spatial_scale, sampling_ratio, aligned)


def _axis_samples(start, length, pooled, grid, limit):
    """Per-axis bilinear sample table: for each (p, i) -> (low, high, w_low, w_high, valid).

    torchvision cpu/roi_align_kernel.cpp pre_calc_for_bilinear_interpolate, restated per
    axis (the 2-D validity test and the clamps are separable).  fp32 arithmetic.
    """
    f = np.float32
    bin_sz = f(length) / f(pooled)
    lows = np.zeros((pooled, grid), np.int64)
    highs = np.zeros((pooled, grid), np.int64)
    wl = np.zeros((pooled, grid), np.float32)
    wh = np.zeros((pooled, grid), np.float32)
    ok = np.zeros((pooled, grid), bool)
    for p in range(pooled):
        for i in range(grid):
            c = f(f(start) + f(f(p) * bin_sz)) + f(f(f(i) + f(0.5)) * bin_sz) / f(grid)
            c = f(c)
            if c < -1.0 or c > limit:
                continue
            ok[p, i] = True
            if c <= 0:
                c = f(0)
            lo = int(c)
            if lo >= limit - 1:
                hi = lo = limit - 1
                c = f(lo)
            else:
                hi = lo + 1
            l_ = f(c - f(lo))
            lows[p, i], highs[p, i] = lo, hi
            wh[p, i] = l_                 # weight of the 'high' pixel
            wl[p, i] = f(f(1.0) - l_)     # weight of the 'low' pixel
    return lows, highs, wl, wh, ok


def _roi_geometry(roi, spatial_scale, pooled, sampling_ratio, aligned):
    f = np.float32
    off = f(0.5) if aligned else f(0.0)
    x1 = f(f(roi[1]) * f(spatial_scale)) - off
    y1 = f(f(roi[2]) * f(spatial_scale)) - off
    x2 = f(f(roi[3]) * f(spatial_scale)) - off
    y2 = f(f(roi[4]) * f(spatial_scale)) - off
    rw, rh = f(x2 - x1), f(y2 - y1)
    if not aligned:
        rw, rh = max(rw, f(1.0)), max(rh, f(1.0))
    gh = sampling_ratio if sampling_ratio > 0 else int(math.ceil(rh / f(pooled[0])))
    gw = sampling_ratio if sampling_ratio > 0 else int(math.ceil(rw / f(pooled[1])))
    return x1, y1, rw, rh, gh, gw


def roi_align_np(feat: np.ndarray, rois: np.ndarray, output_size=(7, 7), spatial_scale=1.0 / 16,
                 sampling_ratio=2, aligned=False) -> np.ndarray:
    """numpy restatement of RoIAlign forward (small cases). feat [B,C,H,W], rois [R,5]."""
    f = np.float32
    B, C, H, W = feat.shape
    PH, PW = output_size
    out = np.zeros((rois.shape[0], C, PH, PW), np.float32)
    for r, roi in enumerate(rois):
        b = int(roi[0])
        x1, y1, rw, rh, gh, gw = _roi_geometry(roi, spatial_scale, (PH, PW), sampling_ratio, aligned)
        ylo, yhi, wyl, wyh, yok = _axis_samples(y1, rh, PH, gh, H)
        xlo, xhi, wxl, wxh, xok = _axis_samples(x1, rw, PW, gw, W)
        count = f(max(gh * gw, 1))
        for ph in range(PH):
            for pw in range(PW):
                acc = np.zeros(C, np.float32)
                for iy in range(gh):
                    for ix in range(gw):
                        if not (yok[ph, iy] and xok[pw, ix]):
                            continue
                        hy, ly = wyl[ph, iy], wyh[ph, iy]
                        hx, lx = wxl[pw, ix], wxh[pw, ix]
                        a, bb, c_, d = feat[b, :, ylo[ph, iy], xlo[pw, ix]], feat[b, :, ylo[ph, iy], xhi[pw, ix]], \
                            feat[b, :, yhi[ph, iy], xlo[pw, ix]], feat[b, :, yhi[ph, iy], xhi[pw, ix]]
                        acc = acc + (f(hy * hx) * a + f(hy * lx) * bb + f(ly * hx) * c_ + f(ly * lx) * d)
                out[r, :, ph, pw] = acc / count
    return out


def roi_align_backward_np(grad_out: np.ndarray, rois: np.ndarray, feat_shape, spatial_scale=1.0 / 16,
                          sampling_ratio=2, aligned=False) -> np.ndarray:
    """numpy restatement of RoIAlign backward (small cases); fp64 accumulation."""
    f = np.float32
    B, C, H, W = feat_shape
    R, _, PH, PW = grad_out.shape
    gin = np.zeros(feat_shape, np.float64)
    for r, roi in enumerate(rois):
        b = int(roi[0])
        x1, y1, rw, rh, gh, gw = _roi_geometry(roi, spatial_scale, (PH, PW), sampling_ratio, aligned)
        ylo, yhi, wyl, wyh, yok = _axis_samples(y1, rh, PH, gh, H)
        xlo, xhi, wxl, wxh, xok = _axis_samples(x1, rw, PW, gw, W)
        count = f(max(gh * gw, 1))
        for ph in range(PH):
            for pw in range(PW):
                g = grad_out[r, :, ph, pw].astype(np.float64) / np.float64(count)
                for iy in range(gh):
                    for ix in range(gw):
                        if not (yok[ph, iy] and xok[pw, ix]):
                            continue
                        hy, ly = np.float64(wyl[ph, iy]), np.float64(wyh[ph, iy])
                        hx, lx = np.float64(wxl[pw, ix]), np.float64(wxh[pw, ix])
                        gin[b, :, ylo[ph, iy], xlo[pw, ix]] += g * hy * hx
                        gin[b, :, ylo[ph, iy], xhi[pw, ix]] += g * hy * lx
                        gin[b, :, yhi[ph, iy], xlo[pw, ix]] += g * ly * hx
                        gin[b, :, yhi[ph, iy], xhi[pw, ix]] += g * ly * lx
    return gin.astype(np.float32)


def fpn_level_map(boxes: torch.Tensor, k_min: int, k_max: int, canonical_scale=224,
                  canonical_level=4, eps=1e-6) -> torch.Tensor:
    """FPN eq.(1) level index minus k_min. torchpoolers.py:39-51."""
    area = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])       # torchvision box_area
    s = torch.sqrt(area)
    lvl = torch.floor(canonical_level + torch.log2(s / canonical_scale) + torch.tensor(eps, dtype=s.dtype))
    lvl = torch.clamp(lvl, min=k_min, max=k_max)
    return (lvl.to(torch.int64) - k_min).to(torch.int64)


def multiscale_roi_align(feats: Sequence[torch.Tensor], boxes: torch.Tensor, image_shape,
                         output_size=(7, 7), sampling_ratio=2) -> torch.Tensor:
    """MultiScaleRoIAlign.forward for one image. torchpoolers.py:107-200."""
    scales = []
    for ft in feats:                                                       # infer_scale :107-117
        cand = []
        for s1, s2 in zip(ft.shape[-2:], image_shape):
            cand.append(2 ** float(torch.tensor(float(s1) / float(s2)).log2().round()))
        assert cand[0] == cand[1]
        scales.append(cand[0])
    rois = torch.cat((boxes.new_zeros(boxes.shape[0], 1), boxes), dim=1)
    if len(feats) == 1:
        return roi_align(feats[0], rois, output_size, scales[0], sampling_ratio)
    k_min = int(-torch.log2(torch.tensor(scales[0], dtype=torch.float32)).item())
    k_max = int(-torch.log2(torch.tensor(scales[-1], dtype=torch.float32)).item())
    lv = fpn_level_map(boxes, k_min, k_max)
    out = torch.zeros((rois.shape[0], feats[0].shape[1]) + tuple(output_size), dtype=feats[0].dtype)
    for level, (ft, sc) in enumerate(zip(feats, scales)):
        idx = (lv == level).nonzero().view(-1)
        out[idx] = roi_align(ft, rois[idx], output_size, sc, sampling_ratio)
    return out


# --------------------------------------------------------------------------------------
# MC-dropout reductions  (utils/loss_utils.py:103-141, datasets/db.py:283-303)
# --------------------------------------------------------------------------------------
def compute_bbox_var(samples: torch.Tensor) -> torch.Tensor:
    """Unbiased single-pass sample variance over dim 0, clamped at 0. loss_utils.py:114-120."""
    n = samples.shape[0]
    sq_of_sum = torch.pow(torch.sum(samples, dim=0), 2)
    var = torch.sum(torch.pow(samples, 2), dim=0)
    var += -sq_of_sum / n
    var = var / (n - 1)
    return var.clamp_min(0.0)


def compute_bbox_cov(samples: torch.Tensor) -> torch.Tensor:
    """Biased diag(E[xx^T]-mu mu^T); loss_utils.py:103-112 without the hard-coded .cuda()."""
    mu = torch.mean(samples, dim=0)
    x = samples.unsqueeze(3)
    second = torch.mean(torch.matmul(x, x.transpose(2, 3)), dim=0)
    mu = mu.unsqueeze(2)
    cov = second - torch.matmul(mu, mu.transpose(1, 2))
    cov = cov * torch.eye(cov.shape[-1])
    return torch.sum(cov, dim=-1).clamp_min(0.0)


def categorical_entropy(cls_prob: torch.Tensor) -> torch.Tensor:
    """loss_utils.py:122-129."""
    return -torch.sum(cls_prob * torch.log2(cls_prob), dim=1)


def categorical_mutual_information(cls_score: torch.Tensor) -> torch.Tensor:
    """loss_utils.py:132-141; input [T,N,C] logits."""
    p = torch.softmax(cls_score, dim=2)
    total = categorical_entropy(torch.mean(p, dim=0))
    mi = torch.mean(torch.sum(p * torch.log2(p), dim=2), dim=0)
    mi += total
    return mi


def sort_by_uncertainty(var: np.ndarray, descending=False) -> np.ndarray:
    """argsort of mean-over-columns variance (datasets/db.py:264-303), tie order pinned
    to lower index first (``kind='stable'``; numpy's default quicksort is unstable)."""
    key = np.mean(var, axis=1) if var.ndim == 2 else var
    return np.argsort(-key if descending else key, kind="stable")


# --------------------------------------------------------------------------------------
# Final per-class detection filter  (utils/filter_predictions.py:23-130, model/test.py:213-221)
# --------------------------------------------------------------------------------------
def nms_hstack(scores: torch.Tensor, mean_boxes: torch.Tensor, thresh: float, c: int, bbox_elem: int, db_type: str,
               nms_thresh: float = 0.6):
    """utils/filter_predictions.py:45-72 (nms_hstack_torch) -> (cls_dets [m,E+1] np, inds, keep np)."""
    inds = torch.where(scores[:, c] > thresh)[0]                                          # :46
    if inds.shape[0] == 0:                                                               # :49-52
        return np.empty(0), [], []
    cls_scores = scores[inds, c]                                                          # :53
    cls_boxes = mean_boxes[inds, c * bbox_elem:(c + 1) * bbox_elem]                       # :54
    if db_type == "lidar":                                                               # :55-62
        x1 = cls_boxes[:, 0:1] - cls_boxes[:, 3:4] / 2.0
        y1 = cls_boxes[:, 1:2] - cls_boxes[:, 4:5] / 2.0
        x2 = cls_boxes[:, 0:1] + cls_boxes[:, 3:4] / 2.0
        y2 = cls_boxes[:, 1:2] + cls_boxes[:, 4:5] / 2.0
        nms_boxes = torch.cat((x1, y1, x2, y2), dim=1)
    else:
        nms_boxes = cls_boxes
    cls_dets = np.hstack((cls_boxes.numpy(), cls_scores.unsqueeze(1).numpy())).astype(np.float32, copy=False)  # :64
    keep = nms(nms_boxes, cls_scores, nms_thresh).numpy()                                 # :66-69
    return cls_dets[keep, :], inds, keep                                                  # :70-72


def clamp_pred_boxes_image(pred_boxes: torch.Tensor, info, bbox_elem: int) -> torch.Tensor:
    """utils/filter_predictions.py:77-91: clamp every class box to [0, w/scale - 1] x [0, h/scale - 1]."""
    info = np.asarray(info, dtype=np.float32)
    frame_width, frame_height, scale = info[1] - info[0], info[3] - info[2], info[6]
    out = pred_boxes.clone()
    out[:, 0::bbox_elem] = torch.clamp_min(out[:, 0::bbox_elem], 0)
    out[:, 1::bbox_elem] = torch.clamp_min(out[:, 1::bbox_elem], 0)
    out[:, 2::bbox_elem] = torch.clamp_max(out[:, 2::bbox_elem], float(frame_width / scale - 1))
    out[:, 3::bbox_elem] = torch.clamp_max(out[:, 3::bbox_elem], float(frame_height / scale - 1))
    return out


def filter_detections(cls_score: torch.Tensor, pred_boxes: torch.Tensor, info, num_classes: int, bbox_elem: int,
                      db_type: str, thresh: float = 0.1, nms_thresh: float = 0.6, max_dets: int = 0,
                      uc_row: Optional[torch.Tensor] = None, uc_cls: Optional[torch.Tensor] = None):
    """filter_and_draw_prep (:75-130) followed by the max-dets filter of model/test.py:213-221, per class,
    with the uncertainty gathers of nms_hstack_var_torch (:23-43) applied to the ORIGINAL tensors for
    every class (the reference re-uses the dict it has just overwritten, which only works for K = 2).
    -> list over classes of dict(dets [m,E+1], roi [m] source roi, uc_row [m,U], uc_cls [m,U2,E])."""
    boxes = clamp_pred_boxes_image(pred_boxes, info, bbox_elem) if db_type == "image" else pred_boxes
    out = [None] * num_classes
    for j in range(1, num_classes):
        dets, inds, keep = nms_hstack(cls_score, boxes, thresh, j, bbox_elem, db_type, nms_thresh)
        if len(keep) == 0:
            out[j] = dict(dets=np.zeros((0, bbox_elem + 1), np.float32), roi=np.zeros(0, np.int64),
                          uc_row=None, uc_cls=None)
            continue
        roi = inds.numpy()[keep]
        if max_dets > 0 and len(dets) > max_dets:                                         # test.py:213-221
            filter_thresh = np.sort(dets[:, -1])[-max_dets]
            sel = np.where(dets[:, -1] >= filter_thresh)[0]
            dets, roi = dets[sel, :], roi[sel]
        out[j] = dict(dets=dets, roi=roi,
                      uc_row=None if uc_row is None else uc_row.numpy()[roi],
                      uc_cls=None if uc_cls is None else
                      uc_cls.numpy()[roi][:, :, j * bbox_elem:(j + 1) * bbox_elem])
    return out


# --------------------------------------------------------------------------------------
# Tail of the detection head over the MC-dropout stack
# --------------------------------------------------------------------------------------
def head_tail_decode(bbox_pred: torch.Tensor, cls_score: torch.Tensor, rois: torch.Tensor, anchors_3d, info,
                     net_type: str, a_bbox_var: Optional[torch.Tensor] = None, use_scale: bool = False,
                     cfg: GlueCfg = DEFAULT_CFG):
    """One frame: bbox_pred [T,R,K*E], cls_score [T,R,K], rois [R,5] -> dict of the tensors the final per-class filter
    consumes.  The composition is [INFERRED] (Network.test_frame is in the missing lib/nets/network.py); each step is
    the reference's: de-normalisation config.py:219-223, torch.mean, compute_bbox_var loss_utils.py:114-120, the
    decoders bbox_transform.py:75-105,174-233,235-257, lidar_3d_uncertainty_transform_inv :132-169, softmax mean,
    categorical_entropy / categorical_mutual_information loss_utils.py:122-141."""
    lidar = net_type == "lidar"
    E = 7 if lidar else 4
    K = bbox_pred.shape[2] // E
    stds = torch.tensor(cfg.lidar_stds if lidar else cfg.image_stds, dtype=torch.float32).repeat(K)
    means = torch.tensor(cfg.lidar_means if lidar else cfg.image_means, dtype=torch.float32).repeat(K)
    info = np.asarray(info, dtype=np.float32)
    x = bbox_pred * stds + means
    mean_pred = torch.mean(x, dim=0)
    T = x.shape[0]
    var = compute_bbox_var(x) if T > 1 else torch.zeros_like(mean_pred)
    sc = float(info[6]) if use_scale else None
    out = {}
    if lidar:
        out["boxes"] = lidar_3d_bbox_transform_inv(rois[:, 1:5], anchors_3d.clone(), mean_pred, scales=sc)
        out["e_bbox_var"] = lidar_3d_uncertainty_transform_inv(rois[:, 1:5], anchors_3d.clone(), mean_pred, var, scales=sc)
        if a_bbox_var is not None:
            out["a_bbox_var"] = lidar_3d_uncertainty_transform_inv(rois[:, 1:5], anchors_3d.clone(), mean_pred,
                                                                   a_bbox_var, scales=sc)
    else:
        out["boxes"] = clip_boxes(bbox_transform_inv(rois[:, 1:5], mean_pred, scales=sc), info)
        out["e_bbox_var"] = var
        if a_bbox_var is not None:
            out["a_bbox_var"] = a_bbox_var
    probs = torch.mean(torch.softmax(cls_score, dim=2), dim=0)
    out["probs"] = probs
    out["e_entropy"] = categorical_entropy(probs)
    out["e_mutual_info"] = categorical_mutual_information(cls_score)
    return out
