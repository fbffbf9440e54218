"""Box codecs with the reference's signatures (lib/model/bbox_transform.py), one CUDA launch each.

The fork normalises centre deltas by the box *diagonal* sqrt(w^2+h^2), not by w/h
(bbox_transform.py:55,82).  Kernels: csrc/codecs.cu.
"""
import numpy as np
import torch

from .._lib import B2DError, check, f32c, lib, ptr, require_cuda, stream_ptr


def _info_dev(info, device):
    if isinstance(info, torch.Tensor):
        t = info.to(device=device, dtype=torch.float32).reshape(-1)
    else:
        t = torch.as_tensor(np.asarray(info, dtype=np.float32).reshape(-1), device=device)
    if t.numel() < 4:
        raise B2DError("info/shape needs [x_min, x_max, y_min, y_max]")
    return t.contiguous()


def bbox_transform(ex_rois, gt_rois):
    """bbox_transform.py:52-70 -> targets [n,4]."""
    require_cuda(ex_rois, gt_rois)
    ex, gt = f32c(ex_rois), f32c(gt_rois)
    n = ex.shape[0]
    out = torch.empty(n, 4, device=ex.device)
    check(lib(ex.device).b2d_bbox_transform(n, ptr(ex), ex.stride(0) if n else 4, ptr(gt), gt.stride(0) if n else 4, ptr(out),
                                   stream_ptr(ex.device)), "b2d_bbox_transform")
    return out


def bbox_transform_inv(boxes, deltas, scales=None):
    """bbox_transform.py:75-105: boxes [n,4], deltas [n,4K] -> [n,4K]."""
    if len(boxes) == 0:
        return deltas.detach() * 0                                  # :79-80
    require_cuda(boxes, deltas)
    b, d = f32c(boxes), f32c(deltas)
    n = b.shape[0]
    k = d.shape[1] // 4
    out = torch.empty(n, 4 * k, device=b.device)
    use_scale = scales is not None
    check(lib(b.device).b2d_bbox_transform_inv(n, k, ptr(b), b.stride(0), ptr(d), int(use_scale),
                                       float(scales) if use_scale else 1.0, 0, None, ptr(out),
                                       stream_ptr(b.device)), "b2d_bbox_transform_inv")
    return out


def clip_boxes(boxes, shape):
    """bbox_transform.py:235-257: clamp x to [shape0, shape1-1], y to [shape2, shape3-1]."""
    require_cuda(boxes)
    b = f32c(boxes)
    n = b.shape[0]
    k = (b.numel() // max(n, 1)) // 4
    out = torch.empty_like(b)
    if n:
        info = _info_dev(shape, b.device)
        check(lib(b.device).b2d_clip_boxes(n, k, ptr(b), ptr(info), ptr(out), stream_ptr(b.device)), "b2d_clip_boxes")
    return out.view(n, -1)


def lidar_3d_bbox_transform(ex_rois, ex_anchors, gt_rois):
    """bbox_transform.py:16-49 -> targets [n,7]."""
    require_cuda(ex_rois, ex_anchors, gt_rois)
    r, a, g = f32c(ex_rois), f32c(ex_anchors), f32c(gt_rois)
    n = r.shape[0]
    out = torch.empty(n, 7, device=r.device)
    check(lib(r.device).b2d_lidar_bbox_transform(n, ptr(r), r.stride(0) if n else 4, ptr(a), ptr(g),
                                         g.stride(0) if n else 7, ptr(out), stream_ptr(r.device)),
          "b2d_lidar_bbox_transform")
    return out


def _lidar_inv(rois, boxes, deltas, scales, mode):
    require_cuda(rois, boxes, deltas)
    if scales is not None:
        # the reference mutates `boxes` in place here (bbox_transform.py:178-180 / :134-136)
        boxes[:, 0:2] = boxes[:, 0:2] / scales
        boxes[:, 3:5] = boxes[:, 3:5] / scales
        rois = rois / scales
    r, b, d = f32c(rois), f32c(boxes), f32c(deltas)
    n = b.shape[0]
    k = d.shape[1] // 7
    out = torch.empty(n, 7 * k, device=b.device)
    check(lib(b.device).b2d_lidar_bbox_transform_inv(n, k, ptr(r), r.stride(0) if n else 4, ptr(b), ptr(d), mode, ptr(out),
                                             stream_ptr(b.device)), "b2d_lidar_bbox_transform_inv")
    return out


def lidar_3d_bbox_transform_inv(rois, boxes, deltas, scales=None):
    """bbox_transform.py:174-233."""
    if scales is None and len(boxes) == 0:
        return deltas.detach() * 0
    return _lidar_inv(rois, boxes, deltas, scales, 0)


def lidar_3d_uncertainty_transform_inv(rois, boxes, deltas, uncertainty, scales=None):
    """bbox_transform.py:132-169 (``deltas`` is unused there as well)."""
    return _lidar_inv(rois, boxes, uncertainty, scales, 1)
