"""IoU and rotated-box helpers with the reference's signatures (lib/utils/bbox.py)."""
import numpy as np
import torch

from .._lib import check, f32c, lib, ptr, require_cuda, stream_ptr


def bbox_overlaps(boxes, query_boxes):
    """utils/bbox.py:5-33: dense [N,K] IoU, +1 pixel convention.  ndarray in -> ndarray out
    (computed on the current CUDA device in fp32; the reference would keep fp64 for fp64 arrays)."""
    as_np = isinstance(boxes, np.ndarray)
    if as_np:
        boxes = torch.as_tensor(np.ascontiguousarray(boxes), dtype=torch.float32, device="cuda")
        query_boxes = torch.as_tensor(np.ascontiguousarray(query_boxes), dtype=torch.float32, device="cuda")
    require_cuda(boxes, query_boxes)
    b, q = f32c(boxes), f32c(query_boxes)
    n, k = b.shape[0], q.shape[0]
    out = torch.empty(n, k, device=b.device)
    check(lib(b.device).b2d_bbox_overlaps(n, k, ptr(b), b.stride(0) if n else 4, ptr(q), q.stride(0) if k else 4, ptr(out),
                                  stream_ptr(b.device)), "b2d_bbox_overlaps")
    return out.cpu().numpy() if as_np else out


def bbaa_graphics_gems_torch(bboxes, width, height, clip=True):
    """utils/bbox.py:296-336: rotated BEV box [x,y,z,l,w,h,ry] -> enclosing AABB [x1,y1,x2,y2]."""
    require_cuda(bboxes)
    b = f32c(bboxes)
    n = b.shape[0]
    out = torch.empty(n, 4, device=b.device)
    check(lib(b.device).b2d_bbaa(n, ptr(b), int(bool(clip)), float(width), float(height), ptr(out), stream_ptr(b.device)),
          "b2d_bbaa")
    return out


def bbaa_graphics_gems(bboxes, width, height, clip=True):
    """utils/bbox.py:256-293 (numpy in / numpy out; fp32 on the device)."""
    t = torch.as_tensor(np.ascontiguousarray(bboxes), dtype=torch.float32, device="cuda")
    return bbaa_graphics_gems_torch(t, width, height, clip).cpu().numpy()
