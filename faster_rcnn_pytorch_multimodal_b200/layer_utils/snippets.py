"""generate_anchors_pre (lib/layer_utils/snippets.py:13-40): tile the A base anchors over the grid."""
import ctypes as C

import numpy as np
import torch

from .._lib import check, lib, ptr, stream_ptr
from .generate_anchors import generate_anchors


def generate_anchors_pre(height, width, feat_stride, anchor_scales=(8, 16, 32), anchor_ratios=(0.5, 1, 2),
                         frame_scale=1.0, device=None):
    """-> (anchors [H*W*A, 4] fp32 CUDA tensor in (h, w, a) order, length int32).

    The reference returns a numpy array that the caller uploads every frame; this returns the
    device tensor directly (``.cpu().numpy()`` gives the reference's array bit for bit).
    """
    base = generate_anchors(ratios=np.asarray(anchor_ratios, np.float64),
                            scales=np.asarray(anchor_scales) * frame_scale)
    A = base.shape[0]
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    out = torch.empty(int(height) * int(width) * A, 4, device=device)
    flat = np.ascontiguousarray(base.reshape(-1), dtype=np.float64)
    check(lib(device).b2d_generate_anchors(int(height), int(width), int(feat_stride), A,
                                     flat.ctypes.data_as(C.POINTER(C.c_double)), ptr(out), stream_ptr(device)),
          "b2d_generate_anchors")
    return out, np.int32(out.shape[0])
