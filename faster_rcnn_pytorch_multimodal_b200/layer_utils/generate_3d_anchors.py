"""3-D BEV anchor grid (lib/layer_utils/generate_3d_anchors.py:15-118).

Generated once per (H, W) and cached by the caller; a few thousand rows of closed-form values,
so it is assembled with torch ops on the device (no kernel needed), in (y, x, size, rot) order.
"""
import math

import numpy as np
import torch

from ..model.config import cfg


class GridAnchor3dGenerator(object):
    def name_scope(self):
        return 'GridAnchor3dGenerator'

    def _generate(self, height, width, feature_stride, anchor_scales, anchor_rotations, frame_scale, device=None):
        scales = np.asarray(anchor_scales).reshape(-1)
        assert len(scales) == 1                                           # :31
        x_max = width * feature_stride - 1
        y_max = height * feature_stride - 1
        voxel_len = cfg.LIDAR.VOXEL_LEN / frame_scale
        sizes = np.asarray(cfg.LIDAR.ANCHORS, np.float64) / np.array([voxel_len, voxel_len, 1.0]) * scales[0]
        return tile_anchors_3d([[0, x_max], [0, y_max], [0, 0]], sizes, [feature_stride, feature_stride],
                               np.asarray(anchor_rotations, np.float64), device=device)


def tile_anchors_3d(area_extents, anchor_3d_sizes, anchor_stride, anchor_rotations, device=None):
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    xc = torch.arange(area_extents[0][0], area_extents[0][1], anchor_stride[0], dtype=torch.float32, device=device)
    yc = torch.arange(area_extents[1][0], area_extents[1][1], anchor_stride[1], dtype=torch.float32, device=device)
    sizes = torch.as_tensor(np.asarray(anchor_3d_sizes, np.float32), device=device)      # fp64 -> fp32 store (:109)
    rots = torch.as_tensor(np.asarray(anchor_rotations, np.float32), device=device)      # (:113)
    Y, X, S, R = len(yc), len(xc), sizes.shape[0], rots.shape[0]
    out = torch.zeros(Y, X, S, R, 7, device=device)
    out[..., 0] = xc.view(1, X, 1, 1)
    out[..., 1] = yc.view(Y, 1, 1, 1)
    out[..., 2] = float(np.float32(np.float32(0.0) + np.float64(anchor_3d_sizes[0][2]) / 2.0))   # (:99)
    out[..., 3:6] = sizes.view(1, 1, S, 1, 3)
    out[..., 6] = rots.view(1, 1, 1, R)
    out = out.reshape(-1, 7)
    return out.shape[0], out
