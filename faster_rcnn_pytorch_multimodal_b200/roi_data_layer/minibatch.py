"""Device-side BEV rasterisation of a LiDAR sweep: the arithmetic of the reference's
``_get_lidar_blob`` (``lib/roi_data_layer/minibatch.py:428-512``) after file loading and augmentation.

``lidar_bev_map(source_bin, scale)`` takes the augmented point array the reference calls ``source_bin``
(``[Np, 4|5]``: x, y, z, intensity[, elongation]) as a CUDA tensor and returns what the reference appends to
``infos`` / ``processed_frames``: ``info = [0, num_x, 0, num_y, 0, num_z, scale]`` and the map
``[num_y, num_x, NUM_SLICES + NUM_META_CHANNEL]`` (``None`` when no point survives ``filter_points``, as the
reference returns ``None`` for the blob, ``:430-432``).  ``lidar_bev_blob`` stacks frames like
``bev_map_list_to_blob`` (``lib/utils/blob.py:57-70``).  Configuration is read from ``cfg.LIDAR`` /
``cfg.DB_NAME`` exactly as the reference does.  No CPU fallback: CPU tensors raise.
"""
import numpy as np
import torch

from .._lib import check, f32c, lib, ptr, require_cuda, stream_ptr, workspaces, B2DError
from ..model.config import cfg


def _grid(scale):
    L = cfg.LIDAR
    voxel_len = L.VOXEL_LEN / scale                                                    # :434
    num_x = int((L.X_RANGE[1] - L.X_RANGE[0]) * (1 / voxel_len))                       # :435
    num_y = int((L.Y_RANGE[1] - L.Y_RANGE[0]) * (1 / voxel_len))                       # :436
    num_z = int(L.NUM_SLICES)                                                          # :437
    # the voxel generator derives its own grid from the fp32 extents (VoxelGeneratorV2.__init__); the
    # reference indexes a [num_x, num_y, C] map with its coordinates, so the two must agree
    vs = np.array([voxel_len, voxel_len, L.VOXEL_HEIGHT], dtype=np.float32)
    rng = np.array([L.X_RANGE[0], L.Y_RANGE[0], 0, L.X_RANGE[1], L.Y_RANGE[1], L.Z_RANGE[1] - L.Z_RANGE[0]], dtype=np.float32)
    grid = np.round((rng[3:] - rng[:3]) / vs).astype(np.int64)
    if tuple(grid) != (num_x, num_y, num_z):
        raise B2DError(f"voxel grid {tuple(grid)} differs from the BEV map {(num_x, num_y, num_z)}: "
                       "the reference would index out of bounds with this cfg.LIDAR")
    return float(vs[0]), num_x, num_y, num_z


def lidar_bev_map(source_bin, scale=1.0, return_num_voxels=False):
    require_cuda(source_bin)
    pts = f32c(source_bin)
    if pts.dim() != 2 or pts.shape[1] < 4:
        raise B2DError("source_bin must be [Np, >=4] (x, y, z, intensity[, elongation])")
    L = cfg.LIDAR
    voxel_len, nx, ny, nz = _grid(scale)
    info = [0, nx, 0, ny, 0, nz, scale]                                                # :438
    n_meta = int(L.NUM_META_CHANNEL)
    elong = 1 if (n_meta >= 3 and cfg.DB_NAME == 'waymo') else 0                       # :501-505
    if elong and pts.shape[1] < 5:
        raise B2DError("cfg.DB_NAME == 'waymo' reads the elongation column: source_bin needs 5 features")
    n = pts.shape[0]
    dev = pts.device
    out = torch.empty(ny, nx, nz + n_meta, device=dev)
    nvox = torch.empty(1, dtype=torch.int32, device=dev)
    nbytes = lib(dev).b2d_bev_workspace_bytes(max(n, 1), nx, ny, nz)
    ws = workspaces.get(dev, "bev", nbytes)
    check(lib(dev).b2d_bev_rasterize(n, pts.shape[1], ptr(pts), L.X_RANGE[0], L.X_RANGE[1], L.Y_RANGE[0], L.Y_RANGE[1],
                                  L.Z_RANGE[0], L.Z_RANGE[1], voxel_len, L.VOXEL_HEIGHT, nx, ny, nz,
                                  int(L.MAX_PTS_PER_VOXEL), int(L.MAX_NUM_VOXEL), n_meta, elong, ptr(out), ptr(nvox),
                                  ptr(ws), ws.numel(), stream_ptr(dev)), "b2d_bev_rasterize")
    if return_num_voxels:
        return info, out, nvox
    return info, out


def lidar_bev_blob(frames, scale=1.0):
    """`_get_lidar_blob` for a list of point arrays: (infos, blob [F, num_y, num_x, C]); like the reference, a frame
    without any point inside the ranges ends the minibatch: (infos so far, None)  (:429-432)."""
    infos, maps = [], []
    for pts in frames:
        info, m, nv = lidar_bev_map(pts, scale, return_num_voxels=True)
        if int(nv) == 0 and not bool(_any_inside(pts)):
            return infos, None
        infos.append(info)
        maps.append(m)
    return infos, torch.stack(maps)                                                    # utils/blob.py:57-70


def _any_inside(pts):
    L = cfg.LIDAR
    p = pts[:, :3]
    lo = torch.tensor([L.X_RANGE[0], L.Y_RANGE[0], L.Z_RANGE[0]], device=pts.device, dtype=pts.dtype)
    hi = torch.tensor([L.X_RANGE[1], L.Y_RANGE[1], L.Z_RANGE[1]], device=pts.device, dtype=pts.dtype)
    return ((p >= lo) & (p < hi)).all(1).any()
