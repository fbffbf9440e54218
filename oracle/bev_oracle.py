"""CPU restatement of the reference's LiDAR BEV rasterisation (oracle; tests only).  SURVEY.md §8f rank 2.

Path: ``roi_data_layer/minibatch.py:428-512`` (`_get_lidar_blob`, everything after the augmentation):
``filter_points`` (`:232-235`) -> shift z by ``-Z_RANGE[0]`` -> ``spconv.utils.VoxelGeneratorV2.generate``
-> per-voxel max height / density / tanh(mean intensity) / tanh(mean elongation) scattered into a
``[num_x, num_y, NUM_SLICES + NUM_META_CHANNEL]`` map -> transpose to ``[num_y, num_x, C]``.

Third-party arithmetic: the voxeliser is **spconv==1.0** (``req.txt:261``), which is neither vendored under
``/root/reference`` nor installable here.  `points_to_voxel_loop` restates its published algorithm
(``spconv/utils/__init__.py: points_to_voxel`` -> C++ ``points_to_voxel_3d_np`` in
``include/spconv/point2voxel.h``): points are visited in input order; the voxel coordinate is
``floor((p - range_min) / voxel_size)`` in the dtype of the points (float32); a point outside the grid
is skipped; a voxel is created on first visit while fewer than ``max_voxels`` exist (later new voxels
are skipped, the scan continues); a voxel keeps its first ``max_points`` points.  **Parity of that step
is unpinned** (no spconv to run); everything downstream of it is pinned by
``tests/golden/bev.npz``, which `oracle/gen_golden.py` makes by running the reference's own
`_get_lidar_blob` with `spconv.utils.VoxelGeneratorV2` bound to this restatement.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Tuple

import numpy as np


@dataclass
class LidarCfg:
    """model/config.py:397-406 (defaults of the reference)."""
    x_range: Tuple[float, float] = (0, 70)
    y_range: Tuple[float, float] = (-40, 40)
    z_range: Tuple[float, float] = (-3, 3)
    voxel_len: float = 0.1
    voxel_height: float = 0.5
    num_slices: int = 12
    num_meta_channel: int = 3
    max_pts_per_voxel: int = 32
    max_num_voxel: int = 25000
    db_name: str = "waymo"                       # elongation channel only for waymo (minibatch.py:494-497)

    @property
    def num_channel(self) -> int:
        return self.num_slices + self.num_meta_channel

    def pc_extents(self) -> np.ndarray:
        return np.array([self.x_range[0], self.y_range[0], self.z_range[0],
                         self.x_range[1], self.y_range[1], self.z_range[1]], dtype=np.float64)


def filter_points(pc: np.ndarray, c: LidarCfg) -> np.ndarray:
    """minibatch.py:232-235."""
    lo = pc[(pc[:, 0] >= c.x_range[0]) & (pc[:, 1] >= c.y_range[0]) & (pc[:, 2] >= c.z_range[0])]
    return lo[(lo[:, 0] < c.x_range[1]) & (lo[:, 1] < c.y_range[1]) & (lo[:, 2] < c.z_range[1])]


def _grid(voxel_size, coors_range):
    voxel_size = np.asarray(voxel_size, dtype=np.float32)
    coors_range = np.asarray(coors_range, dtype=np.float32)
    grid = np.round((coors_range[3:] - coors_range[:3]) / voxel_size).astype(np.int64)   # VoxelGeneratorV2.__init__
    return voxel_size, coors_range, grid


def points_to_voxel_loop(points, voxel_size, coors_range, max_points, max_voxels):
    """spconv 1.0 ``points_to_voxel_3d_np`` restated literally (one Python iteration per point)."""
    voxel_size, coors_range, grid = _grid(voxel_size, coors_range)
    points = np.ascontiguousarray(points, dtype=np.float32)
    nfeat = points.shape[1]
    voxels = np.zeros((max_voxels, max_points, nfeat), dtype=np.float32)
    coors = np.zeros((max_voxels, 3), dtype=np.int32)
    num = np.zeros((max_voxels,), dtype=np.int32)
    lut = {}
    voxel_num = 0
    for i in range(points.shape[0]):
        coor = [0, 0, 0]
        failed = False
        for j in range(3):
            c = int(np.floor((points[i, j] - coors_range[j]) / voxel_size[j]))      # float32 arithmetic
            if c < 0 or c >= grid[j]:
                failed = True
                break
            coor[2 - j] = c                                                          # stored zyx
        if failed:
            continue
        key = tuple(coor)
        v = lut.get(key, -1)
        if v == -1:
            if voxel_num >= max_voxels:
                continue
            v = voxel_num
            voxel_num += 1
            lut[key] = v
            coors[v] = coor
        if num[v] < max_points:
            voxels[v, num[v]] = points[i]
            num[v] += 1
    return voxels[:voxel_num], coors[:voxel_num], num[:voxel_num]


def points_to_voxel(points, voxel_size, coors_range, max_points, max_voxels):
    """Vectorised equivalent of `points_to_voxel_loop` (same outputs, checked in tests/test_oracle.py)."""
    voxel_size, coors_range, grid = _grid(voxel_size, coors_range)
    points = np.ascontiguousarray(points, dtype=np.float32)
    n, nfeat = points.shape
    c = np.floor((points[:, :3] - coors_range[:3]) / voxel_size).astype(np.int64)     # float32 arithmetic
    ok = np.all((c >= 0) & (c < grid), axis=1)
    idx = np.nonzero(ok)[0]
    c = c[idx]
    key = (c[:, 0] * grid[1] + c[:, 1]) * grid[2] + c[:, 2]
    uniq, first, inv = np.unique(key, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")                 # voxels in order of first appearance
    rank_of_uniq = np.empty_like(order)
    rank_of_uniq[order] = np.arange(order.size)
    vox = rank_of_uniq[inv]                                   # voxel index of every valid point
    keep = vox < max_voxels
    idx, vox, c = idx[keep], vox[keep], c[keep]
    # position of a point inside its voxel = number of earlier points of the same voxel
    o = np.argsort(vox, kind="stable")
    sv = vox[o]
    start = np.r_[0, np.nonzero(sv[1:] != sv[:-1])[0] + 1]
    pos_sorted = np.arange(sv.size) - np.repeat(start, np.diff(np.r_[start, sv.size]))
    pos = np.empty_like(pos_sorted)
    pos[o] = pos_sorted
    sel = pos < max_points
    nv = int(min(order.size, max_voxels))
    voxels = np.zeros((nv, max_points, nfeat), dtype=np.float32)
    voxels[vox[sel], pos[sel]] = points[idx[sel]]
    num = np.bincount(vox[sel], minlength=nv).astype(np.int32)
    coors = np.zeros((nv, 3), dtype=np.int32)
    coors[vox, 0], coors[vox, 1], coors[vox, 2] = c[:, 2], c[:, 1], c[:, 0]           # zyx
    return voxels, coors, num


class VoxelGeneratorV2:
    """Interface of ``spconv.utils.VoxelGeneratorV2`` as minibatch.py:445-453 uses it."""

    def __init__(self, voxel_size, point_cloud_range, max_num_points, max_voxels=20000, loop=False, **_):
        self._voxel_size = np.array(voxel_size, dtype=np.float32)
        self._range = np.array(point_cloud_range, dtype=np.float32)
        self._max_points, self._max_voxels, self._loop = max_num_points, max_voxels, loop

    def generate(self, points, max_voxels=None):
        fn = points_to_voxel_loop if self._loop else points_to_voxel
        voxels, coors, num = fn(points, self._voxel_size, self._range, self._max_points, max_voxels or self._max_voxels)
        return {"voxels": voxels, "coordinates": coors, "num_points_per_voxel": num}


def lidar_bev_map(source_bin: np.ndarray, scale: float = 1.0, c: LidarCfg = None, loop: bool = False):
    """minibatch.py:428-507 for one frame.  `source_bin` [Np, >=4(5)] float32 after augmentation.

    Returns (info[7], bev_map [num_y, num_x, C] float32) or (info, None) when no point survives the filter."""
    c = c or LidarCfg()
    source_bin = filter_points(np.array(source_bin, dtype=np.float32, copy=True), c)          # :428
    voxel_len = c.voxel_len / scale                                                           # :434
    num_x = int((c.x_range[1] - c.x_range[0]) * (1 / voxel_len))                              # :435
    num_y = int((c.y_range[1] - c.y_range[0]) * (1 / voxel_len))                              # :436
    num_z = int(c.num_slices)                                                                 # :437
    info = [0, num_x, 0, num_y, 0, num_z, scale]                                              # :438
    if source_bin.shape[0] <= 0:                                                              # :430-432
        return info, None
    ext = c.pc_extents()
    ext[5] -= ext[2]                                                                          # :442
    ext[2] = 0                                                                                # :443
    gen = VoxelGeneratorV2([voxel_len, voxel_len, c.voxel_height], ext, c.max_pts_per_voxel, c.max_num_voxel, loop=loop)
    source_bin[:, 2] -= c.z_range[0]                                                          # :454
    res = gen.generate(source_bin)
    voxels, coords, npv = res["voxels"], res["coordinates"].copy(), res["num_points_per_voxel"]
    bev = np.zeros((num_x, num_y, c.num_channel), dtype=np.float32)                           # :460
    coords[:, [2, 1, 0]] = coords[:, [0, 1, 2]]                                               # :462  zyx -> xyz
    xy = coords[:, 0:2]
    vmax = np.amax(voxels[:, :, 2], axis=1) - coords[:, 2] * c.voxel_height                   # :468
    bev[tuple(zip(*coords))] = vmax                                                           # :478-479
    S = c.num_slices
    if c.num_meta_channel >= 1:                                                               # :482-489
        dens = npv / c.max_pts_per_voxel
        bev[tuple(zip(*np.hstack((xy, np.full((xy.shape[0], 1), S)))))] = dens
    if c.num_meta_channel >= 2:                                                               # :491-498
        inten = np.sum(voxels[:, :, 3], axis=1) / npv
        bev[tuple(zip(*np.hstack((xy, np.full((xy.shape[0], 1), S + 1)))))] = np.tanh(inten)
    if c.num_meta_channel >= 3:                                                               # :500-509
        if c.db_name == "waymo":
            elong = np.sum(voxels[:, :, 4], axis=1) / npv
        else:
            elong = np.zeros((voxels.shape[0]))
        bev[tuple(zip(*np.hstack((xy, np.full((xy.shape[0], 1), S + 2)))))] = np.tanh(elong)
    return info, np.transpose(bev, axes=[1, 0, 2])                                            # :512


def synth_point_cloud(seed: int, n: int, c: LidarCfg = None, nfeat: int = 5) -> np.ndarray:
    """Lidar-like synthetic sweep: dense near the sensor (1/r), a ground plane plus vertical clutter,
    a few percent of points outside the ranges, intensity / elongation in [0, 1.5)."""
    c = c or LidarCfg()
    g = np.random.default_rng(seed)
    xr, yr, zr = c.x_range, c.y_range, c.z_range
    rmax = float(np.hypot(xr[1] - xr[0], max(abs(yr[0]), abs(yr[1]))))
    r = rmax * g.random(n) ** 2 + 0.5                     # quadratic: many near returns
    th = (g.random(n) - 0.5) * np.pi
    x = xr[0] + r * np.cos(th)
    y = r * np.sin(th)
    ground = g.random(n) < 0.6
    z = np.where(ground, zr[0] + 1.2 + 0.05 * g.standard_normal(n), zr[0] + (zr[1] - zr[0]) * g.random(n) * 1.05)
    pts = np.stack([x, y, z] + [1.5 * g.random(n) for _ in range(nfeat - 3)], axis=1)
    return pts.astype(np.float32)
