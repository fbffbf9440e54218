"""LiDAR BEV rasterisation: oracle vs the reference's own `_get_lidar_blob` output
(tests/golden/bev.npz, made by oracle/gen_golden.py with spconv bound to the restated voxeliser), and the CUDA
path vs the oracle.  Bars: which points / voxels are kept and which voxel wins a column - exact (a wrong choice
changes values by far more than the tolerance); map values 1e-5 relative (fp32 sums, tanh)."""
import numpy as np
import pytest
import torch

from oracle import bev_oracle as B


def _case(g, name):
    x0, x1, y0, y1, mp, mv, nmeta, waymo = g[f"{name}_cfg"]
    c = B.LidarCfg(x_range=(x0, x1), y_range=(y0, y1), max_pts_per_voxel=int(mp), max_num_voxel=int(mv),
                   num_meta_channel=int(nmeta), db_name="waymo" if waymo else "nuscenes")
    want = np.zeros(int(np.prod(g[f"{name}_shape"])), dtype=np.float32)
    want[g[f"{name}_nz_idx"]] = g[f"{name}_nz_val"]
    return c, g[f"{name}_points"], g[f"{name}_info"], want.reshape(g[f"{name}_shape"])


CASES = ["caps", "roomy", "kitti4"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_output(golden, name):
    c, pts, info, want = _case(golden("bev"), name)
    for loop in (False, True):
        got_info, got = B.lidar_bev_map(pts, 1.0, c, loop=loop)
        assert np.array_equal(np.asarray(got_info, dtype=np.float64), info)
        assert np.array_equal(got, want)


def test_voxeliser_vectorised_equals_literal_loop():
    c = B.LidarCfg(x_range=(0, 6), y_range=(-3, 3), max_pts_per_voxel=3, max_num_voxel=150)
    ext = c.pc_extents()
    ext[5] -= ext[2]
    ext[2] = 0
    for seed in range(4):
        p = B.filter_points(B.synth_point_cloud(seed, 4000, c), c)
        p[:, 2] -= c.z_range[0]
        a = B.points_to_voxel_loop(p, [0.1, 0.1, 0.5], ext, 3, 150)
        b = B.points_to_voxel(p, [0.1, 0.1, 0.5], ext, 3, 150)
        assert a[0].shape[0] == 150 and int(a[2].max()) == 3           # both caps bind
        for x, y in zip(a, b):
            assert np.array_equal(x, y)


def test_empty_and_outside_points():
    c = B.LidarCfg(x_range=(0, 4), y_range=(-2, 2))
    far = np.full((10, 5), 100.0, dtype=np.float32)
    info, m = B.lidar_bev_map(far, 1.0, c)
    assert m is None and info[:6] == [0, 40, 0, 40, 0, 12]


# ------------------------------------------------------------------------------------------ GPU
def _set_cfg(c):
    from faster_rcnn_pytorch_multimodal_b200.model.config import cfg
    saved = {k: cfg.LIDAR[k] for k in ("X_RANGE", "Y_RANGE", "MAX_PTS_PER_VOXEL", "MAX_NUM_VOXEL", "NUM_META_CHANNEL")}
    saved_db = cfg.DB_NAME
    cfg.LIDAR.X_RANGE, cfg.LIDAR.Y_RANGE = list(c.x_range), list(c.y_range)
    cfg.LIDAR.MAX_PTS_PER_VOXEL, cfg.LIDAR.MAX_NUM_VOXEL = c.max_pts_per_voxel, c.max_num_voxel
    cfg.LIDAR.NUM_META_CHANNEL, cfg.DB_NAME = c.num_meta_channel, c.db_name

    def restore():
        for k, v in saved.items():
            cfg.LIDAR[k] = v
        cfg.DB_NAME = saved_db
    return restore


def _gpu_map(pts, c, scale=1.0):
    from faster_rcnn_pytorch_multimodal_b200.roi_data_layer.minibatch import lidar_bev_map
    restore = _set_cfg(c)
    try:
        info, m, nv = lidar_bev_map(torch.from_numpy(pts).cuda(), scale, return_num_voxels=True)
        return info, m.cpu().numpy(), int(nv)
    finally:
        restore()


def _assert_map(got, want):
    # same support (a wrong voxel / point / column-winner choice shows up here or as a large error) ...
    assert np.array_equal(got != 0, want != 0)
    # ... and values to 1e-5 relative
    assert np.allclose(got, want, rtol=1e-5, atol=1e-6), float(np.abs(got - want).max())


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_matches_reference_output(golden, name):
    c, pts, info, want = _case(golden("bev"), name)
    got_info, got, nv = _gpu_map(pts, c)
    assert np.array_equal(np.asarray(got_info, dtype=np.float64), info)
    _assert_map(got, want)
    assert nv == B.points_to_voxel(*_voxel_args(pts, c))[0].shape[0]


def _voxel_args(pts, c):
    ext = c.pc_extents()
    ext[5] -= ext[2]
    ext[2] = 0
    p = B.filter_points(np.array(pts, dtype=np.float32), c)
    p[:, 2] -= c.z_range[0]
    return p, [c.voxel_len, c.voxel_len, c.voxel_height], ext, c.max_pts_per_voxel, c.max_num_voxel


@pytest.mark.gpu
@pytest.mark.parametrize("n,seed", [(180000, 7), (40000, 8)])
def test_gpu_full_size_waymo_grid(n, seed):
    """cfg.LIDAR defaults: 700 x 800 x 12 grid, 32 points / voxel, 25 000 voxels (the cap binds at 180 k points)."""
    c = B.LidarCfg()
    pts = B.synth_point_cloud(seed, n, c)
    _, want = B.lidar_bev_map(pts, 1.0, c)
    info, got, nv = _gpu_map(pts, c)
    assert info == [0, 700, 0, 800, 0, 12, 1.0] and got.shape == (800, 700, 15)
    _assert_map(got, want)
    assert nv == min(25000, nv) and (n < 100000 or nv == 25000)


@pytest.mark.gpu
def test_gpu_dense_voxels_and_scale():
    """Hundreds of points in a few voxels (rank-by-counting path), duplicates of one point, and scale 2 (5 cm voxels)."""
    c = B.LidarCfg(x_range=(0, 4), y_range=(-2, 2), max_pts_per_voxel=8, max_num_voxel=500)
    g = np.random.default_rng(5)
    blob = np.concatenate([np.tile(np.array([[1.03, 0.52, -1.9, 0.7, 0.2]], np.float32), (300, 1)),
                           (np.array([[2.0, -1.0, 0.0, 0, 0]]) + g.random((900, 5)) * [0.25, 0.25, 1.0, 1.5, 1.5]).astype(np.float32),
                           B.synth_point_cloud(6, 3000, c)])
    blob = blob[g.permutation(blob.shape[0])]
    for scale in (1.0, 2.0):
        _, want = B.lidar_bev_map(blob, scale, c)
        _, got, _ = _gpu_map(blob, c, scale)
        _assert_map(got, want)


@pytest.mark.gpu
def test_gpu_degenerate_cloud_all_points_in_one_voxel():
    """Zero-padded sweeps put tens of thousands of points into one voxel: must stay exact and fast."""
    import time
    c = B.LidarCfg()
    pts = np.zeros((60000, 5), dtype=np.float32)
    pts[:, 3] = np.linspace(0, 1, 60000, dtype=np.float32)
    pts[::1000, :3] = [[10.0, 3.0, -1.0]]
    _, want = B.lidar_bev_map(pts, 1.0, c)
    _gpu_map(pts, c)                                   # warm-up (workspace allocation)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _, got, nv = _gpu_map(pts, c)
    assert time.perf_counter() - t0 < 0.5
    assert nv == 2
    _assert_map(got, want)


@pytest.mark.gpu
def test_gpu_no_points_inside():
    c = B.LidarCfg(x_range=(0, 4), y_range=(-2, 2))
    far = np.full((10, 5), 100.0, dtype=np.float32)
    info, got, nv = _gpu_map(far, c)
    assert nv == 0 and not got.any()
    from faster_rcnn_pytorch_multimodal_b200.roi_data_layer.minibatch import lidar_bev_blob
    restore = _set_cfg(c)
    try:
        infos, blob = lidar_bev_blob([torch.from_numpy(far).cuda()])
        assert blob is None and infos == []
    finally:
        restore()
