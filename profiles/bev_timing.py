"""Device time of the BEV rasterisation (CUDA events) at the reference's cfg.LIDAR sizes, with the CPU oracle
(numpy restatement of minibatch.py:428-512, vectorised voxeliser) timed beside it."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import bev_oracle as B
from faster_rcnn_pytorch_multimodal_b200.model.config import cfg
from faster_rcnn_pytorch_multimodal_b200.roi_data_layer.minibatch import lidar_bev_map

cfg.DB_NAME = "waymo"
out = {}
for n in (60000, 180000):
    pts = B.synth_point_cloud(3, n)
    d = torch.from_numpy(pts).cuda()
    for _ in range(50):          # also lets the clocks ramp: the first timed loop of a fresh process is 10x slow otherwise
        lidar_bev_map(d)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        info, m, nv = lidar_bev_map(d, return_num_voxels=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    t0 = time.perf_counter()
    for _ in range(3):
        B.lidar_bev_map(pts)
    cpu_ms = (time.perf_counter() - t0) / 3 * 1e3
    alg = pts.nbytes + m.numel() * 4
    out[f"points_{n}"] = {"gpu_ms": ms, "cpu_oracle_ms": cpu_ms, "voxels_kept": int(nv), "algorithmic_MB": alg / 1e6,
                          "GBps": alg / ms / 1e6}
print(json.dumps(out, indent=1))
