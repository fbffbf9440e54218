"""Device timings (CUDA events, 20 iterations after 3 warm-ups) of the hot-path pieces that are NOT the
bench.py headline: the other BASELINE.json configs and the train-mode / uncertainty / final-filter rows of
SURVEY.md §8a.  Prints one JSON object; run on a GPU box:  python profiles/secondary_timings.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
from faster_rcnn_pytorch_multimodal_b200 import ops
from faster_rcnn_pytorch_multimodal_b200.layer_utils.anchor_target_layer import anchor_target_layer_torch
from faster_rcnn_pytorch_multimodal_b200.layer_utils.proposal_target_layer import proposal_target_layer
from faster_rcnn_pytorch_multimodal_b200.layer_utils.snippets import generate_anchors_pre
from faster_rcnn_pytorch_multimodal_b200.model.config import cfg
from faster_rcnn_pytorch_multimodal_b200.utils import loss_utils

dev = torch.device("cuda", 0)


def timed(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def inference_stage(name, Hf, Wf, A, C, F, scales, ratios, frame_hw, pre=6000, post=300):
    c = dict(bench.CFG, Hf=Hf, Wf=Wf, A=A, C=C, frame_hw=frame_hw)
    anchors, _ = generate_anchors_pre(Hf, Wf, 16, scales, ratios, 1.0, device=dev)
    prob, deltas, feat, info = bench.synth_frames(c, F, dev, 0)
    pooled = torch.empty(F * post, C, 7, 7, device=dev)

    def step():
        rois, _, _, _, num = ops.proposal_batched(prob, deltas, info, anchors, None, A, pre, post, 0.7, batch_index_stride=1)
        ops._roi_align_forward(feat, rois.view(-1, 5), (7, 7), 1.0 / 16, 2, False, seg_count=num, seg_stride=post, out=pooled)
    ms = timed(step)
    _, _, b_crop = bench.algorithmic_bytes(c, post)
    b_prop = Hf * Wf * A * 20
    return {"config": name, "frames_per_step": F, "ms_per_step": ms, "frames_per_s": F / ms * 1e3,
            "fused_stage_GBps": F * (b_prop + b_crop) / ms / 1e6}


out = {}
out["kitti_image_24x78"] = inference_stage("configs[0] KITTI 375x1242", 24, 78, 25, 1024, 64, bench.SCALES, bench.RATIOS, (375, 1242))
out["bev_lidar_50x44"] = inference_stage("configs[2] BEV 800x700 (A=2, image codec on the AABB anchors)", 50, 44, 2, 1024, 64, [1], [1, 2], (800, 700))

# ---- train mode, Waymo image: targets + RoIAlign forward/backward on the 256 sampled RoIs
Hf, Wf, A, C = 80, 120, 25, 1024
anchors, _ = generate_anchors_pre(Hf, Wf, 16, bench.SCALES, bench.RATIOS, 1.0, device=dev)
g = torch.Generator().manual_seed(5)
G = 32
wh = torch.exp(torch.rand(G, 2, generator=g) * 3.2 + 2.8)
xy = torch.rand(G, 2, generator=g) * torch.tensor([1920.0 - 420, 1280.0 - 420])
gt = torch.cat((xy, xy + wh.clamp(max=400), torch.ones(G, 1)), 1).to(dev)
info = [0, 1920, 0, 1280, 0, 0, 1.0]
cfg.NET_TYPE = "image"
out["anchor_target_waymo_ms"] = timed(lambda: anchor_target_layer_torch(gt, None, info, anchors, A, Hf, Wf, dev))
c = dict(bench.CFG)
prob, deltas, feat, info_t = bench.synth_frames(c, 1, dev, 0)
rois, scores, _, _, num = ops.proposal_batched(prob, deltas, info_t, anchors, None, A, 12000, 2000, 0.7)
n = int(num[0])
r2000, s2000 = rois[0, :n].contiguous(), scores[0, :n].contiguous().view(-1, 1)
a3 = torch.zeros(n, 7, device=dev)
true_gt = torch.zeros(G, 8, device=dev)
cfg.TRAIN.BG_MODE = "intended"
out["proposal_target_waymo_ms"] = timed(lambda: proposal_target_layer(r2000, s2000, a3, gt, true_gt, None, 2, 4))
out["proposal_layer_train_12000_2000_ms"] = timed(lambda: ops.proposal_batched(prob, deltas, info_t, anchors, None, A, 12000, 2000, 0.7))
samp = r2000[:256].contiguous()
fz = feat.clone().requires_grad_(True)
gout = torch.randn(256, C, 7, 7, device=dev)
out["roi_align_fwd_R256_ms"] = timed(lambda: ops._roi_align_forward(feat, samp, (7, 7), 1.0 / 16, 2, False))
out["roi_align_bwd_R256_ms"] = timed(lambda: ops._roi_align_backward(gout, samp, tuple(feat.shape), (7, 7), 1.0 / 16, 2, False))
out["roi_align_fwd_R2000_ms"] = timed(lambda: ops._roi_align_forward(feat, r2000, (7, 7), 1.0 / 16, 2, False))

# ---- FPN: four levels, C = 256, 300 RoIs
from collections import OrderedDict
from faster_rcnn_pytorch_multimodal_b200.utils.torchpoolers import MultiScaleRoIAlign
feats = OrderedDict((f"p{i + 2}", torch.randn(1, 256, 320 >> i, 480 >> i, device=dev)) for i in range(4))
m = MultiScaleRoIAlign(list(feats), 7, 2)
boxes = rois[0, :300, 1:].contiguous()
out["fpn_multiscale_crop_300_ms"] = timed(lambda: m(feats, [boxes], [(1280, 1920)]))

# ---- configs[3]: RPN on FPN level p2 of a Waymo frame (320 x 480 x 25 = 3.84 M anchors at stride 4), 6000 -> 300
c2 = dict(bench.CFG, Hf=320, Wf=480, C=1)
anchors_p2, _ = generate_anchors_pre(320, 480, 4, bench.SCALES, bench.RATIOS, 1.0, device=dev)
prob2, deltas2, _, info2 = bench.synth_frames(c2, 1, dev, 0)
out["proposal_layer_fpn_p2_3p84M_anchors_ms"] = timed(lambda: ops.proposal_batched(prob2, deltas2, info2, anchors_p2, None, 25, 6000, 300, 0.7))
del prob2, deltas2, anchors_p2

# ---- uncertainty: T = 20 MC samples, 300 RoIs, K*E = 14; final filter for 64 frames
mc = torch.randn(20, 300, 14, device=dev)
out["mc_bbox_var_T20_ms"] = timed(lambda: loss_utils.compute_bbox_var(mc))
logits = torch.randn(20, 300, 4, device=dev)
out["mc_mutual_info_T20_ms"] = timed(lambda: loss_utils.categorical_mutual_information(logits))
F, R, K, E = 64, 300, 2, 4
sc = torch.softmax(torch.randn(F, R, K, device=dev) * 2, 2)
ctr = torch.rand(F, R, 1, 2, device=dev) * torch.tensor([1920.0, 1280.0], device=dev)
whb = torch.exp(torch.rand(F, R, K, 2, device=dev) * 2 + 3)
pb = torch.cat((ctr - whb / 2, ctr + whb / 2), 3).reshape(F, R, K * E).contiguous()
inf = torch.tensor([[0, 1920, 0, 1280, 0, 0, 1.0]], device=dev).repeat(F, 1)
out["final_detections_64frames_ms"] = timed(lambda: ops.final_detections(sc, pb, inf, E, "image", 0.1, 0.6, max_dets=100))
print(json.dumps(out, indent=1))
