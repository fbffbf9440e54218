"""Launches of one anchor_target_layer_torch and one proposal_target_layer call (Waymo train frame) for an ncu launch list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from faster_rcnn_pytorch_multimodal_b200 import ops
from faster_rcnn_pytorch_multimodal_b200.layer_utils.anchor_target_layer import anchor_target_layer_torch
from faster_rcnn_pytorch_multimodal_b200.layer_utils.proposal_target_layer import proposal_target_layer
from faster_rcnn_pytorch_multimodal_b200.layer_utils.snippets import generate_anchors_pre
from faster_rcnn_pytorch_multimodal_b200.model.config import cfg
dev = torch.device("cuda", 0)
Hf, Wf, A = 80, 120, 25
anchors, _ = generate_anchors_pre(Hf, Wf, 16, bench.SCALES, bench.RATIOS, 1.0, device=dev)
g = torch.Generator().manual_seed(5)
G = 32
wh = torch.exp(torch.rand(G, 2, generator=g) * 3.2 + 2.8)
xy = torch.rand(G, 2, generator=g) * torch.tensor([1920.0 - 420, 1280.0 - 420])
gt = torch.cat((xy, xy + wh.clamp(max=400), torch.ones(G, 1)), 1).to(dev)
info = [0, 1920, 0, 1280, 0, 0, 1.0]
cfg.NET_TYPE = "image"
prob, deltas, feat, info_t = bench.synth_frames(bench.CFG, 1, dev, 0)
rois, scores, _, _, num = ops.proposal_batched(prob, deltas, info_t, anchors, None, A, 12000, 2000, 0.7)
n = int(num[0])
r2000, s2000 = rois[0, :n].contiguous(), scores[0, :n].contiguous().view(-1, 1)
a3 = torch.zeros(n, 7, device=dev)
true_gt = torch.zeros(G, 8, device=dev)
cfg.TRAIN.BG_MODE = "intended"
import time
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    anchor_target_layer_torch(gt, None, info, anchors, A, Hf, Wf, dev)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    proposal_target_layer(r2000, s2000, a3, gt, true_gt, None, 2, 4)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"wall: anchor_target {1e3*(t1-t0):.3f} ms, proposal_target {1e3*(t2-t1):.3f} ms")
