"""Launches of (a) one train-mode proposal_layer call (12000 -> 2000) and (b) one 4-level FPN crop of 300 RoIs,
for an ncu launch list:  ncu --metrics gpu__time_duration.sum -k regex:... python profiles/train_fpn_launches.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from collections import OrderedDict
import torch, bench
from faster_rcnn_pytorch_multimodal_b200 import ops
from faster_rcnn_pytorch_multimodal_b200.layer_utils.snippets import generate_anchors_pre
from faster_rcnn_pytorch_multimodal_b200.utils.torchpoolers import MultiScaleRoIAlign
dev = torch.device("cuda", 0)
cfg = bench.CFG
anchors, _ = generate_anchors_pre(cfg["Hf"], cfg["Wf"], 16, bench.SCALES, bench.RATIOS, 1.0, device=dev)
prob, deltas, feat, info = bench.synth_frames(cfg, 1, dev, 0)
for _ in range(3):
    rois, scores, _, _, num = ops.proposal_batched(prob, deltas, info, anchors, None, cfg["A"], 12000, 2000, 0.7)
feats = OrderedDict((f"p{i + 2}", torch.randn(1, 256, 320 >> i, 480 >> i, device=dev)) for i in range(4))
m = MultiScaleRoIAlign(list(feats), 7, 2)
boxes = rois[0, :300, 1:].contiguous()
for _ in range(3):
    out = m(feats, [boxes], [(1280, 1920)])
torch.cuda.synchronize()
