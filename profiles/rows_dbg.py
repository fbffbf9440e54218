"""Smallest rows-kernel launch (debug aid): one Waymo-size frame, 64 channels, 20 RoIs, checked against the gather kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from faster_rcnn_pytorch_multimodal_b200 import ops
dev = torch.device('cuda', 0)
g = torch.Generator().manual_seed(1)
H, W, C, R = 80, 120, 64, 20
feat = torch.randn(1, C, H, W, generator=g).to(dev)
xy = torch.rand(R, 2, generator=g) * torch.tensor([W * 16 - 200.0, H * 16 - 200.0])
wh = torch.rand(R, 2, generator=g) * 180 + 8
rois = torch.cat((torch.zeros(R, 1), xy, xy + wh), 1).to(dev)
out = ops._roi_align_forward(feat, rois, (7, 7), 1.0 / 16, 2, False)
torch.cuda.synchronize()
os.environ["B2D_ROWS_COOP_FILL"] = "1"
ref = ops._roi_align_forward(feat, rois, (7, 7), 1.0 / 16, 2, False)
torch.cuda.synchronize()
print("max abs diff vs cp.async fill:", (out - ref).abs().max().item(), "nonzero", (out != 0).float().mean().item())
