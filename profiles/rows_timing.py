"""Cycle breakdown of rows::fwd_kernel (consumer waits / run_item / producer): build the library with -DB2D_ROWS_TIMING first
(see csrc/roi_align_rows.cu), then run this on a GPU box."""
import ctypes, sys, subprocess, os, json
sys.path.insert(0, '/root/repo')
import torch, bench
from faster_rcnn_pytorch_multimodal_b200 import _lib, ops
from faster_rcnn_pytorch_multimodal_b200.layer_utils.snippets import generate_anchors_pre
L = _lib.lib()
dev = torch.device('cuda', 0)
cfg = bench.CFG; F = 16
anchors, _ = generate_anchors_pre(cfg["Hf"], cfg["Wf"], cfg["stride"], bench.SCALES, bench.RATIOS, 1.0, device=dev)
prob, deltas, feat, info = bench.synth_frames(cfg, F, dev, 0)
rois, scores, _, _, num = ops.proposal_batched(prob, deltas, info, anchors, None, cfg["A"], cfg["pre_nms"], cfg["post_nms"], cfg["nms_thresh"], batch_index_stride=1)
pooled = torch.empty(F * 300, 1024, 7, 7, device=dev)
out = (ctypes.c_ulonglong * 16)()
for it in range(3):
    L.b2d_rows_debug(out, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops._roi_align_forward(feat, rois.view(-1, 5), (7, 7), 1.0 / 16, 2, False, seg_count=num, seg_stride=300, out=pooled)
    e1.record(); torch.cuda.synchronize()
    L.b2d_rows_debug(out, 0)
    v = list(out)
    ncons = 8 * 32 * F
    print(f"ms {e0.elapsed_time(e1):.3f} per consumer-warp avg cycles: total {v[0]/ncons:.0f} wait_full {v[1]/ncons:.0f} wait_done {v[2]/ncons:.0f} run_item {v[3]/ncons:.0f} fetch(prod) {v[4]/(64*F):.0f} ri_setup {v[8]/ncons:.0f} ri_rows {v[9]/ncons:.0f} ri_out {v[10]/ncons:.0f} | producer: total {v[7]/(64*F):.0f} wait_done {v[5]/(64*F):.0f} wait_stg {v[6]/(64*F):.0f}")
