"""A/B timing of the RoIAlign forward launch on the bench workload (device time, CUDA events).

  python profiles/rows_ab.py [frames] [reps] lib1.so lib2.so ...    # each library in its own process
  python profiles/rows_ab.py --one frames reps                       # (internal) time the library in B2D_LIB_PATH

A/B libraries are built with `python -m faster_rcnn_pytorch_multimodal_b200.build -DB2D_...=1 -o_ab/x.so`.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(F, reps):
    import torch
    import bench
    from faster_rcnn_pytorch_multimodal_b200 import ops
    dev = torch.device('cuda', 0)
    cfg = bench.WORKLOADS[os.environ.get("AB_WORKLOAD", "waymo_test")]
    anchors, a3d = bench.anchors_for(cfg, dev)
    prob, deltas, feat, info = bench.synth_frames(cfg, F, dev, 0)
    M = cfg["post_nms"]
    rois, scores, _, _, num = ops.proposal_batched(prob, deltas, info, anchors, None, cfg["A"], cfg["pre_nms"], M,
                                                   cfg["nms_thresh"], batch_index_stride=1)
    pooled = torch.empty(F * M, cfg["C"], 7, 7, device=dev)
    r = rois.view(-1, 5)
    ts = []
    for it in range(reps + 3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops._roi_align_forward(feat, r, (7, 7), 1.0 / 16, 2, False, seg_count=num, seg_stride=M, out=pooled)
        e1.record()
        torch.cuda.synchronize()
        if it >= 3:
            ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"{os.path.basename(os.environ.get('B2D_LIB_PATH', 'default'))}: median {ts[len(ts) // 2] * 1e3 / F:.2f} us/frame  "
          f"min {ts[0] * 1e3 / F:.2f}  (F={F})", flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "--one":
        one(int(sys.argv[2]), int(sys.argv[3]))
    else:
        F, reps = sys.argv[1], sys.argv[2]
        for lib in sys.argv[3:]:
            env = dict(os.environ)
            if lib != "default":
                env["B2D_LIB_PATH"] = os.path.abspath(lib)
            subprocess.run([sys.executable, os.path.abspath(__file__), "--one", F, reps], env=env)
