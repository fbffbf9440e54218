#!/usr/bin/env python
"""Benchmark of the hot path: proposal + NMS + RoI-crop frames/s (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU torch path

Workload = BASELINE.json configs[1], inference stage: synthetic Waymo camera frames
(1280x1920 -> 80x120 stride-16 grid, A=25, N=240 000 anchors, 6000 -> 300 proposals, res101 C4
feature map C=1024, 7x7 RoIAlign with sampling_ratio 2).  A "step" is one pass of the hot path
over one batch of `--frames` independent frames per GPU; frames are sharded one stream per GPU
with no data-path collective (weak scaling); the only NCCL call is the end-of-stream gather of
the detection records.  Inputs live in HBM before the timed region (`value`); `e2e` repeats the
measurement through the host-buffer C-ABI entry (H2D + kernels + D2H inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

SCALES, RATIOS = [2, 4, 8, 16, 32], [0.5, 0.75, 1, 1.25, 2]
CFG = dict(name="waymo_image_test_1280x1920_res101_c4", frame_hw=(1280, 1920), Hf=80, Wf=120, A=25, C=1024,
           pre_nms=6000, post_nms=300, nms_thresh=0.7, pooled=7, sampling_ratio=2, stride=16)
METRIC = "proposal+NMS+RoI-crop frames/s"


def algorithmic_bytes(cfg, rois_per_frame):
    """per-frame algorithmic bytes of the fused inference stage (fp32)."""
    N = cfg["Hf"] * cfg["Wf"] * cfg["A"]
    b_prop = N * 4 + N * 16
    b_nms = 2 * cfg["pre_nms"] * 20 + cfg["post_nms"] * 24
    b_crop = cfg["C"] * cfg["Hf"] * cfg["Wf"] * 4 + rois_per_frame * cfg["C"] * cfg["pooled"] ** 2 * 4
    return b_prop, b_nms, b_crop


def synth_frames(cfg, F, device, first_frame):
    """SURVEY.md §8d inputs, generator seeded 3 + frame index (model/config.py:346)."""
    Hf, Wf, A, C = cfg["Hf"], cfg["Wf"], cfg["A"], cfg["C"]
    prob = torch.empty(F, Hf, Wf, 2 * A, device=device)
    deltas = torch.empty(F, Hf, Wf, 4 * A, device=device)
    feat = torch.empty(F, C, Hf, Wf, device=device)
    for i in range(F):
        g = torch.Generator(device=device).manual_seed(3 + first_frame + i)
        logits = torch.randn(Hf, Wf, 2 * A, generator=g, device=device)
        pair = torch.stack((logits[..., :A], logits[..., A:]), -1).softmax(-1)
        prob[i] = torch.cat((pair[..., 0], pair[..., 1]), -1)
        d = torch.randn(Hf, Wf, A, 4, generator=g, device=device)
        d[..., :2] *= 0.1
        d[..., 2:] *= 0.2
        deltas[i] = d.reshape(Hf, Wf, 4 * A)
        feat[i] = torch.randn(C, Hf, Wf, generator=g, device=device)
    H, W = cfg["frame_hw"]
    info = torch.tensor([[0, W, 0, H, 0, 0, 1.0]], device=device).repeat(F, 1)
    return prob, deltas, feat, info


class ClockSampler:
    """nvidia-smi clocks + throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t0=None, t1=None):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for t, r in self.rows if t0 is None or (t0 <= t <= t1)]
        window = "timed region"
        if not rows:            # region shorter than nvidia-smi's sampling period: use the whole loaded run
            rows, window = [r for _, r in self.rows], "whole run (timed region shorter than one sample)"
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for n, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def cpu_reference_frames_per_s(cfg, n_frames, warmup, threads):
    """The reference's CPU torch path (oracle port): proposal_layer -> torchvision roi_align."""
    from oracle import glue_oracle as O
    torch.set_num_threads(threads)
    anchors = torch.from_numpy(O.generate_anchors_pre(cfg["Hf"], cfg["Wf"], cfg["stride"], SCALES, RATIOS, 1.0)[0])
    a3 = torch.zeros(anchors.shape[0], 7)
    ocfg = O.GlueCfg(test_pre_nms=cfg["pre_nms"], test_post_nms=cfg["post_nms"], test_nms_thresh=cfg["nms_thresh"])
    times = []
    for i in range(warmup + n_frames):
        prob, deltas, feat, info = synth_frames(cfg, 1, torch.device("cpu"), i)
        t0 = time.perf_counter()
        blob, _, _ = O.proposal_layer(prob, deltas, info[0].numpy(), "TEST", anchors, a3, cfg["A"], cfg=ocfg,
                                      stable_sort=False)
        O.roi_align(feat, blob, (cfg["pooled"],) * 2, 1.0 / cfg["stride"], cfg["sampling_ratio"], False)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return len(times) / sum(times), float(np.median(times))


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    per_step = max(1, args.ref_frames_per_step)
    fps, med = cpu_reference_frames_per_s(CFG, args.steps * per_step, args.warmup, threads)
    sample = f"{args.steps * per_step} frames of {CFG['name']} after {args.warmup} warm-up frames, {threads} torch threads"
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * per_step / fps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": CFG["name"], "frames_per_step": per_step, "pre_nms": CFG["pre_nms"],
                       "post_nms": CFG["post_nms"], "channels": CFG["C"], "device": "host CPU"},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample,
                             "median_ms_per_frame": 1e3 * med},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=128, help="independent frames per step per GPU")
    ap.add_argument("--e2e-frames", type=int, default=16, help="frames per host-buffer call")
    ap.add_argument("--ref-frames-per-step", type=int, default=1)
    ap.add_argument("--cpu-baseline-frames", type=int, default=12)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    from faster_rcnn_pytorch_multimodal_b200 import _lib, ops
    from faster_rcnn_pytorch_multimodal_b200.layer_utils.snippets import generate_anchors_pre
    L = _lib.lib()
    cfg, F = CFG, args.frames
    A, P = cfg["A"], cfg["pooled"]
    n_loc = cfg["Hf"] * cfg["Wf"]
    anchors, _ = generate_anchors_pre(cfg["Hf"], cfg["Wf"], cfg["stride"], SCALES, RATIOS, 1.0, device=dev)
    prob, deltas, feat, info = synth_frames(cfg, F, dev, first_frame=rank * F)
    M = cfg["post_nms"]
    pooled = torch.empty(F * M, cfg["C"], P, P, device=dev)

    def step(timed_events=None):
        rois, scores, _, _, num = ops.proposal_batched(prob, deltas, info, anchors, None, A, cfg["pre_nms"],
                                                       cfg["post_nms"], cfg["nms_thresh"], batch_index_stride=1)
        if timed_events is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        ops._roi_align_forward(feat, rois.view(-1, 5), (P, P), 1.0 / cfg["stride"], cfg["sampling_ratio"], False,
                               seg_count=num, seg_stride=M, out=pooled)
        if timed_events is not None:
            e1.record()
            timed_events.append((e0, e1))
        return rois, scores, num

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        rois, scores, num = step()
    if dist is not None:          # warm the communicator: the first NCCL collective builds its channels
        rec = torch.cat((rois.view(F, -1), scores, num.view(F, 1).float()), dim=1)
        dist.all_gather([torch.empty_like(rec) for _ in range(world)], rec)
    barrier()
    crop_events = []
    wall0 = time.perf_counter()
    launches0 = L.b2d_launch_count()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for _ in range(args.steps):
        rois, scores, num = step(crop_events)
    if dist is not None:
        # end-of-stream gather of the detection records (the path's only collective)
        rec = torch.cat((rois.view(F, -1), scores, num.view(F, 1).float()), dim=1)
        gathered = [torch.empty_like(rec) for _ in range(world)]
        dist.all_gather(gathered, rec)
    t_end.record()
    barrier()
    wall1 = time.perf_counter()
    launches = L.b2d_launch_count() - launches0
    elapsed_ms = t_start.elapsed_time(t_end)
    crop_ms = float(np.mean([a.elapsed_time(b) for a, b in crop_events]))
    if dist is not None:
        t = torch.tensor([elapsed_ms, crop_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms, crop_ms = float(t[0]), float(t[1])
    n_rois = int(num.sum().item())
    value = world * F * args.steps / (elapsed_ms * 1e-3)

    # ---- e2e: host buffers through the C ABI (H2D + kernels + D2H inside the timed region)
    Fe = max(1, min(args.e2e_frames, F))
    h = lambda t: t.cpu().pin_memory()
    hp, hd, hf, hi = h(prob[:Fe]), h(deltas[:Fe]), h(feat[:Fe]), h(info[:Fe])
    o_rois = torch.empty(Fe, M, 5).pin_memory()
    o_sc = torch.empty(Fe, M).pin_memory()
    o_num = torch.empty(Fe, dtype=torch.int32).pin_memory()
    o_pool = torch.empty(Fe * M, cfg["C"], P, P).pin_memory()
    ws_bytes = L.b2d_pipeline_device_bytes(Fe, n_loc, A, cfg["C"], cfg["Hf"], cfg["Wf"], cfg["pre_nms"], M, P)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)

    def e2e_step():
        _lib.check(L.b2d_proposal_crop_host(Fe, n_loc, A, cfg["C"], cfg["Hf"], cfg["Wf"], _lib.ptr(hp), _lib.ptr(hd),
                                            _lib.ptr(hi), _lib.ptr(anchors), _lib.ptr(hf), cfg["pre_nms"], M,
                                            cfg["nms_thresh"], P, 1.0 / cfg["stride"], cfg["sampling_ratio"],
                                            _lib.ptr(o_rois), _lib.ptr(o_sc), _lib.ptr(o_num), _lib.ptr(o_pool),
                                            _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)), "b2d_proposal_crop_host")

    for _ in range(3):
        e2e_step()
    barrier()
    e_steps = max(3, args.steps // 2)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(e_steps):
        e2e_step()
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t[0])
    e2e_value = world * Fe * e_steps / (e2e_ms * 1e-3)
    h2d = sum(t.numel() * t.element_size() for t in (hp, hd, hf, hi))
    d2h = sum(t.numel() * t.element_size() for t in (o_rois, o_sc, o_num, o_pool))
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None
    # device and host paths must agree
    assert torch.equal(o_num, num[:Fe].cpu()) and torch.equal(o_rois, rois[:Fe].cpu())

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
        else:
            peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
        b_prop, b_nms, b_crop = algorithmic_bytes(cfg, n_rois / F)
        crop_gbs = F * b_crop / (crop_ms * 1e-3) / 1e9
        step_gbs = F * (b_prop + b_nms + b_crop) / (elapsed_ms / args.steps * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roi_align_traffic.json")
        if os.path.exists(tp):
            tj = json.load(open(tp))            # one ncu --set full capture, scaled to this run's frames per launch
            traffic = tj["dram_bytes_per_launch"] * F / tj["frames_in_launch"]
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["name"], "frames_per_step_per_gpu": F, "anchors_per_frame": n_loc * A,
                       "pre_nms": cfg["pre_nms"], "post_nms": M, "channels": cfg["C"], "pooled": P,
                       "sampling_ratio": cfg["sampling_ratio"], "rois_per_frame": n_rois / F,
                       "l2_policy": f"inputs larger than L2 ({F * (b_prop + cfg['C'] * n_loc * 4) / 1e6:.0f} MB read, "
                                    f"{F * M * cfg['C'] * P * P * 4 / 1e6:.0f} MB written per step)",
                       "parallelism": f"frame-stream x{world}, no data-path collective"},
            "roofline": {"bound": "hbm", "kernel": "rows::fwd_kernel<2,true> (RoIAlign forward)", "achieved": crop_gbs, "peak": peak,
                         "unit": "GB/s", "frac": crop_gbs / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": F * b_crop, "kernel_ms": crop_ms,
                         "kernel_share_of_step": crop_ms / (elapsed_ms / args.steps),
                         "fused_stage": {"achieved": step_gbs, "frac": step_gbs / peak,
                                         "algorithmic_bytes_per_frame": b_prop + b_nms + b_crop}},
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "frames_per_call": Fe, "ms_per_call": e2e_ms / e_steps,
                    "note": "PCIe-bound: 45 MB up and 60 MB down per frame; frames pipelined over H2D / compute / D2H streams"},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            fps, med = cpu_reference_frames_per_s(cfg, args.cpu_baseline_frames, 1, threads)
            line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                                    "sample": f"{args.cpu_baseline_frames} frames of {cfg['name']} after 1 warm-up, "
                                              f"oracle port of proposal_layer + torchvision roi_align, {threads} threads",
                                    "median_ms_per_frame": 1e3 * med}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
