"""Parity of the MEASURED paths (the ones bench.py times) against the oracle:
  * the host-buffer C-ABI entry `b2d_proposal_crop_host` (bench `e2e`), every output including the pooled features;
  * the rows RoIAlign kernel at the bench configuration (C = 1024, 80x120, RoIs from the proposal stage);
  * device guard / per-stream scratch of the Python wrappers.
"""
import numpy as np
import pytest
import torch

from oracle import glue_oracle as O

pytestmark = pytest.mark.gpu

SCALES, RATIOS = [2, 4, 8, 16, 32], [0.5, 0.75, 1, 1.25, 2]


def dev(i=0):
    return torch.device("cuda", i)


def synth(seed, F, Hf, Wf, A, C):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(F, Hf, Wf, 2 * A, generator=g)
    pair = torch.stack((logits[..., :A], logits[..., A:]), dim=-1).softmax(-1)
    prob = torch.cat((pair[..., 0], pair[..., 1]), dim=-1).contiguous()
    d = torch.randn(F, Hf, Wf, A, 4, generator=g)
    d[..., :2] *= 0.1
    d[..., 2:] *= 0.2
    feat = torch.randn(F, C, Hf, Wf, generator=g)
    return prob, d.reshape(F, Hf, Wf, 4 * A).contiguous(), feat


def run_host_entry(prob, deltas, info, anchors_dev, feat, A, pre, post, thr, P=7, stride=16, sr=2):
    from faster_rcnn_pytorch_multimodal_b200 import _lib
    L = _lib.lib()
    F, Hf, Wf = prob.shape[:3]
    C = feat.shape[1]
    n_loc = Hf * Wf
    pin = lambda t: t.contiguous().pin_memory()
    hp, hd, hi, hf = pin(prob), pin(deltas), pin(info), pin(feat)
    o_rois = torch.full((F, post, 5), -7.0).pin_memory()
    o_sc = torch.full((F, post), -7.0).pin_memory()
    o_num = torch.full((F,), -7, dtype=torch.int32).pin_memory()
    o_pool = torch.full((F * post, C, P, P), -7.0).pin_memory()
    ws = torch.empty(L.b2d_pipeline_device_bytes(F, n_loc, A, C, Hf, Wf, pre, post, P), dtype=torch.uint8, device=dev())
    _lib.check(L.b2d_proposal_crop_host(F, n_loc, A, C, Hf, Wf, _lib.ptr(hp), _lib.ptr(hd), _lib.ptr(hi),
                                        _lib.ptr(anchors_dev), _lib.ptr(hf), pre, post, thr, P, 1.0 / stride, sr,
                                        _lib.ptr(o_rois), _lib.ptr(o_sc), _lib.ptr(o_num), _lib.ptr(o_pool),
                                        _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev())), "b2d_proposal_crop_host")
    return o_rois, o_sc, o_num, o_pool


@pytest.mark.parametrize("name,Hf,Wf,W,H,C", [("kitti", 24, 78, 1242, 375, 64), ("waymo", 80, 120, 1920, 1280, 96)])
def test_host_entry_every_output_vs_device_path_and_oracle(name, Hf, Wf, W, H, C):
    """`b2d_proposal_crop_host` with pinned host buffers, F = 3, one frame keeping fewer than post_nms RoIs:
    rois / scores / num_out bit-exact vs the device path, pooled features vs torchvision-CPU to 1e-5."""
    from faster_rcnn_pytorch_multimodal_b200 import ops
    from faster_rcnn_pytorch_multimodal_b200.layer_utils.snippets import generate_anchors_pre
    F, A, pre, post, thr = 3, 25, 6000, 300, 0.7
    prob, deltas, feat = synth(31, F, Hf, Wf, A, C)
    info = torch.tensor([[0, W, 0, H, 0, 0, 1.0]]).repeat(F, 1)
    d1 = deltas[1].view(Hf, Wf, A, 4)
    d1[..., 2:] = 5.0                                # frame 1: every box 148x its anchor -> all clip to the whole frame -> NMS keeps one
    anchors, _ = generate_anchors_pre(Hf, Wf, 16, SCALES, RATIOS, 1.0, device=dev())
    o_rois, o_sc, o_num, o_pool = run_host_entry(prob, deltas, info, anchors, feat, A, pre, post, thr)
    rois, sc, _, _, num = ops.proposal_batched(prob.to(dev()), deltas.to(dev()), info.to(dev()), anchors, None, A, pre,
                                               post, thr, batch_index_stride=1)
    assert torch.equal(o_num, num.cpu())
    assert 0 < int(o_num[1]) < 10 and int(o_num[0]) == post
    assert torch.equal(o_rois, rois.cpu()) and torch.equal(o_sc, sc.cpu())
    pooled = o_pool.view(F, post, C, 7, 7)
    for f in range(F):
        n = int(o_num[f])
        blob = o_rois[f, :n].clone()
        blob[:, 0] = 0
        want = O.roi_align(feat[f:f + 1], blob, (7, 7), 1.0 / 16, 2, False)
        assert torch.allclose(pooled[f, :n], want, rtol=1e-5, atol=1e-5), float((pooled[f, :n] - want).abs().max())
        assert float(pooled[f, n:].abs().max()) == 0.0 if n < post else True


def test_rows_kernel_at_the_bench_configuration_vs_torchvision():
    """C = 1024, 80x120, F = 2, the RoIs the proposal stage produces (the bench workload): three 32-channel
    groups, including the last, against torchvision-CPU; every other group against channel independence."""
    from faster_rcnn_pytorch_multimodal_b200 import ops
    from faster_rcnn_pytorch_multimodal_b200.layer_utils.snippets import generate_anchors_pre
    F, Hf, Wf, A, C, M = 2, 80, 120, 25, 1024, 300
    prob, deltas, feat = synth(5, F, Hf, Wf, A, C)
    info = torch.tensor([[0, 1920, 0, 1280, 0, 0, 1.0]]).repeat(F, 1)
    anchors, _ = generate_anchors_pre(Hf, Wf, 16, SCALES, RATIOS, 1.0, device=dev())
    rois, _, _, _, num = ops.proposal_batched(prob.to(dev()), deltas.to(dev()), info.to(dev()), anchors, None, A, 6000,
                                              M, 0.7, batch_index_stride=1)
    assert num.tolist() == [M, M]
    got = ops.roi_align(feat.to(dev()), rois.view(-1, 5), (7, 7), 1.0 / 16, 2, False, seg_count=num, seg_stride=M)
    assert got.shape == (F * M, C, 7, 7)
    rc = rois.view(-1, 5).cpu()
    tall = ((rc[:, 4] - rc[:, 2]) / 16 > 12).float().mean()
    assert 0.1 < float(tall) < 0.9, "the workload must exercise whole-RoI and split-RoI items"
    for g in (0, 13, 31):
        want = O.roi_align(feat[:, g * 32:(g + 1) * 32], rc, (7, 7), 1.0 / 16, 2, False)
        g_got = got[:, g * 32:(g + 1) * 32].cpu()
        assert torch.allclose(g_got, want, rtol=1e-5, atol=1e-5), (g, float((g_got - want).abs().max()))
    # all groups run the same code on different planes: feed every group the planes of group 13
    rep = feat[:, 13 * 32:14 * 32].repeat(1, 32, 1, 1)
    got_rep = ops.roi_align(rep.to(dev()), rois.view(-1, 5), (7, 7), 1.0 / 16, 2, False, seg_count=num, seg_stride=M)
    ref13 = got[:, 13 * 32:14 * 32]
    for g in range(32):
        assert torch.equal(got_rep[:, g * 32:(g + 1) * 32], ref13), g


def test_rows_kernel_split_rois_every_item_count(monkeypatch):
    """RoIs from 1 to 70 feature rows tall (1..7 items each, incl. bin-rows taller than the window) on the 80x120
    ring, mixed with small ones, F = 3 with ragged counts: the assembled slices must equal torchvision's."""
    from faster_rcnn_pytorch_multimodal_b200 import ops
    monkeypatch.setattr(ops, "ROI_ROUTE", "rows")        # (the automatic dispatch sends calls this small to the gather kernel)
    F, C, H, W, M = 3, 64, 80, 120, 120
    g = torch.Generator().manual_seed(17)
    feat = torch.randn(F, C, H, W, generator=g)
    rois = torch.zeros(F, M, 5)
    for f in range(F):
        hh = torch.linspace(8.0, 70.0 * 16, M)[torch.randperm(M, generator=g)]
        ww = torch.rand(M, generator=g) * 600 + 8
        x1 = torch.rand(M, generator=g) * (W * 16 - ww).clamp(min=1)
        y1 = torch.rand(M, generator=g) * (H * 16 - hh).clamp(min=1)
        rois[f] = torch.stack((torch.full((M,), float(f)), x1, y1, x1 + ww, y1 + hh), 1)
    cnt = torch.tensor([M, M - 31, 1], dtype=torch.int32)
    got = ops.roi_align(feat.to(dev()), rois.view(-1, 5).to(dev()), (7, 7), 1.0 / 16, 2, False, seg_count=cnt.to(dev()),
                        seg_stride=M).view(F, M, C, 7, 7).cpu()
    want = O.roi_align(feat, rois.view(-1, 5), (7, 7), 1.0 / 16, 2, False).view(F, M, C, 7, 7)
    for f in range(F):
        want[f, int(cnt[f]):] = 0
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-5), float((got - want).abs().max())
    # plain [R,5] list (frames filter by col0) incl. a RoI of a frame that does not exist -> zero row
    flat = rois.view(-1, 5).clone()
    flat[5, 0] = 9.0
    got2 = ops.roi_align(feat.to(dev()), flat.to(dev()), (7, 7), 1.0 / 16, 2, False).cpu()
    want2 = O.roi_align(feat, rois.view(-1, 5), (7, 7), 1.0 / 16, 2, False)
    want2[5] = 0
    assert torch.allclose(got2, want2, rtol=1e-5, atol=1e-5)


def test_roi_align_int64_counts_and_half_features(monkeypatch):
    """seg_count given as int64 is converted (the kernels read int32); half features come back as half."""
    from faster_rcnn_pytorch_multimodal_b200 import ops
    monkeypatch.setattr(ops, "ROI_ROUTE", "rows")
    F, C, H, W, M = 2, 32, 24, 40, 20
    g = torch.Generator().manual_seed(2)
    feat = torch.randn(F, C, H, W, generator=g)
    xy = torch.rand(F * M, 2, generator=g) * 200
    wh = torch.rand(F * M, 2, generator=g) * 150 + 4
    rois = torch.cat((torch.arange(F).repeat_interleave(M).float().view(-1, 1), xy, xy + wh), 1)
    cnt64 = torch.tensor([M, 7], dtype=torch.int64, device=dev())
    got = ops.roi_align(feat.to(dev()), rois.to(dev()), (7, 7), 1.0 / 16, 2, False, seg_count=cnt64, seg_stride=M)
    want = O.roi_align(feat, rois, (7, 7), 1.0 / 16, 2, False).view(F, M, C, 7, 7)
    want[1, 7:] = 0
    assert torch.allclose(got.cpu().view(F, M, C, 7, 7), want, rtol=1e-5, atol=1e-5)
    half = ops.roi_align(feat.to(dev()).half(), rois.to(dev()), (7, 7), 1.0 / 16, 2, False)
    assert half.dtype == torch.float16


def test_wrappers_use_per_stream_scratch():
    """Two proposal calls overlapped on two CUDA streams must not share a workspace."""
    from faster_rcnn_pytorch_multimodal_b200 import ops
    from faster_rcnn_pytorch_multimodal_b200.layer_utils.snippets import generate_anchors_pre
    Hf, Wf, A = 40, 60, 25
    anchors, _ = generate_anchors_pre(Hf, Wf, 16, SCALES, RATIOS, 1.0, device=dev())
    info = torch.tensor([[0, 960, 0, 640, 0, 0, 1.0]], device=dev())
    ins = [tuple(t.to(dev()) for t in synth(40 + i, 1, Hf, Wf, A, 1)[:2]) for i in range(2)]
    serial = [ops.proposal_batched(p, d, info, anchors, None, A, 6000, 300, 0.7) for p, d in ins]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for _ in range(5):
        outs = []
        for (p, d), s in zip(ins, streams):
            with torch.cuda.stream(s):
                outs.append(ops.proposal_batched(p, d, info, anchors, None, A, 6000, 300, 0.7))
        torch.cuda.synchronize()
        for a, b in zip(serial, outs):
            assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[4], b[4])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_wrappers_switch_to_the_tensors_device():
    """Tensors on cuda:1 while cuda:0 is current (torchvision's ops guard the device; so do these)."""
    from faster_rcnn_pytorch_multimodal_b200 import ops
    assert torch.cuda.current_device() == 0
    g = torch.Generator().manual_seed(9)
    boxes = torch.rand(500, 2, generator=g) * 300
    boxes = torch.cat((boxes, boxes + torch.rand(500, 2, generator=g) * 80 + 1), 1)
    scores = torch.rand(500, generator=g)
    keep = ops.nms(boxes.to(dev(1)), scores.to(dev(1)), 0.5)
    assert keep.device == dev(1) and torch.equal(keep.cpu(), O.nms(boxes, scores, 0.5))
    feat = torch.randn(1, 32, 24, 40, generator=g)
    rois = torch.cat((torch.zeros(500, 1), boxes), 1)
    got = ops.roi_align(feat.to(dev(1)), rois.to(dev(1)), (7, 7), 1.0 / 16, 2, False)
    assert torch.allclose(got.cpu(), O.roi_align(feat, rois, (7, 7), 1.0 / 16, 2, False), rtol=1e-5, atol=1e-5)
    assert torch.cuda.current_device() == 0
    # both devices from one process, one after the other: kernel attributes (cluster size, dynamic shared memory, the
    # cooperative launch capacity) are per device
    Hf, Wf, A = 24, 78, 25
    gen = torch.Generator().manual_seed(10)
    logits = torch.randn(2, Hf, Wf, 2 * A, generator=gen)
    prob = torch.cat((logits[..., :A], logits[..., A:]), -1).sigmoid()
    deltas = torch.randn(2, Hf, Wf, 4 * A, generator=gen) * 0.1
    anchors = torch.from_numpy(O.generate_anchors_pre(Hf, Wf, 16, np.array([2, 4, 8, 16, 32]), np.array([0.5, 0.75, 1, 1.25, 2]), 1.0)[0])
    info = torch.tensor([[0, 1242, 0, 375, 0, 0, 1.0]]).repeat(2, 1)
    outs = []
    for d in (0, 1, 0):
        r, s, _, _, n = ops.proposal_batched(prob.to(dev(d)), deltas.to(dev(d)), info.to(dev(d)), anchors.to(dev(d)), None, A,
                                             6000, 300, 0.7)
        k, kn = ops.nms_sorted(r[:, :, 1:].contiguous(), 0.5)
        k = k.cpu()
        for f in range(2):
            k[f, int(kn[f]):] = -1                     # (rows past the count are not written)
        outs.append((r.cpu(), s.cpu(), n.cpu(), k, kn.cpu()))
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            assert torch.equal(a, b)
    assert torch.cuda.current_device() == 0
