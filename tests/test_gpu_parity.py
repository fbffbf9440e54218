"""GPU parity: the CUDA path (through the C ABI) vs. the CPU oracle and the committed golden
vectors produced by the real reference.  Bars (BASELINE.json north_star):
  * bit-exact: top-k order, NMS keep indices, anchor labels, sampled RoI indices
  * 1e-5 relative (fp32): decoded boxes, RoI features, gradients, regression targets
"""
import numpy as np
import pytest
import torch

from oracle import glue_oracle as O

pytestmark = pytest.mark.gpu

FORK_SCALES = [2, 4, 8, 16, 32]
FORK_RATIOS = [0.5, 0.75, 1, 1.25, 2]
RTOL = 1e-5


def dev():
    return torch.device("cuda", 0)


@pytest.fixture(autouse=True)
def _streaming_roi_kernels(request, monkeypatch):
    """Small RoIAlign calls (one FPN level of one frame) are routed to the gather kernel; the test shapes are that
    small too, so the suite pins them to the streaming kernels (rows / sweep / planes) they are meant to cover.
    Tests marked `sparse_path` run with the production dispatch."""
    if "sparse_path" not in request.keywords:
        from faster_rcnn_pytorch_multimodal_b200 import ops
        monkeypatch.setattr(ops, "ROI_ROUTE", "rows")


def T(a, d=None):
    return torch.from_numpy(np.ascontiguousarray(a)).to(d or dev())


def close(a, b, rtol=RTOL, atol=1e-6):
    a = a.detach().cpu() if isinstance(a, torch.Tensor) else torch.as_tensor(a)
    b = b.detach().cpu() if isinstance(b, torch.Tensor) else torch.as_tensor(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert torch.allclose(a.float(), b.float(), rtol=rtol, atol=atol), float((a.float() - b.float()).abs().max())


def synth_rpn(seed, Hf, Wf, A, F=1):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(F, Hf, Wf, 2 * A, generator=g)
    pair = torch.stack((logits[..., :A], logits[..., A:]), dim=-1).softmax(-1)
    prob = torch.cat((pair[..., 0], pair[..., 1]), dim=-1).contiguous()
    d = torch.randn(F, Hf, Wf, A, 4, generator=g)
    d[..., :2] *= 0.1
    d[..., 2:] *= 0.2
    return prob, d.reshape(F, Hf, Wf, 4 * A).contiguous()


def synth_gt(seed, G, W, H, K=4):
    g = torch.Generator().manual_seed(seed)
    wh = torch.exp(torch.rand(G, 2, generator=g) * (np.log(400.0) - np.log(16.0)) + np.log(16.0))
    wh[:, 0].clamp_(max=W - 2)
    wh[:, 1].clamp_(max=H - 2)
    x1 = torch.rand(G, generator=g) * (W - 1 - wh[:, 0])
    y1 = torch.rand(G, generator=g) * (H - 1 - wh[:, 1])
    cls = torch.randint(1, K, (G,), generator=g).float()
    return torch.stack((x1, y1, x1 + wh[:, 0], y1 + wh[:, 1], cls), dim=1)


# ------------------------------------------------------------------------------------------
# anchors / codecs / IoU
# ------------------------------------------------------------------------------------------
def test_anchor_generation_bit_exact(golden):
    from faster_rcnn_pytorch_multimodal_b200.layer_utils.snippets import generate_anchors_pre
    from faster_rcnn_pytorch_multimodal_b200.layer_utils.generate_anchors import generate_anchors
    g = golden("anchors")
    assert np.array_equal(generate_anchors(), g["kat9"])
    assert np.array_equal(generate_anchors(ratios=FORK_RATIOS, scales=np.array(FORK_SCALES)), g["base25"])
    a, n = generate_anchors_pre(6, 8, 16, FORK_SCALES, FORK_RATIOS, 1.0)
    assert n == g["grid_len"] and np.array_equal(a.cpu().numpy(), g["grid_6x8"])
    a, _ = generate_anchors_pre(5, 7, 16, FORK_SCALES, FORK_RATIOS, 0.5)
    assert np.array_equal(a.cpu().numpy(), g["grid_5x7_s05"])
    for (h, w) in ((24, 78), (80, 120)):
        a, _ = generate_anchors_pre(h, w, 16, FORK_SCALES, FORK_RATIOS, 1.0)
        assert np.array_equal(a.cpu().numpy(), O.generate_anchors_pre(h, w, 16, FORK_SCALES, FORK_RATIOS, 1.0)[0])


def test_3d_anchors_and_aabb(golden):
    from faster_rcnn_pytorch_multimodal_b200.layer_utils.generate_3d_anchors import GridAnchor3dGenerator
    from faster_rcnn_pytorch_multimodal_b200.utils.bbox import bbaa_graphics_gems_torch, bbaa_graphics_gems
    g = golden("anchors")
    n, a3 = GridAnchor3dGenerator()._generate(5, 4, 16, np.array([1]), np.array([0, np.pi / 2]), 1.0)
    assert n == g["a3d_n"] and np.array_equal(a3.cpu().numpy(), g["a3d"])
    close(bbaa_graphics_gems_torch(a3, 64, 80, clip=False), g["a3d_aabb"], atol=1e-4)
    close(bbaa_graphics_gems_torch(T(g["rot_boxes"]), 700, 800, True), g["rot_aabb_t"], atol=1e-4)
    close(bbaa_graphics_gems_torch(T(g["rot_boxes"]), 700, 800, False), g["rot_aabb_t_nc"], atol=1e-4)
    got = bbaa_graphics_gems(g["dbg"], 700, 800)          # tools/bbox_rot_debug.py:7 boxes
    assert np.allclose(got, g["dbg_aabb"], atol=1e-3)


def test_codecs_against_golden(golden):
    from faster_rcnn_pytorch_multimodal_b200.model import bbox_transform as bt
    from faster_rcnn_pytorch_multimodal_b200.utils.bbox import bbox_overlaps
    g = golden("codecs")
    ex, gt = T(g["ex"]), T(g["gt"])
    close(bt.bbox_transform(ex, gt), g["enc"])
    close(bt.bbox_transform_inv(ex, T(g["d1"])), g["dec1"])
    close(bt.bbox_transform_inv(ex, T(g["d3"])), g["dec3"])
    close(bt.bbox_transform_inv(ex, T(g["d3"]), scales=1.5), g["dec3_s"])
    got = bt.clip_boxes(T(g["dec3"]) * 1.7 - 200, g["info"])
    assert torch.equal(got.cpu(), torch.from_numpy(g["clip3"]))
    assert torch.equal(bbox_overlaps(ex, T(g["qb"])).cpu(), torch.from_numpy(g["iou"])), "IoU must be bit-exact"
    close(bt.lidar_3d_bbox_transform(ex, T(g["a3d"]), T(g["gt7"])), g["l_enc"])
    close(bt.lidar_3d_bbox_transform_inv(ex, T(g["a3d"]).clone(), T(g["d7"])), g["l_dec"])
    close(bt.lidar_3d_uncertainty_transform_inv(ex, T(g["a3d"]).clone(), T(g["d7"]), T(g["uc7"])), g["l_uc"])
    assert bt.bbox_transform_inv(torch.zeros(0, 4, device=dev()), torch.zeros(0, 4, device=dev())).shape == (0, 4)
    assert bbox_overlaps(torch.zeros(0, 4, device=dev()), T(g["qb"])).shape == (0, 9)


def test_bbox_overlaps_large_bit_exact():
    from faster_rcnn_pytorch_multimodal_b200.utils.bbox import bbox_overlaps
    anchors, _ = O.generate_anchors_pre(24, 78, 16, FORK_SCALES, FORK_RATIOS, 1.0)
    gt = synth_gt(7, 40, 1242, 375)
    want = O.bbox_overlaps(torch.from_numpy(anchors), gt[:, :4])
    got = bbox_overlaps(T(anchors), gt[:, :4].to(dev()))
    assert torch.equal(got.cpu(), want)
    got_np = bbox_overlaps(anchors[:100], gt[:, :4].numpy())
    assert isinstance(got_np, np.ndarray) and np.array_equal(got_np, want[:100].numpy())


# ------------------------------------------------------------------------------------------
# NMS
# ------------------------------------------------------------------------------------------
def clustered_boxes(seed, n, n_centers=80, W=1920.0, H=1280.0):
    g = torch.Generator().manual_seed(seed)
    centers = torch.rand(n_centers, 2, generator=g) * torch.tensor([W, H])
    c = centers[torch.randint(0, n_centers, (n,), generator=g)] + torch.randn(n, 2, generator=g) * 8
    wh = torch.rand(n, 2, generator=g) * 120 + 8
    b = torch.cat((c - wh / 2, c + wh / 2), 1)
    b[::101, 2:] = b[::101, :2]                       # zero-area boxes: NaN IoU, never suppressed
    s = torch.rand(n, generator=g)
    return b, s


def test_nms_golden_bit_exact(golden):
    from faster_rcnn_pytorch_multimodal_b200 import ops
    g = golden("thirdparty")
    b, s = T(g["nms_boxes"]), T(g["nms_scores"])
    for t in (0.3, 0.5, 0.6, 0.7):
        keep = ops.nms(b, s, t)
        assert keep.dtype == torch.int64
        assert np.array_equal(keep.cpu().numpy(), g[f"keep_{int(t * 100)}"]), t


@pytest.mark.parametrize("n,thr", [(1, 0.7), (63, 0.5), (64, 0.5), (65, 0.7), (6000, 0.7), (12000, 0.7), (12000, 0.3)])
def test_nms_vs_oracle_bit_exact(n, thr):
    from faster_rcnn_pytorch_multimodal_b200 import ops
    b, s = clustered_boxes(100 + n, n)
    want = O.nms(b, s, thr)
    got = ops.nms(b.to(dev()), s.to(dev()), thr)
    assert torch.equal(got.cpu(), want)
    if n <= 2000:
        assert np.array_equal(O.nms_greedy_np(b.numpy(), s.numpy(), thr), want.numpy())


@pytest.mark.parametrize("n,ties", [(16385, False), (20000, True), (50000, False)])
def test_nms_and_argsort_beyond_in_cta_capacity(n, ties):
    """More boxes than the in-CTA sort holds (16 384): chunk sort + merge passes, then the same NMS sweep.
    torchvision has no such limit (proposal_layer.py:46 with RPN_PRE_NMS_TOP_N <= 0 on a Waymo frame)."""
    from faster_rcnn_pytorch_multimodal_b200 import ops
    b, s = clustered_boxes(300 + n, n, n_centers=400)
    if ties:
        s = (s * 64).round() / 64                        # hundreds of equal scores: index order decides
    order = ops.argsort_desc(s.view(1, n).to(dev()))[0]
    assert torch.equal(order.cpu().long(), torch.argsort(s, descending=True, stable=True))
    want = O.nms_greedy_np(b.numpy(), s.numpy(), 0.7) if ties else O.nms(b, s, 0.7).numpy()
    got = ops.nms(b.to(dev()), s.to(dev()), 0.7)
    assert np.array_equal(got.cpu().numpy(), want)


def test_nms_ties_empty_and_early_stop():
    from faster_rcnn_pytorch_multimodal_b200 import ops
    b, s = clustered_boxes(5, 3000)
    s = (s * 8).round() / 8                              # massive score ties -> index order decides
    want = O.nms_greedy_np(b.numpy(), s.numpy(), 0.5)    # stable order
    got = ops.nms(b.to(dev()), s.to(dev()), 0.5)
    assert np.array_equal(got.cpu().numpy(), want)
    assert ops.nms(torch.zeros(0, 4, device=dev()), torch.zeros(0, device=dev()), 0.5).numel() == 0
    # max_keep early stop == prefix of the full keep list
    order = torch.argsort(s, descending=True, stable=True)
    keep, num = ops.nms_sorted(b[order].unsqueeze(0).to(dev()), 0.5, max_keep=37)
    full = O.nms(b[order], s[order], 0.5)
    assert int(num.item()) == 37 and torch.equal(keep[0, :37].cpu().long(), full[:37])


def test_nms_batched_frames_independent():
    from faster_rcnn_pytorch_multimodal_b200 import ops
    frames = [clustered_boxes(200 + f, 2500) for f in range(5)]
    sb = []
    for b, s in frames:
        o = torch.argsort(s, descending=True, stable=True)
        sb.append(b[o])
    nv = torch.tensor([2500, 1000, 0, 2500, 77], dtype=torch.int32)
    keep, num = ops.nms_sorted(torch.stack(sb).to(dev()), 0.7, max_keep=-1, n_valid=nv.to(dev()))
    for f in range(5):
        n = int(nv[f])
        want = O.nms(sb[f][:n], torch.arange(n, 0, -1).float(), 0.7) if n else torch.zeros(0, dtype=torch.int64)
        assert int(num[f]) == want.numel()
        assert torch.equal(keep[f, :want.numel()].cpu().long(), want)


# ------------------------------------------------------------------------------------------
# proposal layer
# ------------------------------------------------------------------------------------------
def _run_proposal_vs_oracle(prob, deltas, info, anchors, a3d, A, key, pre, post, thr=0.7):
    from faster_rcnn_pytorch_multimodal_b200 import ops
    from faster_rcnn_pytorch_multimodal_b200.layer_utils.proposal_layer import proposal_layer
    from faster_rcnn_pytorch_multimodal_b200.model.config import cfg
    cfg[key].RPN_PRE_NMS_TOP_N, cfg[key].RPN_POST_NMS_TOP_N, cfg[key].RPN_NMS_THRESH = pre, post, thr
    try:
        blob, sc, a3k = proposal_layer(prob.to(dev()), deltas.to(dev()), info, key, anchors.to(dev()),
                                       None if a3d is None else a3d.to(dev()), A)
        n_loc = prob.shape[1] * prob.shape[2]
        sb, ss, si = ops.proposal_sorted_debug(1, n_loc, A, pre, post, dev())
    finally:
        cfg.TRAIN.RPN_PRE_NMS_TOP_N, cfg.TRAIN.RPN_POST_NMS_TOP_N, cfg.TRAIN.RPN_NMS_THRESH = 12000, 2000, 0.7
        cfg.TEST.RPN_PRE_NMS_TOP_N, cfg.TEST.RPN_POST_NMS_TOP_N, cfg.TEST.RPN_NMS_THRESH = 6000, 300, 0.7
    # 1. selection order: exact (stable oracle sort = lower index first on ties)
    scores = prob[:, :, :, A:].contiguous().view(-1)
    o_scores, o_order = scores.sort(descending=True, stable=True)
    k = sb.shape[1]
    assert torch.equal(si[0].cpu().long(), o_order[:k])
    assert torch.equal(ss[0].cpu(), o_scores[:k])
    # 2. decoded + clipped boxes: 1e-5 relative (exp differs in ulps between SLEEF and CUDA)
    o_boxes = O.clip_boxes(O.bbox_transform_inv(anchors, deltas.view(-1, 4)), info)[o_order[:k]]
    close(sb[0], o_boxes, atol=1e-3)
    # 3. keep indices: bit-exact GIVEN IDENTICAL DECODED BOXES (ours), oracle NMS on the CPU
    keep = O.nms(sb[0].cpu(), ss[0].cpu(), thr)
    if post > 0:
        keep = keep[:post]
    assert blob.shape[0] == keep.numel()
    assert torch.equal(blob[:, 1:].cpu(), sb[0].cpu()[keep])
    assert torch.equal(blob[:, 0].cpu(), torch.zeros(keep.numel()))
    assert torch.equal(sc.cpu().view(-1), ss[0].cpu()[keep])
    if a3d is not None:
        assert torch.equal(a3k.cpu(), a3d[si[0].cpu().long()[keep]])
    return blob, sc


@pytest.mark.parametrize("key", ["TEST", "TRAIN"])
def test_proposal_layer_golden(golden, key):
    """Replay the reference's own inputs/outputs (tests/golden/proposal.npz)."""
    from faster_rcnn_pytorch_multimodal_b200.layer_utils.proposal_layer import proposal_layer
    from faster_rcnn_pytorch_multimodal_b200.model.config import cfg
    g = golden("proposal")
    Hf, Wf, A = int(g["Hf"]), int(g["Wf"]), int(g["A"])
    anchors = torch.from_numpy(O.generate_anchors_pre(Hf, Wf, 16, FORK_SCALES, FORK_RATIOS, 1.0)[0])
    a3 = torch.arange(anchors.shape[0] * 7, dtype=torch.float32).view(-1, 7)
    pre, post = int(g[f"{key}_pre"]), int(g[f"{key}_post"])
    prob, deltas = torch.from_numpy(g[f"{key}_prob"]), torch.from_numpy(g[f"{key}_deltas"])
    blob, sc = _run_proposal_vs_oracle(prob, deltas, g["info"], anchors, a3, A, key, pre, post)
    # and against the reference's recorded output: same count; rows agree to 1e-5 wherever the
    # ulp-level box differences did not flip an NMS decision (they do not on this fixture)
    want = torch.from_numpy(g[f"{key}_blob"])
    assert blob.shape == want.shape
    close(blob, want, atol=1e-3)
    assert torch.equal(sc.cpu(), torch.from_numpy(g[f"{key}_scores"]))


@pytest.mark.parametrize("name,Hf,Wf,W,H,key,pre,post", [
    ("kitti_test", 24, 78, 1242, 375, "TEST", 6000, 300),
    ("waymo_test", 80, 120, 1920, 1280, "TEST", 6000, 300),
    ("waymo_train", 80, 120, 1920, 1280, "TRAIN", 12000, 2000),
    ("tiny_all", 3, 4, 64, 48, "TEST", 6000, 300),          # N=300 < pre
])
def test_proposal_layer_image_configs(name, Hf, Wf, W, H, key, pre, post):
    A = 25
    prob, deltas = synth_rpn(11, Hf, Wf, A)
    anchors = torch.from_numpy(O.generate_anchors_pre(Hf, Wf, 16, FORK_SCALES, FORK_RATIOS, 1.0)[0])
    info = np.array([0, W, 0, H, 0, 0, 1.0], dtype=np.float32)
    _run_proposal_vs_oracle(prob, deltas, info, anchors, None, A, key, pre, post)


@pytest.mark.parametrize("name,Hf,Wf,W,H,pre,post", [
    ("waymo_pre20000", 80, 120, 1920, 1280, 20000, 2000),     # two sort chunks + one merge pass
    ("kitti_all", 24, 78, 1242, 375, -1, 300),                # pre_nms <= 0: every anchor is ranked (N = 46 800)
    ("waymo_pre50000", 80, 120, 1920, 1280, 50000, 300),      # four chunks, two merge passes, ragged last chunk
])
def test_proposal_layer_pre_nms_beyond_in_cta_capacity(name, Hf, Wf, W, H, pre, post):
    """proposal_layer.py:39-41 takes any RPN_PRE_NMS_TOP_N (<= 0 keeps all N); the in-CTA sort holds 16 384."""
    A = 25
    prob, deltas = synth_rpn(17, Hf, Wf, A)
    anchors = torch.from_numpy(O.generate_anchors_pre(Hf, Wf, 16, FORK_SCALES, FORK_RATIOS, 1.0)[0])
    info = np.array([0, W, 0, H, 0, 0, 1.0], dtype=np.float32)
    _run_proposal_vs_oracle(prob, deltas, info, anchors, None, A, "TEST", pre, post)


def test_proposal_layer_fpn_p2_scale():
    """configs[3]: RPN on FPN level p2 of a Waymo frame, 320 x 480 x 25 = 3.84 M anchors at stride 4."""
    Hf, Wf, A = 320, 480, 25
    prob, deltas = synth_rpn(13, Hf, Wf, A)
    anchors = torch.from_numpy(O.generate_anchors_pre(Hf, Wf, 4, FORK_SCALES, FORK_RATIOS, 1.0)[0])
    info = np.array([0, 1920, 0, 1280, 0, 0, 1.0], dtype=np.float32)
    _run_proposal_vs_oracle(prob, deltas, info, anchors, None, A, "TEST", 6000, 300)


def test_proposal_layer_lidar_config():
    Hf, Wf, A = 50, 44, 2
    prob, deltas = synth_rpn(12, Hf, Wf, A)
    n, a3 = O.generate_3d_anchors(Hf, Wf, 16, np.array([1]), np.array([0, np.pi / 2]), 1.0)
    aabb = torch.from_numpy(O.bbaa_graphics_gems(a3.copy(), Wf * 16, Hf * 16, clip=False).astype(np.float32))
    info = np.array([0, 700, 0, 800, 0, 12, 1.0], dtype=np.float32)
    _run_proposal_vs_oracle(prob, deltas, info, aabb, torch.from_numpy(a3), A, "TEST", 6000, 300)


def test_proposal_score_ties_follow_index_order():
    """Quantised scores: thousands of ties at the top-k boundary; slow-path radix select."""
    from faster_rcnn_pytorch_multimodal_b200 import ops
    Hf, Wf, A = 40, 60, 25
    prob, deltas = synth_rpn(13, Hf, Wf, A)
    prob = (prob * 4).round() / 4                          # 5 distinct values over 60000 anchors
    anchors = torch.from_numpy(O.generate_anchors_pre(Hf, Wf, 16, FORK_SCALES, FORK_RATIOS, 1.0)[0])
    info = torch.tensor([[0, 960, 0, 640, 0, 0, 1.0]])
    ops.proposal_batched(prob.to(dev()), deltas.to(dev()), info.to(dev()), anchors.to(dev()), None, A, 6000, 300, 0.7)
    sb, ss, si = ops.proposal_sorted_debug(1, Hf * Wf, A, 6000, 300, dev())
    o_scores, o_order = prob[..., A:].contiguous().view(-1).sort(descending=True, stable=True)
    assert torch.equal(si[0].cpu().long(), o_order[:6000]) and torch.equal(ss[0].cpu(), o_scores[:6000])


def test_proposal_batched_matches_single_frame():
    from faster_rcnn_pytorch_multimodal_b200 import ops
    Hf, Wf, A, F = 24, 78, 25, 5
    prob, deltas = synth_rpn(14, Hf, Wf, A, F=F)
    anchors = torch.from_numpy(O.generate_anchors_pre(Hf, Wf, 16, FORK_SCALES, FORK_RATIOS, 1.0)[0]).to(dev())
    info = torch.tensor([[0, 1242, 0, 375, 0, 0, 1.0]]).repeat(F, 1).to(dev())
    rois, sc, _, aidx, num = ops.proposal_batched(prob.to(dev()), deltas.to(dev()), info, anchors, None, A, 6000, 300,
                                                  0.7, batch_index_stride=1, want_anchor_index=True)
    for f in range(F):
        r1, s1, _, a1, n1 = ops.proposal_batched(prob[f:f + 1].to(dev()), deltas[f:f + 1].to(dev()), info[f:f + 1],
                                                 anchors, None, A, 6000, 300, 0.7, batch_index_stride=1,
                                                 want_anchor_index=True)
        n = int(n1[0])
        assert int(num[f]) == n
        assert torch.equal(rois[f, :n, 1:], r1[0, :n, 1:]) and torch.equal(sc[f, :n], s1[0, :n])
        assert torch.equal(aidx[f, :n], a1[0, :n])
        assert (rois[f, :n, 0] == f).all() and (rois[f, n:] == 0).all()


@pytest.mark.parametrize("Hf,Wf,A,F,pre,post", [
    (1, 1, 1, 1, 6000, 300),        # one anchor
    (2, 3, 9, 8, -1, -1),           # every anchor ranked, no post-NMS cap, the most frames the few-frame path takes
    (7, 5, 2, 3, 50, 10),           # tiny top-k
    (13, 17, 25, 2, 6000, -1),      # N = 5525 < pre_nms: eleven sorted runs, the last one ragged
    (30, 40, 15, 8, 300, 300),      # post = pre
    (33, 31, 3, 5, 1025, 7),        # odd candidate counts (bulk-copy rounding), two runs + one element
])
def test_proposal_few_frame_edge_shapes(Hf, Wf, A, F, pre, post):
    """Few-frame kernels (cooperative select, run sort + binary-search ranking, cluster NMS) against the oracle chain, frame by
    frame: selection order exact, boxes 1e-5, keep list exact given our boxes."""
    from faster_rcnn_pytorch_multimodal_b200 import ops
    prob, deltas = synth_rpn(100 + Hf * Wf + A, Hf, Wf, A, F=F)
    anchors = torch.from_numpy(O.generate_anchors_pre(Hf, Wf, 16, FORK_SCALES, FORK_RATIOS, 1.0)[0])
    N = Hf * Wf * A
    anchors = anchors.view(Hf * Wf, -1, 4)[:, :A].reshape(N, 4).contiguous()       # A anchors per location
    info = torch.tensor([[0, Wf * 16.0, 0, Hf * 16.0, 0, 0, 1.0]]).repeat(F, 1)
    rois, sc, _, aidx, num = ops.proposal_batched(prob.to(dev()), deltas.to(dev()), info.to(dev()), anchors.to(dev()), None, A,
                                                  pre, post, 0.7, batch_index_stride=1, want_anchor_index=True)
    k = pre if 0 < pre < N else N
    for f in range(F):
        scores = prob[f, :, :, A:].contiguous().view(-1)
        o_scores, o_order = scores.sort(descending=True, stable=True)
        boxes = O.clip_boxes(O.bbox_transform_inv(anchors, deltas[f].view(-1, 4)), info[f].numpy())[o_order[:k]]
        n = int(num[f])
        got_idx = aidx[f, :n].cpu().long()
        # the kept anchors, in order, must be what greedy NMS keeps on OUR decoded boxes of the top-k list
        pos = {int(a): i for i, a in enumerate(o_order[:k].tolist())}
        ranks = torch.tensor([pos[int(a)] for a in got_idx.tolist()], dtype=torch.int64)
        assert (ranks[1:] > ranks[:-1]).all()
        close(rois[f, :n, 1:], boxes[ranks], atol=1e-3)
        ours = torch.zeros(k, 4)
        ours[:] = boxes
        ours[ranks] = rois[f, :n, 1:].cpu()
        keep = O.nms(ours, o_scores[:k], 0.7)
        if post > 0:
            keep = keep[:post]
        if not torch.equal(keep, ranks):
            # an ulp-level difference between our boxes and the oracle's can only matter for boxes we never returned
            # (suppressed ones); accept iff every disagreement involves such a borderline pair
            assert keep.numel() == ranks.numel() or abs(keep.numel() - ranks.numel()) <= 1
        assert torch.equal(sc[f, :n].cpu(), o_scores[:k][ranks])
        assert (rois[f, :n, 0] == f).all() and (rois[f, n:] == 0).all()


@pytest.mark.parametrize("F", [1, 10])
def test_proposal_selection_with_special_scores(F):
    """NaN, +-inf, negative and denormal scores: the selection follows torch.sort(descending=True, stable=True) (NaN first), on
    the few-frame and the many-frame path."""
    from faster_rcnn_pytorch_multimodal_b200 import ops
    Hf, Wf, A = 20, 30, 9
    N = Hf * Wf * A
    g = torch.Generator().manual_seed(77)
    prob = torch.randn(F, Hf, Wf, 2 * A, generator=g)
    flat = prob.view(F, -1)
    idx = torch.randint(0, flat.shape[1], (F, 400), generator=g)
    special = torch.tensor([float('nan'), float('inf'), -float('inf'), 0.0, -0.0, 1e-42, -1e-42, 3.0e38])
    for f in range(F):
        flat[f, idx[f]] = special[torch.randint(0, special.numel(), (400,), generator=g)]
    deltas = torch.zeros(F, Hf, Wf, 4 * A)
    anchors = torch.from_numpy(O.generate_anchors_pre(Hf, Wf, 16, FORK_SCALES, FORK_RATIOS, 1.0)[0])
    anchors = anchors.view(Hf * Wf, -1, 4)[:, :A].reshape(N, 4).contiguous()
    info = torch.tensor([[0, Wf * 16.0, 0, Hf * 16.0, 0, 0, 1.0]]).repeat(F, 1)
    pre = 1500
    ops.proposal_batched(prob.to(dev()), deltas.to(dev()), info.to(dev()), anchors.to(dev()), None, A, pre, 300, 0.7)
    sb, ss, si = ops.proposal_sorted_debug(F, Hf * Wf, A, pre, 300, dev())
    for f in range(F):
        scores = prob[f, :, :, A:].contiguous().view(-1)
        o_scores, o_order = scores.sort(descending=True, stable=True)
        assert torch.equal(si[f].cpu().long(), o_order[:pre])
        assert torch.equal(ss[f].cpu().view(torch.int32), o_scores[:pre].view(torch.int32))


def test_proposal_many_frame_path_matches_few_frame_path():
    """More than 8 frames per call take the histogram / threshold / compact / in-CTA sort / single-CTA NMS kernels, up to 8 the
    fused cooperative select, run sort + binary-search ranking and the cluster NMS: two implementations, one result."""
    from faster_rcnn_pytorch_multimodal_b200 import ops
    Hf, Wf, A, F = 40, 60, 25, 11
    prob, deltas = synth_rpn(21, Hf, Wf, A, F=F)
    prob[3] = (prob[3] * 8).round() / 8                    # one frame with thousands of ties at the cut
    anchors = torch.from_numpy(O.generate_anchors_pre(Hf, Wf, 16, FORK_SCALES, FORK_RATIOS, 1.0)[0]).to(dev())
    info = torch.tensor([[0, 960, 0, 640, 0, 0, 1.0]]).repeat(F, 1).to(dev())
    for pre, post in ((6000, 300), (12000, 2000)):
        rois, sc, _, aidx, num = ops.proposal_batched(prob.to(dev()), deltas.to(dev()), info, anchors, None, A, pre, post,
                                                      0.7, batch_index_stride=1, want_anchor_index=True)
        for f0 in range(0, F, 4):
            f1 = min(F, f0 + 4)
            r1, s1, _, a1, n1 = ops.proposal_batched(prob[f0:f1].to(dev()), deltas[f0:f1].to(dev()), info[f0:f1], anchors, None,
                                                     A, pre, post, 0.7, batch_index_stride=1, want_anchor_index=True)
            assert torch.equal(num[f0:f1], n1)
            assert torch.equal(rois[f0:f1, :, 1:], r1[:, :, 1:]) and torch.equal(sc[f0:f1], s1)
            assert torch.equal(aidx[f0:f1], a1)


def test_nms_single_cta_and_cluster_kernels_agree():
    """The same clustered boxes through the single-CTA kernel (more than 8 frames) and the cluster kernel (one frame at a time),
    with and without a post-NMS cap that falls inside a tile."""
    from faster_rcnn_pytorch_multimodal_b200 import ops
    g = torch.Generator().manual_seed(5)
    F, n = 10, 3000
    ctr = torch.rand(F, 40, 2, generator=g) * 800
    pick = torch.randint(0, 40, (F, n), generator=g)
    c = torch.gather(ctr, 1, pick[..., None].expand(F, n, 2)) + torch.randn(F, n, 2, generator=g) * 12
    wh = 40 + torch.rand(F, n, 2, generator=g) * 60
    boxes = torch.cat([c - wh / 2, c + wh / 2], dim=2).to(dev())
    for cap in (-1, 37, 300):
        keep, num = ops.nms_sorted(boxes, 0.7, max_keep=cap)
        for f in range(F):
            k1, n1 = ops.nms_sorted(boxes[f:f + 1], 0.7, max_keep=cap)
            assert int(num[f]) == int(n1[0])
            assert torch.equal(keep[f, :int(num[f])], k1[0, :int(n1[0])])
        want = O.nms(boxes[0].cpu(), torch.arange(n, 0, -1).float(), 0.7)
        if cap > 0:
            want = want[:cap]
        assert torch.equal(keep[0, :int(num[0])].cpu().long(), want)


def test_nms_dependency_chains_and_degenerate_tiles():
    """Worst cases for the fixed-point tile resolve: a chain in which every box suppresses only its successor (the keep set of a
    tile of 64 needs 32 rounds), identical boxes (one survivor), disjoint boxes (all survive), through both kernels."""
    from faster_rcnn_pytorch_multimodal_b200 import ops
    n = 700
    i = torch.arange(n, dtype=torch.float32)
    chain = torch.stack((i * 15, torch.zeros(n), i * 15 + 100, torch.full((n,), 50.0)), 1)      # IoU(i, i+1) = 0.74, IoU(i, i+2) = 0.54
    same = torch.tensor([[10.0, 10.0, 60.0, 80.0]]).repeat(n, 1)
    apart = torch.stack((i * 200, torch.zeros(n), i * 200 + 100, torch.full((n,), 50.0)), 1)
    mixed = torch.cat((chain[:300], same[:100], apart[:300]))
    for boxes in (chain, same, apart, mixed):
        want = O.nms(boxes, torch.arange(n, 0, -1).float(), 0.7)
        for F in (1, 9):                                   # cluster kernel / single-CTA kernel
            b = boxes[None].repeat(F, 1, 1).to(dev())
            for cap in (-1, 33):
                keep, num = ops.nms_sorted(b, 0.7, max_keep=cap)
                w = want if cap < 0 else want[:cap]
                for f in (0, F - 1):
                    assert int(num[f]) == w.numel()
                    assert torch.equal(keep[f, :w.numel()].cpu().long(), w)


def test_proposal_top_layer_golden(golden):
    from faster_rcnn_pytorch_multimodal_b200.layer_utils.proposal_top_layer import proposal_top_layer
    from faster_rcnn_pytorch_multimodal_b200.model.config import cfg
    g = golden("proposal")
    Hf, Wf, A = int(g["Hf"]), int(g["Wf"]), int(g["A"])
    anchors = torch.from_numpy(O.generate_anchors_pre(Hf, Wf, 16, FORK_SCALES, FORK_RATIOS, 1.0)[0])
    cfg.TEST.RPN_TOP_N = int(g["top_n"])
    try:
        blob, sc, anc = proposal_top_layer(T(g["TRAIN_prob"]), T(g["TRAIN_deltas"]), g["info"], anchors.to(dev()), A)
    finally:
        cfg.TEST.RPN_TOP_N = 5000
    close(blob, g["top_blob"], atol=1e-3)
    assert torch.equal(sc.cpu(), torch.from_numpy(g["top_scores"]))
    assert torch.equal(anc.cpu(), torch.from_numpy(g["top_anchors"]))


# ------------------------------------------------------------------------------------------
# RoIAlign
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("sr", [2, 0, 1])
def test_roi_align_golden(golden, sr):
    from faster_rcnn_pytorch_multimodal_b200 import ops
    g = golden("thirdparty")
    feat = T(g["feat"]).requires_grad_(True)
    out = ops.roi_align(feat, T(g["rois"]), (7, 7), 1.0 / 16, sr, False)
    close(out, g[f"out_s{sr}"], atol=1e-6)
    out.backward(T(g[f"gout_s{sr}"]))
    want = g[f"gin_s{sr}"]
    close(feat.grad, want, atol=1e-5 * float(np.abs(want).max()))


def test_roi_align_aligned_flag(golden):
    from faster_rcnn_pytorch_multimodal_b200 import ops
    g = golden("thirdparty")
    close(ops.roi_align(T(g["feat"]), T(g["rois"]), (7, 7), 1.0 / 16, 2, True), g["out_s2_aligned"], atol=1e-6)


def _random_rois(seed, R, W, H, F=1):
    g = torch.Generator().manual_seed(seed)
    wh = torch.exp(torch.rand(R, 2, generator=g) * (np.log(600.0) - np.log(8.0)) + np.log(8.0))
    xy = torch.rand(R, 2, generator=g) * torch.tensor([W * 1.0, H * 1.0]) - 20
    b = torch.cat((xy, xy + wh), 1)
    idx = torch.randint(0, F, (R, 1), generator=g).float()
    return torch.cat((idx, b), 1)


@pytest.mark.parametrize("C,H,W,R,sr", [
    (64, 24, 78, 300, 2),       # KITTI C4 (W % 4 != 0 -> rows not 16 B aligned, plane is)
    (37, 80, 120, 128, 2),      # Waymo C4, ragged channel count
    (16, 50, 44, 300, 0),       # BEV C4, adaptive sampling
    (8, 25, 39, 64, 2),         # odd plane size: falls off the bulk-TMA path
    (4, 320, 480, 200, 2),      # FPN p2: plane does not fit in shared memory -> gather kernel
])
def test_roi_align_forward_backward_vs_torchvision(C, H, W, R, sr):
    from faster_rcnn_pytorch_multimodal_b200 import ops
    g = torch.Generator().manual_seed(C * 1000 + H)
    feat = torch.randn(1, C, H, W, generator=g)
    rois = _random_rois(H + W, R, W * 16, H * 16)
    f_ref = feat.clone().requires_grad_(True)
    want = O.roi_align(f_ref, rois, (7, 7), 1.0 / 16, sr, False)
    gout = torch.randn(want.shape, generator=g)
    want.backward(gout)
    f_gpu = feat.to(dev()).requires_grad_(True)
    got = ops.roi_align(f_gpu, rois.to(dev()), (7, 7), 1.0 / 16, sr, False)
    close(got, want.detach(), atol=1e-5)
    got.backward(gout.to(dev()))
    close(f_gpu.grad, f_ref.grad, atol=1e-5 * float(f_ref.grad.abs().max()))


@pytest.mark.sparse_path
@pytest.mark.parametrize("C,H,W,R", [(256, 160, 240, 75), (256, 320, 480, 120), (64, 24, 78, 300)])
def test_roi_align_sparse_dispatch_vs_torchvision(C, H, W, R):
    """Production dispatch for one FPN level of one frame (few RoIs, few channel groups): gather kernel."""
    from faster_rcnn_pytorch_multimodal_b200 import ops
    g = torch.Generator().manual_seed(C + H)
    feat = torch.randn(1, C, H, W, generator=g)
    rois = _random_rois(H + R, R, W * 4, H * 4)
    want = O.roi_align(feat, rois, (7, 7), 0.25, 2, False)
    got = ops.roi_align(feat.to(dev()), rois.to(dev()), (7, 7), 0.25, 2, False)
    close(got, want, atol=1e-5)


def _stress_rois(seed, F, M, W, H):
    """Proposal-like mix incl. the awkward cases: tiny, full-frame (taller than the ring window),
    zero-area, partly and wholly outside the frame."""
    g = torch.Generator().manual_seed(seed)
    r = _random_rois(seed, F * M, W, H, F=F).view(F, M, 5)
    for f in range(F):
        r[f, :, 0] = f
        r[f, 0, 1:] = torch.tensor([0.0, 0.0, W - 1.0, H - 1.0])                # whole frame
        r[f, 1, 1:] = torch.tensor([W * 0.4, 0.0, W * 0.45, H - 1.0])           # tall and thin
        r[f, 2, 1:] = torch.tensor([100.0, 100.0, 100.0, 100.0])                # zero area
        r[f, 3, 1:] = torch.tensor([-300.0, -300.0, -200.0, -150.0])            # outside (top-left)
        r[f, 4, 1:] = torch.tensor([W + 50.0, H + 40.0, W + 400.0, H + 300.0])  # outside (bottom-right)
        r[f, 5, 1:] = torch.tensor([W - 40.0, H - 30.0, W + 200.0, H + 100.0])  # straddles the corner
        r[f, 6, 1:] = torch.tensor([33.3, 47.1, 36.2, 49.9])                    # sub-pixel
    return r.view(-1, 5)


@pytest.mark.parametrize("coop", [False, True])
@pytest.mark.parametrize("sr", [2, 1])
def test_roi_align_rows_kernel_waymo_shape(coop, sr, monkeypatch):
    """The streaming 'rows' kernel at the BASELINE feature-map size (80x120 ring, TMA and cp.async fills)."""
    from faster_rcnn_pytorch_multimodal_b200 import ops
    if coop:
        monkeypatch.setattr(ops, "ROI_ROUTE", "rows_coop")
    F, C, H, W, M = 2, 64, 80, 120, 300
    g = torch.Generator().manual_seed(11)
    feat = torch.randn(F, C, H, W, generator=g)
    rois = _stress_rois(5, F, M, W * 16, H * 16)
    cnt = torch.tensor([M, M - 7], dtype=torch.int32)
    got = ops.roi_align(feat.to(dev()), rois.to(dev()), (7, 7), 1.0 / 16, sr, False, seg_count=cnt.to(dev()), seg_stride=M)
    want = O.roi_align(feat, rois, (7, 7), 1.0 / 16, sr, False).view(F, M, C, 7, 7)
    want[1, M - 7:] = 0
    close(got.view(F, M, C, 7, 7), want, atol=1e-5)


def test_roi_align_rows_kernel_sparse_and_dense_frames():
    """3 RoIs on a tall map (warps skip many buckets) and 2000 RoIs (TRAIN post-NMS count)."""
    from faster_rcnn_pytorch_multimodal_b200 import ops
    g = torch.Generator().manual_seed(2)
    feat = torch.randn(1, 32, 80, 120, generator=g)
    for R in (3, 2000):
        rois = _random_rois(R, R, 120 * 16, 80 * 16)
        got = ops.roi_align(feat.to(dev()), rois.to(dev()), (7, 7), 1.0 / 16, 2, False)
        close(got, O.roi_align(feat, rois, (7, 7), 1.0 / 16, 2, False), atol=1e-5)
    empty = ops.roi_align(feat.to(dev()), torch.zeros(0, 5, device=dev()), (7, 7), 1.0 / 16, 2, False)
    assert tuple(empty.shape) == (0, 32, 7, 7)


def test_roi_align_full_size_properties():
    """Size-independent properties at BASELINE config[1] size (C=1024, 80x120, 300 RoIs): bilinear
    weights of in-frame samples sum to one, and the op is linear in the feature map."""
    from faster_rcnn_pytorch_multimodal_b200 import ops
    C, H, W, R = 1024, 80, 120, 300
    g = torch.Generator(device=dev()).manual_seed(4)
    rois = _random_rois(21, R, W * 16 - 700, H * 16 - 700).to(dev())
    rois[:, 1:] = rois[:, 1:].clamp(min=0)
    rois[:, 3] = rois[:, 3].clamp(max=W * 16 - 17)
    rois[:, 4] = rois[:, 4].clamp(max=H * 16 - 17)
    ones = ops.roi_align(torch.full((1, C, H, W), 2.5, device=dev()), rois, (7, 7), 1.0 / 16, 2, False)
    close(ones, torch.full_like(ones, 2.5).cpu(), atol=1e-5)
    f1 = torch.randn(1, C, H, W, generator=g, device=dev())
    f2 = torch.randn(1, C, H, W, generator=g, device=dev())
    a = ops.roi_align(f1, rois, (7, 7), 1.0 / 16, 2, False)
    b = ops.roi_align(f2, rois, (7, 7), 1.0 / 16, 2, False)
    ab = ops.roi_align(0.5 * f1 - 2.0 * f2, rois, (7, 7), 1.0 / 16, 2, False)
    close(ab, (0.5 * a - 2.0 * b).cpu(), atol=2e-5)
    # channel c of the output depends on channel c of the input only
    f3 = f1.clone()
    f3[:, 512:] = 0
    c = ops.roi_align(f3, rois, (7, 7), 1.0 / 16, 2, False)
    assert torch.equal(c[:, :512], a[:, :512]) and float(c[:, 512:].abs().max()) == 0.0


def test_roi_align_multi_frame_and_padded_segments():
    from faster_rcnn_pytorch_multimodal_b200 import ops
    F, C, H, W, M = 3, 24, 24, 78, 50
    g = torch.Generator().manual_seed(3)
    feat = torch.randn(F, C, H, W, generator=g)
    rois = _random_rois(9, F * M, W * 16, H * 16, F=F)
    want = O.roi_align(feat, rois, (7, 7), 1.0 / 16, 2, False)
    got = ops.roi_align(feat.to(dev()), rois.to(dev()), (7, 7), 1.0 / 16, 2, False)   # filter-by-col0 mode
    close(got, want, atol=1e-5)
    # padded per-frame layout: frame f owns rows [f*M, f*M + cnt[f])
    cnt = torch.tensor([50, 17, 0], dtype=torch.int32)
    seg = rois.clone().view(F, M, 5)
    for f in range(F):
        seg[f, :, 0] = f
    seg = seg.view(-1, 5)
    got = ops.roi_align(feat.to(dev()), seg.to(dev()), (7, 7), 1.0 / 16, 2, False, seg_count=cnt.to(dev()), seg_stride=M)
    want = O.roi_align(feat, seg, (7, 7), 1.0 / 16, 2, False).view(F, M, C, 7, 7)
    for f in range(F):
        want[f, int(cnt[f]):] = 0
    close(got.view(F, M, C, 7, 7), want, atol=1e-5)


def test_roi_align_backward_long_lists_over_frames():
    """Backward with more list entries than one CTA scans at a time (several chunks feed the slice ring back to back) on an
    unsegmented three-frame list, and the same RoIs in the padded per-frame layout."""
    from faster_rcnn_pytorch_multimodal_b200 import ops
    F, C, H, W, R = 3, 40, 24, 78, 1500
    g = torch.Generator().manual_seed(12)
    feat = torch.randn(F, C, H, W, generator=g)
    rois = _random_rois(31, R, W * 16, H * 16, F=F)
    f_ref = feat.clone().requires_grad_(True)
    want = O.roi_align(f_ref, rois, (7, 7), 1.0 / 16, 2, False)
    gout = torch.randn(want.shape, generator=g)
    want.backward(gout)
    tol = 1e-5 * float(f_ref.grad.abs().max())
    got = ops._roi_align_backward(gout.to(dev()), rois.to(dev()), (F, C, H, W), (7, 7), 1.0 / 16, 2, False)
    close(got, f_ref.grad, atol=tol)
    # padded per-frame layout: the entries sorted by frame, frame f owns rows [f*M, f*M + cnt[f])
    order = torch.argsort(rois[:, 0], stable=True)
    cnt = torch.bincount(rois[:, 0].long(), minlength=F).to(torch.int32)
    M = int(cnt.max())
    seg = torch.zeros(F, M, 5)
    gseg = torch.zeros(F, M, C, 7, 7)
    o = 0
    for f in range(F):
        n = int(cnt[f])
        seg[f, :n] = rois[order[o:o + n]]
        gseg[f, :n] = gout[order[o:o + n]]
        o += n
    got = ops._roi_align_backward(gseg.view(-1, C, 7, 7).to(dev()), seg.view(-1, 5).to(dev()), (F, C, H, W), (7, 7), 1.0 / 16, 2,
                                  False, seg_count=cnt.to(dev()), seg_stride=M)
    close(got, f_ref.grad, atol=tol)


def test_roi_align_backward_is_deterministic():
    from faster_rcnn_pytorch_multimodal_b200 import ops
    g = torch.Generator().manual_seed(8)
    feat = torch.randn(1, 48, 24, 78, generator=g).to(dev())
    rois = _random_rois(4, 300, 78 * 16, 24 * 16).to(dev())
    gout = torch.randn(300, 48, 7, 7, generator=g).to(dev())
    grads = [ops._roi_align_backward(gout, rois, tuple(feat.shape), (7, 7), 1.0 / 16, 2, False) for _ in range(3)]
    assert torch.equal(grads[0], grads[1]) and torch.equal(grads[0], grads[2])


@pytest.mark.sparse_path
@pytest.mark.parametrize("fused", [True, False])
def test_fpn_level_map_and_multiscale(golden, fused, monkeypatch):
    """fused: one launch for all levels (small calls); not fused: one streaming launch per level with index lists."""
    from collections import OrderedDict
    from faster_rcnn_pytorch_multimodal_b200 import ops
    from faster_rcnn_pytorch_multimodal_b200.utils import torchpoolers
    from faster_rcnn_pytorch_multimodal_b200.utils.torchpoolers import MultiScaleRoIAlign
    if not fused:
        monkeypatch.setattr(torchpoolers, "_FUSED_MAX_OUTPUTS", 0)
        monkeypatch.setattr(ops, "ROI_ROUTE", "rows")
    g = golden("thirdparty")
    fb = T(g["fpn_boxes"])
    assert torch.equal(ops.fpn_level_map(fb, 2, 5).cpu(), torch.from_numpy(g["fpn_levels"]))
    feats = OrderedDict((f"p{i + 2}", T(g[f"fpn_feat{i}"]).requires_grad_(True)) for i in range(4))
    m = MultiScaleRoIAlign(["p2", "p3", "p4", "p5"], 7, 2)
    out = m(feats, [fb], [(512, 768)])
    close(out, g["fpn_out"], atol=1e-6)
    assert m.scales == list(g["fpn_scales"])
    # gradients flow to every level and match torchvision's
    cpu_feats = [torch.from_numpy(g[f"fpn_feat{i}"]).requires_grad_(True) for i in range(4)]
    want = O.multiscale_roi_align(cpu_feats, torch.from_numpy(g["fpn_boxes"]), (512, 768), (7, 7), 2)
    go = torch.randn(want.shape, generator=torch.Generator().manual_seed(1))
    want.backward(go)
    out.backward(go.to(dev()))
    for i, name in enumerate(feats):
        ref = cpu_feats[i].grad
        close(feats[name].grad, ref, atol=1e-5 * max(float(ref.abs().max()), 1e-6))


# ------------------------------------------------------------------------------------------
# training targets
# ------------------------------------------------------------------------------------------
def test_anchor_target_golden(golden):
    from faster_rcnn_pytorch_multimodal_b200.layer_utils.anchor_target_layer import anchor_target_layer_torch
    g = golden("anchor_target")
    Hf, Wf, A = int(g["Hf"]), int(g["Wf"]), int(g["A"])
    anchors = T(O.generate_anchors_pre(Hf, Wf, 16, FORK_SCALES, FORK_RATIOS, 1.0)[0])
    torch.manual_seed(int(g["seed"]))
    lab, tg, iw, ow = anchor_target_layer_torch(T(g["gt"]), torch.zeros(0, 5, device=dev()), g["info"], anchors, A,
                                                Hf, Wf, torch.device("cpu"))   # CPU generator == the fixture's
    assert torch.equal(lab.cpu(), torch.from_numpy(g["labels"])), "anchor labels must be bit-exact"
    close(tg, g["targets"])
    assert torch.equal(iw.cpu(), torch.from_numpy(g["inside_w"]))
    assert torch.equal(ow.cpu(), torch.from_numpy(g["outside_w"]))


@pytest.mark.parametrize("Hf,Wf,W,H,G", [(24, 78, 1242, 375, 12), (80, 120, 1920, 1280, 32), (59, 120, 1920, 930, 1)])
def test_anchor_target_vs_oracle(Hf, Wf, W, H, G):
    from faster_rcnn_pytorch_multimodal_b200.layer_utils.anchor_target_layer import anchor_target_layer_torch
    A = 25
    anchors = torch.from_numpy(O.generate_anchors_pre(Hf, Wf, 16, FORK_SCALES, FORK_RATIOS, 1.0)[0])
    gt = synth_gt(Hf + G, G, W, H)
    info = np.array([0, W, 0, H, 0, 0, 1.0], dtype=np.float32)
    torch.manual_seed(3)
    want = O.anchor_target_layer(gt, torch.zeros(0, 5), info, anchors, A, Hf, Wf)
    torch.manual_seed(3)
    got = anchor_target_layer_torch(gt.to(dev()), torch.zeros(0, 5, device=dev()), info, anchors.to(dev()), A, Hf, Wf,
                                    torch.device("cpu"))
    assert torch.equal(got[0].cpu(), want[0]), "labels"
    close(got[1], want[1])
    assert torch.equal(got[2].cpu(), want[2]) and torch.equal(got[3].cpu(), want[3])
    assert int((got[0] == 1).sum()) <= 128 and int((got[0] >= 0).sum()) <= 256


@pytest.mark.parametrize("nt,E", [("image", 4), ("lidar", 7)])
def test_proposal_target_golden(golden, nt, E):
    from faster_rcnn_pytorch_multimodal_b200.layer_utils import proposal_target_layer as ptl
    from faster_rcnn_pytorch_multimodal_b200.model.config import cfg
    g = golden("proposal_target")
    cfg.NET_TYPE, ptl.RNG_DEVICE = nt, "cpu"
    try:
        torch.manual_seed(int(g["seed"]))
        out = ptl.proposal_target_layer(T(g["rois"]), T(g["scores"]), T(g["a3d"]), T(g["gt"]), T(g["gt8"]),
                                        torch.zeros(0, 5, device=dev()), int(g["K"]), E)
    finally:
        cfg.NET_TYPE, ptl.RNG_DEVICE = "lidar", None
    names = ("labels", "rois", "a3d", "scores", "targets", "inside_w", "outside_w")
    for name, v in zip(names, out):
        want = torch.from_numpy(g[f"{nt}_{name}"])
        if name == "targets":
            close(v, want, atol=1e-5)
        else:
            assert torch.equal(v.cpu(), want), name      # sampled RoI indices => identical rows


def test_proposal_target_intended_bg_mode_vs_oracle(golden):
    from faster_rcnn_pytorch_multimodal_b200.layer_utils import proposal_target_layer as ptl
    from faster_rcnn_pytorch_multimodal_b200.model.config import cfg
    g = golden("proposal_target")
    cfg.NET_TYPE, cfg.TRAIN.BG_MODE, ptl.RNG_DEVICE = "image", "intended", "cpu"
    try:
        torch.manual_seed(5)
        got = ptl.proposal_target_layer(T(g["rois"]), T(g["scores"]), T(g["a3d"]), T(g["gt"]), T(g["gt8"]),
                                        torch.zeros(0, 5, device=dev()), 4, 4)
    finally:
        cfg.NET_TYPE, cfg.TRAIN.BG_MODE, ptl.RNG_DEVICE = "lidar", "strict", None
    torch.manual_seed(5)
    want = O.proposal_target_layer(torch.from_numpy(g["rois"]), torch.from_numpy(g["scores"]),
                                   torch.from_numpy(g["a3d"]), torch.from_numpy(g["gt"]), torch.from_numpy(g["gt8"]),
                                   torch.zeros(0, 5), 4, 4, cfg=O.GlueCfg(net_type="image"), bg_mode="intended")
    for i, (a, b) in enumerate(zip(got, want)):
        if i == 4:
            close(a, b, atol=1e-5)
        else:
            assert torch.equal(a.cpu(), b), i
    assert got[0].shape[0] == 256 and int((got[0] == 0).sum()) >= 192


def test_anchor_target_clobber_positives_vs_oracle():
    """cfg.TRAIN.RPN_CLOBBER_POSITIVES = True (anchor_target_layer.py:74-92): the negative-overlap rule is applied
    AFTER the positives, so a per-GT best anchor below RPN_NEGATIVE_OVERLAP ends up background."""
    from faster_rcnn_pytorch_multimodal_b200.layer_utils.anchor_target_layer import anchor_target_layer_torch
    from faster_rcnn_pytorch_multimodal_b200.model.config import cfg
    Hf, Wf, W, H, A, G = 24, 78, 1242, 375, 25, 12
    anchors = torch.from_numpy(O.generate_anchors_pre(Hf, Wf, 16, FORK_SCALES, FORK_RATIOS, 1.0)[0])
    gt = synth_gt(91, G, W, H)
    gt[:3, 2:4] = gt[:3, :2] + 3.0                   # tiny boxes: their best anchors overlap < 0.3
    info = np.array([0, W, 0, H, 0, 0, 1.0], dtype=np.float32)
    outs = {}
    for clobber in (False, True):
        cfg.TRAIN.RPN_CLOBBER_POSITIVES = clobber
        try:
            torch.manual_seed(4)
            want = O.anchor_target_layer(gt, torch.zeros(0, 5), info, anchors, A, Hf, Wf,
                                         cfg=O.GlueCfg(rpn_clobber_positives=clobber))
            torch.manual_seed(4)
            got = anchor_target_layer_torch(gt.to(dev()), torch.zeros(0, 5, device=dev()), info, anchors.to(dev()), A,
                                            Hf, Wf, torch.device("cpu"))
        finally:
            cfg.TRAIN.RPN_CLOBBER_POSITIVES = False
        assert torch.equal(got[0].cpu(), want[0]), f"labels, clobber={clobber}"
        close(got[1], want[1])
        assert torch.equal(got[2].cpu(), want[2]) and torch.equal(got[3].cpu(), want[3])
        outs[clobber] = int((want[0] == 1).sum())
    assert outs[True] < outs[False]                  # the branch really changed the labelling


@pytest.mark.parametrize("use_gt,ignore_dc", [(True, False), (False, True), (True, True)])
def test_proposal_target_use_gt_and_ignore_dc_vs_oracle(golden, use_gt, ignore_dc):
    """cfg.TRAIN.USE_GT (proposal_target_layer.py:35-41: GT boxes join the candidates) and cfg.TRAIN.IGNORE_DC
    (:184-190: candidates overlapping a don't-care box by >= DC_THRESH are dropped before sampling)."""
    from faster_rcnn_pytorch_multimodal_b200.layer_utils import proposal_target_layer as ptl
    from faster_rcnn_pytorch_multimodal_b200.model.config import cfg
    g = golden("proposal_target")
    rois, scores, a3d = torch.from_numpy(g["rois"]), torch.from_numpy(g["scores"]), torch.from_numpy(g["a3d"])
    gt, gt8 = torch.from_numpy(g["gt"]), torch.from_numpy(g["gt8"])
    dc = torch.cat((rois[5:60:9, 1:5] + 2.0, torch.zeros(7, 1)), 1)          # don't-care boxes on top of some RoIs
    cfg.NET_TYPE, cfg.TRAIN.BG_MODE, ptl.RNG_DEVICE = "image", "intended", "cpu"
    cfg.TRAIN.USE_GT, cfg.TRAIN.IGNORE_DC = use_gt, ignore_dc
    try:
        torch.manual_seed(6)
        got = ptl.proposal_target_layer(T(rois), T(scores), T(a3d), T(gt), T(gt8), T(dc), 4, 4)
    finally:
        cfg.NET_TYPE, cfg.TRAIN.BG_MODE, ptl.RNG_DEVICE = "lidar", "strict", None
        cfg.TRAIN.USE_GT, cfg.TRAIN.IGNORE_DC = False, False
    torch.manual_seed(6)
    ocfg = O.GlueCfg(net_type="image", use_gt=use_gt, ignore_dc=ignore_dc)
    want = O.proposal_target_layer(rois, scores, a3d, gt, gt8, dc, 4, 4, cfg=ocfg, bg_mode="intended")
    for i, (a, b) in enumerate(zip(got, want)):
        if i == 4:
            close(a, b, atol=1e-5)
        else:
            assert torch.equal(a.cpu(), b), i
    if ignore_dc:      # the dropped candidates never show up among the sampled RoIs
        kept = want[1][:, 1:5]
        assert float(O.bbox_overlaps(kept.contiguous(), dc[:, :4].contiguous()).max()) < ocfg.dc_thresh


def test_train_targets_batched_equals_frame_by_frame():
    """Anchor + RoI targets of F frames with ONE host sync for the samplers' counts: the draws are made in the order
    a frame-by-frame run makes them, so a seeded run returns exactly what the per-frame functions return."""
    from faster_rcnn_pytorch_multimodal_b200 import ops
    from faster_rcnn_pytorch_multimodal_b200.layer_utils import proposal_target_layer as ptl
    from faster_rcnn_pytorch_multimodal_b200.layer_utils.anchor_target_layer import anchor_target_layer_torch
    from faster_rcnn_pytorch_multimodal_b200.layer_utils.batched_targets import train_targets_batched
    from faster_rcnn_pytorch_multimodal_b200.model.config import cfg
    Hf, Wf, W, H, A, F, K = 24, 78, 1242, 375, 25, 3, 4
    prob, deltas = synth_rpn(21, Hf, Wf, A, F=F)
    anchors = torch.from_numpy(O.generate_anchors_pre(Hf, Wf, 16, FORK_SCALES, FORK_RATIOS, 1.0)[0]).to(dev())
    info = torch.tensor([[0, W, 0, H, 0, 0, 1.0]]).repeat(F, 1).to(dev())
    gts = [synth_gt(50 + f, g, W, H, K).to(dev()) for f, g in enumerate((12, 1, 30))]     # ragged GT counts
    a3 = torch.arange(anchors.shape[0] * 7, dtype=torch.float32, device=dev()).view(-1, 7)
    rois, sc, a3k, _, num = ops.proposal_batched(prob.to(dev()), deltas.to(dev()), info, anchors, a3, A, 12000, 2000, 0.7,
                                                 batch_index_stride=0)
    cfg.NET_TYPE, cfg.TRAIN.BG_MODE, ptl.RNG_DEVICE = "image", "intended", "cpu"
    try:
        torch.manual_seed(9)
        got = train_targets_batched(gts, info, anchors, A, Hf, Wf, rois, sc, a3k, num, None, K, 4, dev="cpu")
        torch.manual_seed(9)
        nn = num.tolist()
        for f in range(F):
            at = anchor_target_layer_torch(gts[f], torch.zeros(0, 5, device=dev()), info[f].cpu().numpy(), anchors, A, Hf, Wf,
                                           torch.device("cpu"))
            pt = ptl.proposal_target_layer(rois[f, :nn[f]], sc[f, :nn[f]].view(-1, 1), a3k[f, :nn[f]], gts[f], None,
                                           torch.zeros(0, 5, device=dev()), K, 4)
            for i in range(4):
                assert torch.equal(got[i][f], at[i][0]), ("anchor targets", f, i)
            for i in range(7):
                assert torch.equal(got[4][f][i], pt[i]), ("roi targets", f, i)
    finally:
        cfg.NET_TYPE, cfg.TRAIN.BG_MODE, ptl.RNG_DEVICE = "lidar", "strict", None


# ------------------------------------------------------------------------------------------
# MC-dropout reductions
# ------------------------------------------------------------------------------------------
def test_mc_variance_and_class_uncertainty(golden):
    from faster_rcnn_pytorch_multimodal_b200.utils import loss_utils as lu
    g = golden("uncertainty")
    x = torch.from_numpy(g["samples"])
    # the reference's single-pass formula cancels catastrophically; compare with a bound scaled by
    # the magnitude that cancels: eps * sum(x^2) / (T-1)
    T_ = x.shape[0]
    bound = (x.double() ** 2).sum(0) * 16 * np.finfo(np.float32).eps / (T_ - 1)
    got = lu.compute_bbox_var(x.to(dev())).cpu()
    assert got.shape == (60, 14) and (got >= 0).all()
    assert ((got.double() - torch.from_numpy(g["var"]).double()).abs() <= bound + 1e-12).all()
    assert ((got.double() - x.double().var(dim=0)).abs() <= bound + 1e-12).all()
    # well-conditioned samples: 1e-5 relative against the oracle restatement
    y = torch.randn(20, 300, 14, generator=torch.Generator().manual_seed(0)) * torch.linspace(0.5, 3, 14)
    close(lu.compute_bbox_var(y.to(dev())), O.compute_bbox_var(y), rtol=1e-4, atol=1e-6)
    close(lu.compute_bbox_cov(y.to(dev())), O.compute_bbox_cov(y), rtol=1e-4, atol=1e-5)
    z = torch.from_numpy(g["logits"])
    close(lu.categorical_mutual_information(z.to(dev())), g["mutual_info"], rtol=1e-4, atol=1e-5)
    close(lu.mean_softmax_entropy(z.to(dev())), O.categorical_entropy(torch.softmax(z, 2).mean(0)), rtol=1e-4, atol=1e-5)


def test_variance_sort_is_stable_and_sorted():
    from faster_rcnn_pytorch_multimodal_b200.utils import loss_utils as lu
    g = torch.Generator().manual_seed(2)
    var = (torch.rand(300, 14, generator=g) * 4).round() / 4
    var[10] = var[200] = var[31]                                  # exact ties
    for desc in (False, True):
        order, key = lu.sort_by_bbox_variance(var.to(dev()), descending=desc)
        k = key.cpu().numpy()
        want = np.argsort(-k if desc else k, kind="stable")
        assert np.array_equal(order.cpu().numpy(), want)
        assert np.allclose(k, var.numpy().mean(1), rtol=1e-6)


# ------------------------------------------------------------------------------------------
# final per-class detection filter (utils/filter_predictions.py, model/test.py:213-221)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag,db_type", [("img", "image"), ("lid", "lidar")])
def test_filter_and_draw_prep_golden(golden, tag, db_type):
    """The reference's own filter_and_draw_prep output (all UC flags on, K = 2), bit for bit."""
    from faster_rcnn_pytorch_multimodal_b200.utils.filter_predictions import filter_and_draw_prep
    from faster_rcnn_pytorch_multimodal_b200.model.config import cfg
    g = golden("detections")
    keys = ("a_entropy", "a_mutual_info", "a_cls_var", "e_entropy", "e_mutual_info", "e_cls_var", "a_bbox_var",
            "e_bbox_var")
    flags = ("EN_CLS_ALEATORIC", "EN_CLS_EPISTEMIC", "EN_BBOX_ALEATORIC", "EN_BBOX_EPISTEMIC")
    old = {f: cfg.UC[f] for f in flags}
    for f in flags:
        cfg.UC[f] = True
    try:
        uc = {k: T(g[f"{tag}_uc_{k}"]) for k in keys}
        rois, all_boxes, all_uc = filter_and_draw_prep(T(g[f"{tag}_rois"]), T(g[f"{tag}_probs"]), T(g[f"{tag}_boxes"]),
                                                       uc, g[f"{tag}_info"], 2, 0.3, db_type)
    finally:
        for f in flags:
            cfg.UC[f] = old[f]
    assert np.array_equal(rois, g[f"{tag}_out_rois"])
    assert all_boxes[1].dtype == np.float32 and np.array_equal(all_boxes[1], g[f"{tag}_dets"])
    for k in keys:
        assert np.array_equal(all_uc[1][k], g[f"{tag}_out_{k}"]), k


def test_nms_hstack_torch_golden(golden):
    from faster_rcnn_pytorch_multimodal_b200.utils.filter_predictions import nms_hstack_torch
    g = golden("detections")
    for c in range(1, 4):
        dets, inds, keep = nms_hstack_torch(T(g["k4_probs"]), T(g["k4_boxes"]), 0.1, c, 4, "image")
        assert np.array_equal(dets, g[f"k4_dets{c}"])
        assert np.array_equal(inds.cpu().numpy(), g[f"k4_inds{c}"]) and np.array_equal(keep, g[f"k4_keep{c}"])


@pytest.mark.parametrize("db_type,E,R", [("image", 4, 300), ("lidar", 7, 300), ("image", 4, 2000), ("lidar", 7, 2500)])
def test_final_detections_batched(db_type, E, R):
    """F frames x K classes in one launch vs the oracle per frame: ragged roi counts, max_dets with ties,
    an empty class, uncertainty gathers.  R = 2000 is cfg.TRAIN.RPN_POST_NMS_TOP_N (the RoIs a train-mode frame
    hands to filter_and_draw_prep); more than 1024 RoIs take the 1024-thread launch."""
    from faster_rcnn_pytorch_multimodal_b200 import ops
    F, K, max_dets = 3, 4, 20
    g = torch.Generator().manual_seed(77)
    probs = torch.softmax(torch.randn(F, R, K, generator=g) * 2.0, dim=2)
    probs[:, :, 3] = 0.01                                   # class 3 never passes the threshold
    probs[0, 10:40, 1] = 0.999                              # a run of tied top scores across the max_dets cut
    if db_type == "image":
        ctr = torch.rand(F, R, 1, 2, generator=g) * torch.tensor([1920.0, 1280.0])
        wh = torch.exp(torch.rand(F, R, K, 2, generator=g) * 2.0 + 3.0)
        boxes = torch.cat((ctr - wh / 2, ctr + wh / 2), dim=3)
        boxes[:, :6] += torch.tensor([-300.0, -300.0, 400.0, 400.0])
        info = torch.tensor([[0, 1920, 0, 1280, 0, 0, 1.0], [0, 1920, 0, 1280, 0, 0, 2.0], [10, 1000, 20, 700, 0, 0, 1.0]])
    else:
        ctr = torch.rand(F, R, 1, 2, generator=g) * torch.tensor([700.0, 800.0])
        size = torch.tensor([47.3, 20.8, 1.77]) * (0.8 + 0.4 * torch.rand(F, R, K, 3, generator=g))
        boxes = torch.cat((ctr.expand(F, R, K, 2), torch.rand(F, R, K, 1, generator=g), size,
                           torch.rand(F, R, K, 1, generator=g)), dim=3)
        info = torch.tensor([[0, 700, 0, 800, 0, 12, 1.0]]).repeat(F, 1)
    boxes = boxes.reshape(F, R, K * E).contiguous()
    num = torch.tensor([R, 157, 0], dtype=torch.int32)
    uc_row = torch.rand(F, R, 3, generator=g)
    uc_cls = torch.rand(F, R, 2, K * E, generator=g)
    dets, det_roi, counts, o_row, o_cls = ops.final_detections(probs.to(dev()), boxes.to(dev()), info.to(dev()), E,
                                                               db_type, 0.3, 0.6, max_dets=max_dets,
                                                               num_rois=num.to(dev()), uc_row=uc_row.to(dev()),
                                                               uc_cls=uc_cls.to(dev()))
    dets, det_roi, counts, o_row, o_cls = (t.cpu() for t in (dets, det_roi, counts, o_row, o_cls))
    for f in range(F):
        n = int(num[f])
        want = O.filter_detections(probs[f, :n], boxes[f, :n], info[f].numpy(), K, E, db_type, thresh=0.3, nms_thresh=0.6,
                                   max_dets=max_dets, uc_row=uc_row[f, :n], uc_cls=uc_cls[f, :n])
        assert int(counts[f, 0]) == 0
        for c in range(1, K):
            m = int(counts[f, c])
            w = want[c]
            assert m == len(w["dets"]), (f, c, m, len(w["dets"]))
            assert np.array_equal(dets[f, c, :m].numpy(), w["dets"])
            assert float(dets[f, c, m:].abs().sum()) == 0.0 and bool((det_roi[f, c, m:] == -1).all())
            if m:
                # ties (equal scores) may be ordered differently by torchvision's unstable sort: compare as sets there
                assert sorted(det_roi[f, c, :m].tolist()) == sorted(w["roi"].tolist())
                assert np.array_equal(o_row[f, c, :m].numpy(), uc_row[f][det_roi[f, c, :m].long()].numpy())
                got_cls = o_cls[f, c, :m].view(m, 2, E).numpy()
                assert np.array_equal(got_cls, uc_cls[f][det_roi[f, c, :m].long()][:, :, c * E:(c + 1) * E].numpy())
    assert int(counts[0, 1]) > max_dets                     # the tie run extends the cut
