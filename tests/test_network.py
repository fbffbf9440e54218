"""The `Network` base: a tiny host subclass - conv-stub `_input_to_head` / `_head_to_tail`, the
reference's module names - runs TRAIN and TEST, image and lidar, tensor and FPN-dict `bottom`, through
`forward` / `test_frame` / `train_step` / `run_eval`, and every glue product is compared with the oracle chain
fed with the same RPN / head tensors."""
from collections import OrderedDict

import numpy as np
import pytest
import torch
from torch import nn

from oracle import glue_oracle as O

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda", 0)


def make_net(net_type, fpn=False, K=3, C=32):
    from faster_rcnn_pytorch_multimodal_b200.model.config import cfg
    from faster_rcnn_pytorch_multimodal_b200.nets.network import Network
    cfg.NET_TYPE = net_type

    class Tiny(Network):
        def _init_head_tail(self):
            E = self._bbox_elem
            cin = 3 if net_type == "image" else 15
            self.stem = nn.Conv2d(cin, C, 3, padding=1)
            self.rpn_net = nn.Conv2d(C, C, 3, padding=1)
            self.rpn_cls_score_net = nn.Conv2d(C, 2 * self._num_anchors, 1)
            self.rpn_bbox_pred_net = nn.Conv2d(C, 4 * self._num_anchors, 1)
            self.fc = nn.Linear(C * 49, 64)
            self.drop = nn.Dropout(0.3)
            self.cls_score_net = nn.Linear(64, K)
            self.bbox_pred_net = nn.Linear(64, K * E)

        def init_weights(self):
            g = torch.Generator().manual_seed(3)
            for p in self.parameters():
                p.data = (torch.randn(p.shape, generator=g) * (0.05 if p.dim() > 1 else 0.01))
            self.rpn_bbox_pred_net.weight.data *= 0.2

        def _input_to_head(self, data):
            x = torch.relu(self.stem(data))
            c4 = nn.functional.avg_pool2d(x, 16)
            if not fpn:
                return c4
            return OrderedDict(p4=c4, p5=nn.functional.avg_pool2d(x, 32))

        def _head_to_tail(self, pool5):
            return self.drop(torch.relu(self.fc(pool5.flatten(1))))

    net = Tiny()
    if fpn:
        net._feat_stride = 16
    net.create_architecture(K, tag="t", anchor_scales=(2, 4, 8, 16, 32), anchor_ratios=(0.5, 0.75, 1, 1.25, 2))
    return net.to(dev()), cfg


def frame(net_type, seed=0):
    g = torch.Generator().manual_seed(seed)
    if net_type == "image":
        H, W = 384, 640
        data = torch.randn(1, 3, H, W, generator=g)
        info = np.array([0, W, 0, H, 0, 0, 1.0], dtype=np.float32)
    else:
        H, W = 480, 352
        data = torch.rand(1, 15, H, W, generator=g)
        info = np.array([0, W, 0, H, 0, 12, 1.0], dtype=np.float32)
    xy = torch.rand(6, 2, generator=g) * torch.tensor([W - 140.0, H - 140.0])
    wh = torch.rand(6, 2, generator=g) * 100 + 30
    gt = torch.cat((xy, xy + wh, torch.randint(1, 3, (6, 1), generator=g).float()), 1)
    true_gt = torch.cat((xy + wh / 2, torch.full((6, 1), 0.9), wh, torch.full((6, 1), 1.7), torch.zeros(6, 1), gt[:, 4:]), 1)
    return data, info, gt, true_gt


@pytest.mark.parametrize("net_type,fpn", [("image", False), ("lidar", False), ("image", True)])
def test_test_frame_matches_the_oracle_chain(net_type, fpn):
    net, cfg = make_net(net_type, fpn)
    try:
        data, info, _, _ = frame(net_type)
        net.set_e_num_sample(1)
        _, probs, boxes, rois, unc = net.test_frame(data.to(dev()), info)
        p = net._predictions
        A, E = net._num_anchors, net._bbox_elem
        # proposal layer: oracle on the SAME RPN tensors (stable sort = our tie rule)
        anchors, a3d = net._anchors.cpu(), net._anchors_3d.cpu()
        ocfg = O.GlueCfg(net_type=net_type)
        blob, sc, a3k = O.proposal_layer(p["rpn_cls_prob"].cpu(), p["rpn_bbox_pred"].cpu(), info, "TEST", anchors, a3d, A,
                                         cfg=ocfg, stable_sort=True)
        assert rois.shape == blob.shape and rois.shape[0] > 10
        assert torch.allclose(rois.cpu(), blob, rtol=1e-5, atol=1e-3)
        if net_type == "lidar":
            assert torch.equal(p["anchors_3d"].cpu(), a3k)
        # RoI crop on OUR rois
        conv = net._input_to_head(data.to(dev()))
        if fpn:
            want = O.multiscale_roi_align([v.cpu() for v in conv.values()], rois[:, 1:5].cpu(), (int(info[3]), int(info[1])))
        else:
            want = O.roi_align(conv.cpu(), rois.cpu(), (7, 7), 1.0 / 16, cfg.POOLING_SAMPLING_RATIO, False)
        assert torch.allclose(p["pool5"].cpu(), want, rtol=1e-5, atol=1e-5)
        # head tail on OUR head outputs
        o = O.head_tail_decode(p["bbox_pred"].cpu(), p["cls_score"].cpu(), rois.cpu(),
                               p["anchors_3d"].cpu() if net_type == "lidar" else None, info, net_type, use_scale=True, cfg=ocfg)
        assert torch.allclose(boxes.cpu(), o["boxes"], rtol=1e-5, atol=1e-3)
        assert torch.allclose(probs.cpu(), o["probs"], rtol=1e-5, atol=1e-6)
        assert set(unc) == {"a_entropy", "a_mutual_info", "a_cls_var", "e_entropy", "e_mutual_info", "e_cls_var",
                            "a_bbox_var", "e_bbox_var"}
        assert boxes.shape == (rois.shape[0], 3 * E) and float(unc["e_bbox_var"].abs().max()) == 0.0
        # MC-dropout: T = 8 passes, epistemic outputs against the oracle on the stacked samples
        cfg.UC.EN_BBOX_EPISTEMIC = True
        net.set_e_num_sample(8)
        torch.manual_seed(11)
        _, probs, boxes, rois, unc = net.test_frame(data.to(dev()), info)
        p = net._predictions
        assert p["bbox_pred"].shape[0] == 8 and float(p["bbox_pred"].std(0).mean()) > 0
        o = O.head_tail_decode(p["bbox_pred"].cpu(), p["cls_score"].cpu(), rois.cpu(),
                               p["anchors_3d"].cpu() if net_type == "lidar" else None, info, net_type, use_scale=True, cfg=ocfg)
        assert torch.allclose(boxes.cpu(), o["boxes"], rtol=1e-5, atol=1e-3)
        assert torch.allclose(unc["e_bbox_var"].cpu(), o["e_bbox_var"], rtol=1e-4, atol=1e-6)
        assert torch.allclose(unc["e_mutual_info"].cpu(), o["e_mutual_info"], rtol=1e-4, atol=1e-5)
        assert torch.allclose(unc["e_entropy"].cpu(), o["e_entropy"], rtol=1e-4, atol=1e-5)
        # and the reference's consumer of test_frame's outputs runs on them
        from faster_rcnn_pytorch_multimodal_b200.utils.filter_predictions import filter_and_draw_prep
        r2, all_boxes, all_unc = filter_and_draw_prep(rois, probs, boxes, unc, info, 3, 0.05, net_type)
        assert len(all_boxes) == 3
    finally:
        cfg.NET_TYPE, cfg.UC.EN_BBOX_EPISTEMIC = "lidar", False


@pytest.mark.parametrize("net_type", ["image", "lidar"])
def test_train_step_targets_match_the_oracle(net_type):
    from faster_rcnn_pytorch_multimodal_b200.layer_utils import proposal_target_layer as ptl
    net, cfg = make_net(net_type)
    ptl.RNG_DEVICE = "cpu"
    net._rng_device = torch.device("cpu")
    cfg.TRAIN.USE_GT = True          # GT boxes join the RoIs (proposal_target_layer.py:35-41): guarantees foreground samples
    try:
        data, info, gt, true_gt = frame(net_type, 1)
        opt = torch.optim.SGD(net.parameters(), lr=1e-3)
        blobs = dict(data=data, info=info, gt_boxes=gt, true_gt_boxes=true_gt, gt_boxes_dc=np.zeros((0, 5), np.float32))
        w0 = net.fc.weight.detach().clone()
        torch.manual_seed(7)
        loss = net.train_step(blobs, opt, True)
        assert np.isfinite(loss) and not torch.equal(w0, net.fc.weight.detach())
        p, at, pt = net._predictions, net._anchor_targets, net._proposal_targets
        A, E, K = net._num_anchors, net._bbox_elem, 3
        Hf, Wf = p["rpn_cls_prob"].shape[1:3]
        ocfg = O.GlueCfg(net_type=net_type, use_gt=True)
        torch.manual_seed(7)
        blob, sc, a3k = O.proposal_layer(p["rpn_cls_prob"].detach().cpu(), p["rpn_bbox_pred"].detach().cpu(), info, "TRAIN",
                                         net._anchors.cpu(), net._anchors_3d.cpu(), A, cfg=ocfg, stable_sort=True)
        lab, tg, iw, ow = O.anchor_target_layer(gt, torch.zeros(0, 5), info, net._anchors.cpu(), A, Hf, Wf, cfg=ocfg)
        assert torch.equal(at["rpn_labels"].cpu(), lab.long()), "anchor labels"
        assert torch.allclose(at["rpn_bbox_targets"].cpu(), tg, rtol=1e-5, atol=1e-5)
        out = O.proposal_target_layer(blob, sc, a3k, gt, true_gt, torch.zeros(0, 5), K, E, cfg=ocfg)
        assert torch.equal(pt["labels"].cpu().view(-1), out[0].long().view(-1)), "sampled RoI labels"
        assert torch.allclose(pt["rois"].cpu(), out[1], rtol=1e-5, atol=1e-3)
        assert torch.allclose(pt["bbox_targets"].cpu(), out[4], rtol=1e-4, atol=1e-4)
        # run_eval / train_step_with_summary keep the reference's return shapes
        l2, summ = net.train_step_with_summary(blobs, opt, 1, False)
        assert np.isfinite(l2) and summ == []
        s, rois, roi_labels, cls_prob, bbox_pred, unc = net.run_eval(blobs, 1, False)
        assert rois.shape[1] == 5 and cls_prob.shape == (rois.shape[0], K) and bbox_pred.shape == (rois.shape[0], K * E)
    finally:
        ptl.RNG_DEVICE = None
        cfg.NET_TYPE, cfg.TRAIN.USE_GT = "lidar", False
