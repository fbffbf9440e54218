"""The detection-glue part of the reference's ``Network`` base class.

``lib/nets/network.py`` is MISSING from the reference snapshot: it is imported at
nets/imagenet.py:11, nets/lidarnet.py:10, nets/fpn.py:5 but not shipped.  What can be pinned is the
contract of its four glue methods, from their callees' signatures and their call sites
(model/train_val.py:173-183,411-458; model/test.py:74-86); members marked [INFERRED] follow the
upstream project the reference's README names (ruotianluo/pytorch-faster-rcnn).

``Network`` here is an ``nn.Module`` base that owns exactly that glue: the anchor cache, the four
``_*_layer`` methods, the MC-dropout tail of ``test_frame`` and thin ``forward`` / ``test_frame`` /
``train_step`` / ``train_step_with_summary`` / ``run_eval`` drivers with the signatures the reference's
drivers call (model/test.py:74-86, model/train_val.py:411-412,449,458).  Backbone, RPN / head modules and
their weights stay with the host model (cuDNN / torch), which subclasses this and provides
``_input_to_head`` / ``_head_to_tail`` and the modules ``rpn_net, rpn_cls_score_net, rpn_bbox_pred_net,
cls_score_net, bbox_pred_net`` exactly as the reference's subclasses do (nets/vgg16.py:49-59,
nets/imagenet.py:65-91).
"""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from .. import ops
from ..layer_utils.anchor_target_layer import anchor_target_layer_torch
from ..layer_utils.generate_3d_anchors import GridAnchor3dGenerator
from ..layer_utils.proposal_layer import proposal_layer
from ..layer_utils.proposal_target_layer import proposal_target_layer
from ..layer_utils.proposal_top_layer import proposal_top_layer
from ..layer_utils.snippets import generate_anchors_pre
from ..model.config import cfg
from ..utils.bbox import bbaa_graphics_gems_torch
from ..utils.loss_utils import categorical_mutual_information, compute_bbox_var, mean_softmax_entropy
from ..utils.torchpoolers import MultiScaleRoIAlign


class Network(nn.Module):
    def __init__(self):
        super().__init__()
        self._predictions, self._anchor_targets, self._proposal_targets = {}, {}, {}
        self._layers = {}
        self._feat_stride = 16                       # imagenet.py:34-44 / lidarnet.py:31-48 override
        self._num_classes = None
        self._anchor_scales, self._anchor_ratios, self._num_anchors = None, None, None
        self._mode = 'TEST'
        self._device = 'cuda'
        self._fpn_en = False
        self._e_num_sample = 1
        self._roi_sampling_ratio = None              # None: cfg.POOLING_SAMPLING_RATIO (see _crop_pool_layer)
        self._rng_device = None                      # device of the samplers' torch generator (None: the tensors')
        self._losses = {}
        self._anchor_cache = {}
        self._msra = None
        self.timers = None                           # train_val.py:361

    # -- evidenced API ---------------------------------------------------------------------
    def create_architecture(self, num_classes, tag=None, anchor_scales=(8, 16, 32), anchor_ratios=(0.5, 1, 2)):
        """train_val.py:173-183, test_net.py:267-277.  Stores the anchor set; subclasses build modules."""
        self._tag = tag
        self._num_classes = num_classes
        self._anchor_scales = anchor_scales
        self._anchor_ratios = anchor_ratios
        if cfg.NET_TYPE == 'lidar':
            self._num_anchors = len(np.asarray(cfg.LIDAR.ANCHOR_SCALES).reshape(-1)) * len(cfg.LIDAR.ANCHOR_ANGLES)
            self._bbox_elem = cfg.LIDAR.NUM_BBOX_ELEM
        else:
            self._num_anchors = len(anchor_scales) * len(anchor_ratios)
            self._bbox_elem = cfg.IMAGE.NUM_BBOX_ELEM
        if hasattr(self, '_init_head_tail'):
            self._init_head_tail()
        if hasattr(self, 'init_weights'):
            self.init_weights()

    def set_e_num_sample(self, n):                   # test.py:74,77
        self._e_num_sample = int(n)

    # -- glue ------------------------------------------------------------------------------
    def _set_frame(self, info, gt_boxes=None, true_gt_boxes=None, gt_boxes_dc=None, mode='TEST'):
        """Per-frame state the glue methods read ([INFERRED] attribute names)."""
        self._info = info
        self._gt_boxes, self._true_gt_boxes, self._gt_boxes_dc = gt_boxes, true_gt_boxes, gt_boxes_dc
        self._mode = mode

    def _anchor_component(self, height, width):
        """Anchors for an (Hf, Wf) grid, cached per shape (the reference regenerates them on the host
        and uploads 0.75-3.8 MB every frame: snippets.py:13-40 / generate_3d_anchors.py:15-118)."""
        scale = float(self._info[6]) if len(self._info) > 6 else 1.0
        key = (cfg.NET_TYPE, int(height), int(width), self._feat_stride, scale)
        hit = self._anchor_cache.get(key)
        if hit is None:
            dev = torch.device(self._device)
            if cfg.NET_TYPE == 'lidar':
                n, a3d = GridAnchor3dGenerator()._generate(height, width, self._feat_stride,
                                                           np.asarray(cfg.LIDAR.ANCHOR_SCALES).reshape(-1),
                                                           cfg.LIDAR.ANCHOR_ANGLES, scale, device=dev)
                anchors = bbaa_graphics_gems_torch(a3d, width * self._feat_stride, height * self._feat_stride,
                                                   clip=False)
            else:
                anchors, n = generate_anchors_pre(height, width, self._feat_stride, self._anchor_scales,
                                                  self._anchor_ratios, scale, device=dev)
                a3d = torch.zeros(int(n), 7, device=dev)     # placeholder: proposal_layer.py:44 indexes it
            hit = (anchors, a3d, int(n))
            self._anchor_cache[key] = hit
        self._anchors, self._anchors_3d, self._anchor_length = hit
        return hit

    def _proposal_layer(self, rpn_cls_prob, rpn_bbox_pred):
        rois, rpn_scores, anchors_3d = proposal_layer(rpn_cls_prob, rpn_bbox_pred, self._info, self._mode,
                                                      self._anchors, self._anchors_3d, self._num_anchors)
        return rois, rpn_scores, anchors_3d

    def _proposal_top_layer(self, rpn_cls_prob, rpn_bbox_pred):
        rois, rpn_scores, anchors = proposal_top_layer(rpn_cls_prob, rpn_bbox_pred, self._info, self._anchors,
                                                       self._num_anchors)
        return rois, rpn_scores, anchors

    def _anchor_target_layer(self, rpn_cls_score):
        h, w = rpn_cls_score.shape[1], rpn_cls_score.shape[2]
        labels, targets, inside_w, outside_w = anchor_target_layer_torch(
            self._gt_boxes, self._gt_boxes_dc, self._info, self._anchors, self._num_anchors, h, w,
            self._rng_device or rpn_cls_score.device)
        self._anchor_targets.update(rpn_labels=labels.long(), rpn_bbox_targets=targets,
                                    rpn_bbox_inside_weights=inside_w, rpn_bbox_outside_weights=outside_w)
        return labels.long()

    def _proposal_target_layer(self, rois, roi_scores, anchors_3d):
        labels, rois, anchors_3d, roi_scores, targets, inside_w, outside_w = proposal_target_layer(
            rois, roi_scores, anchors_3d, self._gt_boxes, self._true_gt_boxes, self._gt_boxes_dc,
            self._num_classes, self._bbox_elem)
        self._proposal_targets.update(rois=rois, labels=labels.long(), bbox_targets=targets,
                                      bbox_inside_weights=inside_w, bbox_outside_weights=outside_w)
        return rois, roi_scores, anchors_3d

    def _crop_pool_layer(self, bottom, rois):
        """RoIAlign (aligned=False) to POOLING_SIZE^2.  The exact op of the missing method is not
        recoverable; RoIAlign is what the subclasses import (imagenet.py:15,
        lidarnet.py:16) and ``cfg.POOLING_MODE == 'multiscale'`` selects MultiScaleRoIAlign
        (trainval_net.py:326-330).  ``bottom`` is a tensor or an OrderedDict of FPN levels."""
        p = cfg.POOLING_SIZE
        sr = self._roi_sampling_ratio if self._roi_sampling_ratio is not None else cfg.POOLING_SAMPLING_RATIO
        if isinstance(bottom, (dict, OrderedDict)):
            if self._msra is None:
                self._msra = MultiScaleRoIAlign(list(bottom.keys()), p, sr)
            h = float(self._info[3] - self._info[2])
            w = float(self._info[1] - self._info[0])
            return self._msra(bottom, [rois[:, 1:5]], [(int(h), int(w))])
        return ops.roi_align(bottom, rois, (p, p), 1.0 / self._feat_stride, sr, False)

    def _region_proposal(self, rpn_cls_prob, rpn_bbox_pred, rpn_cls_score=None):
        """[INFERRED] ordering of the glue inside the reference's ``_region_proposal``."""
        if self._mode == 'TRAIN':
            rois, roi_scores, a3d = self._proposal_layer(rpn_cls_prob, rpn_bbox_pred)
            self._anchor_target_layer(rpn_cls_score if rpn_cls_score is not None else rpn_cls_prob)
            rois, roi_scores, a3d = self._proposal_target_layer(rois, roi_scores, a3d)
        elif cfg.TEST.MODE == 'top':
            rois, roi_scores, a3d = self._proposal_top_layer(rpn_cls_prob, rpn_bbox_pred)
        else:
            rois, roi_scores, a3d = self._proposal_layer(rpn_cls_prob, rpn_bbox_pred)
        self._predictions.update(rois=rois, anchors_3d=a3d)
        return rois, roi_scores, a3d

    # -- drivers ([INFERRED] bodies; signatures and return values per the reference's call sites) -------------
    def _run_rpn(self, rpn_feat):
        """RPN head on one feature map -> (rpn_cls_score, rpn_cls_prob [1,H,W,2A], rpn_bbox_pred [1,H,W,4A]).
        Channel layout of the reference's RPN: A background scores then A foreground scores
        (proposal_layer.py:32 reads the back half), softmax over each (bg, fg) pair."""
        rpn = F.relu(self.rpn_net(rpn_feat))
        score = self.rpn_cls_score_net(rpn)                                   # [1, 2A, H, W]
        n, _, h, w = score.shape
        pair = score.view(n, 2, self._num_anchors, h, w).softmax(dim=1)
        prob = pair.reshape(n, 2 * self._num_anchors, h, w).permute(0, 2, 3, 1).contiguous()
        pred = self.rpn_bbox_pred_net(rpn).permute(0, 2, 3, 1).contiguous()
        return score.permute(0, 2, 3, 1).contiguous(), prob, pred

    def forward(self, data, info, gt_boxes=None, true_gt_boxes=None, gt_boxes_dc=None, mode='TEST'):
        """One frame through head -> RPN -> proposal / target layers -> RoI crop -> T tail passes.

        Returns nothing; fills ``self._predictions`` (rois, anchors_3d, cls_score [T,R,K], bbox_pred [T,R,K*E],
        pool5) and, in TRAIN mode, ``self._anchor_targets`` / ``self._proposal_targets``."""
        dev = data.device
        info_t = torch.as_tensor(np.asarray(info, dtype=np.float32)) if not isinstance(info, torch.Tensor) else info
        self._set_frame(info_t.reshape(-1).float().cpu(), gt_boxes, true_gt_boxes,
                        gt_boxes_dc if gt_boxes_dc is not None else torch.zeros(0, 5, device=dev), mode)
        self._device = dev
        net_conv = self._input_to_head(data)
        rpn_feat = net_conv if torch.is_tensor(net_conv) else next(iter(net_conv.values()))
        self._anchor_component(rpn_feat.shape[2], rpn_feat.shape[3])
        rpn_cls_score, rpn_cls_prob, rpn_bbox_pred = self._run_rpn(rpn_feat)
        with torch.no_grad():                                                # the reference detaches here too (.data)
            rois, roi_scores, a3d = self._region_proposal(rpn_cls_prob, rpn_bbox_pred, rpn_cls_score)
        pool5 = self._crop_pool_layer(net_conv, rois)
        cls, box = [], []
        for _ in range(max(1, self._e_num_sample if mode == 'TEST' else 1)):  # MC-dropout: T tail passes
            fc7 = self._head_to_tail(pool5)
            cls.append(self.cls_score_net(fc7))
            box.append(self.bbox_pred_net(fc7))
        self._predictions.update(rpn_cls_score=rpn_cls_score, rpn_cls_prob=rpn_cls_prob, rpn_bbox_pred=rpn_bbox_pred,
                                 rois=rois, anchors_3d=a3d, pool5=pool5, cls_score=torch.stack(cls),
                                 bbox_pred=torch.stack(box))

    def test_frame(self, data, info):
        """model/test.py:75: -> (cls_score, probs [R,K], bbox_pred [R,K*E] ALREADY DECODED, rois [R,5],
        uncertainties{a_entropy, a_mutual_info, a_cls_var, e_entropy, e_mutual_info, e_cls_var, a_bbox_var,
        e_bbox_var}) (keys per filter_predictions.py:113-124).  Decode, variance and entropy / MI are ONE launch."""
        self.eval()
        if self._dropout_en_at_test():
            for m in self.modules():
                if isinstance(m, (nn.Dropout, nn.Dropout2d)):
                    m.train()
        with torch.no_grad():
            self.forward(data, info, mode='TEST')
            p = self._predictions
            rois, a3d = p['rois'], p['anchors_3d']
            dev = rois.device
            info_t = self._info.to(dev).view(1, -1)
            if info_t.shape[1] < 7:
                info_t = torch.cat((info_t, info_t.new_zeros(1, 7 - info_t.shape[1])), 1)
                info_t[0, 6] = 1.0
            lidar = cfg.NET_TYPE == 'lidar'
            a_var = p.get('a_bbox_var')
            out = ops.head_tail_decode(p['bbox_pred'].unsqueeze(1), p['cls_score'].unsqueeze(1), rois.unsqueeze(0),
                                       a3d.unsqueeze(0) if lidar else None, info_t, 'lidar' if lidar else 'image',
                                       a_bbox_var=None if a_var is None else a_var.unsqueeze(0), use_scale=True)
            zeros = rois.new_zeros(rois.shape[0])
            T = p['cls_score'].shape[0]
            probs_t = p['cls_score'].softmax(dim=2)
            unc = {'a_entropy': zeros, 'a_mutual_info': zeros, 'a_cls_var': torch.zeros_like(out['probs'][0]),
                   'e_entropy': out['e_entropy'][0], 'e_mutual_info': out['e_mutual_info'][0],
                   'e_cls_var': compute_bbox_var(probs_t) if T > 1 else torch.zeros_like(out['probs'][0]),
                   'a_bbox_var': out['a_bbox_var'][0] if 'a_bbox_var' in out else torch.zeros_like(out['boxes'][0]),
                   'e_bbox_var': out['e_bbox_var'][0]}
        return p['cls_score'].mean(0), out['probs'][0], out['boxes'][0], rois, unc

    def _dropout_en_at_test(self):
        return self._e_num_sample > 1 and (cfg.UC.EN_BBOX_EPISTEMIC or cfg.UC.EN_CLS_EPISTEMIC)

    def _add_losses(self):
        """The four standard Faster R-CNN losses over the tensors the glue produced (the reference's own loss
        assembly is in the missing file; losses are host-model territory, this default keeps train_step usable)."""
        p, at, pt = self._predictions, self._anchor_targets, self._proposal_targets
        A = self._num_anchors
        score = p['rpn_cls_score']                                            # [1,H,W,2A]: A bg then A fg
        n, h, w, _ = score.shape
        logits = torch.stack((score[..., :A], score[..., A:]), -1).reshape(-1, 2)
        labels = at['rpn_labels'].permute(0, 2, 3, 1).reshape(-1)            # [1,A,H,W] -> (h, w, a)
        sel = labels >= 0
        rpn_ce = F.cross_entropy(logits[sel], labels[sel])
        diff = at['rpn_bbox_inside_weights'] * (p['rpn_bbox_pred'] - at['rpn_bbox_targets'])
        rpn_box = (at['rpn_bbox_outside_weights'] * F.smooth_l1_loss(diff, torch.zeros_like(diff), reduction='none',
                                                                     beta=1.0 / 9)).sum()
        cls_score, bbox_pred = p['cls_score'][0], p['bbox_pred'][0]
        ce = F.cross_entropy(cls_score, pt['labels'].view(-1))
        d2 = pt['bbox_inside_weights'] * (bbox_pred - pt['bbox_targets'])
        box = (pt['bbox_outside_weights'] * F.smooth_l1_loss(d2, torch.zeros_like(d2), reduction='none')).sum(1).mean()
        self._losses = dict(rpn_cross_entropy=rpn_ce, rpn_loss_box=rpn_box, cross_entropy=ce, loss_box=box,
                            total_loss=rpn_ce + rpn_box + ce + box)
        return self._losses['total_loss']

    def _blobs(self, blobs):
        dev = next(self.parameters()).device
        t = lambda v: v if torch.is_tensor(v) else torch.as_tensor(np.ascontiguousarray(v), dtype=torch.float32)
        g = lambda k: t(blobs[k]).to(dev) if blobs.get(k) is not None else None
        return g('data'), blobs['info'], g('gt_boxes'), g('true_gt_boxes') if 'true_gt_boxes' in blobs else None, \
            g('gt_boxes_dc')

    def train_step(self, blobs, optimizer, update_weights=True):
        """model/train_val.py:458: -> total loss (python float)."""
        self.train()
        data, info, gt, tgt, dc = self._blobs(blobs)
        if tgt is None:
            tgt = gt.new_zeros(gt.shape[0], 8)
        self.forward(data, info, gt, tgt, dc, mode='TRAIN')
        loss = self._add_losses()
        loss.backward()
        if update_weights:
            optimizer.step()
            optimizer.zero_grad()
        return float(loss.item())

    def train_step_with_summary(self, blobs, optimizer, sum_size=1, update_weights=True):
        """model/train_val.py:449: -> (total loss, [summaries]); summaries are the host model's business."""
        return self.train_step(blobs, optimizer, update_weights), []

    def run_eval(self, blobs, val_batch_size=1, update_summaries=False):
        """model/train_val.py:411-412: -> (summary, rois, roi_labels, cls_prob, bbox_pred, uncertainties)."""
        data, info, gt, _, _ = self._blobs(blobs)
        _, probs, boxes, rois, unc = self.test_frame(data, info)
        return [], rois, gt, probs, boxes, unc

    # -- MC-dropout reductions of test_frame (test.py:75; keys per filter_predictions.py:113-124) -----
    @staticmethod
    def epistemic_uncertainties(bbox_samples, cls_score_samples):
        """bbox_samples [T,R,K*E], cls_score_samples [T,R,K] -> dict with the reference's key names."""
        return {'e_bbox_var': compute_bbox_var(bbox_samples),
                'e_mutual_info': categorical_mutual_information(cls_score_samples),
                'e_entropy': mean_softmax_entropy(cls_score_samples)}
