"""The detection-glue part of the reference's ``Network`` base class.

``lib/nets/network.py`` is MISSING from the reference snapshot: it is imported at
nets/imagenet.py:11, nets/lidarnet.py:10, nets/fpn.py:5 but not shipped.  What can be pinned is the
contract of its four glue methods, from their callees' signatures and their call sites
(model/train_val.py:173-183,411-458; model/test.py:74-86); members marked [INFERRED] follow the
upstream project the reference's README names (ruotianluo/pytorch-faster-rcnn).

``Network`` here is an ``nn.Module`` mix-in that owns exactly that glue: the anchor cache, the four
``_*_layer`` methods, and the MC-dropout reductions of ``test_frame``.  Backbone, RPN/head modules,
losses and the train/test drivers stay with the host model (cuDNN / torch), which subclasses this
and provides ``_input_to_head`` / ``_head_to_tail`` exactly as the reference's subclasses do
(nets/vgg16.py:49-59).
"""
from collections import OrderedDict

import numpy as np
import torch
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
        self._roi_sampling_ratio = 2                 # [INFERRED] see _crop_pool_layer
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
            rpn_cls_score.device)
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
        if isinstance(bottom, (dict, OrderedDict)):
            if self._msra is None:
                self._msra = MultiScaleRoIAlign(list(bottom.keys()), p, self._roi_sampling_ratio)
            h = float(self._info[3] - self._info[2])
            w = float(self._info[1] - self._info[0])
            return self._msra(bottom, [rois[:, 1:5]], [(int(h), int(w))])
        return ops.roi_align(bottom, rois, (p, p), 1.0 / self._feat_stride, self._roi_sampling_ratio, False)

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

    # -- MC-dropout reductions of test_frame (test.py:75; keys per filter_predictions.py:113-124) -----
    @staticmethod
    def epistemic_uncertainties(bbox_samples, cls_score_samples):
        """bbox_samples [T,R,K*E], cls_score_samples [T,R,K] -> dict with the reference's key names."""
        return {'e_bbox_var': compute_bbox_var(bbox_samples),
                'e_mutual_info': categorical_mutual_information(cls_score_samples),
                'e_entropy': mean_softmax_entropy(cls_score_samples)}
