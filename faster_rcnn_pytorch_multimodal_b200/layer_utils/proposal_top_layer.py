"""proposal_top_layer (lib/layer_utils/proposal_top_layer.py:18-59): TEST.MODE == 'top'."""
import numpy.random as npr
import torch

from ..model.bbox_transform import bbox_transform_inv, clip_boxes
from ..model.config import cfg
from ..ops import proposal_top_batched
from .proposal_layer import _info_tensor


def proposal_top_layer(rpn_cls_prob, rpn_bbox_pred, info, anchors, num_anchors):
    rpn_top_n = cfg.TEST.RPN_TOP_N
    dev = rpn_cls_prob.device
    length = rpn_cls_prob[..., num_anchors:].numel()
    if length < rpn_top_n:
        # rare branch (:33-38): random fill with replacement from the host RNG, as the reference
        scores = rpn_cls_prob[:, :, :, num_anchors:].contiguous().view(-1, 1)
        top = torch.from_numpy(npr.choice(length, size=rpn_top_n, replace=True)).long().to(dev)
        anc = anchors[top, :].contiguous()
        props = clip_boxes(bbox_transform_inv(anc, rpn_bbox_pred.view(-1, 4)[top, :].contiguous()), info)
        blob = torch.cat([props.new_zeros(props.size(0), 1), props], 1)
        return blob, scores[top].contiguous(), anc
    rois, scores, anc = proposal_top_batched(rpn_cls_prob, rpn_bbox_pred, _info_tensor(info, dev), anchors,
                                             num_anchors, rpn_top_n, batch_index_stride=0)
    return rois[0], scores[0].view(-1, 1), anc[0]
