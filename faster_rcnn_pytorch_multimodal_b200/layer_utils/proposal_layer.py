"""proposal_layer with the reference's signature (lib/layer_utils/proposal_layer.py:18-57).

One C-ABI call (b2d_proposal) replaces: fg-score slice, bbox_transform_inv + clip_boxes over
all N anchors, the full descending sort, torchvision.ops.nms and the three post-NMS gathers.
"""
import numpy as np
import torch

from ..model.config import cfg
from ..ops import proposal_batched


def _info_tensor(info, device, frames=1):
    if isinstance(info, torch.Tensor):
        t = info.to(device=device, dtype=torch.float32)
    else:
        t = torch.as_tensor(np.asarray(info, dtype=np.float32), device=device)
    return t.reshape(frames, -1)


def proposal_layer(rpn_cls_prob, rpn_bbox_pred, info, cfg_key, anchors, anchors_3d, num_anchors):
    """rpn_cls_prob [1,H,W,2A], rpn_bbox_pred [1,H,W,4A] -> (blob [R,5], scores [R,1], anchors_3d [R,7]).

    R <= RPN_POST_NMS_TOP_N is data dependent, so this wrapper reads one int back (the
    reference syncs here as well, inside torchvision.ops.nms).  Equal scores are ordered by
    lower flat anchor index first (the reference's unstable sort leaves that order undefined).
    """
    if type(cfg_key) == bytes:                                   # :25-26
        cfg_key = cfg_key.decode('utf-8')
    pre = cfg[cfg_key].RPN_PRE_NMS_TOP_N
    post = cfg[cfg_key].RPN_POST_NMS_TOP_N
    thr = cfg[cfg_key].RPN_NMS_THRESH
    dev = rpn_cls_prob.device
    info_t = _info_tensor(info, dev)
    rois, scores, a3d, _, num = proposal_batched(rpn_cls_prob, rpn_bbox_pred, info_t, anchors, anchors_3d,
                                                 num_anchors, pre, post, thr, batch_index_stride=0)
    r = int(num[0].item())
    blob = rois[0, :r]
    out_a3d = a3d[0, :r] if a3d is not None else None
    return blob, scores[0, :r].view(-1, 1), out_a3d
