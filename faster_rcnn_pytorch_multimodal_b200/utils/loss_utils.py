"""MC-dropout reductions with the reference's signatures (lib/utils/loss_utils.py:103-141).
Only the reductions are here; the losses of that file are autograd elementwise code left to torch."""
import torch

from .._lib import check, f32c, lib, ptr, require_cuda, stream_ptr


def _variance(bbox_samples, mode):
    require_cuda(bbox_samples)
    x = f32c(bbox_samples)
    T = x.shape[0]
    m = x[0].numel()
    out = torch.empty(x.shape[1:], device=x.device)
    check(lib(x.device).b2d_mc_variance(T, m, ptr(x), mode, ptr(out), stream_ptr(x.device)), "b2d_mc_variance")
    return out


def compute_bbox_var(bbox_samples):
    """loss_utils.py:114-120: unbiased single-pass variance over dim 0 of [T,R,K*E], clamp_min(0)."""
    return _variance(bbox_samples, 0)


def compute_bbox_cov(bbox_samples):
    """loss_utils.py:103-112: diagonal of E[xx^T] - mu mu^T (biased), clamp_min(0)."""
    return _variance(bbox_samples, 1)


def _class_uncertainty(cls_score, want_mi, want_ent):
    require_cuda(cls_score)
    z = f32c(cls_score)
    T, n, K = z.shape
    mi = torch.empty(n, device=z.device) if want_mi else None
    ent = torch.empty(n, device=z.device) if want_ent else None
    check(lib(z.device).b2d_mc_class_uncertainty(T, n, K, ptr(z), ptr(mi), ptr(ent), stream_ptr(z.device)),
          "b2d_mc_class_uncertainty")
    return mi, ent


def categorical_mutual_information(cls_score):
    """loss_utils.py:132-141: [T,N,C] logits -> [N]."""
    return _class_uncertainty(cls_score, True, False)[0]


def categorical_entropy(cls_prob):
    """loss_utils.py:122-129: [N,C] probabilities -> [N] (elementwise; torch is already one fused pass)."""
    return -torch.sum(cls_prob * torch.log2(cls_prob), dim=1)


def mean_softmax_entropy(cls_score):
    """Entropy of the T-averaged softmax, the `total_entropy` term of :135-136, as its own output."""
    return _class_uncertainty(cls_score, False, True)[1]


def sort_by_bbox_variance(var, descending=False):
    """argsort of the row-mean variance (datasets/db.py:264-303: np.mean(uc, axis=1) then argsort),
    ties broken by lower index.  -> (order int64 [n], key fp32 [n])."""
    require_cuda(var)
    v = f32c(var.reshape(var.shape[0], -1))
    n, cols = v.shape
    key = torch.empty(n, device=v.device)
    order = torch.empty(n, dtype=torch.int32, device=v.device)
    check(lib(v.device).b2d_var_sort(n, cols, ptr(v), int(bool(descending)), ptr(key), ptr(order), stream_ptr(v.device)),
          "b2d_var_sort")
    return order.long(), key
