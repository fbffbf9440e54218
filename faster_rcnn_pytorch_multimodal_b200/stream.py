"""Frame-stream sharding across GPUs and the end-of-stream detection gather.

The reference processes one frame at a time on one GPU (``assert num_frames == 1``,
roi_data_layer/minibatch.py:111) and has no distributed code at all.  Frames are independent, so the
B200 design is one process per GPU, each owning a contiguous-stride stream of frames (rank r takes
frames r, r+W, ...), batching ``frames_per_call`` of them into every C-ABI call.  Nothing is
exchanged on the data path; the only collective is one all_gather of fixed-size padded detection
records that rebuilds the reference's ``all_boxes[cls][frame]`` order (model/test.py:162-163,226).
"""
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_frames(num_frames: int, rank: int, world_size: int) -> List[int]:
    """Frame indices owned by ``rank`` (round-robin keeps per-rank load even for ragged tails)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return list(range(rank, num_frames, world_size))


def batches(indices: Sequence[int], frames_per_call: int) -> List[List[int]]:
    return [list(indices[i:i + frames_per_call]) for i in range(0, len(indices), frames_per_call)]


def pack_records(rois: torch.Tensor, scores: torch.Tensor, num_out: torch.Tensor,
                 frame_ids: Sequence[int]) -> torch.Tensor:
    """[F,M,5] rois + [F,M] scores + [F] counts + frame ids -> one fp32 record tensor [F, M*6 + 2]."""
    F, M = scores.shape
    ids = torch.as_tensor(list(frame_ids), dtype=torch.float32, device=rois.device).view(F, 1)
    return torch.cat((ids, num_out.view(F, 1).float(), rois.reshape(F, M * 5), scores.reshape(F, M)), dim=1)


def unpack_records(rec: torch.Tensor, M: int) -> List[Tuple[int, torch.Tensor, torch.Tensor]]:
    out = []
    for row in rec:
        fid, n = int(row[0].item()), int(row[1].item())
        if fid < 0:
            continue                                   # padding row of a rank with fewer frames
        rois = row[2:2 + M * 5].view(M, 5)[:n]
        scores = row[2 + M * 5:2 + M * 6][:n]
        out.append((fid, rois, scores))
    return out


def pack_detection_records(dets: torch.Tensor, counts: torch.Tensor, frame_ids: Sequence[int],
                           uc_row: Optional[torch.Tensor] = None, uc_cls: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The wire format of the end-of-stream gather: one fixed-size fp32 record per frame holding what
    ``ops.final_detections`` emits - dets [F,K,D,E+1] (box, score), counts [F,K], and the gathered uncertainty
    columns uc_row [F,K,D,U] / uc_cls [F,K,D,U2*E] in the reference's hstack order (model/test.py:260-270) ->
    [F, 2 + K + K*D*(E+1+U+U2*E)].  Column 0 is the frame id, column 1 the row width per detection."""
    F, K, D, E1 = dets.shape
    parts = [dets]
    if uc_row is not None:
        parts.append(uc_row)
    if uc_cls is not None:
        parts.append(uc_cls)
    rows = torch.cat(parts, dim=3)                                     # [F,K,D,width]: bbox | score | uncertainties
    width = rows.shape[3]
    ids = torch.as_tensor(list(frame_ids), dtype=torch.float32, device=dets.device).view(F, 1)
    return torch.cat((ids, torch.full((F, 1), float(width), device=dets.device), counts.view(F, K).float(),
                      rows.reshape(F, K * D * width)), dim=1)


def unpack_detection_records(rec: torch.Tensor, num_classes: int, num_frames: int):
    """Gathered records -> the reference's ``all_boxes[cls][frame]`` (model/test.py:162-163,226): a list over
    classes of lists over frames of numpy arrays [m, E+1+uncertainties] (``np.empty(0)`` where a class has no
    detection, as the reference stores it).  Class 0 (background) stays empty lists."""
    import numpy as np
    K = num_classes
    all_boxes = [[np.empty(0) for _ in range(num_frames)] for _ in range(K)]
    rec = rec.cpu()
    for row in rec:
        fid = int(row[0].item())
        if fid < 0 or fid >= num_frames:
            continue
        width = int(row[1].item())
        counts = row[2:2 + K].to(torch.int64)
        body = row[2 + K:].view(K, -1, width)
        for j in range(1, K):
            m = int(counts[j])
            if m > 0:
                all_boxes[j][fid] = body[j, :m].numpy().copy()
    return all_boxes


def gather_detections(local_records: torch.Tensor, max_frames_per_rank: int,
                      group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """all_gather of per-rank record blocks, padded to ``max_frames_per_rank`` rows (frame id -1);
    returns the records of every rank, sorted by frame id, on every rank."""
    width = local_records.shape[1]
    pad = torch.full((max_frames_per_rank, width), -1.0, dtype=local_records.dtype, device=local_records.device)
    pad[:local_records.shape[0]] = local_records
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        world = dist.get_world_size(group)
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        allrec = torch.cat(parts, dim=0)
    else:
        allrec = pad
    allrec = allrec[allrec[:, 0] >= 0]
    return allrec[torch.argsort(allrec[:, 0], stable=True)]


class FrameStream:
    """Runs proposal + RoI crop over this rank's frames in batches (device tensors in, device out)."""

    def __init__(self, anchors, num_anchors, pre_nms, post_nms, nms_thresh, pooled=7, spatial_scale=1.0 / 16,
                 sampling_ratio=2, rank=0, world_size=1, frames_per_call=8):
        self.anchors, self.A = anchors, num_anchors
        self.pre, self.post, self.thr = pre_nms, post_nms, nms_thresh
        self.pooled, self.scale, self.sr = pooled, spatial_scale, sampling_ratio
        self.rank, self.world, self.fpc = rank, world_size, frames_per_call

    def run_batch(self, cls_prob, bbox_pred, info, feat):
        from . import ops
        rois, scores, _, _, num = ops.proposal_batched(cls_prob, bbox_pred, info, self.anchors, None, self.A,
                                                       self.pre, self.post, self.thr, batch_index_stride=1)
        pooled = ops.roi_align(feat, rois.view(-1, 5), (self.pooled, self.pooled), self.scale, self.sr, False,
                               seg_count=num, seg_stride=rois.shape[1])
        return rois, scores, num, pooled
