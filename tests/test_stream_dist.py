"""Host-side logic of the N>1 path: frame sharding and the detection gather, world_size 2 on gloo."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from faster_rcnn_pytorch_multimodal_b200 import stream


def test_shard_partition_is_exact():
    for n in (0, 1, 7, 16, 33):
        for w in (1, 2, 4, 8):
            seen = sorted(i for r in range(w) for i in stream.shard_frames(n, r, w))
            assert seen == list(range(n))
            sizes = [len(stream.shard_frames(n, r, w)) for r in range(w)]
            assert max(sizes) - min(sizes) <= 1
    assert stream.batches([0, 2, 4, 6, 8], 2) == [[0, 2], [4, 6], [8]]


def test_pack_unpack_roundtrip():
    F, M = 3, 5
    rois = torch.arange(F * M * 5, dtype=torch.float32).view(F, M, 5)
    scores = torch.rand(F, M)
    num = torch.tensor([5, 2, 0], dtype=torch.int32)
    rec = stream.pack_records(rois, scores, num, [4, 6, 8])
    got = stream.unpack_records(rec, M)
    assert [g[0] for g in got] == [4, 6, 8]
    assert torch.equal(got[1][1], rois[1, :2]) and torch.equal(got[1][2], scores[1, :2]) and got[2][1].shape[0] == 0


def test_detection_records_rebuild_all_boxes():
    """final_detections' padded outputs -> wire records -> the reference's all_boxes[cls][frame]."""
    import numpy as np
    F, K, D, E = 3, 3, 4, 7
    g = torch.Generator().manual_seed(1)
    dets = torch.rand(F, K, D, E + 1, generator=g)
    uc_row = torch.rand(F, K, D, 2, generator=g)
    uc_cls = torch.rand(F, K, D, E, generator=g)
    counts = torch.tensor([[0, 2, 0], [0, 4, 1], [0, 0, 0]], dtype=torch.int32)
    rec = stream.pack_detection_records(dets, counts, [5, 0, 2], uc_row, uc_cls)
    assert rec.shape == (F, 2 + K + K * D * (E + 1 + 2 + E))
    allrec = stream.gather_detections(rec, 4)
    ab = stream.unpack_detection_records(allrec, K, 6)
    assert len(ab) == K and all(len(c) == 6 for c in ab)
    assert ab[1][5].shape == (2, E + 1 + 2 + E)
    assert np.array_equal(ab[1][5][:, :E + 1], dets[0, 1, :2].numpy()) and np.array_equal(ab[1][5][:, E + 1:E + 3], uc_row[0, 1, :2].numpy())
    assert np.array_equal(ab[2][0], torch.cat((dets[1, 2, :1], uc_row[1, 2, :1], uc_cls[1, 2, :1]), 1).numpy())
    assert ab[1][2].size == 0 and ab[2][5].size == 0 and ab[0][0].size == 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    M = 4
    mine = stream.shard_frames(n_frames, rank, world)
    F = len(mine)
    rois = torch.stack([torch.full((M, 5), float(i)) for i in mine]) if F else torch.zeros(0, M, 5)
    scores = torch.stack([torch.full((M,), float(i) / 10) for i in mine]) if F else torch.zeros(0, M)
    num = torch.tensor([i % (M + 1) for i in mine], dtype=torch.int32)
    rec = stream.pack_records(rois, scores, num, mine)
    allrec = stream.gather_detections(rec, (n_frames + world - 1) // world)
    got = stream.unpack_records(allrec, M)
    ok = [g[0] for g in got] == list(range(n_frames))
    ok = ok and all(g[1].shape[0] == g[0] % (M + 1) and (g[1] == g[0]).all() for g in got)
    out[rank] = ok
    dist.destroy_process_group()


def test_gather_world_size_2_gloo():
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, 7, out), nprocs=2, join=True)
    assert out[0] and out[1]
