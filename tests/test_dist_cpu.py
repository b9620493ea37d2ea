"""World-size-2 gloo tests of the image-sharded multi-GPU glue (yolov10-3d_b200/dist.py).

The CUDA kernels cannot run here; what is checked is the host-side contract the N>1 path relies on:
  * contiguous image shards cover the batch exactly once;
  * ONE all_reduce(sum) of the per-rank un-normalised loss partials followed by the target_scores_sum normalisation
    (reference loss.py:240-256) reproduces the single-process loss of the GLOBAL batch -- per-shard partials come
    from the CPU oracle here, from y3d_v10_loss_fwd on the GPU box;
  * detections of ragged shards all_gather back in image order.
"""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle
from tests import synth

GAINS = (7.5, 0.5, 1.5)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _case(B=5, hw=(128, 160), M=9, nc=7):
    lv = synth.levels(*hw)
    gt = synth.gt2d(B, M, nc, hw, seed=11)
    xm = synth.train_like_head2d(B, nc, lv, gt, seed=12, frac=0.2)
    xo = synth.train_like_head2d(B, nc, lv, gt, seed=13, frac=0.1)
    return lv, gt, xm, xo, nc


def _partials(x, lv, nc, gt, topk):
    """oracle loss items of a shard -> un-normalised sums (box, cls, dfl, target_scores_sum)."""
    items, tss, _ = oracle.v8_loss(x, lv, synth.STRIDES, nc, gt, topk, gains=GAINS)
    assert tss > 1.0  # below 1 the reference clamps and the raw sum cannot be recovered from the items
    return np.array([items[0] * tss / GAINS[0], items[1] * tss / GAINS[1], items[2] * tss / GAINS[2], tss])


def _normalise(p):  # what y3d_v8_loss_finalize does on the device (loss.py:240-256)
    tss = max(p[3], 1.0)
    return np.array([p[0] / tss * GAINS[0], p[1] / tss * GAINS[1], p[2] / tss * GAINS[2]])


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import yolov10_3d_b200 as y3d

    lv, gt, xm, xo, nc = _case()
    B = gt.shape[0]
    lo, hi = y3d.dist.shard_range(B, rank, world)
    parts = torch.from_numpy(np.concatenate([_partials(xm[lo:hi], lv, nc, gt[lo:hi], 10),
                                             _partials(xo[lo:hi], lv, nc, gt[lo:hi], 1)]))
    y3d.dist.reduce_partials(parts)  # the ONE collective of the loss path: 8 doubles
    items = np.concatenate([_normalise(parts[:4].numpy()), _normalise(parts[4:].numpy())])
    # detections: ragged shards (3 + 2 images), 4 rows of 6 per image, tagged with the global image index
    dets = torch.arange(lo, hi, dtype=torch.float32).view(-1, 1, 1).expand(hi - lo, 4, 6).contiguous()
    gathered = y3d.dist.all_gather_detections(dets)
    q.put((rank, (lo, hi), items, gathered.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions_the_batch():
    import yolov10_3d_b200 as y3d

    for B in (1, 5, 64, 257):
        for world in (1, 2, 3, 8):
            spans = [y3d.dist.shard_range(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_world2_loss_partials_and_detection_gather():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    lv, gt, xm, xo, nc = _case()
    B = gt.shape[0]
    _, full = oracle.v10_loss(xm, xo, lv, synth.STRIDES, nc, gt, gains=GAINS)  # single-process loss of the global batch
    assert [r[1] for r in res] == [(0, 3), (3, 5)]
    for _, _, items, gathered in res:
        np.testing.assert_allclose(items, np.asarray(full, dtype=np.float64), rtol=1e-9)  # identical on every rank
        assert gathered.shape == (B, 4, 6)
        assert np.array_equal(gathered[:, 0, 0], np.arange(B, dtype=np.float32))  # image order restored
