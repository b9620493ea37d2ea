"""Image-sharded multi-GPU glue: one process per GPU, contiguous B/N images per rank.

* detections: each rank decodes + top-ks its images; one ``all_gather`` of [B/N, D, 6] rows (7.2 KB per image).
* loss: every loss term is normalised by the batch-global ``target_scores_sum`` (reference loss.py:240), so each rank
  emits un-normalised partials (4 doubles per branch) and one ``all_reduce(sum)`` of 8 doubles precedes the
  normalisation -- this reproduces the single-process reference on the full batch.
No other collective is issued: assignment and decode are independent per image.
"""
import torch
import torch.distributed as dist

from . import loss as _loss


def shard_range(batch, rank, world):
    """Contiguous image range [lo, hi) of ``rank``; the first ``batch % world`` ranks get one extra image."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_detections(dets_local, group=None):
    """[b_local, D, K] on every rank -> [sum b_local, D, K] on every rank (equal shards: one all_gather_into_tensor;
    ragged shards: padded to the largest)."""
    world = dist.get_world_size(group)
    if world == 1:
        return dets_local
    n = torch.tensor([dets_local.shape[0]], device=dets_local.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s) for s in sizes]
    mx = max(sizes)
    pad = dets_local
    if dets_local.shape[0] < mx:
        pad = torch.cat([dets_local, dets_local.new_zeros((mx - dets_local.shape[0],) + tuple(dets_local.shape[1:]))])
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad.contiguous(), group=group)
    return torch.cat([o[:s] for o, s in zip(out, sizes)])


def reduce_partials(partials, group=None):
    """Sum the per-rank loss partials (any shape, float64) over the group, in place."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(partials, op=dist.ReduceOp.SUM, group=group)
    return partials


def v10_loss_sharded(feats_o2m, feats_o2o, strides, nc, gt_local, gains, global_batch, group=None, prof_events=None):
    """``v10DetectLoss`` (loss.py:727-737) on this rank's image shard, normalised over the GLOBAL batch.

    Returns ``(total, items[6])`` identical on every rank and equal to the single-process result on the full batch.
    One rank: the kernels normalise directly.  Several ranks: un-normalised partials -> one all_reduce of 8 doubles
    -> ``y3d_v8_loss_finalize``."""
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    items, parts, _ = _loss.v10_loss_forward(feats_o2m, feats_o2o, strides, nc, gt_local, gains, normalise=not multi,
                                             prof_events=prof_events)
    if multi:
        items = _loss.finalize_partials(reduce_partials(parts, group), gains)
    items = items.view(2, 4)[:, :3].reshape(6)
    return items.sum() * global_batch, items
