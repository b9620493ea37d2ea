"""Image-sharded multi-GPU glue: one process per GPU, contiguous B/N images per rank.

* detections: each rank decodes + top-ks its images; one ``all_gather`` of [B/N, D, 6] rows (7.2 KB per image).
* loss: every loss term is normalised by the batch-global ``target_scores_sum`` (reference loss.py:240), so each rank
  emits un-normalised partials (4 doubles per branch) and one ``all_reduce(sum)`` of 8 doubles precedes the
  normalisation -- this reproduces the single-process reference on the full batch.
No other collective is issued: assignment and decode are independent per image.
"""
import ctypes as C
import sys

import torch
import torch.distributed as dist

from . import _lib
from . import loss as _loss
from ._util import ptr, stream_ptr


def shard_range(batch, rank, world):
    """Contiguous image range [lo, hi) of ``rank``; the first ``batch % world`` ranks get one extra image."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_detections(dets_local, group=None, equal_shards=None):
    """[b_local, D, K] on every rank -> [sum b_local, D, K] on every rank.

    ``equal_shards=True`` (what ``shard_range`` gives when the batch divides by the world size -- the caller knows):
    ONE ``all_gather_into_tensor``, nothing else, no host synchronisation.  ``None`` / ``False``: the shard sizes are
    exchanged first (one small all_gather and a host read), equal sizes then take the single-call route and ragged
    ones are padded to the largest."""
    world = dist.get_world_size(group)
    if world == 1:
        return dets_local
    dets_local = dets_local.contiguous()
    if equal_shards:
        out = dets_local.new_empty((world * dets_local.shape[0],) + tuple(dets_local.shape[1:]))
        dist.all_gather_into_tensor(out, dets_local, group=group)
        return out
    n = torch.tensor([dets_local.shape[0]], device=dets_local.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s) for s in sizes]
    mx = max(sizes)
    if min(sizes) == mx:
        return all_gather_detections(dets_local, group, equal_shards=True)
    pad = dets_local
    if dets_local.shape[0] < mx:
        pad = torch.cat([dets_local, dets_local.new_zeros((mx - dets_local.shape[0],) + tuple(dets_local.shape[1:]))])
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad.contiguous(), group=group)
    return torch.cat([o[:s] for o, s in zip(out, sizes)])


class PeerDetectionGather:
    """The detection path's one exchange, fused into the last kernels of the fused decode + top-k (csrc/topk.cu,
    ``y3d_decode_topk2d_sharded``): the box-decode kernel stores every finished [D, 6] row straight into every peer's
    gather buffer over NVLink peer memory -- value and call number in one 64-bit word, so no fence and no flag follow the
    data, and no collective is launched; a small kernel (one CTA per remote image) unpacks the peers' words as they arrive.

    The buffers come from ``torch.distributed._symmetric_memory``.  Construction is collective; ``available`` is False when
    symmetric memory cannot be set up, and :func:`detect_sharded` then uses NCCL.  Every rank must make every call with
    the same ``n_local`` and ``max_det``."""

    def __init__(self, device, n_local, max_det, group=None):
        self.available = False
        self.seq = 0
        self.n_local, self.max_det = int(n_local), int(max_det)
        try:
            import torch.distributed._symmetric_memory as symm_mem

            pg = group if group is not None else dist.group.WORLD
            self.world, self.rank = dist.get_world_size(pg), dist.get_rank(pg)
            nbytes = int(_lib.lib().y3d_gather_buffer_bytes(self.world, self.n_local, self.max_det))
            self.buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=device)
            self.buf.zero_()
            self.handle = symm_mem.rendezvous(self.buf, pg.group_name if hasattr(pg, "group_name") else pg)
            self.ptrs = (C.c_void_p * self.world)(*[int(p) for p in self.handle.buffer_ptrs])
            self.status = torch.zeros(1, dtype=torch.int32, device=device)
            self._status_host = torch.zeros(1, dtype=torch.int32).pin_memory()
            self._calls, self.check_every, self._pg = 0, 16, pg
            torch.cuda.synchronize(device)
            dist.barrier(group=pg)  # every buffer is zeroed before anyone's first store can land
            self.available = True
        except Exception as e:  # no peer access / symmetric memory unsupported: NCCL path
            print(f"[yolov10-3d_b200] peer-memory detection gather unavailable ({type(e).__name__}: {e}); using NCCL",
                  file=sys.stderr)

    def view(self):
        """[world * n_local, D, 6] view of this rank's buffer at the parity of the last call.  Valid for consumers
        enqueued on the same stream BEFORE this rank's next call: a peer may run one call ahead of this rank's next post,
        and its rows for that call land in this parity's twin -- but the call after that reuses this parity, and the
        only thing that holds a peer back is this rank's next post."""
        n = self.world * self.n_local * self.max_det * 6
        par = self.seq & 1
        return self.buf[par * n * 4:(par + 1) * n * 4].view(torch.float32).view(self.world * self.n_local, self.max_det, 6)

    def check(self):
        """As :meth:`PeerLossReducer.check`: raises, a few calls late and without synchronising, when a peer never came."""
        if int(self._status_host[0]) != 0:
            raise _lib.Y3DError("peer-memory detection gather timed out: a rank did not make the call within "
                                "Y3D_XRANK_TIMEOUT_S (default 600 s)")
        self._calls += 1
        if self._calls % self.check_every == 0:
            self._status_host.copy_(self.status, non_blocking=True)


def detect_sharded(feats_one2one, strides, nc, max_det=300, group=None, gatherer=None):
    """``v10Detect.forward`` export branch (head.py:526-531: decode + ``v10postprocess``) on this rank's images of a batch
    sharded by image, detections gathered on all ranks: returns [world * B_local, max_det, 6] (x1 y1 x2 y2 score label)
    in rank order, identical on every rank.  With a :class:`PeerDetectionGather` the gather rides in the box-decode
    kernel's epilogue (the result is then a view of the gather buffer, see ``PeerDetectionGather.view``); without one,
    one NCCL ``all_gather_into_tensor`` follows the fused kernel.  One rank: the plain fused call."""
    from . import head as _head
    from ._util import Levels, workspace

    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    if not (multi and gatherer is not None and gatherer.available):
        out = _head.v10detect_export_forward(feats_one2one, strides, nc, max_det)
        return all_gather_detections(out, group, equal_shards=True) if multi else out
    lv = Levels(feats_one2one, [float(v) for v in strides])
    if lv.C != 4 * 16 + nc:
        raise ValueError(f"expected {4 * 16 + nc} channels, got {lv.C}")
    if lv.B != gatherer.n_local or int(max_det) != gatherer.max_det:
        raise ValueError("the gatherer was built for another shard size / max_det")
    ws = workspace(_lib.workspace_bytes(_lib.STAGE_DECODE_TOPK, B=lv.B, A=lv.A, nc=nc, D=int(max_det)), lv.device)
    gatherer.seq += 1
    _lib.check(_lib.lib().y3d_decode_topk2d_sharded(*lv.args(), lv.B, nc, 16, 0, int(max_det), gatherer.rank, gatherer.world,
                                                    gatherer.ptrs, C.c_uint64(gatherer.seq), ptr(gatherer.status), ptr(ws),
                                                    ws.numel(), stream_ptr(lv.device)))
    gatherer.check()
    return gatherer.view()


def reduce_partials(partials, group=None):
    """Sum the per-rank loss partials (any shape, float64) over the group, in place."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(partials, op=dist.ReduceOp.SUM, group=group)
    return partials


def dd_loss_sharded(feats_o2m, feats_o2o, strides, nc, gts_local, calibs_local, mean_sizes, gains, global_batch,
                    topk=(8, 1), group=None, **kw):
    """``DetectLoss3d`` (loss.py:741-771), forward / evaluation only, on this rank's image shard, normalised over the GLOBAL
    batch: the un-normalised sums of both branches (2 x 11 doubles -- foreground counts and L1 sums included, SURVEY.md
    section 8e) -> ONE all_reduce -> ``y3d_dd_loss_finalize``.  Returns ``(total, items[12])`` identical on every rank and equal
    to the single-process result on the whole batch.  (NCCL route only: the 3D path has no peer-memory exchange.)"""
    from . import loss3d as _loss3d

    items, parts, _ = _loss3d.dd_loss_dual_forward(feats_o2m, feats_o2o, strides, nc, gts_local, calibs_local, mean_sizes,
                                                   topk, gains, normalise=False, **kw)
    # one more word rides along: does ANY rank hold targets?  (without any the reference returns zeros, loss.py:873-876)
    buf = torch.cat([parts.reshape(-1), parts.new_tensor([1.0 if gts_local.shape[1] > 0 else 0.0])])
    reduce_partials(buf, group)
    items = _loss3d.finalize_partials3d(buf[:-1], bool(buf[-1] > 0), gains)[:, :6].reshape(12)
    return items.sum() * global_batch, items


class DeferredLoss:
    """Result of ``v10_loss_sharded(..., defer=True)``: the loss items are being collected on a side stream.  ``wait()``
    makes the current stream wait for them (no host synchronisation) and returns ``(total, items[6])``."""

    def __init__(self, items8, total, event):
        self._items8, self._total, self.event = items8, total, event

    def wait(self):
        torch.cuda.current_stream(self._items8.device).wait_event(self.event)
        return self._total[0], self._items8.view(2, 4)[:, :3].reshape(6)


class PeerLossReducer:
    """The loss path's one exchange as ONE kernel over NVLink peer memory (csrc/xrank.cu): every rank stores its 8
    partial sums straight into every peer's exchange buffer, waits for the peers' sequence flags, sums in rank order
    and normalises -- instead of an NCCL all-reduce of 64 bytes followed by a finalize kernel.

    The exchange buffers come from ``torch.distributed._symmetric_memory`` (peer-mapped allocations of one process per
    GPU).  Construction is collective; ``available`` is False when symmetric memory cannot be set up on this system,
    and callers then use the NCCL path.

    Like any collective, EVERY rank must make every call, in the same order; a rank that waits longer than
    ``Y3D_XRANK_TIMEOUT_S`` seconds (default 600) for a peer gives up, gets NaN sums and :meth:`check` raises.  The
    sequence number travels as a kernel argument, so calls must not be captured into a CUDA graph and replayed (a replay
    would see the previous replay's flags as its own)."""

    def __init__(self, device, group=None):
        self.available = False
        self.seq = 0
        try:
            import torch.distributed._symmetric_memory as symm_mem

            pg = group if group is not None else dist.group.WORLD
            self.world, self.rank = dist.get_world_size(pg), dist.get_rank(pg)
            nbytes = int(_lib.lib().y3d_xrank_buffer_bytes(self.world))
            self.buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=device)
            self.buf.zero_()
            self.handle = symm_mem.rendezvous(self.buf, pg.group_name if hasattr(pg, "group_name") else pg)
            self.ptrs = (C.c_void_p * self.world)(*[int(p) for p in self.handle.buffer_ptrs])
            # the same table on the device, for the exchange fused into the loss' last kernel
            self.ptrs_dev = torch.tensor([int(p) for p in self.handle.buffer_ptrs], dtype=torch.int64, device=device)
            self.status = torch.zeros(1, dtype=torch.int32, device=device)
            self._status_host = torch.zeros(1, dtype=torch.int32).pin_memory()
            self._calls, self.check_every, self._pg = 0, 16, pg
            self._side, self._ring, self._resolved = None, None, []
            torch.cuda.synchronize(device)
            dist.barrier(group=pg)  # every buffer is zeroed before anyone's first store can land
            self.available = True
        except Exception as e:  # no peer access / symmetric memory unsupported: NCCL path
            print(f"[yolov10-3d_b200] peer-memory loss reduction unavailable ({type(e).__name__}: {e}); using NCCL",
                  file=sys.stderr)

    def check(self):
        """Raises if an earlier exchange timed out (a peer never made the call; its sums came back as NaN).  Does not
        synchronise: every ``check_every`` calls the device-side status word is copied to pinned host memory
        asynchronously and this looks at the last copy that has landed, so a failure surfaces a few calls late instead
        of silently.  After a failure the ranks' sequence numbers are out of step: call :meth:`reset` on every rank."""
        if not self.available:
            return
        if int(self._status_host[0]) != 0:
            raise _lib.Y3DError("peer-memory loss exchange timed out: a rank did not reach the call within "
                                "Y3D_XRANK_TIMEOUT_S (default 600 s); every rank must call the sharded loss, then reset()")
        self._calls += 1
        if self._calls % self.check_every == 0:
            self._status_host.copy_(self.status, non_blocking=True)

    def reset(self):
        """Collective: brings the exchange back to its initial state (after a failure, or to re-synchronise)."""
        if not self.available:
            return
        torch.cuda.synchronize(self.buf.device)
        dist.barrier(group=self._pg)
        self.buf.zero_()
        self.status.zero_()
        self._status_host.zero_()
        self.seq = 0
        torch.cuda.synchronize(self.buf.device)
        dist.barrier(group=self._pg)

    def next_call(self, defer=False):
        """Arguments (rank, world, peer_bufs_dev, seq, defer, status) of ``y3d_v10_loss_fwd_sharded`` for the next
        collective call; every rank must make that call."""
        self.seq += 1
        return self.rank, self.world, ptr(self.ptrs_dev), C.c_uint64(self.seq), int(bool(defer)), ptr(self.status)

    def resolve_deferred(self, gains, total_scale):
        """The collecting half of the call just posted with ``defer=True``: ``y3d_loss_exchange_resolve`` on this reducer's
        side stream, behind an event on the current stream.  Returns a :class:`DeferredLoss`.

        Flow control (two parity slots in the exchange buffers): before the NEXT post, the caller's stream is made to wait
        for the resolve before this one (``pre_post``), so a rank is never more than one call ahead of its own collects --
        and therefore never overwrites a slot a peer has not read."""
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.buf.device)
            self._ring = [(torch.empty(8, dtype=torch.float32, device=self.buf.device),
                           torch.empty(1, dtype=torch.float32, device=self.buf.device)) for _ in range(4)]
        main = torch.cuda.current_stream(self.buf.device)
        posted = torch.cuda.Event()
        posted.record(main)
        items, total = self._ring[self.seq % len(self._ring)]
        self._side.wait_event(posted)
        _lib.check(_lib.lib().y3d_loss_exchange_resolve(
            2, self.rank, self.world, self.ptrs, C.c_uint64(self.seq), float(gains[0]), float(gains[1]), float(gains[2]),
            float(total_scale), ptr(items), ptr(total), None, ptr(self.status), C.c_void_p(self._side.cuda_stream)))
        done = torch.cuda.Event()
        done.record(self._side)
        self._resolved = (self._resolved + [done])[-2:]
        self.check()
        return DeferredLoss(items, total, done)

    def pre_post(self):
        """Before a deferred post: the current stream waits for the resolve two calls back (see ``resolve_deferred``)."""
        if len(self._resolved) == 2:
            torch.cuda.current_stream(self.buf.device).wait_event(self._resolved[0])

    def __call__(self, partials, gains):
        """partials float64[4n] of this rank -> items float32[4n] of the global batch (same on every rank)."""
        self.seq += 1
        n = partials.numel() // 4
        items = torch.empty(4 * n, dtype=torch.float32, device=partials.device)
        _lib.check(_lib.lib().y3d_loss_allreduce_finalize(
            ptr(partials), n, self.rank, self.world, self.ptrs, C.c_uint64(self.seq), float(gains[0]), float(gains[1]),
            float(gains[2]), ptr(items), None, ptr(self.status), stream_ptr(partials.device)))
        self.check()
        return items


def fused_off():
    """Y3D_XRANK_SEPARATE=1: keep the peer-memory reduction as its own kernel (measurement of the fusion's effect)."""
    import os
    return os.environ.get("Y3D_XRANK_SEPARATE") == "1"


def v10_loss_sharded(feats_o2m, feats_o2o, strides, nc, gt_local, gains, global_batch, group=None, prof_events=None,
                     reducer=None, defer=False):
    """``v10DetectLoss`` (loss.py:727-737) on this rank's image shard, normalised over the GLOBAL batch.

    Returns ``(total, items[6])`` identical on every rank and equal to the single-process result on the full batch -- or,
    with ``defer=True`` and a peer-memory ``reducer``, a :class:`DeferredLoss` whose ``wait()`` returns them: the cross-rank
    exchange is then collected on a side stream while this stream goes on (the next step's kernels, or the host's way to
    the backward call), so neither the NVLink round trip nor the slowest rank is waited for inside the step.
    When a head tensor requires grad (and no ``defer``), ``total`` carries the autograd node of ``loss.v10DetectLoss``:
    ``total.backward()`` gives this rank's share of the gradient of the global-batch loss (one rank, or the fused
    peer-memory route; the NCCL-fallback and deferred routes are forward-only).
    One rank: the kernels normalise directly.  Several ranks with a ``reducer`` (:class:`PeerLossReducer`): the
    loss' last kernel exchanges the partial sums over NVLink peer memory itself (``y3d_v10_loss_fwd_sharded``);
    without one: un-normalised partials -> NCCL all_reduce of 8 doubles -> ``y3d_v8_loss_finalize``."""
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    if defer and multi and reducer is not None and reducer.available and not fused_off():
        # the loss' last kernel only POSTS this rank's sums to the peers; the collecting kernel runs on a side stream, so
        # the NVLink round trip and the wait for the slowest rank overlap whatever this stream does next
        reducer.pre_post()
        _loss.v10_loss_forward(feats_o2m, feats_o2o, strides, nc, gt_local, gains, prof_events=prof_events, xrank=reducer,
                               total_scale=global_batch, return_total=True, defer=True)
        return reducer.resolve_deferred(gains, global_batch)
    if (multi and reducer is not None and reducer.available and not fused_off()) or not multi:
        if torch.is_grad_enabled() and any(f.requires_grad for f in (*feats_o2m, *feats_o2o)):
            # training: the same autograd node as loss.v10DetectLoss; its backward kernel reads the globally normalised
            # target_scores_sum from the items of the sharded forward, so total.backward() yields this rank's share of
            # the global-batch gradient (the NCCL-fallback and deferred routes below are forward / evaluation only)
            loss = _loss._FusedLossFn.apply(([float(v) for v in strides], nc, gt_local, (10, 1), gains,
                                             reducer if multi else None, float(global_batch)), *feats_o2m, *feats_o2o)
            if multi:
                reducer.check()
            return loss.sum() * global_batch, loss.detach()
        # one rank, or the exchange rides in the loss' last kernel: no collective launch at all, and the same kernel
        # writes the total (global_batch * sum of the items)
        items, _, _, total = _loss.v10_loss_forward(feats_o2m, feats_o2o, strides, nc, gt_local, gains,
                                                    prof_events=prof_events, xrank=reducer if multi else None,
                                                    total_scale=global_batch, return_total=True)
        if multi:
            reducer.check()
        return total[0], items.view(2, 4)[:, :3].reshape(6)
    items, parts, _ = _loss.v10_loss_forward(feats_o2m, feats_o2o, strides, nc, gt_local, gains, normalise=False,
                                             prof_events=prof_events)
    if reducer is not None and reducer.available:
        items = reducer(parts, gains)  # one extra kernel over peer memory: all-reduce + normalisation
    else:
        items = _loss.finalize_partials(reduce_partials(parts, group), gains)
    items = items.view(2, 4)[:, :3].reshape(6)
    return items.sum() * global_batch, items
