"""``v8DetectionLoss`` / ``v10DetectLoss`` mirrors (reference ultralytics/utils/loss.py:157-257, 727-737): same
constructor (``model`` with ``.args`` and a head as ``model.model[-1]``), same ``__call__(preds, batch)`` and the
same ``(loss.sum() * batch_size, loss.detach())`` return -- the first element carries an autograd graph whose
backward is csrc/loss_bwd.cu.  The forward (csrc/loss.cu) reads the head tensors once and never materialises
``pd_scores`` / ``target_scores``."""
import torch

from . import _lib
from ._util import Levels, ptr, stream_ptr, workspace

REG_MAX = 16


def pack_targets(batch_idx, cls, bboxes, batch_size, imgsz_hw, device, extra=None):
    """``v8DetectionLoss.preprocess`` / ``DDDetectionLoss.preprocess`` (loss.py:180-195, 795-810) through
    ``y3d_pack_targets``: ragged [N, ...] rows -> padded [B, Mmax, 5(+extra)] (cls, xyxy px, ...); zero rows are
    padding.  Row order inside an image is the order of appearance, as in the reference.  Mmax comes from the host copy
    of ``batch_idx`` when it lives there (the dataloader's case), so no device sync is issued."""
    bi = batch_idx.detach().reshape(-1)
    n = bi.numel()
    E = 0 if extra is None else int(extra.shape[1])
    if n == 0:
        return torch.zeros(batch_size, 0, 5 + E, device=device)
    if bi.is_cuda:
        M = int(torch.bincount(bi.long(), minlength=batch_size).max())  # one sync, like the reference's counts.max()
    else:
        M = int(torch.bincount(bi.long(), minlength=batch_size).max())

    def dev32(t, shape):
        return t.detach().to(device=device, dtype=torch.float32).reshape(shape).contiguous()

    bi_d, cls_d, box_d = dev32(bi, (n,)), dev32(cls, (n,)), dev32(bboxes, (n, 4))
    ex_d = dev32(extra, (n, E)) if E else None
    out = torch.empty((batch_size, M, 5 + E), dtype=torch.float32, device=device)
    counts = torch.empty(batch_size, dtype=torch.int32, device=device)
    h, w = imgsz_hw  # scale = imgsz[[1, 0, 1, 0]] (loss.py:223): x by the width, y by the height
    _lib.check(_lib.lib().y3d_pack_targets(ptr(bi_d), ptr(cls_d), ptr(box_d), ptr(ex_d), E, n, int(batch_size), M,
                                           float(w), float(h), ptr(out), ptr(counts), stream_ptr(device)))
    return out


def _branch_forward(levels, nc, gt_packed, topk, gains, normalise, debug, prof_events, xrank=None, total_scale=None,
                    defer=False):
    """Shared body of the one- and two-branch forward calls.  ``levels``: list of 1 or 2 ``Levels``.  Returns a dict with
    ``items`` (float32[4n] or None), ``partials`` (float64[4n]), ``dbg``, ``total`` (float32[1] = ``total_scale`` * sum
    of the loss items, written by the last kernel, or None) and what the backward pass needs (``ws``, ``gt``, ``M``)."""
    n = len(levels)
    l0 = levels[0]
    for lv in levels:
        if lv.C != 4 * REG_MAX + nc:
            raise ValueError(f"expected {4 * REG_MAX + nc} channels, got {lv.C}")
        if lv.hw != l0.hw or lv.B != l0.B:
            raise ValueError("one2many and one2one heads must share batch size and level shapes")
    dev = l0.device
    gt = gt_packed.to(dev, torch.float32).contiguous()
    M = int(gt.shape[1])
    items = torch.empty(4 * n, dtype=torch.float32, device=dev) if normalise else None
    partials = torch.empty(4 * n, dtype=torch.float64, device=dev)
    total = torch.empty(1, dtype=torch.float32, device=dev) if (normalise and total_scale is not None) else None
    tsc = float(total_scale) if total is not None else 0.0
    dbg = None
    if debug:
        shape = (l0.B, l0.A) if n == 1 else (n, l0.B, l0.A)
        dbg = dict(fg_mask=torch.empty(shape, dtype=torch.bool, device=dev),
                   target_gt_idx=torch.empty(shape, dtype=torch.int32, device=dev))
    ws = workspace(_lib.workspace_bytes(_lib.STAGE_V8_LOSS, B=l0.B, A=l0.A, nc=nc, M=M, k=max(topk)), dev)
    tail = (ptr(gt) if M > 0 else None, M, *[int(k) for k in topk], float(gains[0]), float(gains[1]), float(gains[2]),
            int(normalise), ptr(items), ptr(partials), tsc, ptr(total), ptr(dbg["fg_mask"]) if debug else None,
            ptr(dbg["target_gt_idx"]) if debug else None, prof_events, ptr(ws), ws.numel(), stream_ptr(dev))
    if xrank is not None:  # two branches, cross-rank exchange fused into the last kernel (dist.PeerLossReducer)
        if n != 2 or debug or not normalise:
            raise ValueError("the sharded entry point takes both branches, normalises, and has no debug outputs")
        l1 = levels[1]
        _lib.check(_lib.lib().y3d_v10_loss_fwd_sharded(
            l0.c_ptr, l0.c_sB, l0.c_sC, l1.c_ptr, l1.c_sB, l1.c_sC, l0.c_hw, l0.c_stride, l0.nl, l0.B, nc, REG_MAX,
            ptr(gt) if M > 0 else None, M, int(topk[0]), int(topk[1]), float(gains[0]), float(gains[1]), float(gains[2]),
            ptr(items), ptr(partials), tsc, ptr(total), *xrank.next_call(defer), prof_events, ptr(ws), ws.numel(),
            stream_ptr(dev)))
    elif n == 1:
        _lib.check(_lib.lib().y3d_v8_loss_fwd(*l0.args(), l0.B, nc, REG_MAX, *tail))
    else:
        l1 = levels[1]
        _lib.check(_lib.lib().y3d_v10_loss_fwd(l0.c_ptr, l0.c_sB, l0.c_sC, l1.c_ptr, l1.c_sB, l1.c_sC, l0.c_hw,
                                               l0.c_stride, l0.nl, l0.B, nc, REG_MAX, *tail))
    return dict(items=items, partials=partials, dbg=dbg, ws=ws, gt=gt, M=M, total=total)


def v8_loss_forward(feats, strides, nc, gt_packed, topk, gains, normalise=True, debug=False, prof_events=None):
    """One branch through ``y3d_v8_loss_fwd``.  Returns (items[4] = box, cls, dfl, target_scores_sum  -- or ``None``
    when not normalising --, partials float64[4], debug dict or None).  Nothing synchronises.
    ``prof_events``: optional ctypes array of 4 cudaEvent_t handles (see include/y3d.h), for benchmarks."""
    r = _branch_forward([Levels(feats, strides)], nc, gt_packed, (topk,), gains, normalise, debug, prof_events)
    return r["items"], r["partials"], r["dbg"]


def v10_loss_forward(feats_o2m, feats_o2o, strides, nc, gt_packed, gains, topk=(10, 1), normalise=True, debug=False,
                     prof_events=None, xrank=None, total_scale=None, return_total=False, defer=False):
    """Both branches of ``v10DetectLoss`` through ONE call of ``y3d_v10_loss_fwd`` (same launches for both).

    Returns (items float32[8] = (box, cls, dfl, target_scores_sum) x (one2many, one2one) or ``None`` when not
    normalising, partials float64[8], debug dict or None).  Nothing synchronises.
    ``xrank``: a ``dist.PeerLossReducer`` -- this rank's images are a shard of a batch spread over several GPUs; the
    last kernel then sums the partials over the ranks through NVLink peer memory before normalising
    (``y3d_v10_loss_fwd_sharded``), and items / partials are those of the whole batch."""
    r = _branch_forward([Levels(feats_o2m, strides), Levels(feats_o2o, strides)], nc, gt_packed, topk, gains,
                        normalise, debug, prof_events, xrank, total_scale, defer)
    if return_total:  # ``total_scale`` * (sum of the six loss items), from the last kernel: float32[1]
        return r["items"], r["partials"], r["dbg"], r["total"]
    return r["items"], r["partials"], r["dbg"]


def _branch_backward(levels, nc, fwd, topk, gains, items, grad_items):
    """``y3d_v8_loss_bwd`` / ``y3d_v10_loss_bwd`` on the workspace the forward call left behind.  Returns one fp32
    gradient tensor per level and branch, shaped like the (contiguous) head tensors."""
    n = len(levels)
    l0 = levels[0]
    grads = [[torch.empty_like(f) for f in lv.feats] for lv in levels]
    gl = [Levels(g, l0.strides) for g in grads]
    gi = grad_items.to(l0.device, torch.float32).contiguous()
    tail = (l0.c_hw, l0.c_stride, l0.nl, l0.B, nc, REG_MAX, ptr(fwd["gt"]) if fwd["M"] > 0 else None, fwd["M"],
            *[int(k) for k in topk], float(gains[0]), float(gains[1]), float(gains[2]), ptr(items), ptr(gi),
            ptr(fwd["ws"]), fwd["ws"].numel(), stream_ptr(l0.device))
    if n == 1:
        _lib.check(_lib.lib().y3d_v8_loss_bwd(l0.c_ptr, l0.c_sB, l0.c_sC, gl[0].c_ptr, gl[0].c_sB, gl[0].c_sC, *tail))
    else:
        l1 = levels[1]
        _lib.check(_lib.lib().y3d_v10_loss_bwd(l0.c_ptr, l0.c_sB, l0.c_sC, gl[0].c_ptr, gl[0].c_sB, gl[0].c_sC,
                                               l1.c_ptr, l1.c_sB, l1.c_sC, gl[1].c_ptr, gl[1].c_sB, gl[1].c_sC, *tail))
    return grads


class _FusedLossFn(torch.autograd.Function):
    """autograd node of the fused loss: forward = the three forward kernels, backward = the backward kernel.
    Inputs: the per-level head tensors of every branch; output: the 3 loss items of every branch.

    Everything the backward pass reads -- the head tensors, the workspace the forward kernels left behind, the packed GT
    and the items -- goes through ``ctx.save_for_backward``, so autograd's version counters catch an in-place change of
    a feature map between forward and backward instead of the backward silently using changed logits."""

    @staticmethod
    def forward(ctx, cfg, *feats):
        strides, nc, gt, topk, gains = cfg[:5]
        # image-sharded batch (dist.v10_loss_sharded): cfg[5] = dist.PeerLossReducer or None, cfg[6] = global batch size.
        # The items then are those of the GLOBAL batch (sums exchanged by the last forward kernel), and the backward
        # kernel normalises this rank's gradients by the global target_scores_sum it finds in them.
        xrank, total_scale = (cfg[5], cfg[6]) if len(cfg) > 5 else (None, None)
        n = len(topk)
        nl = len(feats) // n
        levels = [Levels(feats[i * nl:(i + 1) * nl], strides) for i in range(n)]
        fwd = _branch_forward(levels, nc, gt, topk, gains, True, False, None, xrank, total_scale)
        used = [f for lv in levels for f in lv.feats]  # fp32, dense rows: the tensors the kernels actually read
        ctx.save_for_backward(*used, fwd["ws"], fwd["gt"], fwd["items"])
        ctx.cfg = (strides, nc, topk, gains, n, nl, fwd["M"])
        ctx.in_dtypes = [f.dtype for f in feats]
        return fwd["items"].view(n, 4)[:, :3].reshape(3 * n)

    @staticmethod
    def backward(ctx, grad_items):
        strides, nc, topk, gains, n, nl, M = ctx.cfg
        saved = ctx.saved_tensors
        used, (ws, gt, items) = saved[:n * nl], saved[n * nl:]
        levels = [Levels(used[i * nl:(i + 1) * nl], strides) for i in range(n)]
        grads = _branch_backward(levels, nc, dict(ws=ws, gt=gt, M=M), topk, gains, items, grad_items)
        flat = [g for br in grads for g in br]
        return (None, *[g.to(dt) for g, dt in zip(flat, ctx.in_dtypes)])


def finalize_partials(partials, gains):
    """All-reduced partials float64[4*n] -> items float32[4*n] (``y3d_v8_loss_finalize``)."""
    n = partials.numel() // 4
    items = torch.empty(4 * n, dtype=torch.float32, device=partials.device)
    _lib.check(_lib.lib().y3d_v8_loss_finalize(ptr(partials), n, float(gains[0]), float(gains[1]), float(gains[2]),
                                               ptr(items), stream_ptr(partials.device)))
    return items


class v8DetectionLoss:
    """loss.py:157-257.  ``model.args`` must provide ``box, cls, dfl``; ``model.model[-1]`` the head (``stride, nc,
    no, reg_max``)."""

    def __init__(self, model, tal_topk=10):
        device = next(model.parameters()).device
        m = model.model[-1]
        self.hyp = model.args
        self.stride = m.stride
        self.nc = m.nc
        self.no = m.no
        self.reg_max = m.reg_max
        self.device = device
        self.use_dfl = m.reg_max > 1
        self.topk = tal_topk
        if self.reg_max != REG_MAX:
            raise _lib.Y3DError("only reg_max == 16 is compiled (head.py:37)")

    def _targets(self, feats, batch):
        h, w = feats[0].shape[2] * float(self.stride[0]), feats[0].shape[3] * float(self.stride[0])  # loss.py:219
        return pack_targets(batch["batch_idx"], batch["cls"], batch["bboxes"], feats[0].shape[0], (h, w),
                            feats[0].device)

    def __call__(self, preds, batch):
        feats = preds[1] if isinstance(preds, tuple) else preds  # loss.py:209
        gt = self._targets(feats, batch)
        gains = (self.hyp.box, self.hyp.cls, self.hyp.dfl)
        loss = _FusedLossFn.apply(([float(s) for s in self.stride], self.nc, gt, (self.topk,), gains), *feats)
        return loss.sum() * feats[0].shape[0], loss.detach()  # loss.py:257


class v10DetectLoss:
    """loss.py:727-737: consistent dual assignment -- top-k 10 on one2many + top-k 1 on one2one."""

    def __init__(self, model):
        self.one2many = v8DetectionLoss(model, tal_topk=10)
        self.one2one = v8DetectionLoss(model, tal_topk=1)

    def __call__(self, preds, batch):
        one2many, one2one = preds["one2many"], preds["one2one"]
        fm = one2many[1] if isinstance(one2many, tuple) else one2many  # loss.py:209
        fo = one2one[1] if isinstance(one2one, tuple) else one2one
        o = self.one2many
        gt = o._targets(fm, batch)
        loss = _FusedLossFn.apply(([float(s) for s in o.stride], o.nc, gt, (o.topk, self.one2one.topk),
                                   (o.hyp.box, o.hyp.cls, o.hyp.dfl)), *fm, *fo)
        # box_om cls_om dfl_om box_oo cls_oo dfl_oo (yolov10/train.py:10)
        return loss.sum() * fm[0].shape[0], loss.detach()
