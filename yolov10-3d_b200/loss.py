"""``v8DetectionLoss`` / ``v10DetectLoss`` mirrors (reference ultralytics/utils/loss.py:157-257, 727-737): same
constructor (``model`` with ``.args`` and a head as ``model.model[-1]``), same ``__call__(preds, batch)`` and the
same ``(loss.sum() * batch_size, loss.detach())`` return.  One branch = one call into csrc/loss.cu, which reads the
head tensors once and never materialises ``pd_scores`` / ``target_scores``."""
import torch

from . import _lib
from ._util import Levels, ptr, stream_ptr, workspace

REG_MAX = 16


def pack_targets(batch_idx, cls, bboxes, batch_size, imgsz_hw, device, extra=None):
    """``v8DetectionLoss.preprocess`` (loss.py:180-195) without the per-image Python loop: ragged [N, ...] rows ->
    padded [B, Mmax, 5(+extra)] (cls, xyxy px, ...); zero rows are padding.  Row order inside an image is the order of
    appearance, as in the reference.  The per-image counts are taken on the host when ``batch_idx`` lives there (the
    dataloader's case), so no device sync is issued."""
    bi_host = batch_idx.detach().view(-1)
    n = bi_host.numel()
    cols = [cls.view(-1, 1), bboxes.view(-1, 4)] + ([extra] if extra is not None else [])
    rows = torch.cat([c.to(device=device, dtype=torch.float32) for c in cols], 1)
    W = rows.shape[1]
    if n == 0:
        return torch.zeros(batch_size, 0, W, device=device)
    counts = torch.bincount(bi_host.long().cpu(), minlength=batch_size)  # sync only if batch_idx was on the GPU
    M = int(counts.max())
    bi = bi_host.to(device).long()
    order = torch.argsort(bi, stable=True)
    sbi = bi[order]
    counts_d = counts.to(device)
    starts = torch.cumsum(counts_d, 0) - counts_d
    pos = torch.arange(n, device=device) - starts[sbi]
    out = torch.zeros(batch_size, M, W, device=device)
    out[sbi, pos] = rows[order]
    h, w = imgsz_hw
    scale = torch.tensor([w, h, w, h], device=device, dtype=torch.float32)  # imgsz[[1, 0, 1, 0]] loss.py:223
    xywh = out[..., 1:5] * scale
    dw, dh = xywh[..., 2] / 2, xywh[..., 3] / 2  # xywh2xyxy ops.py:403-422
    out[..., 1] = xywh[..., 0] - dw
    out[..., 2] = xywh[..., 1] - dh
    out[..., 3] = xywh[..., 0] + dw
    out[..., 4] = xywh[..., 1] + dh
    return out


def v8_loss_forward(feats, strides, nc, gt_packed, topk, gains, normalise=True, debug=False, prof_events=None):
    """One branch through ``y3d_v8_loss_fwd``.  Returns (items[4] = box, cls, dfl, target_scores_sum  -- or ``None``
    when not normalising --, partials float64[4], debug dict or None).  Nothing synchronises.
    ``prof_events``: optional ctypes array of 4 cudaEvent_t handles (see include/y3d.h), for benchmarks."""
    lv = Levels(feats, strides)
    if lv.C != 4 * REG_MAX + nc:
        raise ValueError(f"expected {4 * REG_MAX + nc} channels, got {lv.C}")
    dev = lv.device
    gt = gt_packed.to(dev, torch.float32).contiguous()
    M = int(gt.shape[1])
    items = torch.empty(4, dtype=torch.float32, device=dev) if normalise else None
    partials = torch.empty(4, dtype=torch.float64, device=dev)
    dbg = None
    if debug:
        dbg = dict(fg_mask=torch.empty((lv.B, lv.A), dtype=torch.bool, device=dev),
                   target_gt_idx=torch.empty((lv.B, lv.A), dtype=torch.int32, device=dev))
    ws = workspace(_lib.workspace_bytes(_lib.STAGE_V8_LOSS, B=lv.B, A=lv.A, nc=nc, M=M, k=topk), dev)
    _lib.check(_lib.lib().y3d_v8_loss_fwd(
        *lv.args(), lv.B, nc, REG_MAX, ptr(gt) if M > 0 else None, M, int(topk), float(gains[0]), float(gains[1]),
        float(gains[2]), int(normalise), ptr(items), ptr(partials), ptr(dbg["fg_mask"]) if debug else None,
        ptr(dbg["target_gt_idx"]) if debug else None, prof_events, ptr(ws), ws.numel(), stream_ptr(dev)))
    return items, partials, dbg


def v10_loss_forward(feats_o2m, feats_o2o, strides, nc, gt_packed, gains, topk=(10, 1), normalise=True, debug=False,
                     prof_events=None):
    """Both branches of ``v10DetectLoss`` through ONE call of ``y3d_v10_loss_fwd`` (same launches for both).

    Returns (items float32[8] = (box, cls, dfl, target_scores_sum) x (one2many, one2one) or ``None`` when not
    normalising, partials float64[8], debug dict or None).  Nothing synchronises."""
    lm, lo = Levels(feats_o2m, strides), Levels(feats_o2o, strides)
    if lm.C != 4 * REG_MAX + nc or lo.C != lm.C:
        raise ValueError(f"expected {4 * REG_MAX + nc} channels, got {lm.C} / {lo.C}")
    if lm.hw != lo.hw or lm.B != lo.B:
        raise ValueError("one2many and one2one heads must share batch size and level shapes")
    dev = lm.device
    gt = gt_packed.to(dev, torch.float32).contiguous()
    M = int(gt.shape[1])
    items = torch.empty(8, dtype=torch.float32, device=dev) if normalise else None
    partials = torch.empty(8, dtype=torch.float64, device=dev)
    dbg = None
    if debug:
        dbg = dict(fg_mask=torch.empty((2, lm.B, lm.A), dtype=torch.bool, device=dev),
                   target_gt_idx=torch.empty((2, lm.B, lm.A), dtype=torch.int32, device=dev))
    ws = workspace(_lib.workspace_bytes(_lib.STAGE_V8_LOSS, B=lm.B, A=lm.A, nc=nc, M=M, k=max(topk)), dev)
    _lib.check(_lib.lib().y3d_v10_loss_fwd(
        lm.c_ptr, lm.c_sB, lm.c_sC, lo.c_ptr, lo.c_sB, lo.c_sC, lm.c_hw, lm.c_stride, lm.nl, lm.B, nc, REG_MAX,
        ptr(gt) if M > 0 else None, M, int(topk[0]), int(topk[1]), float(gains[0]), float(gains[1]), float(gains[2]),
        int(normalise), ptr(items), ptr(partials), ptr(dbg["fg_mask"]) if debug else None,
        ptr(dbg["target_gt_idx"]) if debug else None, prof_events, ptr(ws), ws.numel(), stream_ptr(dev)))
    return items, partials, dbg


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
        items, _, _ = v8_loss_forward(feats, [float(s) for s in self.stride], self.nc, gt, self.topk, gains)
        loss = items[:3]
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
        items, _, _ = v10_loss_forward(fm, fo, [float(s) for s in o.stride], o.nc, gt,
                                       (o.hyp.box, o.hyp.cls, o.hyp.dfl), topk=(o.topk, self.one2one.topk))
        loss = items.view(2, 4)[:, :3].reshape(6)  # box_om cls_om dfl_om box_oo cls_oo dfl_oo (yolov10/train.py:10)
        return loss.sum() * fm[0].shape[0], loss.detach()
