"""``DDDetectionLoss`` / ``DetectLoss3d`` mirrors (reference ultralytics/utils/loss.py:741-900): same constructor
(``model`` with ``.args`` and the 3D head as ``model.model[-1]``), same ``__call__`` and the same
``(loss.sum() * batch_size, loss)`` return.  One branch = one call of ``y3d_dd_loss_fwd`` (csrc/loss3d.cu): the head is
read once, ``pd_scores`` / ``pd_3d`` / the dense targets of the reference are never materialised.

The returned total carries an autograd node whose backward is ``y3d_dd_loss_bwd``.  The fork's distillation and
foreground-depth-map terms (loss.py:745-751, 792, 890-895) are outside the hot path and raise if enabled."""
import ctypes as C

import torch

from . import _lib
from ._util import Levels, ptr, stream_ptr, workspace
from .loss import pack_targets


def dd_loss_forward(feats, strides, nc, gts_packed, calibs, mean_sizes, topk, gains, alpha=0.5, beta=1.0, gamma=1.0,
                    use_2d=True, use_3d=True, kps_dist_metric="l1", constrain_anchors=True, normalise=True, debug=False):
    """One branch through ``y3d_dd_loss_fwd``.  ``gts_packed`` [B, M, 17] (DDDetectionLoss.preprocess layout).  Returns
    (items float32[8] = six loss items, target_scores_sum, n_fg -- or None --, partials float64[11], target_gt_idx
    int32 [B, A] with -1 = background, or None)."""
    lv = Levels(feats, strides)
    if lv.C != nc + 35:
        raise ValueError(f"expected {nc + 35} channels, got {lv.C}")
    if not (use_2d or use_3d):
        raise RuntimeError("Either 2D or 3D assignment or both has to be selected!")  # tal.py:486
    if kps_dist_metric not in ("l1", "l2"):
        raise ValueError("kps_dist_metric must be 'l1' or 'l2'")
    dev = lv.device
    gts = gts_packed.to(dev, torch.float32).contiguous()
    M = int(gts.shape[1])
    cal = calibs.to(dev, torch.float32).contiguous()
    ms = mean_sizes.to(dev, torch.float32).contiguous()
    items = torch.empty(8, dtype=torch.float32, device=dev) if normalise else None
    partials = torch.empty(11, dtype=torch.float64, device=dev)
    tgi = torch.empty((lv.B, lv.A), dtype=torch.int32, device=dev) if debug else None
    ws = workspace(_lib.workspace_bytes(_lib.STAGE_DD_LOSS, B=lv.B, A=lv.A, nc=nc, M=M, k=topk), dev)
    flags = int(use_2d) | int(use_3d) << 1 | int(kps_dist_metric == "l2") << 2 | int(constrain_anchors) << 3
    g = (C.c_float * 6)(*[float(v) for v in gains])
    _lib.check(_lib.lib().y3d_dd_loss_fwd(*lv.args(), lv.B, nc, ptr(gts) if M > 0 else None, M, ptr(cal), ptr(ms),
                                          int(topk), float(alpha), float(beta), float(gamma), flags, g, int(normalise),
                                          ptr(items), ptr(partials), ptr(tgi), ptr(ws), ws.numel(), stream_ptr(dev)))
    return items, partials, tgi


def _dual_call(lm, lo, nc, gts, cal, ms, topk, gains, kw, normalise, debug):
    """``y3d_dd_loss_dual_fwd``: both branches in the same four launches.  Returns (items [2,8] or None, partials [2,11],
    tgi [2,B,A] or None, ws, bytes of one branch's workspace block)."""
    if lm.C != nc + 35 or lo.C != nc + 35:
        raise ValueError(f"expected {nc + 35} channels, got {lm.C} / {lo.C}")
    if lm.hw != lo.hw or lm.B != lo.B:
        raise ValueError("one2many and one2one heads must share batch size and level shapes")
    if not (kw["use_2d"] or kw["use_3d"]):
        raise RuntimeError("Either 2D or 3D assignment or both has to be selected!")  # tal.py:486
    if kw["kps_dist_metric"] not in ("l1", "l2"):
        raise ValueError("kps_dist_metric must be 'l1' or 'l2'")
    dev = lm.device
    M = int(gts.shape[1])
    items = torch.empty((2, 8), dtype=torch.float32, device=dev) if normalise else None
    partials = torch.empty((2, 11), dtype=torch.float64, device=dev)
    tgi = torch.empty((2, lm.B, lm.A), dtype=torch.int32, device=dev) if debug else None
    per = _lib.workspace_bytes(_lib.STAGE_DD_LOSS, B=lm.B, A=lm.A, nc=nc, M=M, k=max(topk)) - 256  # one branch's block
    ws = workspace(2 * per + 256, dev)
    flags = int(kw["use_2d"]) | int(kw["use_3d"]) << 1 | int(kw["kps_dist_metric"] == "l2") << 2 | int(
        kw["constrain_anchors"]) << 3
    g = (C.c_float * 6)(*[float(v) for v in gains])
    _lib.check(_lib.lib().y3d_dd_loss_dual_fwd(lm.c_ptr, lm.c_sB, lm.c_sC, lo.c_ptr, lo.c_sB, lo.c_sC, lm.c_hw, lm.c_stride,
                                               lm.nl, lm.B, nc, ptr(gts) if M > 0 else None, M, ptr(cal), ptr(ms),
                                               int(topk[0]), int(topk[1]), float(kw["alpha"]), float(kw["beta"]),
                                               float(kw["gamma"]), flags, g, int(normalise), ptr(items), ptr(partials),
                                               ptr(tgi), ptr(ws), ws.numel(), stream_ptr(dev)))
    return items, partials, tgi, ws, per, g


_DUAL_KW = dict(alpha=0.5, beta=1.0, gamma=1.0, use_2d=True, use_3d=True, kps_dist_metric="l1", constrain_anchors=True)


def dd_loss_dual_forward(feats_o2m, feats_o2o, strides, nc, gts_packed, calibs, mean_sizes, topk=(8, 1),
                         gains=(1.0, 1.0, 1.0, 1.0, 1.0, 1.0), normalise=True, debug=False, **kw):
    """Both branches of ``DetectLoss3d`` (loss.py:750-771), forward only, through ONE call of ``y3d_dd_loss_dual_fwd``
    (four launches for both branches): one2many with ``topk[0]``, one2one with ``topk[1]``.  Returns (items float32[2, 8]
    = six loss items, target_scores_sum, n_fg per branch -- or None --, partials float64[2, 11], target_gt_idx int32
    [2, B, A] with -1 = background, or None)."""
    lm, lo = Levels(feats_o2m, strides), Levels(feats_o2o, strides)
    dev = lm.device
    gts = gts_packed.to(dev, torch.float32).contiguous()
    cal = calibs.to(dev, torch.float32).contiguous()
    ms = mean_sizes.to(dev, torch.float32).contiguous()
    items, partials, tgi, _, _, _ = _dual_call(lm, lo, nc, gts, cal, ms, topk, gains, {**_DUAL_KW, **kw}, normalise, debug)
    return items, partials, tgi


def finalize_partials3d(partials, has_targets, gains):
    """Un-normalised sums of the 3D loss, float64 [n, 11] (summed over the ranks of an image-sharded batch) -> items
    float32 [n, 8] = six loss items, target_scores_sum, n_fg (``y3d_dd_loss_finalize``, loss.py:879-888).
    ``has_targets`` False: the reference's early return (loss.py:873-876), all items zero."""
    p = partials.reshape(-1, 11).contiguous()
    items = torch.empty((p.shape[0], 8), dtype=torch.float32, device=p.device)
    g = (C.c_float * 6)(*[float(v) for v in gains])
    for z in range(p.shape[0]):
        _lib.check(_lib.lib().y3d_dd_loss_finalize(ptr(p[z]), 1 if has_targets else 0, g, ptr(items[z]), stream_ptr(p.device)))
    return items


class _DDLossDualFn(torch.autograd.Function):
    """autograd node of the dual 3D loss: inputs = the per-level head tensors of both branches, output = the twelve loss
    items (one2many, one2one).  Forward: one ``y3d_dd_loss_dual_fwd``; backward: ``y3d_dd_loss_bwd`` per branch on that
    branch's half of the workspace."""

    @staticmethod
    def forward(ctx, cfg, *feats):
        strides, nc, gts, cal, ms, topk, gains, kw = cfg
        nl = len(feats) // 2
        lm, lo = Levels(feats[:nl], strides), Levels(feats[nl:], strides)
        gts = gts.to(lm.device, torch.float32).contiguous()
        items, _, _, ws, per, g = _dual_call(lm, lo, nc, gts, cal, ms, topk, gains, kw, True, False)
        ctx.save_for_backward(*lm.feats, *lo.feats, gts, ws, items)
        ctx.cfg = (lm.strides, int(gts.shape[1]), g, nc, nl, per)
        ctx.in_dtypes = [f.dtype for f in feats]
        return items[:, :6].reshape(12).clone()

    @staticmethod
    def backward(ctx, grad_items):
        strides, M, g, nc, nl, per = ctx.cfg
        saved = ctx.saved_tensors
        gts, ws, items = saved[2 * nl:]
        gi = grad_items.to(ws.device, torch.float32).contiguous()
        out = []
        for z in range(2):
            lv = Levels(saved[z * nl:(z + 1) * nl], strides)
            grads = [torch.empty_like(f) for f in lv.feats]
            gl = Levels(grads, lv.strides)
            _lib.check(_lib.lib().y3d_dd_loss_bwd(lv.c_ptr, lv.c_sB, lv.c_sC, gl.c_ptr, gl.c_sB, gl.c_sC, lv.c_hw,
                                                  lv.c_stride, lv.nl, lv.B, nc, ptr(gts) if M > 0 else None, M, g,
                                                  ptr(items[z]), ptr(gi[6 * z:6 * z + 6]), C.c_void_p(ws.data_ptr() + z * per),
                                                  per, stream_ptr(lv.device)))
            out += grads
        return (None, *[gr.to(dt) for gr, dt in zip(out, ctx.in_dtypes)])


class _DDLossFn(torch.autograd.Function):
    """autograd node of the fused 3D loss: inputs = the per-level head tensors, output = the six loss items."""

    @staticmethod
    def forward(ctx, cfg, *feats):
        strides, nc, gts, cal, ms, topk, gains, kw = cfg
        lv = Levels(feats, strides)
        dev = lv.device
        gts = gts.to(dev, torch.float32).contiguous()
        M = int(gts.shape[1])
        items = torch.empty(8, dtype=torch.float32, device=dev)
        partials = torch.empty(11, dtype=torch.float64, device=dev)
        ws = workspace(_lib.workspace_bytes(_lib.STAGE_DD_LOSS, B=lv.B, A=lv.A, nc=nc, M=M, k=topk), dev)
        flags = int(kw["use_2d"]) | int(kw["use_3d"]) << 1 | int(kw["kps_dist_metric"] == "l2") << 2 | int(
            kw["constrain_anchors"]) << 3
        g = (C.c_float * 6)(*[float(v) for v in gains])
        _lib.check(_lib.lib().y3d_dd_loss_fwd(*lv.args(), lv.B, nc, ptr(gts) if M > 0 else None, M, ptr(cal), ptr(ms),
                                              int(topk), float(kw["alpha"]), float(kw["beta"]), float(kw["gamma"]),
                                              flags, g, 1, ptr(items), ptr(partials), None, ptr(ws), ws.numel(),
                                              stream_ptr(dev)))
        ctx.save_for_backward(*lv.feats, gts, ws, items)  # version-checked by autograd (see loss._FusedLossFn)
        ctx.cfg = (lv.strides, M, g, nc, len(lv.feats))
        ctx.in_dtypes = [f.dtype for f in feats]
        return items[:6].clone()

    @staticmethod
    def backward(ctx, grad_items):
        strides, M, g, nc, nl = ctx.cfg
        saved = ctx.saved_tensors
        lv = Levels(saved[:nl], strides)
        gts, ws, items = saved[nl:]
        grads = [torch.empty_like(f) for f in lv.feats]
        gl = Levels(grads, lv.strides)
        gi = grad_items.to(lv.device, torch.float32).contiguous()
        _lib.check(_lib.lib().y3d_dd_loss_bwd(lv.c_ptr, lv.c_sB, lv.c_sC, gl.c_ptr, gl.c_sB, gl.c_sC, lv.c_hw,
                                              lv.c_stride, lv.nl, lv.B, nc, ptr(gts) if M > 0 else None, M, g,
                                              ptr(items), ptr(gi), ptr(ws), ws.numel(), stream_ptr(lv.device)))
        return (None, *[gr.to(dt) for gr, dt in zip(grads, ctx.in_dtypes)])


class DDDetectionLoss:
    """loss.py:775-900.  ``model.args`` must provide ``loss2d, cls, depth, offset3d, size3d, heading, tal_alpha,
    tal_beta, tal_gamma, tal_2d, tal_3d, kps_dist_metric, constrain_anchors, distillation``."""

    def __init__(self, model, tal_topk=10):
        h = model.args
        m = model.model[-1]
        self.hyp = h
        self.stride = m.stride
        self.nc = m.nc
        self.no = m.no
        self.device = next(model.parameters()).device
        self.topk = tal_topk
        if getattr(h, "distillation", False):
            raise _lib.Y3DError("distillation (SupervisionLoss, loss.py:792) is outside the B200 hot path")

    def targets(self, feats, batch):
        """GT packing (loss.py:795-810, 844-857) and the assigner settings of this loss: (gts [B,M,17], calib, mean sizes,
        assigner keyword arguments, gains)."""
        B = feats[0].shape[0]
        dev = feats[0].device
        h, w = feats[0].shape[2] * float(self.stride[0]), feats[0].shape[3] * float(self.stride[0])  # loss.py:844
        n = batch["cls"].numel()
        extra = torch.cat([batch[k].reshape(n, wd).float() for k, wd in
                           (("center_2d", 2), ("size_2d", 2), ("center_3d", 2), ("size_3d", 3), ("depth", 1),
                            ("heading_bin", 1), ("heading_res", 1))], 1)
        gts = pack_targets(batch["batch_idx"], batch["cls"], batch["bboxes"], B, (h, w), dev, extra=extra.to(dev))
        hp = self.hyp
        if not (hp.tal_2d or hp.tal_3d):
            raise RuntimeError("Either 2D or 3D assignment or both has to be selected!")  # tal.py:486
        kw = dict(alpha=hp.tal_alpha, beta=hp.tal_beta, gamma=hp.tal_gamma, use_2d=hp.tal_2d, use_3d=hp.tal_3d,
                  kps_dist_metric=hp.kps_dist_metric, constrain_anchors=hp.constrain_anchors)
        cal = batch["calib"].to(dev, torch.float32).contiguous()
        ms = batch["mean_sizes"].to(dev, torch.float32).contiguous()
        return gts, cal, ms, kw, (hp.loss2d, hp.cls, hp.depth, hp.offset3d, hp.size3d, hp.heading)

    def __call__(self, preds, batch, embeddings):
        feats = preds[1] if isinstance(preds, tuple) else preds  # loss.py:824
        gts, cal, ms, kw, gains = self.targets(feats, batch)
        loss = _DDLossFn.apply(([float(s) for s in self.stride], self.nc, gts, cal, ms, self.topk, gains, kw), *feats)
        return loss.sum() * feats[0].shape[0], loss  # loss.py:897 (the reference returns the items attached to the graph as well)


class DetectLoss3d:
    """loss.py:741-771 without the foreground-depth-map branches."""

    def __init__(self, model):
        if getattr(model.args, "fgdm_loss", False) or getattr(model.args, "fgdm_supervision", False):
            raise _lib.Y3DError("fgdm_loss / fgdm_supervision (loss.py:745-748) are outside the B200 hot path")
        self.one2many = DDDetectionLoss(model, tal_topk=model.args.tal_topk)
        self.one2one = DDDetectionLoss(model, tal_topk=1)
        self.model = model

    def __call__(self, preds, batch):
        if preds.get("one2many", None):
            # both branches in the same launches (one y3d_dd_loss_dual_fwd); same return as the reference's two calls:
            # (loss_o2m.sum() * B + loss_o2o.sum() * B, cat(items_o2m, items_o2o))   loss.py:763-770
            o = self.one2many
            fm = preds["one2many"][1] if isinstance(preds["one2many"], tuple) else preds["one2many"]
            fo = preds["one2one"][1] if isinstance(preds["one2one"], tuple) else preds["one2one"]
            gts, cal, ms, kw, gains = o.targets(fm, batch)
            loss = _DDLossDualFn.apply(([float(s) for s in o.stride], o.nc, gts, cal, ms, (o.topk, self.one2one.topk),
                                        gains, kw), *fm, *fo)
            B = fm[0].shape[0]
            return loss[:6].sum() * B + loss[6:].sum() * B, loss
        loss_one2one = self.one2one(preds["one2one"], batch, embeddings=preds.get("o2o_embs"))
        return torch.zeros(1), loss_one2one[1]  # loss.py:771
