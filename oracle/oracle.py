"""numpy front-end of the CPU oracle (oracle/y3d_oracle.c).

TEST INFRASTRUCTURE ONLY: imported from tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s
cpu_baseline / ``--impl reference`` legs.  The product package never imports this module.

Every function takes / returns numpy arrays laid out like the reference tensors and cites the
reference function it restates.
"""
import ctypes as C
import os

import numpy as np

from . import build_oracle

_lib = None

_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)
_u8p = C.POINTER(C.c_uint8)


def lib():
    global _lib
    if _lib is None:
        path = build_oracle.build()
        _lib = C.CDLL(path)
        for name in ("y3d_atanf", "y3d_sinf", "y3d_cosf", "y3d_expf"):
            getattr(_lib, name).restype = C.c_float
            getattr(_lib, name).argtypes = [C.c_float]
        _lib.y3d_atan2f.restype = C.c_float
        _lib.y3d_atan2f.argtypes = [C.c_float, C.c_float]
        _lib.y3d_powf.restype = C.c_float
        _lib.y3d_powf.argtypes = [C.c_float, C.c_float]
        _lib.y3d_o_ciou.restype = C.c_float
    return _lib


def _f32(x):
    return np.ascontiguousarray(x, dtype=np.float32)


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


def set_threads(n: int):
    """OpenMP threads used by the per-image parallel loops (assigners, losses)."""
    os.environ["OMP_NUM_THREADS"] = str(n)
    try:
        C.CDLL("libgomp.so.1").omp_set_num_threads(int(n))
    except OSError:
        pass


# --------------------------------------------------------------------------------------------
def make_anchors(lvl_hw, lvl_stride):
    """tal.py:300-312.  Returns (anchor_points [A,2] grid units, stride_tensor [A])."""
    hw = np.ascontiguousarray(lvl_hw, dtype=np.int32).reshape(-1, 2)
    st = _f32(lvl_stride)
    A = int((hw[:, 0] * hw[:, 1]).sum())
    anc = np.empty((A, 2), np.float32)
    s = np.empty((A,), np.float32)
    n = lib().y3d_o_make_anchors(C.c_int(len(st)), _p(hw, _i32p), _p(st, _f32p), _p(anc, _f32p), _p(s, _f32p))
    assert n == A
    return anc, s


def decode2d(xcat, lvl_hw, lvl_stride, nc, reg_max=16, xywh=True):
    """Detect.inference head.py:53-79: x_cat [B,4R+nc,A] -> y [B,4+nc,A]."""
    xcat = _f32(xcat)
    B, Cc, A = xcat.shape
    assert Cc == 4 * reg_max + nc
    anc, s = make_anchors(lvl_hw, lvl_stride)
    assert anc.shape[0] == A
    y = np.empty((B, 4 + nc, A), np.float32)
    lib().y3d_o_decode2d(_p(xcat, _f32p), C.c_int(B), C.c_int(nc), C.c_int(reg_max), C.c_int(A), _p(anc, _f32p),
                         _p(s, _f32p), C.c_int(int(xywh)), _p(y, _f32p))
    return y


def bbox_decode(pred_dist, anc, reg_max=16):
    """v8DetectionLoss.bbox_decode loss.py:197-204: [B,A,4R] -> xyxy grid units [B,A,4]."""
    pred_dist = _f32(pred_dist)
    B, A, _ = pred_dist.shape
    anc = _f32(anc)
    out = np.empty((B, A, 4), np.float32)
    lib().y3d_o_bbox_decode(_p(pred_dist, _f32p), C.c_int(B), C.c_int(A), C.c_int(reg_max), _p(anc, _f32p),
                            _p(out, _f32p))
    return out


def postprocess(preds, max_det, nc, nreg=4, scores_first=False):
    """ops.v10postprocess (2D: nreg=4, scores after boxes) / ops.v10_3Dpostprocess (3D: nreg=35,
    scores first) ops.py:852-880.  ``preds`` is [B,A,nreg+nc]; any strides (a transposed view works).
    Returns (reg [B,D,nreg], scores [B,D], labels [B,D] int64, anchor_idx [B,D] int32)."""
    assert preds.dtype == np.float32
    B, A, Ch = preds.shape
    assert Ch == nreg + nc
    sB, sA, sC = (s // 4 for s in preds.strides)
    D = int(max_det)
    reg = np.empty((B, D, nreg), np.float32)
    scores = np.empty((B, D), np.float32)
    labels = np.empty((B, D), np.int64)
    aidx = np.empty((B, D), np.int32)
    rc = lib().y3d_o_postprocess(_p(preds, _f32p), C.c_long(sB), C.c_long(sA), C.c_long(sC), C.c_int(B), C.c_int(A),
                                 C.c_int(nc), C.c_int(nreg), C.c_int(int(scores_first)), C.c_int(D), _p(reg, _f32p),
                                 _p(scores, _f32p), _p(labels, _i64p), _p(aidx, _i32p))
    if rc != 0:
        raise ValueError("max_det must be in 1..A")
    return reg, scores, labels, aidx


def ciou(box1, box2):
    """bbox_iou(box1, box2, xywh=False, CIoU=True) metrics.py:78-131, elementwise over [N,4]."""
    b1, b2 = _f32(box1).reshape(-1, 4), _f32(box2).reshape(-1, 4)
    out = np.empty((b1.shape[0],), np.float32)
    f = lib().y3d_o_ciou
    for i in range(b1.shape[0]):
        out[i] = f(_p(b1[i], _f32p), _p(b2[i], _f32p))
    return out


def tal_assign(pd_scores, pd_bboxes, anc_points, gt_labels, gt_bboxes, mask_gt, topk, alpha=0.5, beta=6.0,
               eps=1e-9, debug=False):
    """TaskAlignedAssigner.forward tal.py:44-94 (M > 0).  Returns dict of the five reference outputs
    (+ mask_pos / align_metric / overlaps when debug)."""
    pd_scores, pd_bboxes, anc = _f32(pd_scores), _f32(pd_bboxes), _f32(anc_points)
    B, A, nc = pd_scores.shape
    gl = _f32(gt_labels).reshape(B, -1)
    M = gl.shape[1]
    gb = _f32(gt_bboxes).reshape(B, M, 4)
    mg = _f32(mask_gt).reshape(B, M)
    out = dict(
        target_labels=np.empty((B, A), np.int64), target_bboxes=np.empty((B, A, 4), np.float32),
        target_scores=np.empty((B, A, nc), np.float32), fg_mask=np.empty((B, A), np.uint8),
        target_gt_idx=np.empty((B, A), np.int64))
    dbg = [None, None, None]
    if debug:
        dbg = [np.empty((B, M, A), np.uint8), np.empty((B, M, A), np.float32), np.empty((B, M, A), np.float32)]
    rc = lib().y3d_o_tal_assign(
        _p(pd_scores, _f32p), _p(pd_bboxes, _f32p), _p(anc, _f32p), _p(gl, _f32p), _p(gb, _f32p), _p(mg, _f32p),
        C.c_int(B), C.c_int(A), C.c_int(nc), C.c_int(M), C.c_int(topk), C.c_float(alpha), C.c_float(beta),
        C.c_float(eps), _p(out["target_labels"], _i64p), _p(out["target_bboxes"], _f32p),
        _p(out["target_scores"], _f32p), _p(out["fg_mask"], _u8p), _p(out["target_gt_idx"], _i64p),
        _p(dbg[0], _u8p), _p(dbg[1], _f32p), _p(dbg[2], _f32p))
    if rc != 0:
        raise ValueError(f"oracle tal_assign rc={rc}")
    out["fg_mask"] = out["fg_mask"].astype(bool)
    if debug:
        out.update(mask_pos=dbg[0], align_metric=dbg[1], overlaps=dbg[2])
    return out


def preprocess_targets(batch_idx, cls, bboxes, batch_size, img_hw, extra=None):
    """v8DetectionLoss.preprocess loss.py:180-195 (+ DDDetectionLoss.preprocess :795-810 when ``extra``
    [N,12] is given): ragged [N,...] rows -> padded [B,Mmax,5(+12)], boxes scaled to px and xywh->xyxy."""
    batch_idx = np.asarray(batch_idx).reshape(-1)
    cls = _f32(cls).reshape(-1, 1)
    bboxes = _f32(bboxes).reshape(-1, 4)
    cols = [cls, bboxes] + ([_f32(extra)] if extra is not None else [])
    rows = np.concatenate(cols, 1)
    W = rows.shape[1]
    if rows.shape[0] == 0:
        return np.zeros((batch_size, 0, W), np.float32)
    counts = [int((batch_idx == j).sum()) for j in range(batch_size)]
    out = np.zeros((batch_size, max(counts), W), np.float32)
    for j in range(batch_size):
        if counts[j]:
            out[j, : counts[j]] = rows[batch_idx == j]
    h, w = img_hw
    scale = np.array([w, h, w, h], np.float32)
    xywh = out[..., 1:5] * scale
    dw, dh = xywh[..., 2] / np.float32(2), xywh[..., 3] / np.float32(2)
    out[..., 1] = xywh[..., 0] - dw
    out[..., 2] = xywh[..., 1] - dh
    out[..., 3] = xywh[..., 0] + dw
    out[..., 4] = xywh[..., 1] + dh
    return out


def v8_loss(xcat, lvl_hw, lvl_stride, nc, gt_packed, topk, gains=(7.5, 0.5, 1.5), reg_max=16, debug=False):
    """v8DetectionLoss.__call__ loss.py:206-257 for one branch.  ``gt_packed`` [B,M,5] (cls, xyxy px) is the
    output of :func:`preprocess_targets`.  Returns (loss_items float64[3] = box, cls, dfl after gains,
    target_scores_sum, n_fg); with ``debug`` also the assignment it was computed from: fg_mask bool [B,A] and
    target_gt_idx int64 [B,A].  total = loss_items.sum() * B."""
    xcat = _f32(xcat)
    B, Cc, A = xcat.shape
    anc, s = make_anchors(lvl_hw, lvl_stride)
    gt = _f32(gt_packed)
    M = gt.shape[1]
    loss = np.zeros(3, np.float64)
    tss = C.c_double(0)
    nfg = C.c_int64(0)
    fg = np.zeros((B, A), np.uint8) if debug else None
    tgi = np.zeros((B, A), np.int64) if debug else None
    fn = lib().y3d_o_v8_loss_dbg
    fn.restype = C.c_int
    rc = fn(_p(xcat, _f32p), C.c_int(B), C.c_int(nc), C.c_int(reg_max), C.c_int(A), _p(anc, _f32p), _p(s, _f32p),
            _p(gt, _f32p), C.c_int(M), C.c_int(topk), C.c_float(gains[0]), C.c_float(gains[1]), C.c_float(gains[2]),
            _p(loss, _f64p), C.byref(tss), C.byref(nfg),
            fg.ctypes.data_as(C.c_void_p) if debug else None, tgi.ctypes.data_as(C.c_void_p) if debug else None)
    assert rc == 0
    if debug:
        return loss, tss.value, nfg.value, fg.astype(bool), tgi
    return loss, tss.value, nfg.value


def v10_loss(xcat_o2m, xcat_o2o, lvl_hw, lvl_stride, nc, gt_packed, gains=(7.5, 0.5, 1.5)):
    """v10DetectLoss.__call__ loss.py:727-737: topk=10 on one2many + topk=1 on one2one.
    Returns (total, items[6])."""
    B = xcat_o2m.shape[0]
    lm, _, _ = v8_loss(xcat_o2m, lvl_hw, lvl_stride, nc, gt_packed, 10, gains)
    lo, _, _ = v8_loss(xcat_o2o, lvl_hw, lvl_stride, nc, gt_packed, 1, gains)
    return (lm.sum() + lo.sum()) * B, np.concatenate([lm, lo])


def decode3d(xcat, lvl_hw, lvl_stride, nc):
    """v10Detect3d.decode head.py:755-764: [B,nc+35,A] -> [B,nc+35,A]."""
    xcat = _f32(xcat)
    B, Cc, A = xcat.shape
    assert Cc == nc + 35
    anc, s = make_anchors(lvl_hw, lvl_stride)
    y = np.empty_like(xcat)
    lib().y3d_o_decode3d(_p(xcat, _f32p), C.c_int(B), C.c_int(nc), C.c_int(A), _p(anc, _f32p), _p(s, _f32p),
                         _p(y, _f32p))
    return y


def keypoints(c3d, dep, size3d, hbin, hres, calib):
    """keypoint_utils.get_3d_keypoints :11-18 for one box -> [8,3]."""
    out = np.empty((8, 3), np.float32)
    c3d, size3d, calib = _f32(c3d), _f32(size3d), _f32(calib)
    lib().y3d_o_keypoints(_p(c3d, _f32p), C.c_float(dep), _p(size3d, _f32p), C.c_int(int(hbin)), C.c_float(hres),
                          _p(calib, _f32p), _p(out, _f32p))
    return out


def tal_assign3d(pd_scores, pd_bboxes, pd_3d, anc_points, stride, gts_packed, mask_gt, calibs, mean_sizes, topk,
                 alpha=0.5, beta=1.0, gamma=1.0, eps=1e-9, use_2d=True, use_3d=True, kps_dist_metric="l1",
                 constrain_anchors=True, debug=False):
    """TaskAlignedAssigner3d.forward tal.py:391-452.  ``gts_packed`` [B,M,17] = label, bbox(4), center_2d(2),
    size_2d(2), center_3d(2), size_3d(3), depth, heading_bin, heading_res."""
    pd_scores, pd_bboxes, pd_3d = _f32(pd_scores), _f32(pd_bboxes), _f32(pd_3d)
    B, A, nc = pd_scores.shape
    anc, st = _f32(anc_points), _f32(stride).reshape(-1)
    gts = _f32(gts_packed)
    M = gts.shape[1]
    mg = _f32(mask_gt).reshape(B, M)
    cal, ms = _f32(calibs), _f32(mean_sizes)
    out = dict(
        target_labels=np.empty((B, A), np.int64), target_scores=np.empty((B, A, nc), np.float32),
        target_vals=np.empty((B, A, 12), np.float32), fg_mask=np.empty((B, A), np.uint8),
        target_gt_idx=np.empty((B, A), np.int64), pd_keypoints=np.empty((B, A, 8, 3), np.float32),
        gt_keypoints=np.empty((B, M, 8, 3), np.float32))
    dbg = [None, None, None]
    if debug:
        dbg = [np.empty((B, M, A), np.uint8), np.empty((B, M, A), np.float32), np.empty((B, M, A), np.float32)]
    rc = lib().y3d_o_tal_assign3d(
        _p(pd_scores, _f32p), _p(pd_bboxes, _f32p), _p(pd_3d, _f32p), _p(anc, _f32p), _p(st, _f32p), _p(gts, _f32p),
        _p(mg, _f32p), _p(cal, _f32p), _p(ms, _f32p), C.c_int(B), C.c_int(A), C.c_int(nc), C.c_int(M), C.c_int(topk),
        C.c_float(alpha), C.c_float(beta), C.c_float(gamma), C.c_float(eps), C.c_int(int(use_2d)),
        C.c_int(int(use_3d)), C.c_int(int(kps_dist_metric == "l2")), C.c_int(int(constrain_anchors)),
        _p(out["target_labels"], _i64p), _p(out["target_scores"], _f32p), _p(out["target_vals"], _f32p),
        _p(out["fg_mask"], _u8p), _p(out["target_gt_idx"], _i64p), _p(out["pd_keypoints"], _f32p),
        _p(out["gt_keypoints"], _f32p), _p(dbg[0], _u8p), _p(dbg[1], _f32p), _p(dbg[2], _f32p))
    if rc != 0:
        raise ValueError(f"oracle tal_assign3d rc={rc}")
    out["fg_mask"] = out["fg_mask"].astype(bool)
    if debug:
        out.update(mask_pos=dbg[0], align_metric=dbg[1], overlaps=dbg[2])
    return out


def dd_loss(xcat, lvl_hw, lvl_stride, nc, gts_packed, calibs, mean_sizes, topk, gains=(1.0, 1.0, 1.0, 1.0, 1.0, 1.0),
            alpha=0.5, beta=1.0, gamma=1.0, use_2d=True, use_3d=True, kps_dist_metric="l1", constrain_anchors=True):
    """DDDetectionLoss.__call__ loss.py:821-900 for one branch (distillation off).  ``xcat`` [B, nc+35, A] =
    cls | o2d(2) s2d(2) o3d(2) s3d(3) hd(24) dep dep_un; ``gts_packed`` [B,M,17] is the output of
    DDDetectionLoss.preprocess (loss.py:795-810; bbox xyxy px).  ``gains`` = hyp.loss2d, cls, depth, offset3d, size3d,
    heading.  Returns (items float64[6], target_scores_sum, n_fg, assignment dict).  The assigner is the C oracle; the
    loss terms are float64 numpy restatements of compute_box2d_loss :913-926, compute_box3d_loss :928-963,
    laplacian_aleatoric_uncertainty_loss_new :1112-1119 and compute_heading_loss :1122-1136."""
    x = _f32(xcat)
    B, Cc, A = x.shape
    assert Cc == nc + 35
    gts = _f32(gts_packed)
    M = gts.shape[1]
    if M == 0:  # loss.py:873-876: the 5-tuple of the empty-GT assigner cannot be unpacked -> zero loss
        return np.zeros(6), 1.0, 0, None
    anc_g, st = make_anchors(lvl_hw, lvl_stride)  # grid units, [A,2] / [A]
    st = st.reshape(-1)
    t = np.ascontiguousarray(x.transpose(0, 2, 1))  # [B,A,C]
    logits = t[..., :nc]
    pd_scores = (np.float32(1.0) / (np.float32(1.0) + np.exp(-logits, dtype=np.float32))).astype(np.float32)
    centers = (anc_g[None] + t[..., nc:nc + 2]).astype(np.float32)  # bbox_decode loss.py:812-819
    half = (t[..., nc + 2:nc + 4] * np.float32(0.5)).astype(np.float32)
    pd_bboxes = (np.concatenate([centers - half, centers + half], -1).astype(np.float32) *
                 st[None, :, None]).astype(np.float32)
    pd_3d = np.ascontiguousarray(t[..., nc + 4:])
    anc_px = (anc_g * st[:, None]).astype(np.float32)
    mask_gt = (gts[..., 1:5].sum(-1) > 0).astype(np.float32)
    asg = tal_assign3d(pd_scores, pd_bboxes, pd_3d, anc_px, st, gts, mask_gt, calibs, mean_sizes, topk, alpha=alpha,
                       beta=beta, gamma=gamma, use_2d=use_2d, use_3d=use_3d, kps_dist_metric=kps_dist_metric,
                       constrain_anchors=constrain_anchors)
    fg = asg["fg_mask"]
    ts = asg["target_scores"].astype(np.float64)
    tv = asg["target_vals"].astype(np.float64)  # c2d2 s2d2 c3d2 s3d3 dep hbin hres
    tss = max(ts.sum(), 1.0)
    n_fg = int(fg.sum())
    z = logits.astype(np.float64)
    bce = (np.maximum(z, 0) - z * ts + np.log1p(np.exp(-np.abs(z)))).sum()
    p = t.astype(np.float64)[fg]  # [n_fg, C]
    tvf = tv[fg]
    stf = np.broadcast_to(st[None, :], fg.shape)[fg].astype(np.float64)[:, None]
    ancf = np.broadcast_to(anc_px[None], fg.shape + (2,))[fg].astype(np.float64)
    l1 = lambda a, b: np.abs(a - b)  # noqa: E731
    off2d = l1(p[:, nc:nc + 2] * stf, tvf[:, 0:2] - ancf).mean() if n_fg else np.nan
    size2d = l1(p[:, nc + 2:nc + 4] * stf, tvf[:, 2:4]).mean() if n_fg else np.nan
    dep, un = p[:, nc + 33], p[:, nc + 34]
    depth = (1.4142 * np.exp(-0.5 * un) * np.abs(dep - tvf[:, 9]) + 0.5 * un).sum()
    off3d = l1(p[:, nc + 4:nc + 6] * stf, tvf[:, 4:6] - ancf).mean() if n_fg else np.nan
    size3d = l1(p[:, nc + 6:nc + 9], tvf[:, 6:9]).sum()
    hd = p[:, nc + 9:nc + 33]
    tb = tvf[:, 10].astype(np.int64)
    hb = hd[:, :12]
    mx = hb.max(1, keepdims=True)
    lse = (mx + np.log(np.exp(hb - mx).sum(1, keepdims=True)))[:, 0]
    rows = np.arange(n_fg)
    hd_ce = (lse - hb[rows, tb]).sum()
    hd_l1 = np.abs(hd[rows, 12 + tb] - tvf[:, 11]).sum()
    items = np.array([(size2d + off2d) / tss * gains[0], bce / tss * gains[1], depth / tss * gains[2],
                      off3d / tss * gains[3], size3d / tss * gains[4], (hd_ce + hd_l1) / tss * gains[5]])
    return items, tss, n_fg, asg


def decode_preds(dets, calib, inv_affine, ratio, cls_mean_size, threshold=0.001):
    """KITTIDataset.decode_preds kitti.py:519-576.  dets [B,D,37] float32; calib [B,6], inv_affine [B,2,3],
    ratio [B,2], cls_mean_size [nc,3] float64.  Returns (rows [B,D,14] float64, valid [B,D] bool)."""
    dets = _f32(dets)
    B, D, _ = dets.shape
    calib = np.ascontiguousarray(calib, np.float64)
    inv_affine = np.ascontiguousarray(inv_affine, np.float64)
    ratio = np.ascontiguousarray(ratio, np.float64)
    cms = np.ascontiguousarray(cls_mean_size, np.float64)
    rows = np.empty((B, D, 14), np.float64)
    valid = np.empty((B, D), np.uint8)
    lib().y3d_o_decode_preds(_p(dets, _f32p), C.c_int(B), C.c_int(D), _p(calib, _f64p), _p(inv_affine, _f64p),
                             _p(ratio, _f64p), _p(cms, _f64p), C.c_double(threshold), _p(rows, _f64p), _p(valid, _u8p))
    return rows, valid.astype(bool)


# ------------------------------------------------------------------------------------------------ sparse 3D head glue
def select_candidates(scores, max_det):
    """v10Detect3d.select_candidates head.py:681-687 (+ unravel_index :652-657): [B, nc, H, W] -> [B, K, 2] int64
    (row, col) of the K cells with the largest max-over-classes logit, stable descending order."""
    s = _f32(scores)
    B, nc, H, W = s.shape
    mx = s.max(1).reshape(B, -1)
    out = np.empty((B, max_det, 2), np.int64)
    for b in range(B):
        order = np.argsort(-mx[b], kind="stable")[:max_det]
        out[b, :, 0], out[b, :, 1] = order // W, order % W
    return out


def extract_patches(x, indices, patch_size=5):
    """v10Detect3d.extract_patches head.py:659-679: [B, C, H, W], [B, K, 2] -> [B*K, C, P, P], zero padded."""
    x = _f32(x)
    B, Cc, H, W = x.shape
    K = indices.shape[1]
    pad = patch_size // 2
    xp = np.pad(x, ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    out = np.empty((B * K, Cc, patch_size, patch_size), np.float32)
    for b in range(B):
        for k in range(K):
            r, c = int(indices[b, k, 0]) + pad, int(indices[b, k, 1]) + pad
            out[b * K + k] = xp[b, :, r - pad:r + pad + 1, c - pad:c + pad + 1]
    return out


def scatter_candidates(values, indices, output_shape):
    """head.py:709-714: values [B*K, Cout] -> zero-filled [B, Cout, H, W] with out[b, :, row_k, col_k] = values[b*K+k]."""
    B, Cout, H, W = output_shape
    K = indices.shape[1]
    v = _f32(values).reshape(B, K, Cout)
    out = np.zeros(output_shape, np.float32)
    for b in range(B):
        out[b][:, indices[b, :, 0], indices[b, :, 1]] = v[b].T
    return out


def rotate_iou_eval(boxes, query_boxes, criterion=-1):
    """rotate_iou_gpu_eval kitti_eval.py:309-344: boxes [N,5], query_boxes [K,5] = (cx, cy, dx, dy, angle) -> [N,K]."""
    b, q = _f32(boxes), _f32(query_boxes)
    N, K = b.shape[0], q.shape[0]
    out = np.zeros((N, K), np.float32)
    if N and K:
        lib().y3d_o_rotate_iou_eval(_p(b, _f32p), C.c_int(N), _p(q, _f32p), C.c_int(K), C.c_int(criterion), _p(out, _f32p))
    return out
