"""Generate the committed golden fixtures by running the REAL reference (baldhat/yolov10-3D, imported
read-only from /root/reference through oracle/ref_import.py) on seeded synthetic inputs.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Each fixture ``tests/golden/<case>.npz`` stores the recipe of its inputs (``tests/synth.py`` arguments, as a
JSON string) plus a CRC of the regenerated inputs, and the reference's outputs.  The reference is run with
``torch.topk`` replaced by a stable descending sort (lowest index wins ties, the tie-break BASELINE.json
mandates) -- the fixture also records whether the unpatched reference agreed.
"""
import json
import os
import sys
import types
from functools import partial

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402
from tests import synth  # noqa: E402

ref_import.import_reference()
import torch  # noqa: E402
from ultralytics.nn.modules.block import DFL  # noqa: E402
from ultralytics.nn.modules.head import Detect, v10Detect3d  # noqa: E402
from ultralytics.utils import loss as ref_loss  # noqa: E402
from ultralytics.utils import ops as ref_ops  # noqa: E402
from ultralytics.utils import tal as ref_tal  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(8)
_orig_topk = torch.topk


def stable_topk(x, k, dim=-1, largest=True, sorted=True):
    assert largest
    v, i = torch.sort(x, dim=dim, descending=True, stable=True)
    return v.narrow(dim, 0, k), i.narrow(dim, 0, k)


class patched_topk:
    def __enter__(self):
        torch.topk = stable_topk

    def __exit__(self, *a):
        torch.topk = _orig_topk


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x))


def save(name, recipe, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, recipe=json.dumps(recipe), **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


# ---------------------------------------------------------------------------------------------- 2D decode
def head_ns(nc, strides):
    ns = types.SimpleNamespace(nc=nc, reg_max=16, no=nc + 64, dynamic=False, shape=None, export=False, format=None,
                               stride=torch.tensor(strides), dfl=DFL(16), anchors=torch.empty(0),
                               strides=torch.empty(0))
    ns.decode_bboxes = partial(Detect.decode_bboxes, ns)
    return ns


def case_decode_post(name, B, nc, img_hw, D, seed, quantise=None):
    lv = synth.levels(*img_hw)
    x = synth.head2d(B, nc, lv, seed=seed)
    if quantise:  # coarse logits -> many exactly tied scores
        x[:, 64:] = np.round(x[:, 64:] * quantise) / quantise
    feats = [t(f) for f in synth.split_levels(x, lv)]
    ns = head_ns(nc, synth.STRIDES)
    with torch.no_grad():
        y, _ = Detect.inference(ns, feats)
        ns.export = True  # xyxy variant (head.py:107-108)
        ns.shape = None
        y_xyxy = Detect.inference(ns, feats)
        with patched_topk():
            boxes, scores, labels = ref_ops.v10postprocess(y.permute(0, 2, 1), D, nc)
        b2, s2, l2 = ref_ops.v10postprocess(y.permute(0, 2, 1), D, nc)
        agree = bool(torch.equal(labels, l2) and torch.equal(scores, s2) and torch.equal(boxes, b2))
    recipe = dict(kind="decode_post", B=B, nc=nc, img_hw=img_hw, D=D, seed=seed, quantise=quantise)
    save(name, recipe, in_crc=np.int64(synth.checksum(x)), y=y.numpy(), y_xyxy_box=y_xyxy[:, :4].numpy(),
         boxes=boxes.numpy(), scores=scores.numpy(), labels=labels.numpy(), unpatched_agrees=np.bool_(agree))


# ---------------------------------------------------------------------------------------------- 2D assigner
def sparse_targets(ts):
    """target_scores [B,A,nc] is one-hot * norm: store nonzeros only."""
    idx = np.nonzero(ts)
    return np.stack(idx, 1).astype(np.int32), ts[idx]


def make_assign_inputs(kind, B, nc, img_hw, M, seed, crowd=False):
    lv = synth.levels(*img_hw)
    gt = synth.gt2d(B, M, nc, img_hw, seed=seed + 1, crowd=crowd)
    if kind == "random":
        x = synth.head2d(B, nc, lv, seed=seed)
    else:
        x = synth.train_like_head2d(B, nc, lv, gt, seed=seed + 2, frac=0.05)
    pd_scores, pd_bboxes, anc = synth.assigner_inputs_from_head(x, lv, nc)
    if kind == "ties":
        # adversarial: duplicated GT boxes, quantised scores / boxes, a GT over the top-left corner,
        # one image whose scores are all zero, one GT far outside every prediction (CIoU <= 0)
        gt[:, 1] = gt[:, 0]
        gt[0, 2, 1:5] = (0.0, 0.0, 70.0, 50.0)
        gt[1, 3, 1:5] = (1.0, 1.0, 30.0, 20.0)
        pd_scores = np.round(pd_scores * 8) / 8
        pd_bboxes = np.round(pd_bboxes / 4) * 4
        pd_scores[B - 1] = 0.0
        pd_scores = pd_scores.astype(np.float32)
        pd_bboxes = pd_bboxes.astype(np.float32)
    return gt, pd_scores, pd_bboxes, anc


def case_assign(name, kind, B, nc, img_hw, M, topk, seed, crowd=False, alpha=0.5, beta=6.0):
    gt, pd_scores, pd_bboxes, anc = make_assign_inputs(kind, B, nc, img_hw, M, seed, crowd)
    gl, gb = t(gt[..., :1]), t(gt[..., 1:5])
    mg = gb.sum(2, keepdim=True).gt_(0)
    asg = ref_tal.TaskAlignedAssigner(topk=topk, num_classes=nc, alpha=alpha, beta=beta)
    with patched_topk():
        tl, tb, ts, fg, tgi = asg(t(pd_scores), t(pd_bboxes), t(anc), gl, gb, mg)
    u = asg(t(pd_scores), t(pd_bboxes), t(anc), gl, gb, mg)
    agree = bool(torch.equal(fg, u[3]) and torch.equal(tgi, u[4]))
    ts_idx, ts_val = sparse_targets(ts.numpy())
    recipe = dict(kind="assign", inputs=kind, B=B, nc=nc, img_hw=img_hw, M=M, topk=topk, seed=seed, crowd=crowd,
                  alpha=alpha, beta=beta)
    save(name, recipe, in_crc=np.int64(synth.checksum(gt, pd_scores, pd_bboxes, anc)),
         target_labels=tl.numpy().astype(np.int16), target_bboxes_crc=np.int64(synth.checksum(tb.numpy())),
         ts_idx=ts_idx, ts_val=ts_val, fg_mask=np.packbits(fg.numpy()), target_gt_idx=tgi.numpy().astype(np.int16),
         unpatched_agrees=np.bool_(agree))


# ---------------------------------------------------------------------------------------------- 2D loss
class FakeModel(torch.nn.Module):
    """The three attributes v8DetectionLoss.__init__ reads (loss.py:160-178)."""

    def __init__(self, nc, strides, args):
        super().__init__()
        self.p = torch.nn.Parameter(torch.zeros(1))
        self.args = args
        head = types.SimpleNamespace(stride=torch.tensor(strides), nc=nc, no=nc + 64, reg_max=16)
        self.model = [head]


def case_loss(name, kind, B, nc, img_hw, M, seed, crowd=False, ragged=False):
    lv = synth.levels(*img_hw)
    gt = synth.gt2d(B, M, nc, img_hw, seed=seed + 1, crowd=crowd)
    if ragged:
        synth.make_ragged(gt, img_hw)
    if kind == "random":
        xm, xo = synth.head2d(B, nc, lv, seed=seed), synth.head2d(B, nc, lv, seed=seed + 7)
    else:
        xm = synth.train_like_head2d(B, nc, lv, gt, seed=seed + 2, frac=0.05)
        xo = synth.train_like_head2d(B, nc, lv, gt, seed=seed + 3, frac=0.03)
    batch = {k: t(v) for k, v in synth.batch_dict(gt, img_hw).items()}
    fm = [t(f).requires_grad_(True) for f in synth.split_levels(xm, lv)]
    fo = [t(f).requires_grad_(True) for f in synth.split_levels(xo, lv)]
    model = FakeModel(nc, synth.STRIDES, types.SimpleNamespace(box=7.5, cls=0.5, dfl=1.5))
    crit = ref_loss.v10DetectLoss(model)
    with patched_topk():
        total, items = crit({"one2many": fm, "one2one": fo}, batch)
    total.backward()
    gm = torch.cat([f.grad.view(B, nc + 64, -1) for f in fm], 2).numpy()
    go = torch.cat([f.grad.view(B, nc + 64, -1) for f in fo], 2).numpy()
    g = synth.rng(seed + 99)
    pos = g.integers(0, gm.size, 4096)
    # plus every position with a non-zero box-channel gradient of image 0 (fg anchors), capped
    nzm = np.flatnonzero(gm[0, :64])[:4096]
    nzo = np.flatnonzero(go[0, :64])[:4096]
    recipe = dict(kind="loss", inputs=kind, B=B, nc=nc, img_hw=img_hw, M=M, seed=seed, crowd=crowd, ragged=ragged,
                  gains=[7.5, 0.5, 1.5])
    save(name, recipe, in_crc=np.int64(synth.checksum(gt, xm, xo)), total=np.float64(total.item()),
         items=items.detach().numpy().astype(np.float64), grad_pos=pos, grad_m=gm.reshape(-1)[pos],
         grad_o=go.reshape(-1)[pos], grad_m_abs_sum=np.float64(np.abs(gm).sum(dtype=np.float64)),
         grad_o_abs_sum=np.float64(np.abs(go).sum(dtype=np.float64)), nz_m=nzm, nz_m_val=gm[0, :64].reshape(-1)[nzm],
         nz_o=nzo, nz_o_val=go[0, :64].reshape(-1)[nzo])



def case_loss_assign(name, B, nc, img_hw, M, seed, crowd=False, frac=0.02):
    """v10DetectLoss at a BASELINE shape with the assignment captured INSIDE the real loss: fg_mask / target_gt_idx of
    both branches as TaskAlignedAssigner.forward returned them to v8DetectionLoss.__call__ (loss.py:231)."""
    lv = synth.levels(*img_hw)
    gt = synth.gt2d(B, M, nc, img_hw, seed=seed + 1, crowd=crowd, full=crowd)
    xm = synth.train_like_head2d(B, nc, lv, gt, seed=seed + 2, frac=frac)
    xo = synth.train_like_head2d(B, nc, lv, gt, seed=seed + 3, frac=frac)
    batch = {k: t(v) for k, v in synth.batch_dict(gt, img_hw).items()}
    fm = [t(f) for f in synth.split_levels(xm, lv)]
    fo = [t(f) for f in synth.split_levels(xo, lv)]
    model = FakeModel(nc, synth.STRIDES, types.SimpleNamespace(box=7.5, cls=0.5, dfl=1.5))
    crit = ref_loss.v10DetectLoss(model)
    captured = {}

    def capture(branch, asg):
        inner = asg.forward

        def fwd(*a, **k):
            out = inner(*a, **k)
            captured[branch] = (out[3].numpy().copy(), out[4].numpy().copy())
            return out

        asg.forward = fwd

    capture(0, crit.one2many.assigner)
    capture(1, crit.one2one.assigner)
    with patched_topk(), torch.no_grad():
        total, items = crit({"one2many": fm, "one2one": fo}, batch)
    arrays = {}
    for z in (0, 1):
        fg, tgi = captured[z]
        arrays[f"fg{z}"] = np.packbits(fg)
        arrays[f"tgi{z}"] = tgi[fg].astype(np.int16)  # row-major over (image, anchor) at the foreground positions
    recipe = dict(kind="loss_assign", B=B, nc=nc, img_hw=img_hw, M=M, seed=seed, crowd=crowd, frac=frac,
                  gains=[7.5, 0.5, 1.5])
    save(name, recipe, in_crc=np.int64(synth.checksum(gt, xm, xo)), total=np.float64(total.item()),
         items=items.detach().numpy().astype(np.float64), **arrays)


from tests.golden.sparse_feat import sparse_feat_modules  # noqa: E402


def case_forward_feat(name, B, C, mid, img_hw, K, seed):
    """The real v10Detect3d.inference_forward_feat (head.py:694-716) on seeded feature maps and plain torch heads."""
    out_ch = {"cls": 3, "o2d": 2, "s2d": 2, "o3d": 2, "s3d": 3, "hd": 24, "dep": 1, "dep_un": 1}
    lv = synth.levels(*img_hw)
    g = synth.rng(seed)
    x = [t(g.standard_normal((B, C, h, w), dtype=np.float32)) for h, w in lv]
    heads = sparse_feat_modules(C, mid, out_ch, len(lv), seed)
    ns = types.SimpleNamespace(nl=len(lv), output_channels=out_ch, max_det=K, patch_size=5)
    ns.unravel_index = partial(v10Detect3d.unravel_index, ns)
    ns.select_candidates = partial(v10Detect3d.select_candidates, ns)
    ns.extract_patches = partial(v10Detect3d.extract_patches, ns)
    with patched_topk(), torch.no_grad():
        y = v10Detect3d.inference_forward_feat(ns, x, heads)
    recipe = dict(kind="forward_feat", B=B, C=C, mid=mid, img_hw=img_hw, K=K, seed=seed, out_ch=out_ch)
    save(name, recipe, in_crc=np.int64(synth.checksum(*[v.numpy() for v in x])),
         **{f"y{i}": v.numpy() for i, v in enumerate(y)})

# ---------------------------------------------------------------------------------------------- 3D
def head3d_ns(nc, strides):
    ns = types.SimpleNamespace(nc=nc, no=nc + 35, dynamic=False, shape=None, export=False, format=None,
                               stride=torch.tensor(strides), anchors=torch.empty(0), strides=torch.empty(0))
    ns.decode = partial(v10Detect3d.decode, ns)
    return ns


def case_decode3d(name, B, nc, img_hw, D, seed):
    lv = synth.levels(*img_hw)
    x = synth.head3d(B, nc, lv, seed=seed)
    feats = [t(f) for f in synth.split_levels(x, lv)]
    ns = head3d_ns(nc, synth.STRIDES)
    with torch.no_grad():
        y, _ = v10Detect3d.inference(ns, feats)
        with patched_topk():
            reg, scores, labels = ref_ops.v10_3Dpostprocess(y.transpose(-1, -2), D, nc)
        u = ref_ops.v10_3Dpostprocess(y.transpose(-1, -2), D, nc)
    agree = bool(torch.equal(labels, u[2]) and torch.equal(scores, u[1]))
    recipe = dict(kind="decode3d", B=B, nc=nc, img_hw=img_hw, D=D, seed=seed)
    save(name, recipe, in_crc=np.int64(synth.checksum(x)), y=y.numpy(), reg=reg.numpy(), scores=scores.numpy(),
         labels=labels.numpy(), unpatched_agrees=np.bool_(agree))
    return torch.cat((reg, scores.unsqueeze(-1), labels.unsqueeze(-1)), -1).numpy()  # yolov10_3D/val.py:46-47


def case_decode_preds(name, dets):
    from ultralytics.data.datasets.kitti import KITTIDataset
    from ultralytics.data.datasets.kitti_utils import Calibration

    B, D, _ = dets.shape
    g = synth.rng(77)
    calibs, cal_arr = [], []
    for b in range(B):
        c = object.__new__(Calibration)
        vals = np.array(synth.KITTI_CALIB) * (1 + 0.01 * g.standard_normal(6))
        c.cu, c.cv, c.fu, c.fv, c.tx, c.ty = (float(v) for v in vals)
        calibs.append(c)
        cal_arr.append(vals)
    inv = np.stack([np.array([[1.03 + 0.01 * b, 0.0, -3.0 + b], [0.0, 1.02, 2.5 - b]]) for b in range(B)])
    ratio = np.stack([np.array([1.0 + 0.03 * b, 1.0 + 0.02 * b]) for b in range(B)])
    ratio_pad = [(ratio[b], (0.0, 0.0)) for b in range(B)]
    cms = np.array(synth.KITTI_MEAN_SIZES)
    ds = types.SimpleNamespace(cls_mean_size=cms, use_camera_dis=False)
    files = [f"im{b}" for b in range(B)]
    res = KITTIDataset.decode_preds(ds, t(dets.copy()), calibs, files, ratio_pad, list(inv), undo_augment=True,
                                    threshold=0.001)
    rows = np.full((B, D, 14), np.nan)
    counts = np.zeros(B, np.int32)
    for b in range(B):
        r = res[files[b]]
        counts[b] = len(r)
        if r:
            rows[b, : len(r)] = np.array(r, dtype=np.float64)
    recipe = dict(kind="decode_preds", note="rows are the kept detections in order; NaN padded")
    save(name, recipe, dets=dets, calib=np.stack(cal_arr), inv_affine=inv, ratio=ratio, cls_mean_size=cms,
         rows=rows, counts=counts)


def case_assign3d(name, B, nc, img_hw, M, topk, seed, **kw):
    lv = synth.levels(*img_hw)
    x = synth.head3d(B, nc, lv, seed=seed)
    gts = synth.gt3d(B, M, nc, img_hw, seed=seed + 1)
    pd_scores, pd_bboxes, pd_3d, anc, st = synth.assigner3d_inputs_from_head(x, lv, nc)
    # make ~3% of in-GT anchors 'trained': 2D box, 3D centre, depth, heading and class near the GT
    g = synth.rng(seed + 5)
    for b in range(B):
        for m in range(M):
            row = gts[b, m]
            if row[1:5].sum() <= 0:
                continue
            ins = np.nonzero((anc[:, 0] > row[1]) & (anc[:, 0] < row[3]) & (anc[:, 1] > row[2]) & (anc[:, 1] < row[4]))[0]
            if ins.size == 0:
                continue
            for a in g.choice(ins, size=min(ins.size, max(2, int(0.03 * ins.size))), replace=False):
                pd_bboxes[b, a] = row[1:5] + g.standard_normal(4) * 3
                pd_scores[b, a, int(row[0])] = 0.5 + 0.4 * g.random()
                pd_3d[b, a, 0:2] = (row[9:11] - anc[a]) / st[a] + g.standard_normal(2) * 0.05
                pd_3d[b, a, 2:5] = row[11:14] + g.standard_normal(3) * 0.05
                pd_3d[b, a, 5 + int(row[15])] += 4.0
                pd_3d[b, a, 17 + int(row[15])] = row[16] + g.standard_normal() * 0.05
                pd_3d[b, a, 29] = row[14] + g.standard_normal() * 0.5
    pd_scores, pd_bboxes, pd_3d = (np.ascontiguousarray(v, np.float32) for v in (pd_scores, pd_bboxes, pd_3d))
    calibs = np.tile(np.array(synth.KITTI_CALIB, np.float32), (B, 1))
    calibs[:, 0] += np.arange(B, dtype=np.float32)
    ms = np.array(synth.KITTI_MEAN_SIZES, np.float32)
    parts = np.split(gts, np.cumsum([1, 4, 2, 2, 2, 3, 1, 1])[:], axis=2)
    gt_tuple = tuple(t(p) for p in parts)
    mg = gt_tuple[1].sum(2, keepdim=True).gt_(0)
    asg = ref_tal.TaskAlignedAssigner3d(topk=topk, num_classes=nc, **kw)
    with patched_topk():
        targets, fg, tgi, pk, gk = asg(t(pd_scores), t(pd_bboxes), t(pd_3d), t(anc), gt_tuple, mg, t(st[:, None]),
                                       t(calibs), t(ms))
    u = asg(t(pd_scores), t(pd_bboxes), t(pd_3d), t(anc), gt_tuple, mg, t(st[:, None]), t(calibs), t(ms))
    agree = bool(torch.equal(fg, u[1]) and torch.equal(tgi, u[2]))
    ts_idx, ts_val = sparse_targets(targets[1].numpy())
    tv = torch.cat(targets[2:], -1).numpy()
    recipe = dict(kind="assign3d", B=B, nc=nc, img_hw=img_hw, M=M, topk=topk, seed=seed, kw=kw)
    save(name, recipe, in_crc=np.int64(synth.checksum(gts, pd_scores, pd_bboxes, pd_3d)),
         pd_scores=pd_scores.astype(np.float32), pd_bboxes=pd_bboxes, pd_3d=pd_3d, gts=gts, calibs=calibs,
         target_labels=targets[0].numpy().astype(np.int16), ts_idx=ts_idx, ts_val=ts_val,
         target_vals_crc=np.int64(synth.checksum(tv)), fg_mask=np.packbits(fg.numpy()),
         target_gt_idx=tgi.numpy().astype(np.int16), pd_kps_sample=pk.numpy()[:, ::37], gt_kps=gk.numpy(),
         unpatched_agrees=np.bool_(agree))


def case_loss3d(name, B, nc, img_hw, M, topk, seed, **kw):
    """The REAL DDDetectionLoss (loss.py:775-900) on CPU.  compute_heading_loss calls ``.cuda()`` on a fresh one-hot
    (loss.py:1132); ``torch.Tensor.cuda`` is patched to the identity for the duration of the call -- no reference file
    is touched."""
    lv = synth.levels(*img_hw)
    gts = synth.gt3d(B, M, nc, img_hw, seed=seed + 1)
    x = synth.train_like_head3d(B, nc, lv, gts, seed=seed)
    calibs = np.tile(np.array(synth.KITTI_CALIB, np.float32), (B, 1))
    calibs[:, 0] += np.arange(B, dtype=np.float32)
    ms = np.array(synth.KITTI_MEAN_SIZES, np.float32)
    gains = dict(loss2d=1.3, cls=0.7, depth=1.1, offset3d=0.9, size3d=1.2, heading=0.8)
    args = types.SimpleNamespace(distillation=False, tal_topk=topk, tal_alpha=kw.get("alpha", 0.5),
                                 tal_beta=kw.get("beta", 1.0), tal_gamma=kw.get("gamma", 1.0),
                                 tal_2d=kw.get("use_2d", True), tal_3d=kw.get("use_3d", True),
                                 kps_dist_metric=kw.get("kps_dist_metric", "l1"),
                                 constrain_anchors=kw.get("constrain_anchors", True), **gains)
    model = FakeModel(nc, synth.STRIDES, args)
    model.model[0].no = nc + 35
    crit = ref_loss.DDDetectionLoss(model, tal_topk=topk)
    feats = [t(f).requires_grad_(True) for f in synth.split_levels(x, lv)]
    batch = {k: t(v) for k, v in synth.batch_dict3d(gts, img_hw, calibs, ms).items()}
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        with patched_topk():
            total, items = crit(feats, batch, embeddings=None)
        total.backward()
    finally:
        torch.Tensor.cuda = orig_cuda
    gr = torch.cat([f.grad.view(B, nc + 35, -1) for f in feats], 2).numpy()
    pos = synth.rng(seed + 99).integers(0, gr.size, 4096)
    nz = np.flatnonzero(gr[0, nc:])[:4096]  # regression rows of image 0: the foreground anchors
    feats = [f.detach() for f in feats]
    # the packed targets the loss built internally (preprocess loss.py:795-810), for the GT-packing glue test
    imgsz = torch.tensor(feats[0].shape[2:], dtype=torch.float32) * synth.STRIDES[0]
    g_in = torch.cat((batch["batch_idx"].view(-1, 1), batch["cls"].view(-1, 1), batch["bboxes"], batch["center_2d"],
                      batch["size_2d"], batch["center_3d"], batch["size_3d"], batch["depth"].view(-1, 1),
                      batch["heading_bin"].view(-1, 1), batch["heading_res"].view(-1, 1)), 1)
    packed = crit.preprocess(g_in, B, scale_tensor=imgsz[[1, 0, 1, 0]]).numpy()
    recipe = dict(kind="loss3d", B=B, nc=nc, img_hw=img_hw, M=M, topk=topk, seed=seed, kw=kw, gains=list(gains.values()))
    save(name, recipe, in_crc=np.int64(synth.checksum(gts, x)), calibs=calibs, packed=packed,
         total=np.float64(total.item()), items=items.detach().numpy().astype(np.float64), grad_pos=pos,
         grad=gr.reshape(-1)[pos], grad_abs_sum=np.float64(np.abs(gr).sum(dtype=np.float64)), nz=nz,
         nz_val=gr[0, nc:].reshape(-1)[nz])


def case_rotate_iou(name, N, K, seed):
    """The REAL rotate_iou_gpu_eval (data/datasets/kitti_eval.py:309-344), its numba-CUDA kernel executed by numba's
    CUDA simulator (no GPU in the build container).  The simulator evaluates ``math.cos`` / intermediate products in
    float64 where the device uses float32, so values agree with a float32 evaluation to ~1e-6, not bit for bit."""
    from ultralytics.data.datasets.kitti_eval import rotate_iou_gpu_eval

    boxes, query = synth.bev_boxes(N, seed), synth.bev_boxes(K, seed + 1)
    query[:3] = boxes[:3]  # identical boxes
    query[3] = boxes[3]
    query[3, 4] += np.float32(np.pi / 2)  # same box turned by 90 degrees
    out = {f"iou_c{c if c >= 0 else 'm1'}": rotate_iou_gpu_eval(boxes, query, criterion=c) for c in (-1, 0, 1, 2)}
    recipe = dict(kind="rotate_iou", N=N, K=K, seed=seed)
    save(name, recipe, in_crc=np.int64(synth.checksum(boxes, query)), query=query, **out)


def sparse_head_inputs(B, nc, C, H, W, K, Cout, seed):
    g = synth.rng(seed)
    scores = (g.standard_normal((B, nc, H, W), dtype=np.float32) * 2 - 4).astype(np.float32)
    x = g.standard_normal((B, C, H, W), dtype=np.float32)
    vals = g.standard_normal((B * K, Cout), dtype=np.float32)
    return scores, x, vals


def case_sparse_head(name, B, nc, C, H, W, K, Cout, seed, quantise=None):
    """The REAL v10Detect3d.select_candidates / extract_patches (head.py:659-687), called unbound on a stand-in for
    ``self``, and the scatter-back statements of inference_forward_feat (head.py:709-714)."""
    scores, x, vals = sparse_head_inputs(B, nc, C, H, W, K, Cout, seed)
    if quantise:
        scores = (np.round(scores * quantise) / quantise).astype(np.float32)  # many exact ties
    self_ = types.SimpleNamespace(max_det=K, patch_size=5)
    self_.unravel_index = lambda index, shape: v10Detect3d.unravel_index(self_, index, shape)
    with patched_topk():
        idx = v10Detect3d.select_candidates(self_, t(scores), B)
    uidx = v10Detect3d.select_candidates(self_, t(scores), B)
    patches = v10Detect3d.extract_patches(self_, t(x), idx)
    output_shape = (B, Cout, H, W)
    head_output = torch.zeros(output_shape)
    out = t(vals).view(B * K, Cout, 1, 1)[:, :, 0, 0].view(output_shape[0], K, output_shape[1]).transpose(1, 2)
    for b in range(B):
        head_output[b, :, idx[b, :, 0], idx[b, :, 1]] = out[b]
    recipe = dict(kind="sparse_head", B=B, nc=nc, C=C, H=H, W=W, K=K, Cout=Cout, seed=seed, quantise=quantise)
    save(name, recipe, in_crc=np.int64(synth.checksum(scores, x, vals)), idx=idx.numpy().astype(np.int16),
         patches_crc=np.int64(synth.checksum(patches.numpy())), scatter_crc=np.int64(synth.checksum(head_output.numpy())),
         unpatched_agrees=np.bool_(torch.equal(idx, uidx)))


if __name__ == "__main__":
    only = sys.argv[1:] or None

    def want(n):
        return only is None or any(n.startswith(o) for o in only)

    if want("decode"):
        case_decode_post("decode_post_small", B=2, nc=8, img_hw=(160, 160), D=50, seed=10)
        case_decode_post("decode_post_nc80", B=1, nc=80, img_hw=(256, 320), D=300, seed=11)
        case_decode_post("decode_post_ties", B=2, nc=8, img_hw=(160, 160), D=100, seed=12, quantise=2)
    if want("assign_"):
        case_assign("assign_random_k10", "random", B=3, nc=8, img_hw=(160, 160), M=12, topk=10, seed=20)
        case_assign("assign_trained_k10", "trained", B=3, nc=8, img_hw=(256, 320), M=12, topk=10, seed=21)
        case_assign("assign_trained_k1", "trained", B=3, nc=8, img_hw=(256, 320), M=12, topk=1, seed=22)
        case_assign("assign_trained_k13", "trained", B=2, nc=8, img_hw=(160, 160), M=6, topk=13, seed=23,
                    alpha=1.0, beta=6.0)
        case_assign("assign_ties_k10", "ties", B=3, nc=8, img_hw=(160, 160), M=12, topk=10, seed=24)
        case_assign("assign_ties_k1", "ties", B=3, nc=8, img_hw=(160, 160), M=12, topk=1, seed=25)
        case_assign("assign_crowd_k10", "trained", B=2, nc=8, img_hw=(256, 320), M=60, topk=10, seed=26, crowd=True)
    if want("loss"):
        case_loss("loss_random", "random", B=3, nc=8, img_hw=(160, 160), M=10, seed=30)
        case_loss("loss_trained", "trained", B=3, nc=8, img_hw=(256, 320), M=12, seed=31)
        case_loss("loss_crowd", "trained", B=2, nc=8, img_hw=(256, 320), M=60, seed=32, crowd=True)
        case_loss("loss_ragged", "trained", B=3, nc=8, img_hw=(256, 320), M=12, seed=33, ragged=True)
    if want("lossasg"):  # BASELINE shapes (cfg2 / cfg5: nc 80, 640 x 640, 100 / 500 GT per image)
        case_loss_assign("lossasg_cfg2", B=4, nc=80, img_hw=(640, 640), M=100, seed=90)
        case_loss_assign("lossasg_cfg5", B=2, nc=80, img_hw=(640, 640), M=500, seed=91, crowd=True)
    if want("decode3d") or want("preds3d"):
        dets = case_decode3d("decode3d_small", B=3, nc=3, img_hw=(96, 320), D=50, seed=40)
        case_decode_preds("preds3d_small", dets)
    if want("rotate_iou"):
        case_rotate_iou("rotate_iou_small", N=40, K=30, seed=80)
    if want("sparse_head"):
        case_sparse_head("sparse_head_kitti", B=3, nc=3, C=16, H=12, W=40, K=50, Cout=24, seed=70)
        case_sparse_head("sparse_head_ties", B=2, nc=3, C=8, H=24, W=80, K=50, Cout=3, seed=71, quantise=2)
    if want("forward_feat"):
        case_forward_feat("forward_feat_kitti", B=2, C=16, mid=8, img_hw=(96, 320), K=20, seed=75)
    if want("loss3d"):
        case_loss3d("loss3d_k8", B=2, nc=3, img_hw=(96, 320), M=8, topk=8, seed=60)
        case_loss3d("loss3d_k1", B=2, nc=3, img_hw=(96, 320), M=8, topk=1, seed=61)
        case_loss3d("loss3d_free_l2", B=2, nc=3, img_hw=(96, 320), M=6, topk=8, seed=62, beta=3.0, gamma=3.0,
                    kps_dist_metric="l2", constrain_anchors=False)
    if want("loss3d_cfg3"):  # BASELINE cfg3 shape: KITTI 384 x 1280, 3 classes, up to 50 GT per image, top-k 8
        case_loss3d("loss3d_cfg3", B=2, nc=3, img_hw=(384, 1280), M=50, topk=8, seed=63)
    if want("assign3d"):
        case_assign3d("assign3d_k8", B=2, nc=3, img_hw=(96, 320), M=8, topk=8, seed=50, alpha=0.5, beta=1.0, gamma=1.0)
        case_assign3d("assign3d_k1", B=2, nc=3, img_hw=(96, 320), M=8, topk=1, seed=51, alpha=0.5, beta=1.0, gamma=1.0)
        case_assign3d("assign3d_l2_free", B=2, nc=3, img_hw=(96, 320), M=6, topk=8, seed=52, alpha=0.5, beta=3.0,
                      gamma=3.0, kps_dist_metric="l2", constrain_anchors=False)
        case_assign3d("assign3d_3donly", B=2, nc=3, img_hw=(96, 320), M=6, topk=8, seed=53, use_2d=False)
        case_assign3d("assign3d_2donly", B=2, nc=3, img_hw=(96, 320), M=6, topk=8, seed=54, use_3d=False)
