"""CPU, build container only (needs /root/reference): seeded sweep of the REAL reference's assignment inside v10DetectLoss
against the oracle at the BASELINE shapes -- the evidence that the oracle is faithful at size, in the repo.  Skipped where
the reference is absent (the GPU box); there tests/golden/lossasg_*.npz carry the reference's answers."""
import types

import numpy as np
import pytest

from oracle import oracle, ref_import
from tests import synth

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="/root/reference is not present on this machine")

SEEDS = 50


def _reference_assignments(nc, hw, lv, gt, xm, xo):
    import torch

    from tests.golden import make_golden as mg  # imports the reference, patches nothing by itself

    batch = {k: mg.t(v) for k, v in synth.batch_dict(gt, hw).items()}
    fm = [mg.t(f) for f in synth.split_levels(xm, lv)]
    fo = [mg.t(f) for f in synth.split_levels(xo, lv)]
    model = mg.FakeModel(nc, synth.STRIDES, types.SimpleNamespace(box=7.5, cls=0.5, dfl=1.5))
    crit = mg.ref_loss.v10DetectLoss(model)
    captured = {}
    for branch, asg in ((0, crit.one2many.assigner), (1, crit.one2one.assigner)):
        def wrap(inner, branch=branch):
            def fwd(*a, **k):
                out = inner(*a, **k)
                captured[branch] = (out[3].numpy().copy(), out[4].numpy().copy())
                return out
            return fwd
        asg.forward = wrap(asg.forward)
    with mg.patched_topk(), torch.no_grad():
        _, items = crit({"one2many": fm, "one2one": fo}, batch)
    return captured, items.numpy().astype(np.float64)


def test_reference_vs_oracle_assignment_seed_sweep():
    nc, hw = 80, (640, 640)
    lv = synth.levels(*hw)
    n_fg = n_mism = 0
    worst = 0.0
    for s in range(SEEDS):
        crowd = s % 10 == 9  # every tenth seed is a dense-crowd image (500 GT)
        M = 500 if crowd else 100
        gt = synth.gt2d(1, M, nc, hw, seed=3000 + s, crowd=crowd, full=True)
        xm = synth.train_like_head2d(1, nc, lv, gt, seed=4000 + s, frac=0.02)
        xo = synth.train_like_head2d(1, nc, lv, gt, seed=5000 + s, frac=0.02)
        cap, ref_items = _reference_assignments(nc, hw, lv, gt, xm, xo)
        items = []
        for branch, (x, k) in enumerate(((xm, 10), (xo, 1))):
            it, _, _, fg, tgi = oracle.v8_loss(x, lv, synth.STRIDES, nc, gt, k, debug=True)
            rfg, rtgi = cap[branch]
            both = rfg & fg
            n_mism += int((fg != rfg).sum() + (tgi[both] != rtgi[both]).sum())
            n_fg += int(rfg.sum())
            items.append(it)
        worst = max(worst, float(np.max(np.abs(np.concatenate(items) / ref_items - 1.0))))
    print(f"reference vs oracle: {SEEDS} seeds, {n_fg} foreground anchors, {n_mism} mismatches, worst item deviation {worst:.2e}")
    assert n_fg > 20000 and n_mism == 0 and worst < 2e-5


SEEDS_3D = 16


def test_reference_vs_oracle_3d_assignment_seed_sweep():
    """The same at BASELINE cfg3 shape (KITTI 384 x 1280, 3 classes, up to 50 GT per image, top-k 8 and 1): the assignment of
    the REAL TaskAlignedAssigner3d inside the REAL DDDetectionLoss (loss.py:821-900, tal.py:391-700) against the oracle's."""
    import torch

    from tests.golden import make_golden as mg

    nc, hw, M = 3, (384, 1280), 50
    lv = synth.levels(*hw)
    ms = np.array(synth.KITTI_MEAN_SIZES, np.float32)
    calibs = np.array(synth.KITTI_CALIB, np.float32)[None]
    gains = dict(loss2d=1.3, cls=0.7, depth=1.1, offset3d=0.9, size3d=1.2, heading=0.8)
    n_fg = n_mism = 0
    worst = 0.0
    for s in range(SEEDS_3D):
        topk = 8 if s % 2 == 0 else 1
        gts = synth.gt3d(1, M, nc, hw, seed=6000 + s)
        x = synth.train_like_head3d(1, nc, lv, gts, seed=7000 + s, frac=0.03)
        args = types.SimpleNamespace(distillation=False, tal_topk=topk, tal_alpha=0.5, tal_beta=1.0, tal_gamma=1.0,
                                     tal_2d=True, tal_3d=True, kps_dist_metric="l1", constrain_anchors=True, **gains)
        model = mg.FakeModel(nc, synth.STRIDES, args)
        model.model[0].no = nc + 35
        crit = mg.ref_loss.DDDetectionLoss(model, tal_topk=topk)
        captured = {}
        inner = crit.assigner.forward

        def fwd(*a, _inner=inner, **k):
            out = _inner(*a, **k)
            captured["fg"], captured["tgi"] = out[1].numpy().copy(), out[2].numpy().copy()
            return out

        crit.assigner.forward = fwd
        feats = [mg.t(f) for f in synth.split_levels(x, lv)]
        batch = {k: mg.t(v) for k, v in synth.batch_dict3d(gts, hw, calibs, ms).items()}
        orig_cuda = torch.Tensor.cuda
        torch.Tensor.cuda = lambda self, *a, **k: self  # loss.py:1132 calls .cuda() on a fresh one-hot
        try:
            with mg.patched_topk(), torch.no_grad():
                _, ref_items = crit(feats, batch, embeddings=None)
        finally:
            torch.Tensor.cuda = orig_cuda
        bd = synth.batch_dict3d(gts, hw, calibs, ms)
        extra = np.concatenate([bd["center_2d"], bd["size_2d"], bd["center_3d"], bd["size_3d"], bd["depth"][:, None],
                                bd["heading_bin"][:, None], bd["heading_res"][:, None]], 1)
        packed = oracle.preprocess_targets(bd["batch_idx"], bd["cls"], bd["bboxes"], 1, hw, extra=extra)
        items, _, _, asg = oracle.dd_loss(x, lv, synth.STRIDES, nc, packed, calibs, ms, topk, gains=list(gains.values()))
        rfg, rtgi = captured["fg"].astype(bool), captured["tgi"]
        both = rfg & asg["fg_mask"]
        n_mism += int((asg["fg_mask"] != rfg).sum() + (asg["target_gt_idx"][both] != rtgi[both]).sum())
        n_fg += int(rfg.sum())
        worst = max(worst, float(np.max(np.abs(items / ref_items.numpy().astype(np.float64) - 1.0))))
    print(f"reference vs oracle (3D): {SEEDS_3D} seeds, {n_fg} foreground anchors, {n_mism} mismatches, "
          f"worst item deviation {worst:.2e}")
    assert n_fg > 500 and n_mism == 0 and worst < 3e-5


SEEDS_DET = 12


def test_reference_vs_oracle_decode_topk_seed_sweep():
    """Inference path at BASELINE cfg1 / cfg4 shapes (nc = 80, 640 x 640 / 1280 x 1280, top-300): the REAL Detect.inference
    (head.py:53-79) + ops.v10postprocess (ops.py:852-865) against the oracle's decode + two-stage top-k.  The selection
    runs on the reference's own decoded tensor (so only the selection is compared, bit for bit: labels and scores), and
    the decoded tensor itself is held to 1e-5."""
    import torch

    from tests.golden import make_golden as mg

    nc, D = 80, 300
    n_det = n_mism = 0
    worst = 0.0
    for s in range(SEEDS_DET):
        hw = (1280, 1280) if s % 4 == 3 else (640, 640)
        lv = synth.levels(*hw)
        x = synth.head2d(1, nc, lv, seed=8000 + s)
        feats = [mg.t(f) for f in synth.split_levels(x, lv)]
        ns = mg.head_ns(nc, synth.STRIDES)
        with torch.no_grad():
            y, _ = mg.Detect.inference(ns, feats)
            with mg.patched_topk():
                boxes, scores, labels = mg.ref_ops.v10postprocess(y.permute(0, 2, 1), D, nc)
        y = y.numpy()
        oy = oracle.decode2d(x, lv, synth.STRIDES, nc)
        worst = max(worst, float(np.max(np.abs(oy[:, 4:] - y[:, 4:]) / np.maximum(np.abs(y[:, 4:]), 1e-6))))
        np.testing.assert_allclose(oy[:, :4], y[:, :4], rtol=1e-5, atol=1e-3)
        ob, osc, ol, _ = oracle.postprocess(y.transpose(0, 2, 1), D, nc)
        n_mism += int((ol != labels.numpy()).sum() + (osc != scores.numpy()).sum() + (ob != boxes.numpy()).sum())
        n_det += ol.size
    print(f"reference vs oracle (decode + top-k): {SEEDS_DET} seeds, {n_det} detections, {n_mism} mismatches, "
          f"worst relative score deviation of the decode {worst:.2e}")
    assert n_det == SEEDS_DET * D and n_mism == 0 and worst < 1e-5
