"""CPU: the oracle restatement (oracle/y3d_oracle.c) against fixtures produced by the REAL reference
(tests/golden/make_golden.py).  Index / mask outputs must be identical; values within 1e-5 relative."""
import numpy as np
import pytest

from oracle import oracle
from tests import cases, synth

RTOL, ATOL = 1e-5, 1e-6


@pytest.mark.parametrize("name", cases.names("decode_post_"))
def test_decode2d_and_postprocess(name):
    r, z = cases.load(name)
    lv, x = cases.decode_post_inputs(r, z)
    y = oracle.decode2d(x, lv, synth.STRIDES, r["nc"], xywh=True)
    # box coordinates: cx = ((a-l)+(a+r))/2*stride cancels, so the bar is 1e-5 relative to the box scale (px)
    np.testing.assert_allclose(y[:, :4], z["y"][:, :4], rtol=RTOL, atol=1e-4)
    np.testing.assert_allclose(y[:, 4:], z["y"][:, 4:], rtol=RTOL, atol=1e-7)
    yx = oracle.decode2d(x, lv, synth.STRIDES, r["nc"], xywh=False)
    np.testing.assert_allclose(yx[:, :4], z["y_xyxy_box"], rtol=RTOL, atol=1e-4)
    # top-k on the REFERENCE's decoded tensor (identical inputs => bit-exact selection)
    preds = z["y"].transpose(0, 2, 1)
    boxes, scores, labels, _ = oracle.postprocess(preds, r["D"], r["nc"])
    assert np.array_equal(labels, z["labels"])
    assert np.array_equal(scores, z["scores"])
    assert np.array_equal(boxes, z["boxes"])


@pytest.mark.parametrize("name", cases.names("assign_"))
def test_tal_assign(name):
    r, z = cases.load(name)
    inp = cases.assign_inputs(r, z)
    B, A, nc = inp["pd_scores"].shape
    out = oracle.tal_assign(inp["pd_scores"], inp["pd_bboxes"], inp["anc"], inp["gt_labels"], inp["gt_bboxes"],
                            inp["mask_gt"], r["topk"], alpha=r["alpha"], beta=r["beta"])
    assert np.array_equal(out["fg_mask"], cases.unpack_mask(z, "fg_mask", B, A))
    assert np.array_equal(out["target_gt_idx"], z["target_gt_idx"].astype(np.int64))
    assert np.array_equal(out["target_labels"], z["target_labels"].astype(np.int64))
    assert synth.checksum(out["target_bboxes"]) == int(z["target_bboxes_crc"])
    np.testing.assert_allclose(out["target_scores"], cases.dense_target_scores(z, B, A, nc), rtol=2e-5, atol=1e-7)
    assert out["fg_mask"].sum() > 0


@pytest.mark.parametrize("name", cases.names("loss_"))
def test_v10_loss(name):
    r, z = cases.load(name)
    lv, gt, xm, xo = cases.loss_inputs(r, z)
    bd = synth.batch_dict(gt, r["img_hw"])
    packed = oracle.preprocess_targets(bd["batch_idx"], bd["cls"], bd["bboxes"], r["B"], r["img_hw"])
    total, items = oracle.v10_loss(xm, xo, lv, synth.STRIDES, r["nc"], packed, gains=r["gains"])
    np.testing.assert_allclose(items, z["items"], rtol=2e-5)
    np.testing.assert_allclose(total, float(z["total"]), rtol=2e-5)


@pytest.mark.parametrize("name", cases.names("lossasg_"))
def test_v10_loss_assignment_at_baseline_shapes(name):
    """BASELINE shapes (cfg2: nc 80, 640 x 640, 100 GT / image; cfg5: 500 GT / image): the assignment INSIDE the real
    v10DetectLoss (fg_mask / target_gt_idx of both branches, captured from TaskAlignedAssigner.forward at loss.py:231)
    against the oracle's: zero mismatches, loss items to 2e-5."""
    r, z = cases.load(name)
    lv, gt, xm, xo = cases.loss_assign_inputs(r, z)
    A = synth.num_anchors(lv)
    items = []
    for branch, (x, k) in enumerate(((xm, 10), (xo, 1))):
        it, _, nfg, fg, tgi = oracle.v8_loss(x, lv, synth.STRIDES, r["nc"], gt, k, gains=r["gains"], debug=True)
        efg, etgi = cases.loss_assign_expected(z, branch, r["B"], A)
        mism = int((fg != efg).sum() + (tgi[efg & fg] != etgi[efg & fg]).sum())
        print(f"{name} branch {branch}: {int(efg.sum())} foreground anchors, {mism} mismatches")
        assert mism == 0 and nfg == int(efg.sum()) > 0
        items.append(it)
    np.testing.assert_allclose(np.concatenate(items), z["items"], rtol=2e-5)


@pytest.mark.parametrize("name", cases.names("decode3d_"))
def test_decode3d_and_postprocess(name):
    r, z = cases.load(name)
    lv, x = cases.decode3d_inputs(r, z)
    y = oracle.decode3d(x, lv, synth.STRIDES, r["nc"])
    np.testing.assert_allclose(y, z["y"], rtol=RTOL, atol=1e-4)
    reg, scores, labels, _ = oracle.postprocess(z["y"].transpose(0, 2, 1), r["D"], r["nc"], nreg=35, scores_first=True)
    assert np.array_equal(labels, z["labels"])
    assert np.array_equal(scores, z["scores"])
    assert np.array_equal(reg, z["reg"])


def test_decode_preds():
    r, z = cases.load("preds3d_small")
    rows, valid = oracle.decode_preds(z["dets"], z["calib"], z["inv_affine"], z["ratio"], z["cls_mean_size"])
    B = rows.shape[0]
    for b in range(B):
        kept = rows[b][valid[b]]
        n = int(z["counts"][b])
        assert kept.shape[0] == n
        np.testing.assert_allclose(kept, z["rows"][b, :n], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("name", cases.names("assign3d_"))
def test_tal_assign3d(name):
    r, z = cases.load(name)
    B, nc, M = r["B"], r["nc"], r["M"]
    lv = synth.levels(*r["img_hw"])
    anc, st = synth.anchors_px(lv)
    gts = z["gts"]
    mask_gt = (gts[..., 1:5].sum(-1) > 0).astype(np.float32)
    ms = np.array(synth.KITTI_MEAN_SIZES, np.float32)
    kw = dict(alpha=0.5, beta=3.0, gamma=3.0)  # TaskAlignedAssigner3d defaults tal.py:370
    kw.update(r["kw"])
    out = oracle.tal_assign3d(z["pd_scores"], z["pd_bboxes"], z["pd_3d"], anc, st, gts, mask_gt, z["calibs"], ms,
                              r["topk"], **kw)
    A = anc.shape[0]
    np.testing.assert_allclose(out["gt_keypoints"], z["gt_kps"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(out["pd_keypoints"][:, ::37], z["pd_kps_sample"], rtol=1e-4, atol=2e-5)
    assert np.array_equal(out["fg_mask"], cases.unpack_mask(z, "fg_mask", B, A))
    assert np.array_equal(out["target_gt_idx"], z["target_gt_idx"].astype(np.int64))
    assert np.array_equal(out["target_labels"], z["target_labels"].astype(np.int64))
    assert synth.checksum(out["target_vals"]) == int(z["target_vals_crc"])
    np.testing.assert_allclose(out["target_scores"], cases.dense_target_scores(z, B, A, nc), rtol=5e-5, atol=1e-7)
    assert out["fg_mask"].sum() > 0


@pytest.mark.parametrize("name", cases.names("loss3d_"))
def test_dd_loss_oracle_vs_reference(name):
    """oracle.dd_loss against the REAL DDDetectionLoss (loss.py:821-900) run on CPU by make_golden.py, including the
    GT packing of DDDetectionLoss.preprocess (loss.py:795-810)."""
    r, z = cases.load(name)
    lv, gts, x, calibs, ms = cases.loss3d_inputs(r, z)
    bd = synth.batch_dict3d(gts, r["img_hw"], calibs, ms)
    extra = np.concatenate([bd["center_2d"], bd["size_2d"], bd["center_3d"], bd["size_3d"], bd["depth"][:, None],
                            bd["heading_bin"][:, None], bd["heading_res"][:, None]], 1)
    packed = oracle.preprocess_targets(bd["batch_idx"], bd["cls"], bd["bboxes"], r["B"], r["img_hw"], extra=extra)
    np.testing.assert_allclose(packed, z["packed"], rtol=1e-6, atol=1e-4)
    items, tss, n_fg, asg = oracle.dd_loss(x, lv, synth.STRIDES, r["nc"], z["packed"], calibs, ms, r["topk"],
                                           gains=r["gains"], **cases.loss3d_kwargs(r))
    assert n_fg > 0
    np.testing.assert_allclose(items, z["items"], rtol=2e-5)
    np.testing.assert_allclose(items.sum() * r["B"], float(z["total"]), rtol=2e-5)


@pytest.mark.parametrize("name", cases.names("sparse_head_"))
def test_sparse_head_glue_oracle_vs_reference(name):
    """oracle.select_candidates / extract_patches / scatter_candidates against the REAL v10Detect3d methods
    (head.py:659-687) and the scatter-back of inference_forward_feat (head.py:709-714)."""
    r, z = cases.load(name)
    scores, x, vals = cases.sparse_head_inputs(r, z)
    idx = oracle.select_candidates(scores, r["K"])
    assert np.array_equal(idx, z["idx"].astype(np.int64))
    assert synth.checksum(oracle.extract_patches(x, idx)) == int(z["patches_crc"])
    out = oracle.scatter_candidates(vals, idx, (r["B"], r["Cout"], r["H"], r["W"]))
    assert synth.checksum(out) == int(z["scatter_crc"])


def test_rotate_iou_oracle_vs_reference():
    """oracle.rotate_iou_eval against the REAL rotate_iou_gpu_eval (kitti_eval.py:309-344, numba CUDA simulator)."""
    r, z = cases.load("rotate_iou_small")
    boxes = synth.bev_boxes(r["N"], r["seed"])
    assert synth.checksum(boxes, z["query"]) == int(z["in_crc"])
    for c, key in ((-1, "iou_cm1"), (0, "iou_c0"), (1, "iou_c1"), (2, "iou_c2")):
        got = oracle.rotate_iou_eval(boxes, z["query"], c)
        assert (z[key] > 0).sum() >= 20
        np.testing.assert_allclose(got, z[key], rtol=1e-5, atol=1e-6)
