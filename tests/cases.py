"""Golden-fixture access: load tests/golden/<case>.npz and regenerate its inputs from the stored recipe."""
import glob
import json
import os

import numpy as np

from . import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def names(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.npz")))


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return json.loads(str(z["recipe"])), z


def dense_target_scores(z, B, A, nc):
    ts = np.zeros((B, A, nc), np.float32)
    idx = z["ts_idx"]
    ts[idx[:, 0], idx[:, 1], idx[:, 2]] = z["ts_val"]
    return ts


def unpack_mask(z, key, B, A):
    return np.unpackbits(z[key])[: B * A].reshape(B, A).astype(bool)


def decode_post_inputs(r, z):
    lv = synth.levels(*r["img_hw"])
    x = synth.head2d(r["B"], r["nc"], lv, seed=r["seed"])
    if r.get("quantise"):
        q = r["quantise"]
        x[:, 64:] = np.round(x[:, 64:] * q) / q
    assert synth.checksum(x) == int(z["in_crc"]), "regenerated inputs differ from the fixture's"
    return lv, x


def assign_inputs(r, z):
    """Same construction as tests/golden/make_golden.py::make_assign_inputs."""
    B, nc, img_hw, M, seed = r["B"], r["nc"], r["img_hw"], r["M"], r["seed"]
    lv = synth.levels(*img_hw)
    gt = synth.gt2d(B, M, nc, img_hw, seed=seed + 1, crowd=r["crowd"])
    if r["inputs"] == "random":
        x = synth.head2d(B, nc, lv, seed=seed)
    else:
        x = synth.train_like_head2d(B, nc, lv, gt, seed=seed + 2, frac=0.05)
    pd_scores, pd_bboxes, anc = synth.assigner_inputs_from_head(x, lv, nc)
    if r["inputs"] == "ties":
        gt[:, 1] = gt[:, 0]
        gt[0, 2, 1:5] = (0.0, 0.0, 70.0, 50.0)
        gt[1, 3, 1:5] = (1.0, 1.0, 30.0, 20.0)
        pd_scores = np.round(pd_scores * 8) / 8
        pd_bboxes = np.round(pd_bboxes / 4) * 4
        pd_scores[B - 1] = 0.0
        pd_scores = pd_scores.astype(np.float32)
        pd_bboxes = pd_bboxes.astype(np.float32)
    assert synth.checksum(gt, pd_scores, pd_bboxes, anc) == int(z["in_crc"]), "regenerated inputs differ"
    mask_gt = (gt[..., 1:5].sum(-1, keepdims=True) > 0).astype(np.float32)
    return dict(pd_scores=pd_scores, pd_bboxes=pd_bboxes, anc=anc, gt_labels=gt[..., :1].copy(),
                gt_bboxes=gt[..., 1:5].copy(), mask_gt=mask_gt, lvl_hw=lv)


def loss_inputs(r, z):
    B, nc, img_hw, M, seed = r["B"], r["nc"], r["img_hw"], r["M"], r["seed"]
    lv = synth.levels(*img_hw)
    gt = synth.gt2d(B, M, nc, img_hw, seed=seed + 1, crowd=r["crowd"])
    if r.get("ragged"):
        synth.make_ragged(gt, img_hw)
    if r["inputs"] == "random":
        xm, xo = synth.head2d(B, nc, lv, seed=seed), synth.head2d(B, nc, lv, seed=seed + 7)
    else:
        xm = synth.train_like_head2d(B, nc, lv, gt, seed=seed + 2, frac=0.05)
        xo = synth.train_like_head2d(B, nc, lv, gt, seed=seed + 3, frac=0.03)
    assert synth.checksum(gt, xm, xo) == int(z["in_crc"]), "regenerated inputs differ"
    return lv, gt, xm, xo


def sparse_head_inputs(r, z):
    """Same construction as tests/golden/make_golden.py::sparse_head_inputs."""
    g = synth.rng(r["seed"])
    B, nc, C, H, W, K, Cout = (r[k] for k in ("B", "nc", "C", "H", "W", "K", "Cout"))
    scores = (g.standard_normal((B, nc, H, W), dtype=np.float32) * 2 - 4).astype(np.float32)
    x = g.standard_normal((B, C, H, W), dtype=np.float32)
    vals = g.standard_normal((B * K, Cout), dtype=np.float32)
    if r.get("quantise"):
        scores = (np.round(scores * r["quantise"]) / r["quantise"]).astype(np.float32)
    assert synth.checksum(scores, x, vals) == int(z["in_crc"]), "regenerated inputs differ"
    return scores, x, vals


def loss3d_inputs(r, z):
    """Same construction as tests/golden/make_golden.py::case_loss3d."""
    B, nc, img_hw, M, seed = r["B"], r["nc"], r["img_hw"], r["M"], r["seed"]
    lv = synth.levels(*img_hw)
    gts = synth.gt3d(B, M, nc, img_hw, seed=seed + 1)
    x = synth.train_like_head3d(B, nc, lv, gts, seed=seed)
    assert synth.checksum(gts, x) == int(z["in_crc"]), "regenerated inputs differ"
    ms = np.array(synth.KITTI_MEAN_SIZES, np.float32)
    return lv, gts, x, z["calibs"].astype(np.float32), ms


def loss3d_kwargs(r):
    kw = r["kw"]
    return dict(alpha=kw.get("alpha", 0.5), beta=kw.get("beta", 1.0), gamma=kw.get("gamma", 1.0),
                use_2d=kw.get("use_2d", True), use_3d=kw.get("use_3d", True),
                kps_dist_metric=kw.get("kps_dist_metric", "l1"), constrain_anchors=kw.get("constrain_anchors", True))


def decode3d_inputs(r, z):
    lv = synth.levels(*r["img_hw"])
    x = synth.head3d(r["B"], r["nc"], lv, seed=r["seed"])
    assert synth.checksum(x) == int(z["in_crc"])
    return lv, x


def loss_assign_inputs(r, z):
    """Same construction as tests/golden/make_golden.py::case_loss_assign."""
    B, nc, img_hw, M, seed = r["B"], r["nc"], r["img_hw"], r["M"], r["seed"]
    lv = synth.levels(*img_hw)
    gt = synth.gt2d(B, M, nc, img_hw, seed=seed + 1, crowd=r["crowd"], full=r["crowd"])
    xm = synth.train_like_head2d(B, nc, lv, gt, seed=seed + 2, frac=r["frac"])
    xo = synth.train_like_head2d(B, nc, lv, gt, seed=seed + 3, frac=r["frac"])
    assert synth.checksum(gt, xm, xo) == int(z["in_crc"]), "regenerated inputs differ"
    return lv, gt, xm, xo


def loss_assign_expected(z, branch, B, A):
    """(fg_mask bool [B,A], target_gt_idx int64 [B,A] with 0 at background) of one branch of a lossasg_* fixture."""
    fg = np.unpackbits(z[f"fg{branch}"])[:B * A].reshape(B, A).astype(bool)
    tgi = np.zeros((B, A), np.int64)
    tgi[fg] = z[f"tgi{branch}"].astype(np.int64)
    return fg, tgi
