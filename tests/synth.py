"""Seeded synthetic inputs for the head hot path.

numpy PCG64 streams only (bit-stable across hosts for one numpy version), so the GPU box
regenerates exactly the tensors the golden fixtures were made from; every fixture also stores
a checksum of its regenerated inputs.  Shapes follow the reference tensors:

  head tensor  x_cat [B, 4*16+nc, A]   (Detect.inference head.py:56)    / 3D: [B, nc+35, A]
  GT (packed)  [B, M, 5] = cls, xyxy px, zero rows = padding (v8DetectionLoss.preprocess loss.py:180)
  GT (3D)      [B, M, 17] = cls, xyxy, center_2d, size_2d, center_3d, size_3d(res), depth, hbin, hres
"""
import zlib

import numpy as np

STRIDES = (8.0, 16.0, 32.0)
# KITTI calibration / class mean sizes used by the 3D cases (kitti.py:38-41; values are h, w, l)
KITTI_CALIB = (628.3, 177.0, 743.6, 738.8, -0.06, 0.003)
KITTI_MEAN_SIZES = ((1.76255119, 0.66068622, 0.84422524),
                    (1.52563191, 1.62856739, 3.52588311),
                    (1.73698127, 0.59706367, 1.76282397))


def rng(seed):
    return np.random.Generator(np.random.PCG64(seed))


def levels(img_h, img_w, strides=STRIDES):
    """[(h_l, w_l)] for each stride; A = sum h_l * w_l."""
    return [(int(img_h // s), int(img_w // s)) for s in strides]


def num_anchors(lvl_hw):
    return int(sum(h * w for h, w in lvl_hw))


def checksum(*arrays):
    c = 0
    for a in arrays:
        c = zlib.crc32(np.ascontiguousarray(a).view(np.uint8).reshape(-1), c)
    return c


def anchors_px(lvl_hw, strides=STRIDES):
    pts, st = [], []
    for (h, w), s in zip(lvl_hw, strides):
        ys, xs = np.meshgrid(np.arange(h, dtype=np.float32) + 0.5, np.arange(w, dtype=np.float32) + 0.5, indexing="ij")
        pts.append(np.stack([xs, ys], -1).reshape(-1, 2) * np.float32(s))
        st.append(np.full((h * w,), s, np.float32))
    return np.concatenate(pts), np.concatenate(st)


def head2d(B, nc, lvl_hw, seed=0, reg_max=16):
    """Random head tensor: box channels randn*1.5, class channels randn*2-4 (sparse sigmoid scores)."""
    g = rng(seed)
    A = num_anchors(lvl_hw)
    x = g.standard_normal((B, 4 * reg_max + nc, A), dtype=np.float32)
    x[:, : 4 * reg_max] *= np.float32(1.5)
    x[:, 4 * reg_max:] *= np.float32(2.0)
    x[:, 4 * reg_max:] -= np.float32(4.0)
    return x


def split_levels(xcat, lvl_hw):
    """x_cat [B,C,A] -> list of contiguous [B,C,h_l,w_l] (what the head convs emit)."""
    out, o = [], 0
    B, C, _ = xcat.shape
    for h, w in lvl_hw:
        out.append(np.ascontiguousarray(xcat[:, :, o:o + h * w]).reshape(B, C, h, w))
        o += h * w
    return out


def gt2d(B, M, nc, img_hw, seed=1, crowd=False, full=False):
    """Packed GT [B,M,5]: n_i ~ U{1..M} boxes per image (image 0 forced to M; all images when ``full``)."""
    g = rng(seed)
    H, W = img_hw
    out = np.zeros((B, M, 5), np.float32)
    for b in range(B):
        n = M if (b == 0 or full) else int(g.integers(1, M + 1))
        cx = g.uniform(0.1, 0.9, n) * W
        cy = g.uniform(0.1, 0.9, n) * H
        lo, hi = (0.01, 0.08) if crowd else (0.03, 0.33)
        w = g.uniform(lo, hi, n) * W
        h = g.uniform(lo, hi, n) * H
        out[b, :n, 0] = g.integers(0, nc, n)
        out[b, :n, 1] = cx - w / 2
        out[b, :n, 2] = cy - h / 2
        out[b, :n, 3] = cx + w / 2
        out[b, :n, 4] = cy + h / 2
    return out


def make_ragged(gt, img_hw):
    """Edge cases of the ragged ground truth, in place on a packed [B>=3, M>=3, 5]: image 1 without any box, image 2
    with a single one; in image 0 a box smaller than a stride-8 cell that contains no anchor centre (no candidate at
    all), a box covering most of the image, and a zero-area box (a valid row -- its coordinates sum to > 0 -- that no
    anchor can lie inside)."""
    H, W = img_hw
    gt[1] = 0
    gt[2, 1:] = 0
    gt[0, 0, 1:5] = (17.0, 17.0, 19.5, 19.5)           # between the centres 12 and 20
    gt[0, 1, 1:5] = (0.05 * W, 0.05 * H, 0.95 * W, 0.95 * H)
    gt[0, 2, 1:5] = (0.5 * W, 0.25 * H, 0.5 * W, 0.75 * H)  # zero width
    return gt


def batch_dict(gt_packed, img_hw):
    """Packed GT -> the raw dataloader form v8DetectionLoss consumes (loss.py:222): batch_idx [N], cls [N,1],
    bboxes [N,4] normalised xywh."""
    H, W = img_hw
    valid = gt_packed[..., 1:5].sum(-1) > 0
    bi, mi = np.nonzero(valid)
    rows = gt_packed[bi, mi]
    x1, y1, x2, y2 = rows[:, 1], rows[:, 2], rows[:, 3], rows[:, 4]
    bb = np.stack([(x1 + x2) / 2 / W, (y1 + y2) / 2 / H, (x2 - x1) / W, (y2 - y1) / H], 1).astype(np.float32)
    return dict(batch_idx=bi.astype(np.float32), cls=rows[:, :1].astype(np.float32), bboxes=bb)


def train_like_head2d(B, nc, lvl_hw, gt_packed, seed=2, frac=0.02, reg_max=16):
    """Head tensor whose decoded boxes / class logits look 'trained' around each GT: for ~``frac`` of the anchors
    inside a GT the DFL logits peak at the GT's l/t/r/b distances (+ noise) and the GT-class logit gets +6, so that
    positive alignment metrics and multi-GT conflicts really occur."""
    g = rng(seed)
    x = head2d(B, nc, lvl_hw, seed=seed + 1000, reg_max=reg_max)
    anc, st = anchors_px(lvl_hw)
    bins = np.arange(reg_max, dtype=np.float32)
    for b in range(B):
        for m in range(gt_packed.shape[1]):
            lab, x1, y1, x2, y2 = gt_packed[b, m]
            if x1 + y1 + x2 + y2 <= 0:
                continue
            inside = np.nonzero((anc[:, 0] > x1) & (anc[:, 0] < x2) & (anc[:, 1] > y1) & (anc[:, 1] < y2))[0]
            if inside.size == 0:
                continue
            n = max(2, int(round(frac * inside.size)))
            pick = g.choice(inside, size=min(n, inside.size), replace=False)
            for a in pick:
                s = st[a]
                d = np.array([anc[a, 0] - x1, anc[a, 1] - y1, x2 - anc[a, 0], y2 - anc[a, 1]], np.float32) / s
                d = d + g.standard_normal(4).astype(np.float32) * np.float32(4.0) / s
                d = np.clip(d, 0.0, reg_max - 1.0)
                for side in range(4):
                    x[b, side * reg_max:(side + 1) * reg_max, a] = -2.0 * (bins - d[side]) ** 2
                x[b, 4 * reg_max + int(lab), a] += np.float32(6.0)
    return x


def assigner_inputs_from_head(xcat, lvl_hw, nc, reg_max=16):
    """Training-side decode in plain numpy (loss.py:197-204,231-236) -> pd_scores [B,A,nc] (sigmoid),
    pd_bboxes [B,A,4] xyxy px, anc_points [A,2] px.  Only used to *make inputs* for assigner-level cases."""
    B, C, A = xcat.shape
    anc, st = anchors_px(lvl_hw)
    box = xcat[:, : 4 * reg_max].reshape(B, 4, reg_max, A).astype(np.float32)
    e = np.exp(box - box.max(2, keepdims=True))
    p = e / e.sum(2, keepdims=True)
    d = (p * np.arange(reg_max, dtype=np.float32)[None, None, :, None]).sum(2)  # [B,4,A] grid units
    ag = anc / st[:, None]
    x1 = (ag[None, :, 0] - d[:, 0]) * st
    y1 = (ag[None, :, 1] - d[:, 1]) * st
    x2 = (ag[None, :, 0] + d[:, 2]) * st
    y2 = (ag[None, :, 1] + d[:, 3]) * st
    pd_bboxes = np.stack([x1, y1, x2, y2], -1).astype(np.float32)
    logits = np.ascontiguousarray(xcat[:, 4 * reg_max:].transpose(0, 2, 1))
    pd_scores = (1.0 / (1.0 + np.exp(-logits.astype(np.float64)))).astype(np.float32)
    return pd_scores, pd_bboxes, anc


# ------------------------------------------------------------------------------------------------ 3D
def head3d(B, nc, lvl_hw, seed=0):
    """3D head tensor [B, nc+35, A]: cls, o2d(2), s2d(2), o3d(2), s3d(3), hd(24), dep, dep_un."""
    g = rng(seed)
    A = num_anchors(lvl_hw)
    x = np.empty((B, nc + 35, A), np.float32)
    x[:, :nc] = g.standard_normal((B, nc, A), dtype=np.float32) * 2 - 4
    x[:, nc:nc + 2] = g.standard_normal((B, 2, A), dtype=np.float32) * 0.5
    x[:, nc + 2:nc + 4] = g.uniform(1, 20, (B, 2, A)).astype(np.float32)
    x[:, nc + 4:nc + 6] = g.standard_normal((B, 2, A), dtype=np.float32) * 0.5
    x[:, nc + 6:nc + 9] = g.standard_normal((B, 3, A), dtype=np.float32) * 0.2
    x[:, nc + 9:nc + 33] = g.standard_normal((B, 24, A), dtype=np.float32)
    x[:, nc + 33] = g.uniform(3, 53, (B, A)).astype(np.float32)
    x[:, nc + 34] = g.standard_normal((B, A), dtype=np.float32)
    return x


def gt3d(B, M, nc, img_hw, seed=1, full=False):
    """Packed 3D GT [B,M,17] (see module docstring); zero rows = padding."""
    g = rng(seed)
    base = gt2d(B, M, nc, img_hw, seed=seed + 500, full=full)
    out = np.zeros((B, M, 17), np.float32)
    out[..., :5] = base
    for b in range(B):
        n = int((base[b, :, 1:5].sum(-1) > 0).sum())
        x1, y1, x2, y2 = (base[b, :n, i] for i in (1, 2, 3, 4))
        c2 = np.stack([(x1 + x2) / 2, (y1 + y2) / 2], 1)
        out[b, :n, 5:7] = c2
        out[b, :n, 7:9] = np.stack([x2 - x1, y2 - y1], 1)
        out[b, :n, 9:11] = c2 + g.standard_normal((n, 2)) * 3
        out[b, :n, 11:14] = g.standard_normal((n, 3)) * 0.2
        out[b, :n, 14] = g.uniform(3, 53, n)
        out[b, :n, 15] = g.integers(0, 12, n)
        out[b, :n, 16] = g.uniform(-0.25, 0.25, n)
    return out.astype(np.float32)


def assigner3d_inputs_from_head(xcat, lvl_hw, nc):
    """DDDetectionLoss-side decode in numpy (loss.py:812-819,830-862): pd_scores (sigmoid) [B,A,nc],
    pd_bboxes [B,A,4] px, pd_3d [B,A,31], anc px [A,2], stride [A]."""
    B, C, A = xcat.shape
    anc, st = anchors_px(lvl_hw)
    ag = anc / st[:, None]
    t = np.ascontiguousarray(xcat.transpose(0, 2, 1))
    pd_scores = (1.0 / (1.0 + np.exp(-t[..., :nc].astype(np.float64)))).astype(np.float32)
    centers = ag[None] + t[..., nc:nc + 2]
    size = t[..., nc + 2:nc + 4]
    pd_bboxes = (np.concatenate([centers - size / 2, centers + size / 2], -1) * st[None, :, None]).astype(np.float32)
    pd_3d = np.ascontiguousarray(t[..., nc + 4:])
    return pd_scores, pd_bboxes, pd_3d, anc, st


def train_like_head3d(B, nc, lv, gts, seed, frac=0.05):
    """head3d with ~frac of the in-GT anchors of every GT made 'trained' (2D box, 3D centre, size, depth, heading and
    class close to the GT), so that positive metrics, conflicts and non-trivial loss terms occur."""
    x = head3d(B, nc, lv, seed=seed)
    anc, st = anchors_px(lv)
    g = rng(seed + 5)
    for b in range(B):
        for m in range(gts.shape[1]):
            row = gts[b, m]
            if row[1:5].sum() <= 0:
                continue
            ins = np.nonzero((anc[:, 0] > row[1]) & (anc[:, 0] < row[3]) & (anc[:, 1] > row[2]) & (anc[:, 1] < row[4]))[0]
            if ins.size == 0:
                continue
            for a in g.choice(ins, size=min(ins.size, max(2, int(frac * ins.size))), replace=False):
                s_ = st[a]
                x[b, int(row[0]), a] = 1.0 + 2.0 * g.random()
                x[b, nc + 0:nc + 2, a] = (row[5:7] - anc[a]) / s_ + g.standard_normal(2) * 0.3
                x[b, nc + 2:nc + 4, a] = row[7:9] / s_ + g.standard_normal(2) * 0.3
                x[b, nc + 4:nc + 6, a] = (row[9:11] - anc[a]) / s_ + g.standard_normal(2) * 0.05
                x[b, nc + 6:nc + 9, a] = row[11:14] + g.standard_normal(3) * 0.05
                x[b, nc + 9 + int(row[15]), a] += 4.0
                x[b, nc + 21 + int(row[15]), a] = row[16] + g.standard_normal() * 0.05
                x[b, nc + 33, a] = row[14] + g.standard_normal() * 0.5
    return np.ascontiguousarray(x, np.float32)


def batch_dict3d(gts, img_hw, calibs, mean_sizes):
    """The batch keys DDDetectionLoss.__call__ reads (loss.py:848-856) for packed GT rows [B,M,17] (bbox xyxy px);
    numpy arrays."""
    h, w = img_hw
    rows, bi = [], []
    for b in range(gts.shape[0]):
        for m in range(gts.shape[1]):
            if gts[b, m, 1:5].sum() > 0:
                rows.append(gts[b, m])
                bi.append(b)
    r = np.array(rows, np.float32).reshape(-1, 17)
    x1, y1, x2, y2 = r[:, 1], r[:, 2], r[:, 3], r[:, 4]
    bb = np.stack([(x1 + x2) / 2 / w, (y1 + y2) / 2 / h, (x2 - x1) / w, (y2 - y1) / h], 1).astype(np.float32)
    c = np.ascontiguousarray
    return dict(batch_idx=np.array(bi, np.float32), cls=c(r[:, 0:1]), bboxes=bb, center_2d=c(r[:, 5:7]),
                size_2d=c(r[:, 7:9]), center_3d=c(r[:, 9:11]), size_3d=c(r[:, 11:14]), depth=c(r[:, 14]),
                heading_bin=c(r[:, 15]), heading_res=c(r[:, 16]), calib=np.asarray(calibs, np.float32),
                mean_sizes=np.asarray(mean_sizes, np.float32))


def bev_boxes(n, seed):
    """Bird's-eye-view boxes [n, 5] = (x, z, l, w, ry) like the KITTI evaluator builds them (kitti_eval.py:458-461):
    car-sized rectangles scattered densely enough in a 40 m x 40 m patch that many pairs overlap."""
    g = rng(seed)
    out = np.empty((n, 5), np.float32)
    out[:, 0] = g.uniform(-20, 20, n)
    out[:, 1] = g.uniform(5, 45, n)
    out[:, 2] = g.uniform(2.5, 6.0, n)
    out[:, 3] = g.uniform(1.4, 2.6, n)
    out[:, 4] = g.uniform(-np.pi, np.pi, n)
    return out
