"""``TaskAlignedAssigner`` / ``TaskAlignedAssigner3d`` mirrors (reference ultralytics/utils/tal.py:19-264, 355-700):
same constructor arguments, same ``forward`` signatures, same return tuples and dtypes.  The work is done by
csrc/assign.cu; nothing dense in [B, M, A] is ever materialised."""
import torch
import torch.nn as nn

from . import _lib
from ._util import f32c, geometry, ptr, stream_ptr, workspace


def make_anchors(feats, strides, grid_cell_offset=0.5):
    """Anchor centres and stride column with the semantics of the reference helper (tal.py:300-312): level by level,
    row-major inside a level, ``(x, y) = (col, row) + grid_cell_offset`` in grid units.  Convenience for callers that want
    the tensors -- the kernels derive both from the cell index and never read them."""
    dtype, device = feats[0].dtype, feats[0].device
    points, stride_col = [], []
    for f, s in zip(feats, strides):
        h, w = int(f.shape[-2]), int(f.shape[-1])
        cell = torch.arange(h * w, device=device)
        xy = torch.stack((cell % w, cell // w), dim=1).to(dtype) + grid_cell_offset
        points.append(xy)
        stride_col.append(xy.new_full((h * w, 1), float(s)))
    return torch.cat(points), torch.cat(stride_col)


def _require_cuda(t):
    if not t.is_cuda:
        raise _lib.Y3DError("yolov10-3d_b200 runs on CUDA tensors only (no CPU fallback)")


class TaskAlignedAssigner(nn.Module):
    """tal.py:19-264.  ``grid=(lvl_hw, strides)`` is an optional promise that ``anc_points`` are
    ``make_anchors(...) * stride`` for those levels (what ``v8DetectionLoss`` passes, loss.py:231-238); the kernel
    then walks each GT's rectangle instead of all anchors.  Results are identical either way."""

    def __init__(self, topk=13, num_classes=80, alpha=1.0, beta=6.0, eps=1e-9, grid=None):
        super().__init__()
        self.topk = topk
        self.num_classes = num_classes
        self.bg_idx = num_classes
        self.alpha = alpha
        self.beta = beta
        self.eps = eps
        self.grid = grid

    @torch.no_grad()
    def forward(self, pd_scores, pd_bboxes, anc_points, gt_labels, gt_bboxes, mask_gt):
        self.bs = pd_scores.shape[0]
        self.n_max_boxes = gt_bboxes.shape[1]
        if self.n_max_boxes == 0:  # tal.py:68-76, dtypes as in the reference (all float)
            device = gt_bboxes.device
            return (
                torch.full_like(pd_scores[..., 0], self.bg_idx).to(device),
                torch.zeros_like(pd_bboxes).to(device),
                torch.zeros_like(pd_scores).to(device),
                torch.zeros_like(pd_scores[..., 0]).to(device),
                torch.zeros_like(pd_scores[..., 0]).to(device),
            )
        _require_cuda(pd_scores)
        B, A, nc = pd_scores.shape
        M = self.n_max_boxes
        dev = pd_scores.device
        pd_scores = pd_scores if pd_scores.dtype == torch.float32 else pd_scores.float()
        pd_bboxes, anc = f32c(pd_bboxes), f32c(anc_points)
        gl, gb, mg = f32c(gt_labels).view(B, M), f32c(gt_bboxes).view(B, M, 4), f32c(mask_gt).view(B, M)
        t_lab = torch.empty((B, A), dtype=torch.int64, device=dev)
        t_box = torch.empty((B, A, 4), dtype=torch.float32, device=dev)
        t_sc = torch.empty((B, A, nc), dtype=torch.float32, device=dev)
        fg = torch.empty((B, A), dtype=torch.bool, device=dev)
        t_gi = torch.empty((B, A), dtype=torch.int64, device=dev)
        ws = workspace(_lib.workspace_bytes(_lib.STAGE_TAL_ASSIGN, B=B, A=A, nc=nc, M=M, k=self.topk), dev)
        hw, st, nl = geometry(*self.grid) if self.grid is not None else (None, None, 0)
        sB, sA, sC = pd_scores.stride()
        _lib.check(_lib.lib().y3d_tal_assign(
            ptr(pd_scores), sB, sA, sC, ptr(pd_bboxes), ptr(anc), ptr(gl), ptr(gb), ptr(mg), B, A, nc, M,
            int(self.topk), float(self.alpha), float(self.beta), float(self.eps), hw, st, nl, ptr(t_lab), ptr(t_box),
            ptr(t_sc), ptr(fg), ptr(t_gi), ptr(ws), ws.numel(), stream_ptr(dev)))
        return t_lab, t_box, t_sc, fg, t_gi


class TaskAlignedAssigner3d(nn.Module):
    """tal.py:355-700.  ``forward`` returns ``(targets, fg_mask, target_gt_idx, pd_keypoints, gt_keypoints)`` with
    ``targets`` the reference's 9-element list."""

    def __init__(self, topk=8, num_classes=3, alpha=0.5, beta=3.0, gamma=3.0, eps=1e-9, use_2d=True, use_3d=True,
                 kps_dist_metric="l1", constrain_anchors=True, grid=None):
        super().__init__()
        self.topk = topk
        self.num_classes = num_classes
        self.bg_idx = num_classes
        self.alpha, self.beta, self.gamma, self.eps = alpha, beta, gamma, eps
        self.use_3d, self.use_2d = use_3d, use_2d
        self.kps_dist_metric = kps_dist_metric
        self.constrain_anchors = constrain_anchors
        self.grid = grid

    @torch.no_grad()
    def forward(self, pd_scores, pd_bboxes, pd_3d, anc_points, gts, mask_gt, stride_tensor, calibs, mean_sizes):
        gt_bboxes = gts[1]
        self.bs = pd_scores.shape[0]
        self.n_max_boxes = gt_bboxes.shape[1]
        self.num_anchors = anc_points.shape[0]
        if self.n_max_boxes == 0:  # tal.py:414-422
            device = gt_bboxes.device
            return (
                torch.full_like(pd_scores[..., 0], self.bg_idx).to(device),
                torch.zeros_like(pd_bboxes).to(device),
                torch.zeros_like(pd_scores).to(device),
                torch.zeros_like(pd_scores[..., 0]).to(device),
                torch.zeros_like(pd_scores[..., 0]).to(device),
            )
        if not (self.use_2d or self.use_3d):
            raise RuntimeError("Either 2D or 3D assignment or both has to be selected!")  # tal.py:486
        if self.kps_dist_metric not in ("l1", "l2"):
            raise ValueError("kps_dist_metric must be 'l1' or 'l2'")
        _require_cuda(pd_scores)
        B, A, nc = pd_scores.shape
        M = self.n_max_boxes
        dev = pd_scores.device
        packed = torch.cat([g.to(dev, torch.float32) for g in gts], dim=2).contiguous()  # [B,M,17]
        assert packed.shape[2] == 17
        pd_scores, pd_bboxes, pd_3d = f32c(pd_scores), f32c(pd_bboxes), f32c(pd_3d)
        anc, st = f32c(anc_points), f32c(stride_tensor).view(-1)
        mg = f32c(mask_gt).view(B, M)
        cal, ms = f32c(calibs.to(dev)), f32c(mean_sizes.to(dev))
        t_lab = torch.empty((B, A), dtype=torch.int64, device=dev)
        t_sc = torch.empty((B, A, nc), dtype=torch.float32, device=dev)
        t_vals = torch.empty((B, A, 12), dtype=torch.float32, device=dev)
        fg = torch.empty((B, A), dtype=torch.bool, device=dev)
        t_gi = torch.empty((B, A), dtype=torch.int64, device=dev)
        pk = torch.empty((B, A, 8, 3), dtype=torch.float32, device=dev)
        gk = torch.empty((B, M, 8, 3), dtype=torch.float32, device=dev)
        ws = workspace(_lib.workspace_bytes(_lib.STAGE_TAL_ASSIGN3D, B=B, A=A, nc=nc, M=M, k=self.topk), dev)
        flags = int(self.use_2d) | int(self.use_3d) << 1 | int(self.kps_dist_metric == "l2") << 2 | int(
            self.constrain_anchors) << 3
        hw, gst, nl = geometry(*self.grid) if self.grid is not None else (None, None, 0)
        _lib.check(_lib.lib().y3d_tal_assign3d(
            ptr(pd_scores), ptr(pd_bboxes), ptr(pd_3d), ptr(anc), ptr(st), ptr(packed), ptr(mg), ptr(cal), ptr(ms), B,
            A, nc, M, int(self.topk), float(self.alpha), float(self.beta), float(self.gamma), float(self.eps), flags,
            hw, gst, nl, ptr(t_lab), ptr(t_sc), ptr(t_vals), ptr(fg), ptr(t_gi), ptr(pk), ptr(gk), ptr(ws),
            ws.numel(), stream_ptr(dev)))
        c2d, s2d, c3d, s3d, dep, hb, hr = t_vals.split((2, 2, 2, 3, 1, 1, 1), dim=-1)
        targets = [t_lab, t_sc, c2d, s2d, c3d, s3d, dep, hb, hr]  # tal.py:699-700
        return targets, fg, t_gi, pk, gk
