"""``ops.v10postprocess`` / ``ops.v10_3Dpostprocess`` mirrors (reference ultralytics/utils/ops.py:852-880),
plus ``xywh2xyxy`` (ops.py:403-422).  Same signatures, return types and dtypes; ties break lowest-index-first."""
import torch

from . import _lib
from ._util import ptr, stream_ptr, workspace


def _postprocess(preds, max_det, nc, nreg, scores_first, return_anchor_idx=False):
    if preds.dim() != 3:
        raise ValueError("preds must be [B, A, C]")
    assert nreg + nc == preds.shape[-1]  # ops.py:853 / :868
    if not preds.is_cuda:
        raise _lib.Y3DError("yolov10-3d_b200 runs on CUDA tensors only (no CPU fallback)")
    if preds.dtype != torch.float32:
        preds = preds.float()
    B, A, _ = preds.shape
    D = int(max_det)
    if D > A:  # torch.topk in the reference raises the same way
        raise RuntimeError("selected index k out of range")
    dev = preds.device
    reg = torch.empty((B, D, nreg), dtype=torch.float32, device=dev)
    scores = torch.empty((B, D), dtype=torch.float32, device=dev)
    labels = torch.empty((B, D), dtype=torch.int64, device=dev)
    aidx = torch.empty((B, D), dtype=torch.int32, device=dev) if return_anchor_idx else None
    ws = workspace(_lib.workspace_bytes(_lib.STAGE_POSTPROCESS, B=B, A=A, nc=nc, D=D), dev)
    sB, sA, sC = preds.stride()  # the reference passes a permuted view: strides are honoured, no copy
    _lib.check(_lib.lib().y3d_postprocess(ptr(preds), sB, sA, sC, B, A, nc, nreg, int(scores_first), D, ptr(reg),
                                          ptr(scores), ptr(labels), ptr(aidx), ptr(ws), ws.numel(), stream_ptr(dev)))
    if return_anchor_idx:
        return reg, scores, labels, aidx
    return reg, scores, labels


def v10postprocess(preds, max_det, nc=80):
    """[B, A, 4+nc] (boxes | scores) -> boxes [B,D,4], scores [B,D], labels [B,D] int64."""
    return _postprocess(preds, max_det, nc, 4, False)


def v10_3Dpostprocess(preds, max_det, nc=3):
    """[B, A, nc+35] (score logits | 35 regression channels) -> reg [B,D,35], scores [B,D], labels [B,D] int64."""
    return _postprocess(preds, max_det, nc, preds.shape[-1] - nc, True)


def xywh2xyxy(x):
    """(cx, cy, w, h) -> (x1, y1, x2, y2) on the last axis, torch tensor or numpy array (the contract of ops.py:403-422;
    the validators call it right after ``v10postprocess``).  Halving is exact in binary floating point, so the result
    equals the reference's ``x - w / 2`` bit for bit."""
    if x.shape[-1] != 4:
        raise AssertionError(f"input shape last dimension expected 4 but input shape is {x.shape}")
    centre, half = x[..., :2], x[..., 2:] * 0.5
    if torch.is_tensor(x):
        return torch.cat((centre - half, centre + half), dim=-1)
    import numpy as np

    return np.concatenate((centre - half, centre + half), axis=-1)
