"""Head decode: host-side mirror of the reference's ``Detect.inference`` / ``v10Detect.forward`` decode math
(ultralytics/nn/modules/head.py:53-79, 505-533) and ``v10Detect3d.decode/inference`` (head.py:755-797).

The convolutions stay in PyTorch/cuDNN (out of scope); these functions take what the head convs emit -- the list of
per-level tensors [B, C, H_l, W_l] -- and run the decode as one hand-written CUDA kernel (csrc/decode.cu).
"""
import torch

from . import _lib, ops
from ._util import Levels, ptr, stream_ptr, workspace

REG_MAX = 16  # head.py:37


def detect_inference(x, stride, nc, export=False):
    """``Detect.inference(x)`` (head.py:53-79): list of [B, 4*16+nc, H_l, W_l] -> y [B, 4+nc, A].

    Returns ``(y, x)`` like the reference, or ``y`` alone when ``export`` (which also switches the boxes from
    xywh to xyxy, head.py:105-108).  ``stride`` is the head's ``self.stride`` (one value per level)."""
    lv = Levels(x, [float(s) for s in stride])
    if lv.C != 4 * REG_MAX + nc:
        raise ValueError(f"expected {4 * REG_MAX + nc} channels, got {lv.C}")
    y = torch.empty((lv.B, 4 + nc, lv.A), dtype=torch.float32, device=lv.device)
    _lib.check(_lib.lib().y3d_decode2d(*lv.args(), lv.B, nc, REG_MAX, 0 if export else 1, ptr(y),
                                       stream_ptr(lv.device)))
    return y if export else (y, x)


def v10detect_export_forward(x_one2one, stride, nc, max_det=300, return_anchor_idx=False):
    """``v10Detect.forward`` with ``export=True`` (head.py:526-531): fused decode (xyxy) + ``ops.v10postprocess``.

    ``x_one2one``: the one2one branch's per-level tensors.  Returns [B, max_det, 6] = x1 y1 x2 y2 score label."""
    lv = Levels(x_one2one, [float(s) for s in stride])
    if lv.C != 4 * REG_MAX + nc:
        raise ValueError(f"expected {4 * REG_MAX + nc} channels, got {lv.C}")
    assert max_det != -1  # head.py:529
    D = int(max_det)
    out = torch.empty((lv.B, D, 6), dtype=torch.float32, device=lv.device)
    aidx = torch.empty((lv.B, D), dtype=torch.int32, device=lv.device) if return_anchor_idx else None
    ws = workspace(_lib.workspace_bytes(_lib.STAGE_DECODE_TOPK, B=lv.B, A=lv.A, nc=nc, D=D), lv.device)
    _lib.check(_lib.lib().y3d_decode_topk2d(*lv.args(), lv.B, nc, REG_MAX, 0, D, ptr(out), ptr(aidx), ptr(ws),
                                            ws.numel(), stream_ptr(lv.device)))
    return (out, aidx) if return_anchor_idx else out


class V10DetectDecoder:
    """Drop-in for the decode half of ``v10Detect`` (head.py:505-533): give it the raw branch outputs.

    Mirrors the attributes the reference head exposes to its callers: ``nc, nl, no, reg_max, stride, export,
    max_det``.  ``forward(one2many_feats, one2one_feats)`` reproduces head.py:519-533 for eval mode."""

    max_det = 300  # head.py:507

    def __init__(self, nc=80, stride=(8.0, 16.0, 32.0)):
        self.nc, self.reg_max = nc, REG_MAX
        self.no = nc + 4 * REG_MAX
        self.stride = torch.tensor([float(s) for s in stride])
        self.nl = len(stride)
        self.export = False

    def inference(self, x):
        return detect_inference(x, self.stride.tolist(), self.nc, export=self.export)

    def forward(self, one2many, one2one):
        if self.export:  # head.py:526-531
            return v10detect_export_forward(one2one, self.stride.tolist(), self.nc, self.max_det)
        return {"one2many": self.inference(one2many), "one2one": self.inference(one2one)}  # head.py:522-527

    __call__ = forward


# ------------------------------------------------------------------------------------------------ 3D head
def detect3d_decode(x, stride, nc):
    """``v10Detect3d.inference`` decode (head.py:755-797): list of [B, nc+35, H_l, W_l] -> ([B, nc+35, A], x);
    channels out = cls logits | bbox xyxy px | center3d px | s3d | hd(24) | dep | dep_un."""
    lv = Levels(x, [float(s) for s in stride])
    if lv.C != nc + 35:
        raise ValueError(f"expected {nc + 35} channels, got {lv.C}")
    y = torch.empty((lv.B, nc + 35, lv.A), dtype=torch.float32, device=lv.device)
    _lib.check(_lib.lib().y3d_decode3d(*lv.args(), lv.B, nc, ptr(y), stream_ptr(lv.device)))
    return y, x


def detect3d_postprocess(y, max_det=50, nc=3):
    """``YOLOv10_3DDetectionValidator.postprocess`` (models/yolov10_3D/val.py:33-47): [B, nc+35, A] ->
    [B, max_det, 37] = reg(35) | score (logit) | label."""
    reg, scores, labels = ops.v10_3Dpostprocess(y.transpose(-1, -2), max_det, nc)
    return torch.cat((reg, scores.unsqueeze(-1), labels.unsqueeze(-1).to(reg.dtype)), dim=-1)


# ------------------------------------------------------------------------------------------------ sparse 3D head glue
def _cuda4d(t, name):
    if t.dim() != 4:
        raise ValueError(f"{name} must be [B, C, H, W]")
    if not t.is_cuda:
        raise _lib.Y3DError("yolov10-3d_b200 runs on CUDA tensors only (no CPU fallback)")
    t = t.float() if t.dtype != torch.float32 else t
    if t.stride(3) != 1 or t.stride(2) != t.shape[3]:
        t = t.contiguous()
    return t


def select_candidates(scores, max_det):
    """``v10Detect3d.select_candidates`` (head.py:681-687): class logits of one level [B, nc, H, W] ->
    ``topk_indices`` [B, max_det, 2] int64 = (row, col), best first.  (The reference keeps this tensor on the CPU and
    loops over the batch; here it stays on the device.)"""
    s = _cuda4d(scores, "scores")
    B, nc, H, W = s.shape
    if max_det > H * W:
        raise RuntimeError("selected index k out of range")  # torch.topk
    idx = torch.empty((B, int(max_det), 2), dtype=torch.int64, device=s.device)
    _lib.check(_lib.lib().y3d_select_candidates(ptr(s), s.stride(0), s.stride(1), B, nc, H, W, int(max_det), ptr(idx),
                                                stream_ptr(s.device)))
    return idx


def extract_patches(x, indices, patch_size=5):
    """``v10Detect3d.extract_patches`` (head.py:659-679): x [B, C, H, W], indices [B, K, 2] ->
    [B*K, C, patch_size, patch_size] zero-padded neighbourhoods (no Python loop over B x K)."""
    xx = _cuda4d(x, "x")
    B, C, H, W = xx.shape
    idx = indices.to(device=xx.device, dtype=torch.int64).contiguous()
    K = int(idx.shape[1])
    out = torch.empty((B * K, C, patch_size, patch_size), dtype=torch.float32, device=xx.device)
    _lib.check(_lib.lib().y3d_extract_patches(ptr(xx), xx.stride(0), xx.stride(1), ptr(idx), B, C, H, W, K,
                                              int(patch_size), ptr(out), stream_ptr(xx.device)))
    return out


def scatter_candidates(values, indices, output_shape):
    """head.py:709-714: ``values`` [B*K, Cout] (or [B*K, Cout, 1, 1]: a head evaluated on the patches) scattered into a
    zero-filled ``output_shape`` = (B, Cout, H, W) at the candidate cells."""
    B, Cout, H, W = (int(v) for v in output_shape)
    v = values.reshape(values.shape[0], -1)
    if not v.is_cuda:
        raise _lib.Y3DError("yolov10-3d_b200 runs on CUDA tensors only (no CPU fallback)")
    v = v.float().contiguous()
    idx = indices.to(device=v.device, dtype=torch.int64).contiguous()
    K = int(idx.shape[1])
    if v.shape != (B * K, Cout):
        raise ValueError(f"values must be [{B * K}, {Cout}], got {tuple(v.shape)}")
    out = torch.empty((B, Cout, H, W), dtype=torch.float32, device=v.device)
    _lib.check(_lib.lib().y3d_scatter_candidates(ptr(v), ptr(idx), B, Cout, H, W, K, ptr(out), stream_ptr(v.device)))
    return out


def inference_forward_feat(det, x, heads):
    """``v10Detect3d.inference_forward_feat(x, heads)`` (head.py:694-716) with the head module passed as ``det``: per level,
    the class head runs on the whole feature map, the ``det.max_det`` best cells are selected, and the regression heads
    run on the 5 x 5 patches around them only; their outputs are scattered back into zero maps and concatenated behind
    the class map.  The convolutions stay ``torch`` modules (``heads[j][i]``); candidate selection, patch extraction and
    the scatter -- a ``for b in range(batch)`` loop with CPU index tensors in the reference -- are one kernel each and
    never leave the device.  Returns the list of per-level tensors [B, sum(output_channels), H_l, W_l]."""
    y = []
    head_names = list(det.output_channels.keys())
    out_ch = list(det.output_channels.values())
    patch = getattr(det, "patch_size", 5)
    for i in range(det.nl):
        outputs = {head_names[0]: heads[0][i](x[i])}
        cand = select_candidates(outputs[head_names[0]], det.max_det)
        inputs = extract_patches(x[i], cand, patch)
        for j, module in enumerate(heads[1:]):
            for layer in module[i]:  # head.py:704-706: the patch already carries the receptive field of the first conv
                if hasattr(layer, "conv") and hasattr(layer.conv, "padding") and type(layer).__name__ == "Conv":
                    layer.conv.padding = 0
            shape = (x[i].shape[0], out_ch[j + 1], x[i].shape[2], x[i].shape[3])
            vals = module[i](inputs)[:, :, 0, 0]  # [B * K, Cout]
            outputs[head_names[j + 1]] = scatter_candidates(vals, cand, shape)
        y.append(torch.cat(list(outputs.values()), dim=1))
    return y
