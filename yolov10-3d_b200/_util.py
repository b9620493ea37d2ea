"""Small host-side helpers shared by the API mirrors: level tables, stream, workspace."""
import ctypes as C

import torch

from . import _lib

DEFAULT_STRIDES = (8.0, 16.0, 32.0)


class Levels:
    """Host-side description of the head's per-level tensors (include/y3d.h 'Head geometry')."""

    def __init__(self, feats, strides):
        if len(feats) < 1 or len(feats) > _lib.MAX_LEVELS:
            raise _lib.Y3DError(f"1..{_lib.MAX_LEVELS} head levels supported, got {len(feats)}")
        if len(strides) != len(feats):
            raise ValueError("one stride per level")
        self.feats = []
        for f in feats:
            if f.dim() != 4:
                raise ValueError("head level tensors must be [B, C, H, W]")
            if f.dtype != torch.float32:
                f = f.float()  # the path is fp32 (AMP note)
            if f.stride(3) != 1 or f.stride(2) != f.shape[3]:
                f = f.contiguous()
            self.feats.append(f)
        f0 = self.feats[0]
        if not f0.is_cuda:
            raise _lib.Y3DError("yolov10-3d_b200 runs on CUDA tensors only (no CPU fallback)")
        self.device = f0.device
        self.B, self.C = f0.shape[0], f0.shape[1]
        self.nl = len(self.feats)
        self.hw = [(int(f.shape[2]), int(f.shape[3])) for f in self.feats]
        self.A = sum(h * w for h, w in self.hw)
        self.strides = [float(s) for s in strides]
        n = self.nl
        self.c_ptr = (C.c_void_p * n)(*[f.data_ptr() for f in self.feats])
        self.c_sB = (C.c_int64 * n)(*[f.stride(0) for f in self.feats])
        self.c_sC = (C.c_int64 * n)(*[f.stride(1) for f in self.feats])
        self.c_hw = (C.c_int * (2 * n))(*[v for hw in self.hw for v in hw])
        self.c_stride = (C.c_float * n)(*self.strides)

    def args(self):
        return (self.c_ptr, self.c_sB, self.c_sC, self.c_hw, self.c_stride, self.nl)

    @staticmethod
    def from_cat(xcat, lvl_hw, strides):
        """x_cat [B, C, A] (head.py:56) viewed as levels without copying."""
        xcat = xcat if xcat.is_contiguous() else xcat.contiguous()
        feats, o = [], 0
        B, Cc, _ = xcat.shape
        for h, w in lvl_hw:
            feats.append(xcat[:, :, o:o + h * w].unflatten(2, (h, w)))
            o += h * w
        return Levels(feats, strides)


def geometry(lvl_hw, strides):
    n = len(lvl_hw)
    return (C.c_int * (2 * n))(*[int(v) for hw in lvl_hw for v in hw]), (C.c_float * n)(*[float(s) for s in strides]), n


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def workspace(nbytes, device):
    """Scratch from torch's caching allocator (256-byte aligned by construction: allocations are 512 B aligned)."""
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def f32c(t):
    """fp32 + contiguous without copying when already so."""
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()
