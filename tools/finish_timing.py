#!/usr/bin/env python
"""Developer tool: per-CTA timestamps of loss_finish_ap_kernel (library built with -DY3D_TIMING by
tools/phase_timing.py build).  Run on the GPU box: python tools/finish_timing.py"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import yolov10_3d_b200 as y3d
from yolov10_3d_b200 import _lib

_lib.LIB_PATH = os.path.join(ROOT, "tools", "liby3d_timing.so")
import bench
from tests import synth

lv, gt, xm, xo = bench.make_inputs(seed=0)
dev = torch.device("cuda", 0)
fm = [torch.from_numpy(f).to(dev) for f in synth.split_levels(xm, lv)]
fo = [torch.from_numpy(f).to(dev) for f in synth.split_levels(xo, lv)]
gtd = torch.from_numpy(gt).to(dev)
h = _lib.lib()
h.y3d_debug_read_stamps.argtypes = [ctypes.c_void_p, ctypes.c_int]
h.y3d_debug_read_topk_timeline.argtypes = [ctypes.c_void_p, ctypes.c_int]
for _ in range(8):
    y3d.loss.v10_loss_forward(fm, fo, list(synth.STRIDES), 80, gtd, (7.5, 0.5, 1.5))
torch.cuda.synchronize()
n = 2048 * 8
buf = (ctypes.c_ulonglong * n)()
tk = (ctypes.c_ulonglong * 64)()
h.y3d_debug_read_topk_timeline(tk, 1)
y3d.loss.v10_loss_forward(fm, fo, list(synth.STRIDES), 80, gtd, (7.5, 0.5, 1.5))
torch.cuda.synchronize()
assert h.y3d_debug_read_stamps(buf, n) == 0
h.y3d_debug_read_topk_timeline(tk, 0)
T = np.array(tk, dtype=np.uint64).reshape(16, 4)
row = int(np.argmax(T[:, 1]))
NOT = (1 << 64) - 1
tk_first, tk_last_exit, tk_first_exit = NOT - int(T[row, 0]), int(T[row, 1]), NOT - int(T[row, 2])
t = np.array(buf, dtype=np.uint64).reshape(2048, 8).astype(np.int64)[:1024].reshape(2, 64, 8, 8)
base = tk_first
rel = lambda x: (x - base) / 1e3
print(f"top-k: first warp start 0.0, first exit {rel(tk_first_exit):.1f}, last exit {rel(tk_last_exit):.1f} us")
valid = (gt[..., 1:5].sum(-1) > 0).sum(1)
# chunk-0 CTAs ordered by the time they got past the image wait
c0 = t[:, :, 0, :]
order = np.dstack(np.unravel_index(np.argsort(c0[..., 1], axis=None), c0[..., 1].shape))[0]
print("last 8 images to become ready: (z, b, nGT) | past wait | R end | S end | acc | ticket | final   (us after top-k start)")
for z, b in order[-8:]:
    r = c0[z, b]
    vals = [rel(int(v)) if v > 0 and v >= r[0] else float("nan") for v in r[1:7]]
    print(f"   z={z} b={b:2d} nGT={valid[b]:3d} | " + " | ".join(f"{v:7.1f}" for v in vals))
d = c0[..., 2] - c0[..., 1]
print("chunk-0 R duration mean/max (us):", d.mean() / 1e3, d.max() / 1e3)
d = c0[..., 5] - c0[..., 2]
print("R end -> ticket mean/max (us):", d.mean() / 1e3, d.max() / 1e3)
