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
for _ in range(5):
    y3d.loss.v10_loss_forward(fm, fo, list(synth.STRIDES), 80, gtd, (7.5, 0.5, 1.5))
torch.cuda.synchronize()
h = _lib.lib()
n = 2048 * 8
buf = (ctypes.c_ulonglong * n)()
h.y3d_debug_read_stamps.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert h.y3d_debug_read_stamps(buf, n) == 0
t = np.array(buf, dtype=np.uint64).reshape(2048, 8).astype(np.int64)[:1024]
# CTA index = (z * B + b) * chunks + chunk, chunks = 8
t = t.reshape(2, 64, 8, 8)
t0 = t[..., 0][t[..., 0] > 0].min()
rel = lambda x: np.where(x > 0, (x - t0) / 1e3, np.nan)
print("start (us): min/mean/max", np.nanmin(rel(t[..., 0])), np.nanmean(rel(t[..., 0])), np.nanmax(rel(t[..., 0])))
print("after wait (1): min/mean/max", np.nanmin(rel(t[..., 1])), np.nanmean(rel(t[..., 1])), np.nanmax(rel(t[..., 1])))
valid = (gt[..., 1:5].sum(-1) > 0).sum(1)
for z in range(2):
    for b in (0, 1, 2, 5):
        print(f"z={z} b={b} nGT={valid[b]}")
        for ch in range(8):
            r = rel(t[z, b, ch].astype(np.float64))
            print("   chunk", ch, np.round(r[:7], 1))
s1 = rel(t[..., 1]); s2 = rel(t[..., 2]); s3 = rel(t[..., 3]); s4 = rel(t[..., 4]); s5 = rel(t[..., 5]); s6 = rel(t[..., 6])
act = t[..., 2] >= t[..., 1]
print("phase R duration mean/max:", np.nanmean((s2 - s1)[act]), np.nanmax((s2 - s1)[act]))
print("fence+count mean/max:", np.nanmean((s3 - s2)[act]), np.nanmax((s3 - s2)[act]))
last = t[..., 4] >= t[..., 3]
print("phase S mean/max:", np.nanmean((s4 - s3)[last & act]), np.nanmax((s4 - s3)[last & act]))
print("ticket mean/max:", np.nanmean((s5 - s4)[last & act]), np.nanmax((s5 - s4)[last & act]))
print("latest R end", np.nanmax(s2[act]), " latest S end", np.nanmax(s4[last & act]), " final", np.nanmax(np.where(t[..., 6] >= t[..., 5], s6, np.nan)))
