#!/usr/bin/env python
"""Developer tool: timeline of the fused-loss steps inside a back-to-back train (library built with -DY3D_TIMING by
tools/phase_timing.py build): when do the three kernels start / end relative to each other?  Run on the GPU box."""
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
a = (ctypes.c_ulonglong * 128)()
t = (ctypes.c_ulonglong * 64)()
h.y3d_debug_read_timeline.argtypes = [ctypes.c_void_p, ctypes.c_int]
h.y3d_debug_read_topk_timeline.argtypes = [ctypes.c_void_p, ctypes.c_int]
NOT = (1 << 64) - 1
for _ in range(20):
    y3d.loss.v10_loss_forward(fm, fo, list(synth.STRIDES), 80, gtd, (7.5, 0.5, 1.5))
torch.cuda.synchronize()
h.y3d_debug_read_timeline(a, 1)
h.y3d_debug_read_topk_timeline(t, 1)
N = 12
for _ in range(N):
    y3d.loss.v10_loss_forward(fm, fo, list(synth.STRIDES), 80, gtd, (7.5, 0.5, 1.5))
torch.cuda.synchronize()
h.y3d_debug_read_timeline(a, 0)
h.y3d_debug_read_topk_timeline(t, 0)
A = np.array(a, dtype=np.uint64).reshape(16, 8)
T = np.array(t, dtype=np.uint64).reshape(16, 4)
inv = lambda v: (NOT - int(v))
base = inv(A[0, 0])
names = ["stream first start", "stream last end", "topk first past wait", "topk prologue done (last)", "topk first warp exit",
         "topk last warp exit", "finish first CTA start", "finish first past image wait", "finish last R end", "finish final end"]
print("step |", " | ".join(n[:14] for n in names), "| period")
prev = None
for sidx in range(N):
    v = [inv(A[sidx, 0]), int(A[sidx, 1]), inv(T[sidx, 0]), int(T[sidx, 3]), inv(T[sidx, 2]), int(T[sidx, 1]), inv(A[sidx, 2]),
         inv(A[sidx, 3]), int(A[sidx, 4]), int(A[sidx, 5])]
    s0 = v[0]
    rel = [(x - s0) / 1e3 for x in v]
    per = (s0 - prev) / 1e3 if prev else float("nan")
    prev = s0
    print(f"{sidx:4d} |", " | ".join(f"{r:14.1f}" for r in rel), f"| {per:6.1f}")
