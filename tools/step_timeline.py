#!/usr/bin/env python
"""Developer tool: timeline of one fused-loss step inside a back-to-back sequence (library built with -DY3D_TIMING by
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
a = (ctypes.c_ulonglong * 8)()
t = (ctypes.c_ulonglong * 4)()
h.y3d_debug_read_timeline.argtypes = [ctypes.c_void_p, ctypes.c_int]
h.y3d_debug_read_topk_timeline.argtypes = [ctypes.c_void_p, ctypes.c_int]
NOT = (1 << 64) - 1
for rep in range(3):
    for _ in range(10):
        y3d.loss.v10_loss_forward(fm, fo, list(synth.STRIDES), 80, gtd, (7.5, 0.5, 1.5))
    torch.cuda.synchronize()
    h.y3d_debug_read_timeline(a, 1)
    h.y3d_debug_read_topk_timeline(t, 1)
    # one step in the middle of a sequence: reset, run 3, the stamps keep min-of-first ... so run exactly ONE step between
    # two unstamped neighbours is not possible; instead run one step after a sync and one inside a train of 3
    y3d.loss.v10_loss_forward(fm, fo, list(synth.STRIDES), 80, gtd, (7.5, 0.5, 1.5))
    torch.cuda.synchronize()
    h.y3d_debug_read_timeline(a, 1)
    h.y3d_debug_read_topk_timeline(t, 1)
    s0 = NOT - a[0]
    ev = {"stream first start": 0.0, "stream last end": (a[1] - s0) / 1e3,
          "topk first past wait": (NOT - t[0] - s0) / 1e3, "topk prologue done (last)": (t[3] - s0) / 1e3,
          "topk first warp exit": (NOT - t[2] - s0) / 1e3, "topk last warp exit": (t[1] - s0) / 1e3,
          "finish first CTA start": (NOT - a[2] - s0) / 1e3, "finish first past image wait": (NOT - a[3] - s0) / 1e3,
          "finish last R end": (a[4] - s0) / 1e3, "finish final end": (a[5] - s0) / 1e3}
    print(f"--- isolated step (us from the first stream CTA), rep {rep}")
    for k, v in ev.items():
        print(f"   {k:32s} {v:8.1f}")
