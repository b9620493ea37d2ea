#!/usr/bin/env python
"""Developer tool: phase stamps of topk_select_kernel (CTA 0) inside the fused decode + top-k, at a given batch / size
(library built with -DY3D_TIMING by `python tools/phase_timing.py build`).

    python tools/select_timing.py 32 1280      # cfg4 shape, on the GPU box
"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from yolov10_3d_b200 import _lib  # noqa: E402

_lib.LIB_PATH = os.path.join(ROOT, "tools", "liby3d_timing.so")
import yolov10_3d_b200 as y3d  # noqa: E402
from tests import synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
hw = (int(sys.argv[2]),) * 2 if len(sys.argv) > 2 else (1280, 1280)
lv = synth.levels(*hw)
x = synth.head2d(4, 80, lv, seed=0)
x = np.concatenate([x] * ((B + 3) // 4), 0)[:B]
f = [torch.from_numpy(v).cuda() for v in synth.split_levels(x, lv)]
for _ in range(3):
    y3d.v10detect_export_forward(f, synth.STRIDES, 80, 300)
torch.cuda.synchronize()
h = _lib.lib()
st = (ctypes.c_longlong * 16)()
h.y3d_debug_read_sel_stamps.argtypes = [ctypes.c_void_p]
h.y3d_debug_read_sel_stamps(st)
sa = np.array(st, dtype=np.int64)
mhz = 1965.0
us = lambda c: round(float(c) / mhz, 2)
print(f"B={B} {hw}: select kernel CTA 0 phases (us at {mhz:.0f} MHz): stage-1 top-D", us(sa[1] - sa[0]), " gather+compact",
      us(sa[2] - sa[1]), " stage-2 sort", us(sa[3] - sa[2]), " outputs", us(sa[4] - sa[3]), " total", us(sa[4] - sa[0]))
print("  (last block_topk call) radix passes", us(sa[9] - sa[8]), " collect", us(sa[10] - sa[9]), " sort", us(sa[11] - sa[10]),
      " L", sa[12], " mask", hex(int(sa[13]) & 0xffffffff), " key staging", us(sa[8] - sa[0]))
