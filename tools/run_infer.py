#!/usr/bin/env python
"""Developer tool: a few calls of the inference path (decode2d, v10postprocess, fused decode+top-k) for ncu."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import yolov10_3d_b200 as y3d  # noqa: E402
from tests import synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
hw = (int(sys.argv[2]), int(sys.argv[2])) if len(sys.argv) > 2 else (640, 640)
lv = synth.levels(*hw)
x = synth.head2d(4, 80, lv, seed=0)
x = np.concatenate([x] * ((B + 3) // 4), 0)[:B]
f = [torch.from_numpy(v).cuda() for v in synth.split_levels(x, lv)]
for _ in range(3):
    y, _ = y3d.detect_inference(f, synth.STRIDES, 80)
    y3d.v10postprocess(y.permute(0, 2, 1), 300, 80)
    y3d.v10detect_export_forward(f, synth.STRIDES, 80, 300)
torch.cuda.synchronize()
