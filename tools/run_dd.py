#!/usr/bin/env python
"""Developer tool: a few calls of the fused 3D loss at cfg3 shape (for ncu launch lists)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yolov10_3d_b200 as y3d  # noqa: E402
from tests import synth  # noqa: E402

hw, nc, B, M = (384, 1280), 3, 32, 50
lv = synth.levels(*hw)
gts = synth.gt3d(B, M, nc, hw, seed=1)
x3 = synth.train_like_head3d(B, nc, lv, gts, seed=0, frac=0.03)
f3 = [torch.from_numpy(v).cuda() for v in synth.split_levels(x3, lv)]
cal = torch.from_numpy(np.tile(np.array(synth.KITTI_CALIB, np.float32), (B, 1))).cuda()
ms = torch.tensor(synth.KITTI_MEAN_SIZES, dtype=torch.float32).cuda()
g = torch.from_numpy(gts).cuda()
x3o = synth.train_like_head3d(B, nc, lv, gts, seed=5, frac=0.03)
f3o = [torch.from_numpy(v).cuda() for v in synth.split_levels(x3o, lv)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(4):  # both branches in the same launches (y3d_dd_loss_dual_fwd), L2 flushed in between
    flush.zero_()
    y3d.loss3d.dd_loss_dual_forward(f3, f3o, list(synth.STRIDES), nc, g, cal, ms, (8, 1), (1, 1, 1, 1, 1, 1))
torch.cuda.synchronize()
# timing (CUDA events, L2 flushed before every call), as bench.py's other_configs.cfg3
tot = 0.0
for _ in range(30):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    y3d.loss3d.dd_loss_dual_forward(f3, f3o, list(synth.STRIDES), nc, g, cal, ms, (8, 1), (1, 1, 1, 1, 1, 1))
    e1.record()
    torch.cuda.synchronize()
    tot += e0.elapsed_time(e1)
print(f"cfg3 dual 3D loss forward: {tot / 30 * 1e3:.1f} us per call")
