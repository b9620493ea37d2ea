#!/usr/bin/env python
"""Developer tool: forward and backward of the fused v10DetectLoss at cfg2 (or CROWD=1: cfg5), timed separately with
CUDA events through the host mirror (yolov10-3d_b200/loss.py), for ncu and for the DESIGN.md results table."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yolov10_3d_b200 as y3d  # noqa: E402
from tests import synth  # noqa: E402
from yolov10_3d_b200._util import Levels  # noqa: E402

crowd = os.environ.get("CROWD") == "1"
B, M = (128, 500) if crowd else (64, 100)
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
nc, hw = 80, (640, 640)
lv = synth.levels(*hw)
dev = torch.device("cuda", 0)
nb = 8
gt = synth.gt2d(nb, M, nc, hw, seed=1, crowd=crowd, full=crowd)
rep = lambda x: np.concatenate([x] * (B // nb), 0)
fm = [torch.from_numpy(f).to(dev) for f in synth.split_levels(rep(synth.train_like_head2d(nb, nc, lv, gt, seed=2, frac=0.02)), lv)]
fo = [torch.from_numpy(f).to(dev) for f in synth.split_levels(rep(synth.train_like_head2d(nb, nc, lv, gt, seed=3, frac=0.02)), lv)]
gtd = torch.from_numpy(rep(gt)).to(dev)
levels = [Levels(fm, synth.STRIDES), Levels(fo, synth.STRIDES)]
gains = (7.5, 0.5, 1.5)
gi = torch.ones(6, device=dev)


def ev():
    return torch.cuda.Event(enable_timing=True)


fwd = y3d.loss._branch_forward(levels, nc, gtd, (10, 1), gains, True, False, None)
for _ in range(3):
    y3d.loss._branch_backward(levels, nc, fwd, (10, 1), gains, fwd["items"], gi)
torch.cuda.synchronize()
evs = []
for _ in range(iters):  # queue everything, synchronise once: the event deltas then hold no host launch time
    e = [ev() for _ in range(3)]
    e[0].record()
    fwd = y3d.loss._branch_forward(levels, nc, gtd, (10, 1), gains, True, False, None)
    e[1].record()
    g = y3d.loss._branch_backward(levels, nc, fwd, (10, 1), gains, fwd["items"], gi)
    e[2].record()
    evs.append(e)
    del g
torch.cuda.synchronize()
tf = sum(e[0].elapsed_time(e[1]) for e in evs[iters // 2:]) * iters / (iters - iters // 2)
tb = sum(e[1].elapsed_time(e[2]) for e in evs[iters // 2:]) * iters / (iters - iters // 2)
A = synth.num_anchors(lv)
bwd_bytes = 2 * 4.0 * (nc + 64 + nc) * A * B
print(f"fwd {tf / iters:.4f} ms  bwd {tb / iters:.4f} ms  bwd algorithmic {bwd_bytes / 1e6:.1f} MB -> "
      f"{bwd_bytes / (tb / iters * 1e-3) / 1e9:.0f} GB/s")
