#!/usr/bin/env python
"""Developer tool: is the bench loop bound by the host (Python + launch) or by the GPU?"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import yolov10_3d_b200 as y3d
import bench
from tests import synth

lv, gt, xm, xo = bench.make_inputs(seed=0)
dev = torch.device("cuda", 0)
fm = [torch.from_numpy(f).to(dev) for f in synth.split_levels(xm, lv)]
fo = [torch.from_numpy(f).to(dev) for f in synth.split_levels(xo, lv)]
gtd = torch.from_numpy(gt).to(dev)
st = list(synth.STRIDES)
for _ in range(20):
    y3d.dist.v10_loss_sharded(fm, fo, st, 80, gtd, (7.5, 0.5, 1.5), 64)
torch.cuda.synchronize()
for K in (50, 200, 1000):
    t0 = time.perf_counter()
    for _ in range(K):
        y3d.dist.v10_loss_sharded(fm, fo, st, 80, gtd, (7.5, 0.5, 1.5), 64)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"K={K}: host enqueue {1e6 * (t1 - t0) / K:.1f} us/step, total {1e6 * (t2 - t0) / K:.1f} us/step")
