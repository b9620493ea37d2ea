#!/usr/bin/env python
"""Developer tool: when does every persistent warp of tal_topk_kernel start and leave (library built with
-DY3D_TAILTIME)?  Shows how much of the kernel is the tail of the dynamic work distribution.

    python tools/tail_timing.py build     # here (nvcc)   -> tools/liby3d_tail.so
    python tools/tail_timing.py run       # on the GPU box
"""
import ctypes
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tools", "liby3d_tail.so")

if sys.argv[1] == "build":
    src = sorted(glob.glob(os.path.join(ROOT, "yolov10-3d_b200", "csrc", "*.cu")))
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a",
                           "-lineinfo", "-shared", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-DY3D_TAILTIME",
                           "-o", OUT, *src])
    print(OUT)
    sys.exit(0)

import numpy as np
import torch

import yolov10_3d_b200 as y3d
from yolov10_3d_b200 import _lib

_lib.LIB_PATH = OUT
import bench
from tests import synth

lv, gt, xm, xo = bench.make_inputs(seed=0)
dev = torch.device("cuda", 0)
fm = [torch.from_numpy(f).to(dev) for f in synth.split_levels(xm, lv)]
fo = [torch.from_numpy(f).to(dev) for f in synth.split_levels(xo, lv)]
gtd = torch.from_numpy(gt).to(dev)
for _ in range(6):
    y3d.loss.v10_loss_forward(fm, fo, list(synth.STRIDES), 80, gtd, (7.5, 0.5, 1.5))
torch.cuda.synchronize()
h = _lib.lib()
n = 888 * 4
buf = (ctypes.c_ulonglong * (n * 4))()
h.y3d_debug_read_topk_tail.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert h.y3d_debug_read_topk_tail(buf, n * 4) == 0
t = np.array(buf, dtype=np.uint64).reshape(n, 4).astype(np.int64)
t0 = t[:, 0].min()
start, end = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3
print(f"warps {n}: start min/mean/max {start.min():.1f}/{start.mean():.1f}/{start.max():.1f} us; "
      f"exit min/p10/p50/p90/max {end.min():.1f}/{np.percentile(end, 10):.1f}/{np.percentile(end, 50):.1f}/"
      f"{np.percentile(end, 90):.1f}/{end.max():.1f} us")
print(f"mean busy {(end - start).mean():.1f} us of span {end.max():.1f} us -> idle fraction {1 - (end - start).mean() / end.max():.2f}")
print(f"items per warp min/mean/max {t[:, 2].min()}/{t[:, 2].mean():.2f}/{t[:, 2].max()}, valid per warp "
      f"{t[:, 3].min()}/{t[:, 3].mean():.2f}/{t[:, 3].max()}")
late = np.argsort(end)[-8:]
print("latest warps (exit us, items, valid):", [(round(float(end[i]), 1), int(t[i, 2]), int(t[i, 3])) for i in late])
