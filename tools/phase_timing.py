#!/usr/bin/env python
"""Developer tool: per-CTA phase timestamps of loss_finish_kernel (library built with -DY3D_TIMING).

    python tools/phase_timing.py build     # here (nvcc)   -> tools/liby3d_timing.so
    python tools/phase_timing.py run       # on the GPU box
"""
import ctypes
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tools", "liby3d_timing.so")

if sys.argv[1] == "build":
    src = sorted(glob.glob(os.path.join(ROOT, "yolov10-3d_b200", "csrc", "*.cu")))
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a",
                           "-lineinfo", "-shared", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-DY3D_TIMING",
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

if os.environ.get("CROWD"):
    Bc, Mc = 128, 500
    lv = synth.levels(640, 640)
    gt = np.concatenate([synth.gt2d(8, Mc, 80, (640, 640), seed=1, crowd=True, full=True)] * (Bc // 8))
    xm = np.concatenate([synth.train_like_head2d(8, 80, lv, gt[:8], seed=2, frac=0.02)] * (Bc // 8))
    xo = np.concatenate([synth.train_like_head2d(8, 80, lv, gt[:8], seed=3, frac=0.02)] * (Bc // 8))
else:
    lv, gt, xm, xo = bench.make_inputs(seed=0)

dev = torch.device("cuda", 0)
fm = [torch.from_numpy(f).to(dev) for f in synth.split_levels(xm, lv)]
fo = [torch.from_numpy(f).to(dev) for f in synth.split_levels(xo, lv)]
gtd = torch.from_numpy(gt).to(dev)
for _ in range(5):
    y3d.loss.v10_loss_forward(fm, fo, list(synth.STRIDES), 80, gtd, (7.5, 0.5, 1.5))
torch.cuda.synchronize()
h = _lib.lib()
prof = (ctypes.c_ulonglong * 16)()
h.y3d_debug_read_topk_prof.argtypes = [ctypes.c_void_p, ctypes.c_int]
h.y3d_debug_read_topk_prof(prof, 1)
y3d.loss.v10_loss_forward(fm, fo, list(synth.STRIDES), 80, gtd, (7.5, 0.5, 1.5))
torch.cuda.synchronize()
h.y3d_debug_read_topk_prof(prof, 0)
pv = np.array(prof, dtype=np.float64)
names = ["fetch+valid", "gt load+rects", "phase0", "stage1 trips", "stage2 pops", "list updates", "claims"]
tot = pv[:7].sum()
print("top-k kernel, warp-cycles by phase (one step):")
for i, nme in enumerate(names):
    print(f"  {nme:14s} {pv[i] / 1e6:8.2f} Mcyc  {100 * pv[i] / tot:5.1f}%")
gts_n = max(pv[8], 1)
print(f"  valid GT-warps {int(pv[8])}, cells/GT {pv[9] / gts_n:.1f}, trips/GT {pv[10] / gts_n:.2f}, pops/GT {pv[11] / gts_n:.2f}, "
      f"popped cand/GT {pv[12] / gts_n:.1f}, insertions/GT {pv[13] / gts_n:.1f}, cycles/GT {tot / gts_n:.0f}")
n = 128 * 8
buf = (ctypes.c_ulonglong * n)()
h.y3d_debug_read_stamps.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert h.y3d_debug_read_stamps(buf, n) == 0
t = np.array(buf, dtype=np.uint64).reshape(128, 8).astype(np.int64)
t0 = t[:, 0].min()
valid = (gt[..., 1:5].sum(-1) > 0).sum(1)
print("cta(b,z) nGT  start  gts  resolve  fg  reduce  (us, relative to first CTA start; phases = durations)")
d = np.diff(t[:, :5], axis=1) / 1e3
for i in list(range(0, 6)) + list(range(64, 70)):
    print(i % 64, i // 64, valid[i % 64], round((t[i, 0] - t0) / 1e3, 1), d[i].round(1))
print("mean phase durations (us):", d.mean(0).round(2), " max:", d.max(0).round(2))
print("kernel span (us): first start -> last CTA end", (t[:, :5].max() - t0) / 1e3, " final-reduce stamp", (t[:, 5].max() - t0) / 1e3)

# ---- select kernel (fused decode + top-k), CTA 0
f64 = [f[:64] if f.shape[0] >= 64 else f for f in fo]
for _ in range(3):
    y3d.v10detect_export_forward(fo, synth.STRIDES, 80, 300)
torch.cuda.synchronize()
st = (ctypes.c_longlong * 16)()
h.y3d_debug_read_sel_stamps.argtypes = [ctypes.c_void_p]
h.y3d_debug_read_sel_stamps(st)
sa = np.array(st, dtype=np.int64)
print("  (last block_topk call) passes", sa[9] - sa[8], " collect", sa[10] - sa[9], " sort", sa[11] - sa[10], " L", sa[12], " mask", hex(int(sa[13]) & 0xffffffff))
sv = sa[:5]
print("select kernel CTA 0 phases (cycles): stage1 top-D", sv[1] - sv[0], " gather+compact", sv[2] - sv[1], " stage2 sort", sv[3] - sv[2],
      " outputs", sv[4] - sv[3])
