#!/usr/bin/env python
"""Multi-GPU check of the image-sharded detection path (y3d_decode_topk2d_sharded: the gather of the [B/N, 300, 6]
detections fused into the box-decode kernel's epilogue over NVLink peer memory) against the single-process result and the
NCCL route, plus timings.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29534 \
        tools/check_peer_gather.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yolov10_3d_b200 as y3d  # noqa: E402
from tests import synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok = True
# ---- correctness on a small shape: the global batch is generated identically on every rank, rank r takes its shard
nc, hw, Bl, D = 8, (320, 320), 3, 50
lv = synth.levels(*hw)
x_all = synth.head2d(Bl * world, nc, lv, seed=7)
lo, hi = y3d.dist.shard_range(Bl * world, rank, world)
feats = [torch.from_numpy(f).to(dev) for f in synth.split_levels(x_all[lo:hi], lv)]
feats_all = [torch.from_numpy(f).to(dev) for f in synth.split_levels(x_all, lv)]
single = y3d.v10detect_export_forward(feats_all, synth.STRIDES, nc, D)  # what one process computes on the whole batch
g = y3d.dist.PeerDetectionGather(dev, Bl, D)
assert g.available, "symmetric memory unavailable"
for it in range(6):
    got = y3d.dist.detect_sharded(feats, synth.STRIDES, nc, D, gatherer=g)
    if not torch.equal(got, single):
        ok = False
        print("fused gather != single-process result", rank, it, (got != single).sum().item())
        break
    nccl = y3d.dist.detect_sharded(feats, synth.STRIDES, nc, D, gatherer=None)
    if not torch.equal(nccl, single):
        ok = False
        print("nccl gather != single-process result", rank, it)
        break
assert int(g.status.item()) == 0
# ---- cfg4 shape timing: 32 images per GPU at 1280 x 1280
nc, hw, Bl, D = 80, (1280, 1280), 32, 300
lv = synth.levels(*hw)
x = np.concatenate([synth.head2d(4, nc, lv, seed=100 + rank)] * (Bl // 4), 0)
feats = [torch.from_numpy(f).to(dev) for f in synth.split_levels(x, lv)]
g4 = y3d.dist.PeerDetectionGather(dev, Bl, D)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, fn in (("fused epilogue gather", lambda: y3d.dist.detect_sharded(feats, synth.STRIDES, nc, D, gatherer=g4)),
                 ("nccl all_gather_into_tensor", lambda: y3d.dist.detect_sharded(feats, synth.STRIDES, nc, D, gatherer=None)),
                 ("no gather (local only)", lambda: y3d.v10detect_export_forward(feats, synth.STRIDES, nc, D))):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0.record()
    for _ in range(50):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 50], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"cfg4 {name}: {float(t) * 1e3:.1f} us per step, {Bl * world / (float(t) * 1e-3):,.0f} images/s on {world} GPUs")
a = y3d.dist.detect_sharded(feats, synth.STRIDES, nc, D, gatherer=g4).clone()
b = y3d.dist.detect_sharded(feats, synth.STRIDES, nc, D, gatherer=None)
if not torch.equal(a, b):
    ok = False
    print("cfg4: fused gather != nccl gather", rank)
okt = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(okt, op=dist.ReduceOp.MIN)
if rank == 0:
    print("peer gather == single process == nccl gather:", bool(okt[0] > 0.5))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if okt[0] > 0.5 else 1)
