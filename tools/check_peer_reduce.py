#!/usr/bin/env python
"""Multi-GPU check of the fused peer-memory all-reduce + finalize kernel (csrc/xrank.cu) against the NCCL path.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        tools/check_peer_reduce.py
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yolov10_3d_b200 as y3d  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
red = y3d.dist.PeerLossReducer(dev)
assert red.available, "symmetric memory unavailable"
gains = (7.5, 0.5, 1.5)
g = torch.Generator(device="cpu").manual_seed(1234 + rank)
ok = True
for it in range(200):
    parts = (torch.rand(8, generator=g, dtype=torch.float64) * (1 + it)).to(dev)
    if it % 7 == 0:
        parts[3] = 0.01  # target_scores_sum below 1 on this rank
    items = red(parts, gains)
    ref = parts.clone()
    dist.all_reduce(ref)
    want = y3d.loss.finalize_partials(ref, gains)
    if not torch.allclose(items, want, rtol=1e-6, atol=0):
        ok = False
        print(rank, it, items, want)
        break
assert int(red.status.item()) == 0
# timing: back-to-back calls
torch.cuda.synchronize()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
parts = torch.rand(8, dtype=torch.float64).to(dev)
for name, fn in (("peer kernel", lambda: red(parts, gains)),
                 ("nccl + finalize", lambda: y3d.loss.finalize_partials(dist.all_reduce(parts.clone()) or parts, gains))):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0.record()
    for _ in range(200):
        fn()
    e1.record()
    torch.cuda.synchronize()
    if rank == 0:
        print(f"{name}: {e0.elapsed_time(e1) / 200 * 1e3:.1f} us per call")
if rank == 0:
    print("peer reduce == nccl reduce:", ok)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
