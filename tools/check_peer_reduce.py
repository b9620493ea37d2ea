#!/usr/bin/env python
"""Multi-GPU check of the peer-memory all-reduce + finalize kernel (csrc/xrank.cu) and of the same exchange fused into
the loss' last kernel (y3d_v10_loss_fwd_sharded) against the NCCL path.

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
# the exchange fused into the loss' last kernel (y3d_v10_loss_fwd_sharded) against the NCCL route, on a small shard
import numpy as np  # noqa: E402

from tests import synth  # noqa: E402

nc, hw, Bl, M = 8, (320, 320), 4, 20
lv = synth.levels(*hw)
gt = synth.gt2d(Bl, M, nc, hw, seed=100 + rank)
if rank == world - 1:
    gt[1:] = 0  # a rank with almost no targets
fm = [torch.from_numpy(f).to(dev) for f in synth.split_levels(synth.train_like_head2d(Bl, nc, lv, gt, seed=200 + rank, frac=0.05), lv)]
fo = [torch.from_numpy(f).to(dev) for f in synth.split_levels(synth.train_like_head2d(Bl, nc, lv, gt, seed=300 + rank, frac=0.05), lv)]
gtd = torch.from_numpy(gt).to(dev)
for it in range(20):
    tot_f, it_f = y3d.dist.v10_loss_sharded(fm, fo, list(synth.STRIDES), nc, gtd, gains, Bl * world, reducer=red)
    tot_n, it_n = y3d.dist.v10_loss_sharded(fm, fo, list(synth.STRIDES), nc, gtd, gains, Bl * world, reducer=None)
    if not (torch.allclose(it_f, it_n, rtol=1e-6, atol=0) and torch.isfinite(it_f).all()):
        ok = False
        print("fused sharded loss != nccl route", rank, it, it_f, it_n)
        break
    chk = it_f.clone()
    dist.all_reduce(chk, op=dist.ReduceOp.MAX)
    if not torch.equal(chk, it_f):  # identical on every rank, bit for bit
        ok = False
        print("ranks disagree", rank, it_f, chk)
        break
assert int(red.status.item()) == 0
# training through the sharded call: gradients of the global-batch loss w.r.t. this rank's head tensors == the rows of
# this rank in the gradients one process computes on the whole batch (every rank builds the whole batch for the check)
gts_all = [synth.gt2d(Bl, M, nc, hw, seed=100 + r) for r in range(world)]
gts_all[world - 1][1:] = 0
xm_all = np.concatenate([synth.train_like_head2d(Bl, nc, lv, gts_all[r], seed=200 + r, frac=0.05) for r in range(world)])
xo_all = np.concatenate([synth.train_like_head2d(Bl, nc, lv, gts_all[r], seed=300 + r, frac=0.05) for r in range(world)])
gt_all = torch.from_numpy(np.concatenate(gts_all)).to(dev)
fm_all = [torch.from_numpy(f).to(dev).requires_grad_() for f in synth.split_levels(xm_all, lv)]
fo_all = [torch.from_numpy(f).to(dev).requires_grad_() for f in synth.split_levels(xo_all, lv)]
loss_1 = y3d.loss._FusedLossFn.apply((list(map(float, synth.STRIDES)), nc, gt_all, (10, 1), gains), *fm_all, *fo_all)
(loss_1.sum() * (Bl * world)).backward()
fm_g = [f.clone().requires_grad_() for f in fm]
fo_g = [f.clone().requires_grad_() for f in fo]
tot_s, it_s = y3d.dist.v10_loss_sharded(fm_g, fo_g, list(synth.STRIDES), nc, gtd, gains, Bl * world, reducer=red)
tot_s.backward()
lo = rank * Bl
for name, loc, full in (("one2many", fm_g, fm_all), ("one2one", fo_g, fo_all)):
    for a, b_ in zip(loc, full):
        want = b_.grad[lo:lo + Bl]
        if not torch.allclose(a.grad, want, rtol=1e-5, atol=1e-7):
            ok = False
            print("sharded gradient != single-process gradient", rank, name, (a.grad - want).abs().max().item())
if not torch.allclose(it_s, loss_1.detach(), rtol=1e-6, atol=0):
    ok = False
    print("sharded items (autograd route) != single-process items", rank, it_s, loss_1)
assert int(red.status.item()) == 0
# the deferred variant (post in the loss' last kernel, collect on a side stream, one call of slack): same items, and the
# two-parity flow control holds over a train of back-to-back calls with a rank that is made slow every few steps
handles = []
for it in range(40):
    if it % 5 == rank % 5:
        torch.cuda._sleep(3_000_000)  # ~1.5 ms of skew on this rank
    handles.append(y3d.dist.v10_loss_sharded(fm, fo, list(synth.STRIDES), nc, gtd, gains, Bl * world, reducer=red, defer=True))
    if it >= 2:
        tot_d, it_d = handles[it - 2].wait()
        if not torch.equal(it_d, it_f):
            ok = False
            print("deferred sharded loss != fused", rank, it, it_d, it_f)
            break
tot_d, it_d = handles[-1].wait()
torch.cuda.synchronize()
ok = ok and bool(torch.equal(it_d, it_f)) and abs(float(tot_d) - float(it_f.sum()) * Bl * world) <= 1e-5 * abs(float(tot_d))
assert int(red.status.item()) == 0
if rank == 0:
    print("fused sharded loss == nccl route == deferred:", ok, it_f.tolist())
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
