#!/usr/bin/env python
"""Per-stage measurement of every C-ABI entry at the BASELINE.json configs: CUDA-event time of
the public call, algorithmic bytes, and the fraction of the measured HBM peak.  Run on the GPU box:

    python tools/bench_stages.py > gpurun_out/stages.json      (also prints a markdown table on stderr)

Inputs are synthetic (tests/synth.py), resident on the device; every stage is timed over `iters` back-to-back calls after
warm-up.  Batches are sized so that each stage's working set exceeds the 126 MB L2 where the config allows it.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yolov10_3d_b200 as y3d  # noqa: E402
from tests import synth  # noqa: E402

PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) \
    if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
dev = torch.device("cuda", 0)


def timed(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def feats(x, lv):
    return [torch.from_numpy(f).to(dev) for f in synth.split_levels(x, lv)]


def tile_batch(x, B):
    """[b, ...] -> [B, ...] by repetition (cheap way to build a large synthetic batch from a few seeded images)."""
    reps = (B + x.shape[0] - 1) // x.shape[0]
    return np.concatenate([x] * reps, 0)[:B]


rows = []


def report(stage, cfg, ms, alg_bytes, images):
    gbs = alg_bytes / (ms * 1e-3) / 1e9
    rows.append(dict(stage=stage, config=cfg, ms=ms, images_per_s=images / (ms * 1e-3), algorithmic_bytes=alg_bytes,
                     achieved_gbs=gbs, frac_of_measured_hbm=gbs / PEAK))
    print(f"| {stage} | {cfg} | {ms:.3f} | {images / (ms * 1e-3):,.0f} | {alg_bytes / 1e6:.1f} | {gbs:,.0f} | {gbs / PEAK:.2f} |",
          file=sys.stderr)


print("| stage | config | ms/call | images/s | algorithmic MB | GB/s | frac of measured HBM peak |", file=sys.stderr)
print("|---|---|---|---|---|---|---|", file=sys.stderr)

# ---------------------------------------------------------------- cfg1 / cfg4: inference decode + NMS-free top-k
for name, hw, B, nc, D in (("cfg1 640x640 B=1", (640, 640), 1, 80, 300), ("cfg1-shape B=64", (640, 640), 64, 80, 300),
                           ("cfg4 1280x1280 B=32/GPU", (1280, 1280), 32, 80, 300)):
    lv = synth.levels(*hw)
    A = synth.num_anchors(lv)
    x = tile_batch(synth.head2d(min(B, 4), nc, lv, seed=0), B)
    f = feats(x, lv)
    y, _ = y3d.detect_inference(f, synth.STRIDES, nc)
    report("decode2d (a2-a4)", name, timed(lambda: y3d.detect_inference(f, synth.STRIDES, nc)), 912.0 * A * B, B)
    yp = y.permute(0, 2, 1)
    report("v10postprocess (a6)", name, timed(lambda: y3d.v10postprocess(yp, D, nc)), (4.0 * (4 + nc) * A + 28 * D) * B, B)
    report("decode+top-k fused (a2+a6)", name, timed(lambda: y3d.v10detect_export_forward(f, synth.STRIDES, nc, D)),
           (576.0 * A + 28 * D) * B, B)
    del f, y, yp, x

# ---------------------------------------------------------------- cfg2 / cfg5: assigner (API-faithful) and fused loss
for name, B, M, crowd in (("cfg2 B=64 M<=100", 64, 100, False), ("cfg5 B=128 M=500 crowd", 128, 500, True)):
    hw, nc = (640, 640), 80
    lv = synth.levels(*hw)
    A = synth.num_anchors(lv)
    nb = 8
    gt = tile_batch(synth.gt2d(nb, M, nc, hw, seed=1, crowd=crowd, full=crowd), B)
    xm = tile_batch(synth.train_like_head2d(nb, nc, lv, gt[:nb], seed=2, frac=0.02), B)
    fm = feats(xm, lv)
    gtd = torch.from_numpy(gt).to(dev)
    for k in (10, 1):
        pd_b = torch.empty((B, A, 4), device=dev)
        pd_s = torch.empty((B, A, nc), device=dev)
        lvl = y3d._util.Levels(fm, synth.STRIDES)
        y3d._lib.check(y3d.lib().y3d_train_decode(*lvl.args(), B, nc, 16, y3d._util.ptr(pd_b), y3d._util.ptr(pd_s),
                                                  y3d._util.stream_ptr(dev)))
        anc, st = synth.anchors_px(lv)
        pd_b = pd_b * torch.from_numpy(st).to(dev)[None, :, None]
        ancd = torch.from_numpy(anc).to(dev)
        asg = y3d.TaskAlignedAssigner(topk=k, num_classes=nc, alpha=0.5, beta=6.0, grid=(lv, synth.STRIDES))
        mask = (gtd[..., 1:5].sum(-1, keepdim=True) > 0).float()
        gl, gb = gtd[..., :1].contiguous(), gtd[..., 1:5].contiguous()
        ms = timed(lambda: asg(pd_s, pd_b, ancd, gl, gb, mask), iters=10)
        report(f"TaskAlignedAssigner top-k {k} (a10-a16)", name, ms, (4.0 * A * (nc + 6) + 24 * M + A * (8 * nc + 57) + 24 * M) * B, B)
        del pd_b, pd_s
    xo = tile_batch(synth.train_like_head2d(nb, nc, lv, gt[:nb], seed=3, frac=0.02), B)
    fo = feats(xo, lv)
    ms = timed(lambda: y3d.loss.v10_loss_forward(fm, fo, list(synth.STRIDES), nc, gtd, (7.5, 0.5, 1.5)))
    report("v10DetectLoss fwd fused (a18)", name, ms, (2 * 4.0 * (64 + nc) * A + 20 * M) * B, B)
    # forward + backward through the autograd node
    fmg = [f.requires_grad_(True) for f in fm]
    fog = [f.requires_grad_(True) for f in fo]

    def fb():
        items = y3d.loss._FusedLossFn.apply((list(synth.STRIDES), nc, gtd, (10, 1), (7.5, 0.5, 1.5)), *fmg, *fog)
        items.sum().backward()
        for f in fmg + fog:
            f.grad = None

    ms = timed(fb, iters=10)
    report("v10DetectLoss fwd+bwd (a18 + f1)", name, ms, (2 * 4.0 * (64 + nc) * A * 2 + 2 * 4.0 * nc * A + 20 * M) * B, B)
    del fm, fo, fmg, fog, xm, xo

# ---------------------------------------------------------------- cfg3: KITTI-shape 3D head
hw, nc, B, M, D = (384, 1280), 3, 32, 50, 50
lv = synth.levels(*hw)
A = synth.num_anchors(lv)
gts = synth.gt3d(B, M, nc, hw, seed=1)
x3 = synth.train_like_head3d(B, nc, lv, gts, seed=0, frac=0.03)
f3 = feats(x3, lv)
name = "cfg3 384x1280 B=32 nc=3"
y3, _ = y3d.detect3d_decode(f3, synth.STRIDES, nc)
report("decode3d (a7)", name, timed(lambda: y3d.detect3d_decode(f3, synth.STRIDES, nc)), 2 * 4.0 * (nc + 35) * A * B, B)
y3t = y3.transpose(-1, -2)
report("v10_3Dpostprocess (a8)", name, timed(lambda: y3d.v10_3Dpostprocess(y3t, D, nc)), (4.0 * (nc + 35) * A + 152 * D) * B, B)
dets = y3d.detect3d_postprocess(y3, D, nc)
calibs = np.tile(np.array(synth.KITTI_CALIB, np.float32), (B, 1))
inv = np.tile(np.array([[1, 0, 0], [0, 1, 0]], np.float64), (B, 1, 1))
ratio = np.ones((B, 2))
ms3 = np.array(synth.KITTI_MEAN_SIZES)
report("KITTI decode_preds (a9)", name, timed(lambda: y3d.kitti.decode_preds_tensor(dets, calibs, inv, ratio, ms3)),
       (4.0 * 37 + 8 * 14 + 1) * D * B, B)
gtsd, cald, msd = torch.from_numpy(gts).to(dev), torch.from_numpy(calibs).to(dev), torch.from_numpy(ms3.astype(np.float32)).to(dev)
pd_scores, pd_bboxes, pd_3d, anc, st = synth.assigner3d_inputs_from_head(x3, lv, nc)
tens = [torch.from_numpy(np.ascontiguousarray(v)).to(dev) for v in (pd_scores, pd_bboxes, pd_3d, anc, st)]
gt_tuple = tuple(t.contiguous() for t in gtsd.split((1, 4, 2, 2, 2, 3, 1, 1, 1), 2))
mg = (gt_tuple[1].sum(2, keepdim=True) > 0).float()
for k in (8, 1):
    asg3 = y3d.TaskAlignedAssigner3d(topk=k, num_classes=nc, alpha=0.5, beta=1.0, gamma=1.0, grid=(lv, synth.STRIDES))
    ms = timed(lambda: asg3(tens[0], tens[1], tens[2], tens[3], gt_tuple, mg, tens[4][:, None], cald, msd), iters=10)
    report(f"TaskAlignedAssigner3d top-k {k} (a19)", name, ms,
           (4.0 * A * (nc + 4 + 31 + 2 + 1) + 68 * M + 24 + A * (8 + 4 * nc + 48 + 1 + 8 + 96) + 96 * M) * B, B)
for k in (8, 1):
    ms = timed(lambda: y3d.loss3d.dd_loss_forward(f3, list(synth.STRIDES), nc, gtsd, cald, msd, k, (1, 1, 1, 1, 1, 1)), iters=10)
    report(f"DDDetectionLoss fwd top-k {k} (a20)", name, ms, (4.0 * (nc + 35) * A + 68 * M) * B, B)

print(json.dumps(dict(peak_gbs=PEAK, rows=rows), indent=1))
