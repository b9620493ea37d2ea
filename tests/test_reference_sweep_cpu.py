"""CPU, build container only (needs /root/reference): seeded sweep of the REAL reference's assignment inside v10DetectLoss
against the oracle at the BASELINE shapes -- the evidence that the oracle is faithful at size, in the repo.  Skipped where
the reference is absent (the GPU box); there tests/golden/lossasg_*.npz carry the reference's answers."""
import types

import numpy as np
import pytest

from oracle import oracle, ref_import
from tests import synth

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="/root/reference is not present on this machine")

SEEDS = 50


def _reference_assignments(nc, hw, lv, gt, xm, xo):
    import torch

    from tests.golden import make_golden as mg  # imports the reference, patches nothing by itself

    batch = {k: mg.t(v) for k, v in synth.batch_dict(gt, hw).items()}
    fm = [mg.t(f) for f in synth.split_levels(xm, lv)]
    fo = [mg.t(f) for f in synth.split_levels(xo, lv)]
    model = mg.FakeModel(nc, synth.STRIDES, types.SimpleNamespace(box=7.5, cls=0.5, dfl=1.5))
    crit = mg.ref_loss.v10DetectLoss(model)
    captured = {}
    for branch, asg in ((0, crit.one2many.assigner), (1, crit.one2one.assigner)):
        def wrap(inner, branch=branch):
            def fwd(*a, **k):
                out = inner(*a, **k)
                captured[branch] = (out[3].numpy().copy(), out[4].numpy().copy())
                return out
            return fwd
        asg.forward = wrap(asg.forward)
    with mg.patched_topk(), torch.no_grad():
        _, items = crit({"one2many": fm, "one2one": fo}, batch)
    return captured, items.numpy().astype(np.float64)


def test_reference_vs_oracle_assignment_seed_sweep():
    nc, hw = 80, (640, 640)
    lv = synth.levels(*hw)
    n_fg = n_mism = 0
    worst = 0.0
    for s in range(SEEDS):
        crowd = s % 10 == 9  # every tenth seed is a dense-crowd image (500 GT)
        M = 500 if crowd else 100
        gt = synth.gt2d(1, M, nc, hw, seed=3000 + s, crowd=crowd, full=True)
        xm = synth.train_like_head2d(1, nc, lv, gt, seed=4000 + s, frac=0.02)
        xo = synth.train_like_head2d(1, nc, lv, gt, seed=5000 + s, frac=0.02)
        cap, ref_items = _reference_assignments(nc, hw, lv, gt, xm, xo)
        items = []
        for branch, (x, k) in enumerate(((xm, 10), (xo, 1))):
            it, _, _, fg, tgi = oracle.v8_loss(x, lv, synth.STRIDES, nc, gt, k, debug=True)
            rfg, rtgi = cap[branch]
            both = rfg & fg
            n_mism += int((fg != rfg).sum() + (tgi[both] != rtgi[both]).sum())
            n_fg += int(rfg.sum())
            items.append(it)
        worst = max(worst, float(np.max(np.abs(np.concatenate(items) / ref_items - 1.0))))
    print(f"reference vs oracle: {SEEDS} seeds, {n_fg} foreground anchors, {n_mism} mismatches, worst item deviation {worst:.2e}")
    assert n_fg > 20000 and n_mism == 0 and worst < 2e-5
