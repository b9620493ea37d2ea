"""CPU, build container only (needs /root/reference): the host mirrors keep the reference's call surface.

* every mirrored callable's signature against the reference's (names, order, defaults; a mirror may only ADD optional
  trailing parameters) -- and the committed fixture tests/golden/api_signatures.json is refreshed / checked, so that the
  GPU box (no reference there) can hold the mirrors to it (tests/test_abi_cpu.py);
* the rebinds of INTEGRATION.md section 3 installed into the imported reference: its own ``init_criterion``
  (nn/tasks.py:646,651) then builds OUR loss objects from a real reference model;
* the M == 0 early-out of the assigners (tal.py:68-76) returns what the reference returns.
"""
import inspect
import json
import os
import types

import pytest
import torch

import yolov10_3d_b200 as y3d
from oracle import ref_import

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="/root/reference is not present on this machine")
FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "api_signatures.json")


def _ref():
    ref_import.import_reference()
    from ultralytics.nn import tasks
    from ultralytics.nn.modules import head
    from ultralytics.utils import loss, ops, tal
    return types.SimpleNamespace(tasks=tasks, head=head, loss=loss, ops=ops, tal=tal)


def _params(fn):
    out = []
    for p in inspect.signature(fn).parameters.values():
        if p.name == "self":
            continue
        out.append([p.name, None if p.default is inspect.Parameter.empty else repr(p.default)])
    return out


def pairs(r):
    return {
        "ops.v10postprocess": (r.ops.v10postprocess, y3d.v10postprocess),
        "ops.v10_3Dpostprocess": (r.ops.v10_3Dpostprocess, y3d.v10_3Dpostprocess),
        "ops.xywh2xyxy": (r.ops.xywh2xyxy, y3d.xywh2xyxy),
        "tal.make_anchors": (r.tal.make_anchors, y3d.make_anchors),
        "tal.TaskAlignedAssigner.__init__": (r.tal.TaskAlignedAssigner.__init__, y3d.TaskAlignedAssigner.__init__),
        "tal.TaskAlignedAssigner.forward": (r.tal.TaskAlignedAssigner.forward, y3d.TaskAlignedAssigner.forward),
        "tal.TaskAlignedAssigner3d.__init__": (r.tal.TaskAlignedAssigner3d.__init__, y3d.TaskAlignedAssigner3d.__init__),
        "tal.TaskAlignedAssigner3d.forward": (r.tal.TaskAlignedAssigner3d.forward, y3d.TaskAlignedAssigner3d.forward),
        "loss.v8DetectionLoss.__init__": (r.loss.v8DetectionLoss.__init__, y3d.v8DetectionLoss.__init__),
        "loss.v8DetectionLoss.__call__": (r.loss.v8DetectionLoss.__call__, y3d.v8DetectionLoss.__call__),
        "loss.v10DetectLoss.__init__": (r.loss.v10DetectLoss.__init__, y3d.v10DetectLoss.__init__),
        "loss.v10DetectLoss.__call__": (r.loss.v10DetectLoss.__call__, y3d.v10DetectLoss.__call__),
        "loss.DDDetectionLoss.__init__": (r.loss.DDDetectionLoss.__init__, y3d.DDDetectionLoss.__init__),
        "loss.DDDetectionLoss.__call__": (r.loss.DDDetectionLoss.__call__, y3d.DDDetectionLoss.__call__),
        "loss.DetectLoss3d.__init__": (r.loss.DetectLoss3d.__init__, y3d.DetectLoss3d.__init__),
        "loss.DetectLoss3d.__call__": (r.loss.DetectLoss3d.__call__, y3d.DetectLoss3d.__call__),
        "head.v10Detect3d.inference_forward_feat": (r.head.v10Detect3d.inference_forward_feat, y3d.head.inference_forward_feat),
        "head.v10Detect3d.select_candidates": (r.head.v10Detect3d.select_candidates, None),
        "head.v10Detect3d.extract_patches": (r.head.v10Detect3d.extract_patches, None),
    }


def test_mirror_signatures_match_the_reference_and_fixture_is_current():
    r = _ref()
    fixture = {}
    for name, (ref_fn, mine) in pairs(r).items():
        want = _params(ref_fn)
        fixture[name] = want
        if mine is None:
            continue
        got = _params(mine)
        if name.endswith("inference_forward_feat"):  # a function taking the head module as its first argument
            assert got[0][0] in ("det", "head") and got[1:] == want, (name, got, want)
            continue
        assert got[:len(want)] == want, f"{name}: mirror {got} vs reference {want}"
        assert all(d is not None for _, d in got[len(want):]), f"{name}: extra mirror parameters must be optional"
    if os.environ.get("Y3D_WRITE_FIXTURES"):
        with open(FIXTURE, "w") as f:
            json.dump(fixture, f, indent=1, sort_keys=True)
    assert os.path.exists(FIXTURE), "run once with Y3D_WRITE_FIXTURES=1 to create tests/golden/api_signatures.json"
    assert json.load(open(FIXTURE)) == fixture, "the reference's signatures changed: refresh tests/golden/api_signatures.json"


def test_init_criterion_of_the_real_reference_builds_the_mirrors():
    r = _ref()
    saved = {}

    def rebind(mod, name, obj):
        saved[(mod, name)] = getattr(mod, name)
        setattr(mod, name, obj)

    try:  # INTEGRATION.md section 3 (nn/tasks.py binds the loss classes by name at import: rebind there too)
        for mod in (r.loss, r.tasks):
            rebind(mod, "v10DetectLoss", y3d.v10DetectLoss)
            rebind(mod, "DetectLoss3d", y3d.DetectLoss3d)
        rebind(r.loss, "v8DetectionLoss", y3d.v8DetectionLoss)
        rebind(r.loss, "DDDetectionLoss", y3d.DDDetectionLoss)
        rebind(r.tal, "TaskAlignedAssigner", y3d.TaskAlignedAssigner)
        rebind(r.tal, "TaskAlignedAssigner3d", y3d.TaskAlignedAssigner3d)
        rebind(r.ops, "v10postprocess", y3d.v10postprocess)
        rebind(r.ops, "v10_3Dpostprocess", y3d.v10_3Dpostprocess)
        model = r.tasks.YOLOv10DetectionModel("yolov10n.yaml", verbose=False)
        model.args = types.SimpleNamespace(box=7.5, cls=0.5, dfl=1.5)
        crit = model.init_criterion()  # nn/tasks.py:646
        assert isinstance(crit, y3d.v10DetectLoss)
        assert crit.one2many.topk == 10 and crit.one2one.topk == 1 and crit.one2many.nc == 80
        assert [float(s) for s in crit.one2many.stride] == [8.0, 16.0, 32.0]
        model3 = r.tasks.YOLOv10_3DDetectionModel("yolov10m_3D.yaml", verbose=False)
        model3.args = types.SimpleNamespace(loss2d=1.0, cls=1.0, depth=1.0, offset3d=1.0, size3d=1.0, heading=1.0,
                                            tal_topk=8, tal_alpha=0.5, tal_beta=1.0, tal_gamma=1.0, tal_2d=True,
                                            tal_3d=True, kps_dist_metric="l1", constrain_anchors=True,
                                            distillation=False, fgdm_loss=False, fgdm_supervision=False)
        crit3 = model3.init_criterion()  # nn/tasks.py:651
        assert isinstance(crit3, y3d.DetectLoss3d) and crit3.one2many.topk == 8 and crit3.one2one.topk == 1
    finally:
        for (mod, name), obj in saved.items():
            setattr(mod, name, obj)


def test_no_targets_early_out_equals_the_reference():
    r = _ref()
    B, A, nc = 2, 50, 4
    ps, pb = torch.rand(B, A, nc), torch.rand(B, A, 4)
    anc = torch.rand(A, 2)
    gl, gb, mg = torch.zeros(B, 0, 1), torch.zeros(B, 0, 4), torch.zeros(B, 0, 1)
    want = r.tal.TaskAlignedAssigner(topk=10, num_classes=nc)(ps, pb, anc, gl, gb, mg)
    got = y3d.TaskAlignedAssigner(topk=10, num_classes=nc)(ps, pb, anc, gl, gb, mg)
    assert len(want) == len(got) == 5
    for w, g in zip(want, got):
        assert w.dtype == g.dtype and torch.equal(w, g)
