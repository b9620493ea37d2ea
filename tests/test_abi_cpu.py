"""CPU: the C-ABI library loads and exports every symbol include/y3d.h declares; host-side glue behaves like the
reference's (GT packing, M == 0 early-out, loud failure instead of any CPU fallback).  No compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import yolov10_3d_b200 as y3d
from oracle import oracle
from tests import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "y3d.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(y3d_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = declared_symbols()
    assert len(names) >= 14
    handle = ctypes.CDLL(y3d._lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/y3d.h but not exported"
    assert set(names) == set(y3d._lib.SIGNATURES), "ctypes binding out of sync with include/y3d.h"
    assert y3d.lib().y3d_abi_version() == 2
    assert b"workspace" in y3d.lib().y3d_strerror(-4)
    assert y3d._lib.workspace_bytes(y3d._lib.STAGE_TAL_ASSIGN, B=2, A=100, nc=4, M=3, k=10) > 0


def test_argument_errors_come_back_as_codes_not_crashes():
    lib = y3d.lib()
    assert lib.y3d_postprocess(None, 0, 0, 0, 1, 10, 2, 4, 0, 5, None, None, None, None, None, 0, None) == -1
    assert lib.y3d_v8_loss_finalize(None, 1, 1.0, 1.0, 1.0, None, None) == -1
    assert lib.y3d_decode_preds3d(None, 1, 1, 3, None, None, None, None, 0.1, None, None, None) == -1


def test_sharded_loss_entry_and_peer_exchange_host_side():
    """The multi-GPU entry validates its arguments before touching the device; the exchange buffer is 4 slots of 32
    flag-in-data words per rank; without a process group / peer memory the reducer reports itself unavailable (the
    callers then take the NCCL route) instead of raising."""
    lib = y3d.lib()
    assert lib.y3d_xrank_buffer_bytes(8) == 4 * 8 * 32 * 8 and lib.y3d_xrank_buffer_bytes(0) == 4 * 32 * 8
    null = [None] * 8
    assert lib.y3d_v10_loss_fwd_sharded(*null, 3, 2, 8, 16, None, 0, 10, 1, 7.5, 0.5, 1.5, None, None, 2.0, None, 0, 2,
                                        None, ctypes.c_uint64(1), 0, None, None, None, 0, None) == -1
    assert lib.y3d_loss_exchange_resolve(2, 0, 2, None, ctypes.c_uint64(1), 7.5, 0.5, 1.5, 2.0, None, None, None, None,
                                         None) == -1
    assert lib.y3d_decode_topk2d_sharded(None, None, None, None, None, 3, 2, 8, 16, 0, 50, 0, 2, None, ctypes.c_uint64(1),
                                         None, None, 0, None) == -1
    assert lib.y3d_gather_buffer_bytes(2, 4, 300) >= 2 * 2 * 4 * 300 * 6 * 4 + 2 * 2 * 4 * 4
    assert lib.y3d_loss_allreduce_finalize(None, 2, 0, 2, None, ctypes.c_uint64(1), 7.5, 0.5, 1.5, None, None, None,
                                           None) == -1
    red = y3d.dist.PeerLossReducer(torch.device("cpu"))
    assert red.available is False


def test_sharded_detection_entry_host_side():
    """y3d_gather_buffer_bytes / y3d_decode_topk2d_sharded without a GPU: the buffer holds the two result parities the
    mirror's view() addresses plus as many 64-bit staging words, and argument errors come back as codes."""
    lib = y3d.lib()
    world, n_local, D = 4, 8, 300
    n = world * n_local * D * 6
    need = int(lib.y3d_gather_buffer_bytes(world, n_local, D))
    assert need >= 2 * n * 4 + 2 * n * 8 and need % 256 == 0  # [2][world*B][D][6] floats, then [2][world*B][6 D] u64
    assert int(lib.y3d_gather_buffer_bytes(0, n_local, D)) == 0 and int(lib.y3d_gather_buffer_bytes(world, 0, D)) == 0
    bufs = (ctypes.c_void_p * world)(*[0x10000 * (r + 1) for r in range(world)])
    args = [None] * 5 + [3, 2, 8, 16, 0, 50]
    assert lib.y3d_decode_topk2d_sharded(*args, 0, world, None, ctypes.c_uint64(1), None, None, 0, None) == -1  # no buffers
    assert lib.y3d_decode_topk2d_sharded(*args, world, world, bufs, ctypes.c_uint64(1), None, None, 0, None) == -1  # rank
    assert lib.y3d_decode_topk2d_sharded(*args, 0, world, bufs, ctypes.c_uint64(0), None, None, 0, None) == -1  # seq 0
    assert lib.y3d_decode_topk2d_sharded(*args, 0, 17, bufs, ctypes.c_uint64(1), None, None, 0, None) == -1  # world > 16
    # the sequence number travels in the low 32 bits of every word: a call whose low half is 0 would read fresh memory
    assert lib.y3d_decode_topk2d_sharded(*args, 0, world, bufs, ctypes.c_uint64(1 << 32), None, None, 0, None) == -1


def test_no_cpu_fallback():
    with pytest.raises(y3d.Y3DError):
        y3d.v10postprocess(torch.zeros(1, 10, 6), 5, 2)
    with pytest.raises(y3d.Y3DError):
        y3d.detect_inference([torch.zeros(1, 66, 4, 4)], [8.0], 2)
    asg = y3d.TaskAlignedAssigner(topk=10, num_classes=4)
    with pytest.raises(y3d.Y3DError):
        asg(torch.rand(1, 20, 4), torch.rand(1, 20, 4), torch.rand(20, 2), torch.zeros(1, 2, 1), torch.rand(1, 2, 4),
            torch.ones(1, 2, 1))


def test_empty_gt_early_out_matches_reference_dtypes():
    asg = y3d.TaskAlignedAssigner(topk=10, num_classes=4)
    out = asg(torch.rand(2, 20, 4), torch.rand(2, 20, 4), torch.rand(20, 2), torch.zeros(2, 0, 1), torch.zeros(2, 0, 4),
              torch.zeros(2, 0, 1))
    assert [o.dtype for o in out] == [torch.float32] * 5  # tal.py:68-76
    assert float(out[0][0, 0]) == 4.0 and out[1].shape == (2, 20, 4) and out[2].shape == (2, 20, 4)


def test_pack_targets_host_side():
    """Without rows the packed tensor is [B, 0, 5 (+E)] (loss.py:183, 798); with rows the packing is a CUDA kernel
    (y3d_pack_targets, GPU-tested) and must refuse CPU tensors instead of falling back."""
    empty = y3d.loss.pack_targets(torch.zeros(0), torch.zeros(0, 1), torch.zeros(0, 4), 3, (64, 64), "cpu")
    assert empty.shape == (3, 0, 5)
    empty3 = y3d.loss.pack_targets(torch.zeros(0), torch.zeros(0, 1), torch.zeros(0, 4), 2, (64, 64), "cpu",
                                   extra=torch.zeros(0, 12))
    assert empty3.shape == (2, 0, 17)
    lib = y3d.lib()
    assert lib.y3d_pack_targets(None, None, None, None, 0, 4, 2, 3, 64.0, 64.0, None, None, None) == -1


def test_make_anchors_matches_oracle():
    lv = synth.levels(96, 320)
    feats = [torch.zeros(1, 1, h, w) for h, w in lv]
    anc, st = y3d.make_anchors(feats, synth.STRIDES)
    oanc, ost = oracle.make_anchors(lv, synth.STRIDES)
    assert np.array_equal(anc.numpy(), oanc) and np.array_equal(st.numpy()[:, 0], ost)


def test_mirror_signatures_against_the_committed_reference_signatures():
    """tests/golden/api_signatures.json holds the reference's signatures (written where /root/reference exists, by
    tests/test_reference_api_cpu.py); the mirrors must keep them wherever this runs: same names, order and defaults, extra
    mirror parameters optional."""
    import inspect
    import json

    fixture = json.load(open(os.path.join(ROOT, "tests", "golden", "api_signatures.json")))
    mirrors = {
        "ops.v10postprocess": y3d.v10postprocess, "ops.v10_3Dpostprocess": y3d.v10_3Dpostprocess,
        "ops.xywh2xyxy": y3d.xywh2xyxy, "tal.make_anchors": y3d.make_anchors,
        "tal.TaskAlignedAssigner.__init__": y3d.TaskAlignedAssigner.__init__,
        "tal.TaskAlignedAssigner.forward": y3d.TaskAlignedAssigner.forward,
        "tal.TaskAlignedAssigner3d.__init__": y3d.TaskAlignedAssigner3d.__init__,
        "tal.TaskAlignedAssigner3d.forward": y3d.TaskAlignedAssigner3d.forward,
        "loss.v8DetectionLoss.__init__": y3d.v8DetectionLoss.__init__, "loss.v8DetectionLoss.__call__": y3d.v8DetectionLoss.__call__,
        "loss.v10DetectLoss.__init__": y3d.v10DetectLoss.__init__, "loss.v10DetectLoss.__call__": y3d.v10DetectLoss.__call__,
        "loss.DDDetectionLoss.__init__": y3d.DDDetectionLoss.__init__, "loss.DDDetectionLoss.__call__": y3d.DDDetectionLoss.__call__,
        "loss.DetectLoss3d.__init__": y3d.DetectLoss3d.__init__, "loss.DetectLoss3d.__call__": y3d.DetectLoss3d.__call__,
    }
    for name, fn in mirrors.items():
        got = [[p.name, None if p.default is inspect.Parameter.empty else repr(p.default)]
               for p in inspect.signature(fn).parameters.values() if p.name != "self"]
        want = fixture[name]
        assert got[:len(want)] == want, f"{name}: mirror {got} vs reference {want}"
        assert all(d is not None for _, d in got[len(want):]), f"{name}: extra mirror parameters must be optional"
    got = [p.name for p in inspect.signature(y3d.head.inference_forward_feat).parameters.values()]
    assert got[1:] == [n for n, _ in fixture["head.v10Detect3d.inference_forward_feat"]]
