"""``KITTIDataset.decode_preds`` mirror (reference ultralytics/data/datasets/kitti.py:519-576; identical copies in
waymo.py:480 / omni3d.py:460 differ only in ``cls_mean_size``): the per-detection Python loop becomes one kernel."""
import numpy as np
import torch

from . import _lib
from ._util import f32c, ptr, stream_ptr


def decode_preds_tensor(preds, calib, inv_affine, ratio, cls_mean_size, threshold=0.001):
    """Device-side result: rows [B, D, 14] float64 (cls alpha x1 y1 x2 y2 h w l x y z ry score), valid [B, D] bool.
    ``calib`` [B,6] = cu cv fu fv tx ty; ``inv_affine`` [B,2,3]; ``ratio`` [B,2] = ratio_pad[i][0]."""
    if not preds.is_cuda:
        raise _lib.Y3DError("yolov10-3d_b200 runs on CUDA tensors only (no CPU fallback)")
    preds = f32c(preds)
    B, D, W = preds.shape
    assert W == 37
    dev = preds.device

    def f64(x):
        return torch.as_tensor(np.asarray(x, dtype=np.float64)).to(dev).contiguous() if not torch.is_tensor(x) \
            else x.to(dev, torch.float64).contiguous()

    calib, inv_affine, ratio, cms = f64(calib), f64(inv_affine), f64(ratio), f64(cls_mean_size)
    rows = torch.empty((B, D, 14), dtype=torch.float64, device=dev)
    valid = torch.empty((B, D), dtype=torch.bool, device=dev)
    _lib.check(_lib.lib().y3d_decode_preds3d(ptr(preds), B, D, int(cms.shape[0]), ptr(calib), ptr(inv_affine),
                                             ptr(ratio), ptr(cms), float(threshold), ptr(rows), ptr(valid),
                                             stream_ptr(dev)))
    return rows, valid


def decode_preds(preds, calibs, im_files, ratio_pad, inv_trans, cls_mean_size, undo_augment=True, threshold=0.001,
                 use_camera_dis=False):
    """Reference-shaped result: ``{im_file: [[cls, alpha, x1, y1, x2, y2, h, w, l, x, y, z, ry, score], ...]}``.
    ``calibs``: objects with cu, cv, fu, fv, tx, ty (kitti_utils.Calibration) or [B,6] array.  ``cls_mean_size`` is the
    dataset attribute the reference method reads from ``self`` (kitti.py:38-41).

    Only the branch the validator takes is compiled -- ``undo_augment=True`` with ``use_camera_dis`` off (kitti.py:551-557);
    the fixed 1242/1280 rescale (``undo_augment=False``) and ``camera_dis_to_rect`` raise instead of silently computing
    something else."""
    if not undo_augment:
        raise _lib.Y3DError("decode_preds: undo_augment=False (kitti.py:558-564) is outside the B200 hot path")
    if use_camera_dis:
        raise _lib.Y3DError("decode_preds: use_camera_dis (Calibration.camera_dis_to_rect) is outside the B200 hot path")
    if hasattr(calibs[0], "cu"):
        calib = np.array([[c.cu, c.cv, c.fu, c.fv, c.tx, c.ty] for c in calibs], dtype=np.float64)
    else:
        calib = np.asarray(calibs, dtype=np.float64)
    ratio = np.array([np.asarray(rp[0], dtype=np.float64)[:2] for rp in ratio_pad])
    inv = np.stack([np.asarray(t, dtype=np.float64) for t in inv_trans])
    rows, valid = decode_preds_tensor(preds, calib, inv, ratio, cls_mean_size, threshold)
    rows, valid = rows.cpu().numpy(), valid.cpu().numpy()  # the one device->host boundary (kitti.py:521)
    out = {}
    for i, f in enumerate(im_files):
        r = rows[i][valid[i]]
        out[f] = [[int(x[0])] + x[1:].tolist() for x in r]
    return out


def rotate_iou_gpu_eval(boxes, query_boxes, criterion=-1, device_id=0):
    """``rotate_iou_gpu_eval`` (data/datasets/kitti_eval.py:309-344): numpy (or CUDA tensor) boxes [N,5] / [K,5] =
    (cx, cy, dx, dy, angle) -> overlap matrix [N,K] in the input's dtype / container.  numpy in -> numpy out, like the
    reference (whose numba kernel also makes the host round trip); CUDA tensors in -> CUDA tensor out, no round trip."""
    as_numpy = not torch.is_tensor(boxes)
    dev = torch.device("cuda", device_id) if as_numpy else boxes.device
    if not as_numpy and not boxes.is_cuda:
        raise _lib.Y3DError("yolov10-3d_b200 runs on CUDA tensors only (no CPU fallback)")
    b = torch.as_tensor(np.asarray(boxes, dtype=np.float32) if as_numpy else boxes, device=dev).float().contiguous()
    q = torch.as_tensor(np.asarray(query_boxes, dtype=np.float32) if as_numpy else query_boxes, device=dev).float().contiguous()
    N, K = int(b.shape[0]), int(q.shape[0])
    iou = torch.zeros((N, K), dtype=torch.float32, device=dev)
    if N and K:
        _lib.check(_lib.lib().y3d_rotate_iou_eval(ptr(b), N, ptr(q), K, int(criterion), ptr(iou), stream_ptr(dev)))
    return iou.cpu().numpy().astype(np.asarray(boxes).dtype) if as_numpy else iou
