"""Post-process of the hot path behind the reference's names (reference utils/general.py, detect.py:191-193)."""
import ctypes as C

import torch

from .. import _lib

_ws_cache = {}


def xywh2xyxy(x):  # reference utils/general.py:265-272 (host-side helper kept for callers; NMS does this on the device)
    y = x.clone()
    y[:, 0] = x[:, 0] - x[:, 2] / 2
    y[:, 1] = x[:, 1] - x[:, 3] / 2
    y[:, 2] = x[:, 0] + x[:, 2] / 2
    y[:, 3] = x[:, 1] + x[:, 3] / 2
    return y


def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False, labels=(),
                        max_det=300, return_padded=False):
    """reference utils/general.py:421-509.  prediction: (B,A,5+nc) fp32 CUDA.  Returns list[(n,6)] like the reference
    ([x1,y1,x2,y2,conf,cls], descending conf, <=300 rows).  `labels` (auto-labelling apriori boxes) is not on the hot path."""
    if labels:
        raise NotImplementedError("non_max_suppression(labels=...) (autolabelling) is out of scope (SURVEY.md section 8)")
    if not prediction.is_cuda:
        raise _lib.MyoloError("non_max_suppression needs a CUDA tensor: there is no CPU path in multiyolov5_b200")
    pred = prediction.float().contiguous()
    B, A, no = pred.shape
    L = _lib.lib()
    ml = bool(multi_label) and (no - 5) > 1
    nbytes = int(L.myolo_nms_workspace_bytes(B, A, no, int(ml)))
    key = (pred.device, nbytes)
    ws = _ws_cache.get(key)
    if ws is None:
        _ws_cache.clear()
        ws = _ws_cache[key] = torch.empty(nbytes, dtype=torch.uint8, device=pred.device)
    out = torch.zeros((B, max_det, 6), dtype=torch.float32, device=pred.device)
    cnt = torch.empty((B,), dtype=torch.int32, device=pred.device)
    cls_t = torch.tensor(list(classes), dtype=torch.int32, device=pred.device) if classes is not None else None
    _lib.check(L.myolo_nms(_lib.ptr(pred), B, A, no, float(conf_thres), float(iou_thres), _lib.ptr(cls_t),
                           0 if cls_t is None else cls_t.numel(), int(bool(agnostic)), int(ml), int(max_det), 30000, 4096.0,
                           _lib.ptr(out), _lib.ptr(cnt), _lib.ptr(ws), nbytes, _lib.stream_ptr()))
    if return_padded:
        return out, cnt
    counts = cnt.tolist()  # the one device->host sync, like the reference's shape checks (utils/general.py:458,484)
    return [out[b, :counts[b]] for b in range(B)]


def seg_argmax(seg, out_hw=None, out_dtype=torch.int64):
    """detect.py:191-193: F.interpolate(seg,(H0,W0),bilinear,align_corners=True) then argmax over classes, fused.
    seg: (B,C,h,w) fp32/fp16 CUDA -> (B,H0,W0) int64 (or uint8)."""
    if not seg.is_cuda:
        raise _lib.MyoloError("seg_argmax needs a CUDA tensor")
    seg = seg.contiguous()
    B, Cc, h, w = seg.shape
    H, W = out_hw if out_hw is not None else (h, w)
    out = torch.empty((B, H, W), dtype=out_dtype, device=seg.device)
    _lib.check(_lib.lib().myolo_seg_upsample_argmax(_lib.ptr(seg), _lib.torch_dtype_code(seg.dtype), B, Cc, h, w, H, W, _lib.ptr(out),
                                                    _lib.torch_dtype_code(out_dtype), _lib.stream_ptr()))
    return out


def bilinear_align_corners(seg, out_hw):
    seg = seg.float().contiguous()
    B, Cc, h, w = seg.shape
    out = torch.empty((B, Cc, out_hw[0], out_hw[1]), dtype=torch.float32, device=seg.device)
    _lib.check(_lib.lib().myolo_bilinear_nchw(_lib.ptr(seg), B, Cc, h, w, out_hw[0], out_hw[1], _lib.ptr(out), _lib.stream_ptr()))
    return out
