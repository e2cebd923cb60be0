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


def xyxy2xywh(x):  # reference utils/general.py:254-262
    y = x.clone()
    y[:, 0] = (x[:, 0] + x[:, 2]) / 2
    y[:, 1] = (x[:, 1] + x[:, 3]) / 2
    y[:, 2] = x[:, 2] - x[:, 0]
    y[:, 3] = x[:, 3] - x[:, 1]
    return y


def clip_coords(boxes, img_shape):
    """in place, like the reference (utils/general.py:334-340): xyxy boxes clipped to (height, width)"""
    boxes[:, 0].clamp_(0, img_shape[1])
    boxes[:, 1].clamp_(0, img_shape[0])
    boxes[:, 2].clamp_(0, img_shape[1])
    boxes[:, 3].clamp_(0, img_shape[0])


def scale_coords(img1_shape, coords, img0_shape, ratio_pad=None):
    """xyxy boxes from the letterboxed network input back to the original frame, IN PLACE on the caller's tensor (reference
    utils/general.py:319-331, called on NMS output rows at detect.py:169; the rows returned by non_max_suppression are ordinary
    writable tensors for exactly this reason)."""
    if ratio_pad is None:
        gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
        pad = (img1_shape[1] - img0_shape[1] * gain) / 2, (img1_shape[0] - img0_shape[0] * gain) / 2
    else:
        gain, pad = ratio_pad[0][0], ratio_pad[1]
    coords[:, [0, 2]] -= pad[0]
    coords[:, [1, 3]] -= pad[1]
    coords[:, :4] /= gain
    clip_coords(coords, img0_shape)
    return coords


def box_iou(box1, box2):
    """(N,4) x (M,4) xyxy -> (N,M) IoU (reference utils/general.py:388-410)"""
    area1 = (box1[:, 2] - box1[:, 0]) * (box1[:, 3] - box1[:, 1])
    area2 = (box2[:, 2] - box2[:, 0]) * (box2[:, 3] - box2[:, 1])
    inter = (torch.min(box1[:, None, 2:], box2[:, 2:]) - torch.max(box1[:, None, :2], box2[:, :2])).clamp(0).prod(2)
    return inter / (area1[:, None] + area2 - inter)


def strip_optimizer(f="best.pt", s=""):
    """reference utils/general.py:512-525: finalise a training checkpoint - EMA becomes the model, optimiser state dropped, fp16, frozen"""
    import os
    from ..models.experimental import load_checkpoint
    x = load_checkpoint(f, map_location=torch.device("cpu"))
    if x.get("ema"):
        x["model"] = x["ema"]
    for k in ("optimizer", "training_results", "wandb_id", "ema", "updates"):
        x[k] = None
    x["epoch"] = -1
    x["model"].half()
    for p in x["model"].parameters():
        p.requires_grad = False
    torch.save(x, s or f)
    mb = os.path.getsize(s or f) / 1e6
    print(f"Optimizer stripped from {f},{(' saved as %s,' % s) if s else ''} {mb:.1f}MB")


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


# ---- seg output consumers (reference detect.py:69-77,193-194,206) ----
# standard Cityscapes trainId palette (RGB) and trainId -> labelId table, the data of detect.py:19-61
Cityscapes_COLORMAP = [[128, 64, 128], [244, 35, 232], [70, 70, 70], [102, 102, 156], [190, 153, 153], [153, 153, 153], [250, 170, 30],
                       [220, 220, 0], [107, 142, 35], [152, 251, 152], [0, 130, 180], [220, 20, 60], [255, 0, 0], [0, 0, 142], [0, 0, 70],
                       [0, 60, 100], [0, 80, 100], [0, 0, 230], [119, 11, 32]]
Cityscapes_IDMAP = [[7], [8], [11], [12], [13], [17], [19], [20], [21], [22], [23], [24], [25], [26], [27], [28], [31], [32], [33]]
_lut_cache = {}


def _lut(table, device):
    key = (id(table), str(device))
    if key not in _lut_cache:
        _lut_cache[key] = torch.tensor(table, dtype=torch.uint8, device=device).contiguous()
    return _lut_cache[key]


def _lut_call(pred, table, reverse, image=None, alpha=0.0, beta=0.0, want_out=True):
    assert pred.is_cuda and pred.dtype in (torch.uint8, torch.int64), "class map: CUDA uint8 / int64 tensor"
    pred = pred.contiguous()
    lut = _lut(table, pred.device)
    n_entries, ch = lut.shape
    out = torch.empty(tuple(pred.shape) + (ch,), dtype=torch.uint8, device=pred.device) if want_out else None
    blend = None
    if image is not None:
        image = image.contiguous()
        assert image.dtype == torch.uint8 and tuple(image.shape) == tuple(pred.shape) + (ch,) and image.is_cuda
        blend = torch.empty_like(image)
    _lib.check(_lib.lib().myolo_seg_lut_blend(_lib.ptr(pred), _lib.torch_dtype_code(pred.dtype), pred.numel(), _lib.ptr(lut), n_entries, ch,
                                              int(reverse), _lib.ptr(out), _lib.ptr(image), float(alpha), float(beta), _lib.ptr(blend),
                                              _lib.stream_ptr()))
    return out, blend


def label2image(pred, COLORMAP=Cityscapes_COLORMAP):
    """class ids (H,W) -> (H,W,3) uint8 colours (reference detect.py:69-72), on the device"""
    return _lut_call(pred, COLORMAP, False)[0]


def trainid2id(pred, IDMAP=Cityscapes_IDMAP):
    """trainIds -> Cityscapes label ids (H,W,1) uint8 (reference detect.py:74-77), on the device"""
    return _lut_call(pred, IDMAP, False)[0]


def seg_overlay(pred, im0, COLORMAP=Cityscapes_COLORMAP, alpha=0.4, beta=0.6):
    """detect.py:193-194 in one kernel: mask = label2image(pred)[:, :, ::-1] (BGR) and dst = cv2.addWeighted(mask, alpha, im0, beta, 0).
    pred: (H,W) class map, im0: (H,W,3) uint8 BGR frame, both on the GPU.  Returns (mask, dst)."""
    return _lut_call(pred, COLORMAP, True, image=im0, alpha=alpha, beta=beta)
