"""Segmentation validation metrics behind the reference's names (utils/metrics.py:234-275), computed on the GPU: the class map never
leaves the device, only 2 + 3*nclass counters do.

    batch_pix_accuracy(output, target)                -> (pixel_correct, pixel_labeled)
    batch_intersection_union(output, target, nclass)  -> (area_inter[nclass], area_union[nclass])
    seg_eval_batch(seg, target, nclass)               -> all four from the model's seg output (any resolution): the x8 bilinear upsample,
                                                         the argmax and the counters run without materialising full-resolution logits
`output`: (B,C,H,W) CUDA logits; `target`: (B,H,W) integer labels, -1 = ignore.
"""
import numpy as np
import torch

from .. import _lib
from .general import seg_argmax


def _counters(pred: torch.Tensor, target: torch.Tensor, nclass: int) -> np.ndarray:
    assert pred.is_cuda and pred.dtype in (torch.uint8, torch.int64) and pred.shape == target.shape
    tgt = target.to(device=pred.device, dtype=torch.int64).contiguous()
    out = torch.zeros(2 + 3 * nclass, dtype=torch.int64, device=pred.device)
    _lib.check(_lib.lib().myolo_seg_metrics(_lib.ptr(pred.contiguous()), _lib.torch_dtype_code(pred.dtype), _lib.ptr(tgt), pred.numel(), nclass,
                                            _lib.ptr(out), _lib.stream_ptr()))
    return out.cpu().numpy()


def _class_map(output: torch.Tensor, hw) -> torch.Tensor:
    return seg_argmax(output, tuple(hw), out_dtype=torch.uint8 if output.shape[1] <= 256 else torch.int64)


def seg_eval_batch(seg, target, nclass):
    """(pixel_correct, pixel_labeled, area_inter, area_union) like test.py:33-41 (`eval_batch`) for one batch"""
    c = _counters(_class_map(seg, target.shape[-2:]), target, nclass)
    inter, pred_a, lab_a = c[2:2 + nclass], c[2 + nclass:2 + 2 * nclass], c[2 + 2 * nclass:]
    return int(c[0]), int(c[1]), inter.copy(), pred_a + lab_a - inter


def batch_pix_accuracy(output, target):
    correct, labeled, _, _ = seg_eval_batch(output, target, output.shape[1])
    assert correct <= labeled, "Correct area should be smaller than Labeled"
    return correct, labeled


def batch_intersection_union(output, target, nclass):
    _, _, inter, union = seg_eval_batch(output, target, nclass)
    assert (inter <= union).all(), "Intersection area should be smaller than Union area"
    return inter, union
