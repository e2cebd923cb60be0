"""GPU: post-process parity - NMS rows bit-exact with the reference fixtures / the numpy oracle; seg class ids bit-exact."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import restate, synth

pytestmark = pytest.mark.gpu
GOLD = synth.GOLDEN_DIR


def run_nms(pred_np, **kw):
    from multiyolov5_b200.utils.general import non_max_suppression
    outs = non_max_suppression(torch.from_numpy(pred_np).cuda(), **kw)
    torch.cuda.synchronize()
    return [o.cpu().numpy() for o in outs]


def test_nms_reference_fixtures_bit_exact():
    g = np.load(os.path.join(GOLD, "nms_cases.npz"))
    settings = json.load(open(os.path.join(GOLD, "nms_settings.json")))
    n = 0
    for name, kw in settings.items():
        for pn in ("big", "small"):
            if f"out_{name}_{pn}_0" not in g:
                continue
            outs = run_nms(g[f"pred_{pn}"], **kw)
            for b, o in enumerate(outs):
                ref = g[f"out_{name}_{pn}_{b}"]
                assert o.shape == ref.shape, (name, pn, b, o.shape, ref.shape)
                assert np.array_equal(o, ref), (name, pn, b, np.abs(o - ref).max())
                n += 1
    assert n >= 16


@pytest.mark.parametrize("n,conf,iou", [(10000, 0.25, 0.45), (500, 0.5, 0.3), (33, 0.25, 0.45), (1, 0.25, 0.45)])
def test_nms_vs_oracle_config5(n, conf, iou):
    pred = synth.synth_predictions(2, n, seed=n)
    ref = restate.non_max_suppression(pred, conf, iou)
    got = run_nms(pred, conf_thres=conf, iou_thres=iou)
    for r, o in zip(ref, got):
        assert o.shape == r.shape and np.array_equal(o, r)


def test_nms_multilabel_large_and_properties():
    # 3000 anchors x 10 classes at conf 0.001 -> ~30k candidates: global-memory sort path + max_nms truncation
    pred = synth.synth_predictions(1, 3200, seed=11)
    got = run_nms(pred, conf_thres=0.001, iou_thres=0.6, multi_label=True)[0]
    ref = restate.non_max_suppression(pred, 0.001, 0.6, multi_label=True)[0]
    assert got.shape == ref.shape and np.array_equal(got, ref)
    assert got.shape[0] <= 300 and np.all(np.diff(got[:, 4]) <= 0)   # sorted by descending confidence


def test_nms_edge_cases():
    pred = synth.synth_predictions(2, 64, seed=5)
    assert all(o.shape == (0, 6) for o in run_nms(pred, conf_thres=1.5, iou_thres=0.45))     # nothing passes
    same = np.repeat(pred[:, :1], 64, axis=1)                                                 # 64 identical boxes -> one survivor
    outs = run_nms(same, conf_thres=0.01, iou_thres=0.45)
    refs = restate.non_max_suppression(same, 0.01, 0.45)
    for o, r in zip(outs, refs):
        assert np.array_equal(o, r) and o.shape[0] == 1
    deg = pred.copy(); deg[..., 2:4] = 0.0                                                     # zero-area boxes: NaN IoU never suppresses
    for o, r in zip(run_nms(deg, conf_thres=0.25, iou_thres=0.45), restate.non_max_suppression(deg, 0.25, 0.45)):
        assert np.array_equal(o, r)


def test_seg_argmax_fixtures_bit_exact():
    from multiyolov5_b200.utils.general import bilinear_align_corners, seg_argmax
    g = np.load(os.path.join(GOLD, "segpost_cases.npz"))
    for name, hw in {"x8": (128, 256), "odd": (40, 77), "same": (24, 24), "up2": (64, 128)}.items():
        seg = torch.from_numpy(g[f"in_{name}"]).cuda()
        am = seg_argmax(seg, hw)[0].cpu().numpy()
        assert np.array_equal(am, g[f"argmax_{name}"].astype(np.int64)), name
        am8 = seg_argmax(seg, hw, out_dtype=torch.uint8)[0].cpu().numpy()
        assert np.array_equal(am8, g[f"argmax_{name}"])
        if f"up_{name}" in g:
            up = bilinear_align_corners(seg, hw)[0].cpu().numpy()
            assert np.abs(up - g[f"up_{name}"]).max() <= 1e-5 * np.abs(g[f"up_{name}"]).max()


def test_seg_argmax_full_size_property():
    """config 5 size (19x512x1024): fused upsample+argmax == argmax of the materialised upsample; identity-size == plain argmax."""
    from multiyolov5_b200.utils.general import bilinear_align_corners, seg_argmax
    gsd = torch.Generator().manual_seed(0)
    lo = torch.randn(2, 19, 64, 128, generator=gsd).cuda()
    fused = seg_argmax(lo, (512, 1024))
    mat = bilinear_align_corners(lo, (512, 1024))
    assert torch.equal(fused, mat.argmax(1))
    full = torch.randn(1, 19, 512, 1024, generator=gsd).cuda()
    assert torch.equal(seg_argmax(full), full.argmax(1))
    # oracle on a crop-sized case
    small = lo[:1, :, :8, :16].contiguous()
    assert np.array_equal(seg_argmax(small, (64, 128)).cpu().numpy(), restate.seg_postprocess(small.cpu().numpy(), (64, 128)))


# ---- consumers of the seg output: fixtures from the reference's detect.py / utils/metrics.py ----
def test_seg_consumers_match_reference_fixtures():
    import os
    from multiyolov5_b200.utils.general import label2image, seg_overlay, trainid2id
    from multiyolov5_b200.utils.metrics import batch_intersection_union, batch_pix_accuracy, seg_eval_batch
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "consumer_cases.npz"))
    for dt in (torch.int64, torch.uint8):
        pred = torch.from_numpy(g["pred"]).to(dt).cuda()
        assert np.array_equal(label2image(pred).cpu().numpy()[:, :, ::-1], g["mask_bgr"])
        assert np.array_equal(trainid2id(pred).cpu().numpy(), g["ids"])
        mask, dst = seg_overlay(pred, torch.from_numpy(g["im0"]).cuda())
        assert np.array_equal(mask.cpu().numpy(), g["mask_bgr"]) and np.array_equal(dst.cpu().numpy(), g["blend"])
    out, tgt = torch.from_numpy(g["m_out"]).cuda(), torch.from_numpy(g["m_tgt"]).cuda()
    assert batch_pix_accuracy(out, tgt) == (int(g["m_correct"]), int(g["m_labeled"]))
    inter, union = batch_intersection_union(out, tgt, 19)
    assert np.array_equal(inter, g["m_inter"]) and np.array_equal(union, g["m_union"])
    # fused path from low-resolution logits at full size (test.py:38-41): against the oracle on the same inputs
    from oracle import restate
    rs = np.random.RandomState(2)
    seg = rs.normal(0, 1, (2, 19, 64, 128)).astype(np.float32)
    tg = rs.randint(-1, 19, (2, 512, 1024)).astype(np.int64)
    c, l, inter, union = seg_eval_batch(torch.from_numpy(seg).cuda(), torch.from_numpy(tg).cuda(), 19)
    up = restate.bilinear_align_corners_np(seg, (512, 1024))
    c0, l0, i0, u0 = restate.seg_metrics_np(up, tg, 19)
    assert l == l0 and abs(c - c0) <= 1e-4 * l0 and np.abs(inter - i0).max() <= 1e-3 * max(1, i0.max())   # argmax near-ties only


@pytest.mark.parametrize("out_dtype", [torch.int64, torch.uint8])
def test_fp16_argmax_fast_path_first_maximum_wins(out_dtype):
    """16-pixels-per-thread fp16 kernel (half mode of detect.py:96-103,191-193): exact argmax of the fp16 values, ties -> lowest class id"""
    from multiyolov5_b200.utils.general import seg_argmax
    g = torch.Generator(device="cuda").manual_seed(3)
    seg = (torch.randn((3, 19, 64, 96), device="cuda", generator=g) * 2).half()
    seg[:, 5] = seg[:, 2]                       # exact ties between two planes everywhere
    seg[0, :, 10, 10] = 1.5                     # all classes equal at one pixel -> class 0
    out = seg_argmax(seg, (64, 96), out_dtype=out_dtype)
    ref = seg.float().argmax(1)                 # torch.max over dim returns the first maximal index
    assert out.dtype == out_dtype and torch.equal(out.long(), ref)
    assert int(out[0, 10, 10]) == 0 and not bool((out == 5).any())


def test_half_logits_resized_argmax_compares_fp16_rounded_values():
    """detect.py:191-193 in half mode with a real resize: F.interpolate on a half tensor returns fp16 values and max(0)[1] runs over those.
    Our kernel interpolates in fp32 and rounds to fp16 before comparing; where the fp16 values of two classes tie, the lowest id wins.
    Yardstick: torch's own half bilinear on the same GPU.  Interpolation weights are computed slightly differently (ATen forms the source
    index from a precomputed scale), so a pixel may differ where the top two fp16 values are within one ulp: bounded, not zero."""
    from multiyolov5_b200.utils.general import seg_argmax
    g = torch.Generator(device="cuda").manual_seed(5)
    seg = (torch.randn((2, 19, 32, 48), device="cuda", generator=g) * 3).half()
    out = seg_argmax(seg, (200, 333))
    up = torch.nn.functional.interpolate(seg, (200, 333), mode="bilinear", align_corners=True)
    ref = up.float().argmax(1)
    agree = float((out == ref).float().mean())
    top2 = up.float().topk(2, dim=1).values
    close = (top2[:, 0] - top2[:, 1]) <= 2e-2 * top2[:, 0].abs().clamp_min(1.0)      # ~ a few fp16 ulps
    assert bool(((out == ref) | close).all()), "disagreement away from an fp16 near-tie"
    assert agree > 0.995, agree
