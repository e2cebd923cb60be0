"""TEST INFRASTRUCTURE — generates tests/golden/* by running the UNMODIFIED reference (imported from
/root/reference through oracle/ref_shims.py) in the build container.  The reference ships no golden
vectors, so these fixtures are what pins the oracle restatement and the CUDA path.

    python oracle/make_golden.py            # regenerate everything (needs /root/reference)

Fixtures (all small; weights are NOT stored — they are re-synthesised from the manifest + seed):
  manifest_<tag>.json      reference state_dict keys / shapes / dtypes for each model config
  net_<tag>.npz            input, z, raw x_i, seg (+ low-res logits, a few layer outputs) of Model.forward (eval, fused)
  nms_cases.npz            non_max_suppression inputs/outputs for several flag combinations
  segpost_cases.npz        detect.py:191-193 upsample+argmax inputs/outputs
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_shims, synth  # noqa: E402

GOLD = os.path.join(HERE, "..", "tests", "golden")

# tag -> (yaml in multiyolov5_b200/models, reference yaml, head line to enable, B, H, W)
NET_CASES = {
    "s_psp": ("yolov5s_city_seg.yaml", "SegMaskPSP", 2, 64, 96),
    "m_lab": ("yolov5m_city_seg_lab.yaml", "SegMaskLab", 1, 64, 64),
    "s_bise": ("yolov5s_city_seg_bise.yaml", "SegMaskBiSe", 1, 64, 64),
    "s_base": ("yolov5s_city_seg_base.yaml", "SegMaskBase", 1, 64, 64),
    "m_psp": ("yolov5m_city_seg.yaml", "SegMaskPSP", 1, 64, 64),
}
KEEP_LAYERS = (0, 4, 9, 17, 23)


def build_reference_model(ref_yolo, cfg: dict):
    import copy
    return ref_yolo.Model(copy.deepcopy(cfg))


def gen_nets(ref_yolo):
    for tag, (yml, head, B, H, W) in NET_CASES.items():
        cfg = synth.load_cfg(yml)
        torch.manual_seed(0)
        model = build_reference_model(ref_yolo, cfg)
        sd0 = model.state_dict()
        manifest = [[k, list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in sd0.items()]
        with open(os.path.join(GOLD, f"manifest_{tag}.json"), "w") as f:
            json.dump(manifest, f)
        sd = synth.synth_state_dict(manifest, cfg, seed=1)
        # anchors buffers must equal what the reference itself computed
        for k in sd:
            if k.endswith(".anchors") or k.endswith(".anchor_grid"):
                assert torch.allclose(sd[k], sd0[k]), k
        model.load_state_dict(sd)
        model.fuse().eval()
        x = synth.synth_image(B, H, W, seed=0)
        feats = {}
        hooks = [model.model[i].register_forward_hook(lambda m, a, o, i=i: feats.__setitem__(i, o.detach().clone()))
                 for i in KEEP_LAYERS]
        lowres = {}
        seg_head = model.model[24]
        seq = seg_head.out if hasattr(seg_head, "out") else (seg_head.decoder if hasattr(seg_head, "decoder") else seg_head.m)
        hooks.append(seq[-1].register_forward_hook(lambda m, a, o: lowres.__setitem__("x", a[0].detach().clone())))
        with torch.no_grad():
            out = model(x)
        for h in hooks:
            h.remove()
        (z, raw), seg = out
        arrs = dict(x=x.numpy(), z=z.numpy(), seg=seg.numpy(), seg_lowres=lowres["x"].numpy())
        for i, r in enumerate(raw):
            arrs[f"raw{i}"] = r.numpy()
        for i, t in feats.items():
            arrs[f"layer{i}"] = t.numpy().astype(np.float16)  # layer taps stored as fp16 to keep fixtures small
        np.savez_compressed(os.path.join(GOLD, f"net_{tag}.npz"), **arrs)
        print(tag, "params", sum(v.numel() for v in sd0.values()), {k: v.shape for k, v in arrs.items()},
              "z|max|", float(z.abs().max()), "seg std", float(seg.std()))


def gen_nms(ref_general):
    cases = {}
    pred = synth.synth_predictions(2, 3000, seed=0)
    pred[1, ::7, 4] = 0.1  # some rows below conf
    pred[0, 100:110] = pred[0, 90:100]  # exact duplicates -> score ties, stable order matters
    settings = {
        "default": dict(conf_thres=0.25, iou_thres=0.45),
        "test_ml": dict(conf_thres=0.001, iou_thres=0.6, multi_label=True),
        "agnostic": dict(conf_thres=0.4, iou_thres=0.5, agnostic=True),
        "classes": dict(conf_thres=0.25, iou_thres=0.45, classes=[0, 3, 7]),
        "none_pass": dict(conf_thres=1.5, iou_thres=0.45),
    }
    small = synth.synth_predictions(2, 400, seed=3)
    cases["pred_big"] = pred
    cases["pred_small"] = small
    for name, kw in settings.items():
        for pn, p in (("big", pred), ("small", small)):
            if name == "test_ml" and pn == "big":
                continue  # 30k candidates x greedy numpy oracle is slow; the small case covers the flag
            outs = ref_general.non_max_suppression(torch.from_numpy(p.copy()), **kw)
            for b, o in enumerate(outs):
                cases[f"out_{name}_{pn}_{b}"] = o.numpy()
            print("nms", name, pn, [tuple(o.shape) for o in outs])
    with open(os.path.join(GOLD, "nms_settings.json"), "w") as f:
        json.dump(settings, f)
    np.savez_compressed(os.path.join(GOLD, "nms_cases.npz"), **cases)


def gen_segpost():
    import torch.nn.functional as F
    rs = np.random.RandomState(5)
    cases = {}
    for name, (c, h, w, H, W) in {"x8": (19, 16, 32, 128, 256), "odd": (19, 9, 13, 40, 77), "same": (19, 24, 24, 24, 24),
                                  "up2": (19, 32, 64, 64, 128)}.items():
        seg = rs.normal(0, 2, (1, c, h, w)).astype(np.float32)
        up = F.interpolate(torch.from_numpy(seg), (H, W), mode="bilinear", align_corners=True)[0]
        am = up.max(axis=0)[1]
        cases[f"in_{name}"] = seg
        if name in ("odd", "same"):
            cases[f"up_{name}"] = up.numpy()
        cases[f"argmax_{name}"] = am.numpy().astype(np.uint8)
    np.savez_compressed(os.path.join(GOLD, "segpost_cases.npz"), **cases)
    print("segpost", list(cases))


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    cwd = os.getcwd()
    ref_yolo, ref_general = ref_shims.import_reference()
    gen_nets(ref_yolo)
    gen_nms(ref_general)
    gen_segpost()
