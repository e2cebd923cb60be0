"""Checkpoint round-trips of the drop-in boundary (SURVEY.md section 8b "picklability", 8f-4): reference-written `.pt` files load through
`attempt_load` (reference models/experimental.py:113-134), `deepcopy(model).half()` + torch.save as train.py:482-494 does, `strip_optimizer`
(utils/general.py:512-525), ModelEMA (utils/torch_utils.py:270-304).  The fixture `tests/golden/ref_ckpt_tiny.pt` was written by the
UNMODIFIED reference's classes (oracle/make_golden.py gen_ckpt)."""
import copy
import os

import numpy as np
import pytest
import torch

from oracle import restate, synth

CKPT = os.path.join(synth.GOLDEN_DIR, "ref_ckpt_tiny.pt")
OUT = os.path.join(synth.GOLDEN_DIR, "ref_ckpt_tiny_out.npz")


def test_attempt_load_resolves_reference_classes_and_matches_reference_output():
    from multiyolov5_b200.models.experimental import attempt_load
    from multiyolov5_b200.models.yolo import Model
    from multiyolov5_b200.plan import build_plan
    m = attempt_load(CKPT, map_location="cpu")
    g = np.load(OUT)
    assert isinstance(m, Model) and not m.training and next(m.parameters()).dtype == torch.float32
    assert list(m.names) == list(g["names"]) and np.array_equal(m.stride.numpy(), g["stride"])
    fresh = Model(m.yaml)
    assert list(fresh.state_dict().keys()) == list(m.state_dict().keys())
    assert len(build_plan(m, 1, 64, 64).ops) == len(build_plan(fresh, 1, 64, 64).ops)      # planner reads only attributes the fix-up restores
    # the weights that came out of the pickle, run through the oracle, reproduce what the reference computed from the same file
    sd = {k: v.float() for k, v in m.state_dict().items()}
    out = restate.model_forward(dict(m.yaml), sd, synth.synth_image(1, 64, 64, seed=5))
    assert float((out["seg"] - torch.from_numpy(g["seg"])).abs().max()) < 1e-5 * max(1.0, float(np.abs(g["seg"]).max()))
    assert float((out["z"] - torch.from_numpy(g["z"])).abs().max()) < 1e-5 * float(np.abs(g["z"]).max())


def test_deepcopy_half_save_load_and_strip_optimizer(tmp_path):
    """train.py:482-494 + utils/general.py:512-525 on OUR module"""
    from multiyolov5_b200.models.experimental import attempt_load, load_checkpoint
    from multiyolov5_b200.models.yolo import Model
    from multiyolov5_b200.utils.general import strip_optimizer
    from multiyolov5_b200.utils.torch_utils import ModelEMA
    cfg = synth.load_cfg("yolov5s_city_seg.yaml")
    cfg["width_multiple"] = 0.25
    model = Model(cfg)
    model.engine()                      # a live engine object must not travel with copies / pickles
    ema = ModelEMA(model)
    with torch.no_grad():
        for p in model.parameters():
            p.add_(0.01)
    ema.update(model)
    d = ema.decay(1)
    k0 = "model.0.conv.conv.weight"
    assert torch.allclose(ema.ema.state_dict()[k0], model.state_dict()[k0] - 0.01 * d, atol=1e-6)
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9)
    ckpt = {"epoch": 1, "best_fitness": 0.1, "training_results": "", "model": copy.deepcopy(model).half(), "ema": copy.deepcopy(ema.ema).half(),
            "updates": ema.updates, "optimizer": opt.state_dict(), "wandb_id": None}
    f = str(tmp_path / "last.pt")
    torch.save(ckpt, f)
    assert model._engine is not None and ckpt["model"]._engine is None
    m = attempt_load(f, map_location="cpu")            # picks 'ema'
    assert torch.allclose(m.state_dict()[k0], ema.ema.state_dict()[k0].half().float())
    strip_optimizer(f)
    x = load_checkpoint(f)
    assert x["optimizer"] is None and x["ema"] is None and x["epoch"] == -1
    assert next(x["model"].parameters()).dtype == torch.float16 and not any(p.requires_grad for p in x["model"].parameters())
    assert torch.equal(x["model"].state_dict()[k0], ema.ema.state_dict()[k0].half())


def test_scale_coords_clip_and_box_iou_follow_the_reference_arithmetic():
    from multiyolov5_b200.utils.general import box_iou, clip_coords, scale_coords, xywh2xyxy, xyxy2xywh
    rs = np.random.RandomState(0)
    b = torch.from_numpy(rs.uniform(-20, 1100, (50, 6)).astype(np.float32))
    b[:, 2:4] = b[:, :2] + torch.from_numpy(rs.uniform(1, 300, (50, 2)).astype(np.float32))
    ref = b.clone()
    img1, img0 = (512, 1024), (1024, 2048)          # letterboxed input -> Cityscapes frame: gain 0.5, no padding
    out = scale_coords(img1, b, img0)
    assert out is b                                  # in place on the caller's tensor
    exp = ref.clone()
    exp[:, :4] /= 0.5
    exp[:, [0, 2]] = exp[:, [0, 2]].clamp(0, 2048)
    exp[:, [1, 3]] = exp[:, [1, 3]].clamp(0, 1024)
    assert torch.equal(b, exp)
    img0 = (720, 1280)                               # gain = min(512/720, 1024/1280) = 0.7111, pad x = (1024 - 1280*g)/2
    c = ref.clone()
    scale_coords(img1, c, img0)
    g = min(512 / 720, 1024 / 1280)
    px, py = (1024 - 1280 * g) / 2, (512 - 720 * g) / 2
    e = ref.clone()
    e[:, [0, 2]] -= px
    e[:, [1, 3]] -= py
    e[:, :4] /= g
    clip_coords(e, img0)
    assert torch.equal(c, e)
    a = torch.tensor([[0., 0., 10., 10.], [5., 5., 15., 15.]])
    iou = box_iou(a, a)
    assert torch.allclose(iou, torch.tensor([[1.0, 25.0 / 175.0], [25.0 / 175.0, 1.0]]))
    assert torch.allclose(xywh2xyxy(xyxy2xywh(a)), a)


@pytest.mark.gpu
def test_gpu_forward_deepcopy_half_save_load_forward_equal(tmp_path):
    """the sequence the verdict names: forward -> deepcopy(model).half() -> torch.save / torch.load -> forward gives the same outputs as
    the original model put in the same (half) precision"""
    from multiyolov5_b200.models.experimental import attempt_load
    from multiyolov5_b200.models.yolo import Model
    cfg = synth.load_cfg("yolov5s_city_seg.yaml")
    sd = synth.synth_state_dict(synth.load_manifest("s_psp"), cfg, seed=1)
    model = Model("yolov5s_city_seg.yaml")
    model.load_state_dict(sd)
    model.cuda().eval()
    x = synth.synth_image(1, 128, 256, seed=2).cuda()
    (z0, _), seg0 = model(x)
    cp = copy.deepcopy(model).half()
    assert cp._engine is None and model._engine is not None
    f = str(tmp_path / "ckpt.pt")
    torch.save({"model": cp, "ema": None}, f)
    m2 = attempt_load(f, map_location="cuda")          # .float().fuse().eval()
    (z1, _), seg1 = m2(x)
    ref = copy.deepcopy(model)
    ref.half().float()                                  # same fp16-rounded master weights
    (z2, _), seg2 = ref(x)
    torch.cuda.synchronize()
    assert torch.equal(z1, z2) and torch.equal(seg1, seg2)
    assert float((seg1 - seg0).abs().max()) < 0.05 * float(seg0.abs().max())      # only the fp16 rounding of the stored weights differs
    (z3, _), _ = model(x)                               # the original keeps working after having been copied
    assert torch.equal(z3, z0)


@pytest.mark.gpu
def test_gpu_reference_checkpoint_forward_matches_reference_output():
    from multiyolov5_b200.models.experimental import attempt_load
    m = attempt_load(CKPT, map_location="cuda")
    g = np.load(OUT)
    (z, raw), seg = m(synth.synth_image(1, 64, 64, seed=5).cuda())
    torch.cuda.synchronize()
    rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())  # noqa: E731
    assert rel(seg.cpu().numpy(), g["seg"]) < 1e-2 and rel(raw[0].cpu().numpy(), g["raw0"]) < 1e-2 and rel(z.cpu().numpy(), g["z"]) < 2e-2
