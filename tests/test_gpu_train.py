"""GPU: training row - train-mode forward (batch-statistics BatchNorm) and the hand-written backward against
torch.autograd on the oracle's fp32 restatement of the same graph (oracle.restate.model_forward_train).  fp16 activation / gradient
storage: tolerances are relative Frobenius errors per tensor."""
import numpy as np
import pytest
import torch

from oracle import restate, synth

pytestmark = pytest.mark.gpu


def rel_f(a, b):
    a = a.double(); b = b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def setup(tag="s_psp", yml="yolov5s_city_seg.yaml", B=4, H=128, W=256):
    from multiyolov5_b200.models.yolo import Model
    cfg = synth.load_cfg(yml)
    sd = synth.synth_state_dict(synth.load_manifest(tag), cfg, seed=1, gain=1.0)   # contractive weights: well-conditioned gradients
    model = Model(yml)
    model.load_state_dict(sd)
    model.cuda().train()
    x = synth.synth_image(B, H, W, seed=5)
    return model, cfg, sd, x


def oracle_train(cfg, sd, x, Rs, S):
    sdg = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k and "anchor" not in k else v.clone())
           for k, v in sd.items()}
    raw, seg = restate.model_forward_train(cfg, sdg, x)
    loss = sum((r * R).sum() for r, R in zip(raw, Rs)) + (seg * S).sum()
    loss.backward()
    return raw, seg, sdg


def test_train_forward_and_backward_match_autograd_oracle():
    model, cfg, sd, x = setup()
    gen = torch.Generator().manual_seed(11)
    out = model(x.cuda())
    raws, seg = out
    assert len(raws) == 3 and raws[0].shape == (4, 3, 16, 32, 15) and seg.shape == (4, 19, 128, 256) and seg.requires_grad
    Rs = [torch.randn(r.shape, generator=gen) * 4.0 for r in raws]
    S = torch.randn(seg.shape, generator=gen) * 0.05
    SCALE = 1024.0     # the reference trains under amp.GradScaler (train.py:265,371): activation gradients are fp16, so scale the loss
    loss = sum((r * R.cuda()).sum() for r, R in zip(raws, Rs)) + (seg * S.cuda()).sum()
    (loss * SCALE).backward()
    torch.cuda.synchronize()
    o_raw, o_seg, sdg = oracle_train(cfg, sd, x, Rs, S)
    # forward parity (batch statistics, fp16 storage)
    ef = dict(seg=rel_f(seg.detach().cpu(), o_seg.detach()), raw0=rel_f(raws[0].detach().cpu(), o_raw[0].detach()),
              raw2=rel_f(raws[2].detach().cpu(), o_raw[2].detach()))
    print("\ntrain forward rel err", ef)
    assert max(ef.values()) < 5e-2, ef      # relative Frobenius error; ~0.3 % in max-norm terms
    # gradient parity for every parameter
    errs = {}
    for name, p in model.named_parameters():
        g_ref = sdg[name].grad
        assert p.grad is not None and g_ref is not None, name
        if g_ref.norm() < 1e-8:
            continue
        errs[name] = rel_f(p.grad.detach().cpu() / SCALE, g_ref)
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:8]
    med = float(np.median(list(errs.values())))
    print("gradient rel err: median %.3e, worst %s" % (med, [(k, round(v, 4)) for k, v in worst]))
    assert med < 3e-2 and worst[0][1] < 0.25, (med, worst)


def test_running_stats_and_accumulation():
    model, cfg, sd, x = setup(B=2, H=64, W=128)
    bn0 = model.model[0].conv.bn
    rm0 = bn0.running_mean.clone()
    out = model(x.cuda())
    (out[1].sum() * 1e-3).backward()
    g1 = model.model[1].conv.weight.grad.clone()
    # running_mean <- (1-m)*old + m*batch_mean   (reference utils/torch_utils.py:150-152 momentum 0.03)
    xs = x[..., ::2, ::2], x[..., 1::2, ::2], x[..., ::2, 1::2], x[..., 1::2, 1::2]
    u = torch.nn.functional.conv2d(torch.cat(xs, 1), sd["model.0.conv.conv.weight"], None, 1, 1)
    want = 0.97 * rm0.cpu() + 0.03 * u.mean((0, 2, 3))
    assert rel_f(bn0.running_mean.cpu(), want) < 2e-2
    assert int(bn0.num_batches_tracked) == 1
    # a second forward/backward ACCUMULATES into .grad (det pass + seg pass of one iteration, train.py:371,392)
    out = model(x.cuda())
    (out[1].sum() * 1e-3).backward()
    g2 = model.model[1].conv.weight.grad
    assert rel_f(g2.cpu(), 2 * g1.cpu()) < 5e-2
