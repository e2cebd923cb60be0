"""Diagnostics for the training path: run-to-run determinism and per-parameter gradient error against the autograd oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import restate, synth
from multiyolov5_b200.models.yolo import Model


def rel_f(a, b):
    a = a.double(); b = b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def main(B=4, H=128, W=256, mode="rand", scale=1.0):
    yml, tag = "yolov5s_city_seg.yaml", "s_psp"
    cfg = synth.load_cfg(yml)
    sd = synth.synth_state_dict(synth.load_manifest(tag), cfg, seed=1, gain=1.0)
    model = Model(yml); model.load_state_dict(sd); model.cuda().train()
    x = synth.synth_image(B, H, W, seed=5)
    gen = torch.Generator().manual_seed(11)
    grads = []
    outs = []
    Rs = S = None
    for it in range(2):
        model.zero_grad(set_to_none=False)
        raws, seg = model(x.cuda())
        if Rs is None:
            Rs = [torch.randn(r.shape, generator=gen) * 4.0 for r in raws]
            S = torch.randn(seg.shape, generator=gen) * 0.05
        if mode == "rand":
            loss = sum((r * R.cuda()).sum() for r, R in zip(raws, Rs)) + (seg * S.cuda()).sum()
        else:
            loss = seg.sum() * 1e-3
        outs.append([r.detach().cpu().clone() for r in raws] + [seg.detach().cpu().clone()])
        (loss * scale).backward()
        torch.cuda.synchronize()
        grads.append({n: p.grad.detach().cpu().clone() / scale for n, p in model.named_parameters()})
    fw = [rel_f(a, b) for a, b in zip(outs[0], outs[1])]
    print("== run-to-run relative difference (should be ~1e-6) ==")
    d = {n: rel_f(grads[0][n], grads[1][n]) for n in grads[0] if grads[1][n].norm() > 0}
    for n, v in d.items():
        print("  %-40s %.3e" % (n, v))
    print("== forward run-to-run ==", fw)
    sdg = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k and "anchor" not in k else v.clone())
           for k, v in sd.items()}
    raw, oseg = restate.model_forward_train(cfg, sdg, x)
    if mode == "rand":
        loss = sum((r * R).sum() for r, R in zip(raw, Rs)) + (oseg * S).sum()
    else:
        loss = oseg.sum() * 1e-3
    loss.backward()
    # yardstick: the SAME restated graph through torch's own fp16 autocast on the GPU (what the reference's amp.autocast training computes)
    sda = {k: (v.detach().clone().cuda().requires_grad_(True) if v.requires_grad else v.detach().clone().cuda()) for k, v in sdg.items()}
    with torch.autocast("cuda", dtype=torch.float16):
        araw, aseg = restate.model_forward_train(cfg, sda, x.cuda())
    if mode == "rand":
        aloss = sum((r.float() * R.cuda()).sum() for r, R in zip(araw, Rs)) + (aseg.float() * S.cuda()).sum()
    else:
        aloss = aseg.float().sum() * 1e-3
    aloss.backward()
    print("== forward error vs fp32 oracle: ours %s | torch autocast %s" % (
        [round(rel_f(a, b.detach()), 4) for a, b in zip(outs[0], list(raw) + [oseg])],
        [round(rel_f(a.detach().float().cpu(), b.detach()), 4) for a, b in zip(list(araw) + [aseg], list(raw) + [oseg])]))
    ea = {n: rel_f(sda[n].grad.float().cpu(), sdg[n].grad) for n, _ in model.named_parameters() if sdg[n].grad is not None and sdg[n].grad.norm() > 1e-8}
    eo = {n: rel_f(grads[0][n], sdg[n].grad) for n in ea}
    print("== gradient error vs fp32 oracle: ours median %.3e max %.3e | torch autocast median %.3e max %.3e" % (
        float(np.median(list(eo.values()))), max(eo.values()), float(np.median(list(ea.values()))), max(ea.values())))
    print("== per-parameter error vs oracle (model order) ==")
    for n, p in model.named_parameters():
        g = sdg[n].grad
        if g is None or g.norm() < 1e-8:
            continue
        cos = float((grads[0][n].double().flatten() @ g.double().flatten()) / (grads[0][n].double().norm() * g.double().norm() + 1e-30))
        print("  %-40s rel %.3e cos %.5f |g| %.3e ratio %.4f" % (n, rel_f(grads[0][n], g), cos, float(g.norm()), float(grads[0][n].norm() / g.norm())), "amp %.3e" % ea.get(n, -1))


if __name__ == "__main__":
    a = sys.argv[1:]
    main(B=int(a[0]) if a else 4, H=int(a[1]) if len(a) > 1 else 128, W=int(a[2]) if len(a) > 2 else 256,
         mode=a[3] if len(a) > 3 else "rand", scale=float(a[4]) if len(a) > 4 else 1.0)
