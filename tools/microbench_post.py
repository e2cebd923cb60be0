"""BASELINE.json configs[4]: NMS (10k boxes/img) + seg argmax (19x512x1024) micro-bench, device time via CUDA events.
Prints one JSON line; compares with torch/torchvision CUDA ops (the reference's own GPU backend) on the same inputs."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from multiyolov5_b200.utils.general import non_max_suppression, seg_argmax  # noqa: E402
from oracle import restate, synth  # noqa: E402


def timeit(fn, n=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3  # us


def ref_nms_torch_cuda(pred, conf=0.25, iou=0.45):
    import torchvision
    out = []
    for x in pred:
        x = x[x[:, 4] > conf]
        x = x.clone()
        x[:, 5:] *= x[:, 4:5]
        box = x[:, :4].clone()
        box[:, 0] = x[:, 0] - x[:, 2] / 2; box[:, 1] = x[:, 1] - x[:, 3] / 2
        box[:, 2] = x[:, 0] + x[:, 2] / 2; box[:, 3] = x[:, 1] + x[:, 3] / 2
        c, j = x[:, 5:].max(1, keepdim=True)
        x = torch.cat((box, c, j.float()), 1)[c.view(-1) > conf]
        i = torchvision.ops.nms(x[:, :4] + x[:, 5:6] * 4096, x[:, 4], iou)[:300]
        out.append(x[i])
    return out


B = 16
pred_np = synth.synth_predictions(B, 10000, seed=0)
pred = torch.from_numpy(pred_np).cuda()
ours = non_max_suppression(pred, 0.25, 0.45)
ref = restate.non_max_suppression(pred_np[:2], 0.25, 0.45)
exact = all(np.array_equal(o.cpu().numpy(), r) for o, r in zip(ours[:2], ref))
t_nms = timeit(lambda: non_max_suppression(pred, 0.25, 0.45, return_padded=True))
try:
    t_nms_tv = timeit(lambda: ref_nms_torch_cuda(pred), n=5, warm=2)
    tv_same = all(torch.equal(a, b) for a, b in zip(ours, ref_nms_torch_cuda(pred)))
except Exception as e:  # torchvision CUDA ops may be missing
    t_nms_tv, tv_same = None, str(e)[:80]
logits = torch.randn(B, 19, 512, 1024, device="cuda")
lo = torch.randn(B, 19, 64, 128, device="cuda")
t_arg = timeit(lambda: seg_argmax(logits))
t_arg_torch = timeit(lambda: logits.max(1)[1])
t_fused = timeit(lambda: seg_argmax(lo, (512, 1024)))
t_fused_torch = timeit(lambda: torch.nn.functional.interpolate(lo, (512, 1024), mode="bilinear", align_corners=True).max(1)[1])
same_arg = bool(torch.equal(seg_argmax(logits), logits.max(1)[1]))
hbm = logits.numel() * 4 + B * 512 * 1024 * 8
# ---- pre-process: 16 Cityscapes-size BGR frames 2048x1024 uint8 -> (16,3,512,1024) fp16 in [0,1]; CPU side: cv2 + numpy ----
from multiyolov5_b200.utils.datasets import preprocess  # noqa: E402
from multiyolov5_b200.utils.metrics import seg_eval_batch  # noqa: E402
from multiyolov5_b200.utils.general import seg_overlay  # noqa: E402
frames = torch.randint(0, 256, (B, 1024, 2048, 3), dtype=torch.uint8, device="cuda")
t_pre = timeit(lambda: preprocess(frames, 1024, stride=32, half=True))
pre_bytes = frames.numel() + B * 3 * 512 * 1024 * 2
frames_odd = torch.randint(0, 256, (B, 720, 1280, 3), dtype=torch.uint8, device="cuda")
t_pre_lin = timeit(lambda: preprocess(frames_odd, 1024, stride=32, half=True))          # general bilinear path (1280x720 -> 1024x576)
pre_ok = bool(np.array_equal((preprocess(frames[:1], 1024, 32, half=False)[0][0] * 255).round().byte().cpu().numpy(),
                             restate.preprocess_np(frames[0].cpu().numpy(), 1024, 32)))
import time  # noqa: E402
import cv2  # noqa: E402
f0 = frames[0].cpu().numpy()
t0 = time.time()
for _ in range(5):
    restate_img = cv2.resize(f0, (1024, 512), interpolation=cv2.INTER_LINEAR)[:, :, ::-1].transpose(2, 0, 1)
    _ = np.ascontiguousarray(restate_img)
t_pre_cpu = (time.time() - t0) / 5 * 1e6
# ---- consumers: validation counters from low-res logits, overlay ----
tgt = torch.randint(-1, 19, (B, 512, 1024), device="cuda")
t_eval = timeit(lambda: seg_eval_batch(lo, tgt, 19), n=10, warm=3)
cls = seg_argmax(lo[:1], (1024, 2048), out_dtype=torch.uint8)[0]
t_overlay = timeit(lambda: seg_overlay(cls, frames[0]))
print(json.dumps({"preprocess_us_per_frame": t_pre / B, "preprocess_GBps": pre_bytes / (t_pre * 1e-6) / 1e9, "preprocess_bit_exact_vs_oracle": pre_ok,
                  "preprocess_bilinear_1280x720_us_per_frame": t_pre_lin / B, "preprocess_cv2_cpu_us_per_frame": t_pre_cpu,
                  "seg_eval_batch_us_per_img_incl_host_sync": t_eval / B, "seg_overlay_2048x1024_us": t_overlay,
"nms_us_per_img": t_nms / B, "nms_bit_exact_vs_oracle": exact, "nms_torchvision_cuda_us_per_img": None if t_nms_tv is None else t_nms_tv / B,
                  "nms_equal_torchvision_cuda": tv_same, "argmax_us_per_img": t_arg / B, "argmax_GBps": hbm / (t_arg * 1e-6) / 1e9,
                  "argmax_torch_cuda_us_per_img": t_arg_torch / B, "argmax_equal_torch": same_arg,
                  "upsample_argmax_fused_us_per_img": t_fused / B, "upsample_argmax_torch_cuda_us_per_img": t_fused_torch / B, "batch": B}))
