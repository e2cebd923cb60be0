"""Small pass over every C-ABI family for compute-sanitizer (memcheck / racecheck / initcheck are 10-50x slower than native: sizes are the
smallest that still take the tcgen05 paths).  Run ON THE GPU BOX:
    compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_pass.py > gpurun_out/sanitizer_memcheck.txt 2>&1
Inference forward (tcgen05 + CUDA-core convs, grouped launches, graph replay), NMS, seg upsample / argmax, letterbox, consumers, one Trainer
step (train forward x2 concurrently, fused det loss, fused seg CE, backward incl. tcgen05 wgrad, grouped weight repack, SGD)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from multiyolov5_b200.models.yolo import Model
from multiyolov5_b200.utils.general import non_max_suppression, seg_argmax
from multiyolov5_b200.train import Trainer, scale_hyp
import bench

torch.cuda.set_device(0)
yml, cfg, sd = bench.make_weights("s_psp")
model = Model(yml)
model.load_state_dict(sd)
model.cuda().eval()
B, H, W = 2, 128, 256
torch.manual_seed(0)
x = torch.rand(B, 3, H, W, device="cuda")
for _ in range(3):                               # eager warm-up, graph capture, graph replay
    (z, raw), seg = model(x)
dets = non_max_suppression(z, 0.25, 0.45)
cls_map = seg_argmax(seg, (H, W))
(zh, _), segh = model(x.half())                  # half mode: fp16 logits + fp16 argmax
cls_h = seg_argmax(segh, (H, W))
torch.cuda.synchronize()
print("inference ok:", [int(d.shape[0]) for d in dets], int(cls_map.max()), int(cls_h.max()))

model.train()
hyp = dict(lr0=0.01, momentum=0.937, weight_decay=5e-4, box=0.05, cls=0.5, cls_pw=1.0, obj=1.0, obj_pw=1.0, anchor_t=4.0, fl_gamma=0.0)
hyp = scale_hyp(hyp, nl=3, nc=cfg["nc"], imgsz=256, total_batch_size=2 * B)
tr = Trainer(model, hyp, batch_size=B, init_scale=2.0 ** 10)
rs = np.random.RandomState(0)
t = np.zeros((12, 6), np.float32)
t[:, 0] = rs.randint(0, B, 12); t[:, 1] = rs.randint(0, cfg["nc"], 12)
t[:, 2:4] = rs.uniform(0.1, 0.9, (12, 2)); t[:, 4:6] = rs.uniform(0.05, 0.4, (12, 2))
targets = torch.from_numpy(t).cuda()
mask = torch.from_numpy(rs.randint(-1, 19, (B, H, W)).astype(np.int64)).cuda()
segx = torch.rand(B, 3, H, W, device="cuda")
for it in range(3):                              # eager, captured, replayed; step 2 uses the grouped repack
    items, segloss = tr.step(x, targets, segx, mask)
torch.cuda.synchronize()
print("train ok:", [round(float(v), 4) for v in items], round(float(segloss), 4))
model.eval()
(z2, _), _ = model(x)                            # BN-folded packs re-made from the moved running statistics
torch.cuda.synchronize()
print("eval after train ok:", bool(torch.isfinite(z2).all()))
