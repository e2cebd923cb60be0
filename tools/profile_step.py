"""One benchmark step of the hot path between cudaProfilerStart/Stop, for `ncu --profile-from-start off`.
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/profile_step.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from multiyolov5_b200.models.yolo import Model  # noqa: E402
from multiyolov5_b200.utils.general import non_max_suppression, seg_argmax  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "s_psp"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
yml, cfg, sd = bench.make_weights(tag)
model = Model(yml)
model.load_state_dict(sd)
model.cuda().eval()
model.half()          # same configuration as bench.py's default (the reference's CUDA path: detect.py:96-103)
x = torch.rand(B, 3, bench.H, bench.W, device="cuda").half()


def step():
    (z, raw), seg = model(x)
    det, cnt = non_max_suppression(z, 0.25, 0.45, return_padded=True)
    return seg_argmax(seg, (bench.H, bench.W))


for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step:", model.engine().launches(), "forward launches")
