"""clock64 timelines (CTA 0) of conv_tc_kernel for the layer classes of yolov5s_city_seg at batch 16 x 512 x 1024.
    make -C multiyolov5_b200/csrc TIMELINE=1 BUILD=build_tl OUT=../libmyolo_timeline.so
    MYOLO_LIB=multiyolov5_b200/libmyolo_timeline.so python tools/conv_timeline_set.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["MYOLO_CONV_TIMELINE"] = "1"
import torch
from multiyolov5_b200 import ops
SHAPES = [  # B, H, W, Ci, Co, k, s, d   (input map)
    (16, 64, 128, 64, 64, 1, 1, 1),     # P3 bottleneck cv1 (1024 tiles, 3.5 per CTA)
    (16, 32, 64, 128, 128, 1, 1, 1),    # P4 bottleneck cv1 (256 tiles, one per CTA)
    (16, 16, 32, 512, 256, 1, 1, 1),    # P5 1x1 (C3 cv1/cv2, SPP cv1)
    (16, 16, 32, 256, 256, 3, 1, 1),    # P5 bottleneck 3x3
    (16, 32, 64, 256, 128, 1, 1, 1),    # P4 1x1
    (16, 32, 64, 128, 128, 3, 1, 1),    # P4 bottleneck 3x3
    (16, 64, 128, 128, 64, 1, 1, 1),    # P3 1x1
    (16, 64, 128, 64, 64, 3, 1, 1),     # P3 bottleneck 3x3
    (16, 64, 128, 256, 128, 3, 1, 1),   # FFM 3x3
    (16, 128, 256, 64, 32, 1, 1, 1),    # L2 cv1
    (16, 128, 256, 32, 32, 3, 1, 1),    # L2 bottleneck 3x3
    (16, 256, 512, 32, 64, 3, 2, 1),    # L1
    (16, 256, 512, 16, 32, 3, 1, 1),    # L0
    (16, 64, 128, 128, 256, 3, 2, 1),   # L5
]
for (B, H, W, Ci, Co, k, s, d) in SHAPES:
    x = torch.randn(B, H, W, Ci).half().cuda()
    w = (torch.randn(Co, Ci, k, k) * (2.0 / (Ci * k * k)) ** 0.5).cuda()
    bn = [torch.ones(Co).cuda(), torch.zeros(Co).cuda(), torch.zeros(Co).cuda(), torch.ones(Co).cuda()]
    print(f"== conv B{B} {H}x{W} {Ci}->{Co} k{k} s{s} d{d}", flush=True)
    ops.conv_bn_silu(x, w, bn, stride=s, dil=d, path=1)
    torch.cuda.synchronize()
