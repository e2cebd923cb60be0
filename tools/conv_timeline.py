"""Per-tile pipeline timeline of conv_tc_kernel (clock64 stamps of CTA 0) for one layer shape.
    MYOLO_CONV_TIMELINE=1 python tools/conv_timeline.py B H W Ci Co k stride dil"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["MYOLO_CONV_TIMELINE"] = "1"
import torch
from multiyolov5_b200 import ops
B, H, W, Ci, Co, k, s, d = [int(a) for a in sys.argv[1:9]]
x = torch.randn(B, H, W, Ci).half().cuda()
w = (torch.randn(Co, Ci, k, k) * (2.0 / (Ci * k * k)) ** 0.5).cuda()
bn = [torch.ones(Co).cuda(), torch.zeros(Co).cuda(), torch.zeros(Co).cuda(), torch.ones(Co).cuda()]
print(f"== conv B{B} {H}x{W} {Ci}->{Co} k{k} s{s} d{d}", flush=True)
y = ops.conv_bn_silu(x, w, bn, stride=s, dil=d, path=1)
torch.cuda.synchronize()
