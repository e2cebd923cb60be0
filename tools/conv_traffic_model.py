"""Shared-memory fill traffic (L2 -> SM bytes) of every conv launch of a plan, from the same tiling rules as conv_tc_prepare, next to the
measured per-op device time (bench.py --profile-ops output) and the bounds: L2->SM 5.9 KB/clk (measured: 40 B/clk/SM), HBM, tensor."""
import math, os, re, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiyolov5_b200.models.yolo import Model
from multiyolov5_b200.plan import build_plan
from multiyolov5_b200 import _lib

def choose_tile(W, H):
    best = None
    for t in (128, 64, 32, 16, 8):
        hh = 128 // t
        tiles = math.ceil(W / t) * math.ceil(H / hh)
        if best is None or tiles < best[0]:
            best = (tiles, t, hh)
    return best[1], best[2]

def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "yolov5s_city_seg.yaml"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    perop = {}
    if len(sys.argv) > 3:
        for line in open(sys.argv[3]):
            m = re.match(r"\s*(\d+) kind=\s*(\d+) (\S+)\s+([\d.]+) us", line)
            if m:
                perop[int(m.group(1))] = float(m.group(4))
    model = Model(tag)
    pb = build_plan(model, B, 512, 1024)
    tot = dict(a=0, b=0, t=0.0, flop=0.0, l2=0.0)
    print(f"{'op':>3} {'layer':14s} {'shape':28s} {'mode':6s} {'tiles':>6} {'A MB':>7} {'B MB':>7} {'L2us':>6} {'HBMus':>6} {'TCus':>6} {'meas':>6}")
    for i, o in enumerate(pb.ops):
        if o.kind != _lib.OP_CONV:
            continue
        cv = pb.slots[o.slot].conv
        k, s, d = cv.kernel_size[0], cv.stride[0], cv.dilation[0]
        ci = (cv.in_channels + 15) // 16 * 16
        co16 = (cv.out_channels + 15) // 16 * 16
        Ho, Wo = o.out.h, o.out.w
        if Wo < 8 or Ho < 2 or Wo * Ho < 128:
            continue
        tw, th = choose_tile(Wo, Ho)
        mt = B * math.ceil(Wo / tw) * math.ceil(Ho / th)
        BN = co16 if co16 <= 128 else 128
        if co16 > 128 and co16 % 128:
            BN = max(bn for bn in range(16, 129, 16) if co16 % bn == 0)
        ntn = math.ceil(co16 / BN)
        tiles = mt * ntn
        wbytes = k * k * ci * BN * 2
        ws = ntn == 1 and wbytes <= 40 * 1024
        strip = k == 3 and s == 1 and th == 1 and tw + 2 * d <= 256
        G = 1
        if ntn == 1:
            for g in (4, 2):
                if g * BN <= 128 and tiles // g >= 2 * 296:
                    G = g
                    break
        vr = strip and ws and d == 1 and G >= 2
        if vr:
            a = tiles / G * (G + 2) * (tw + 2) * ci * 2
            mode = "vround"
        elif strip:
            a = tiles * 3 * (tw + 2 * d) * ci * 2
            mode = "strip"
        else:
            a = tiles * k * k * 128 * ci * 2
            mode = "taps"
        b = (min(tiles, 296) if ws else tiles) * wbytes
        flop = 2.0 * B * Ho * Wo * cv.out_channels * cv.in_channels * k * k
        hbm = (B * o.in_.h * o.in_.w * ci + B * Ho * Wo * cv.out_channels) * 2
        l2us = (a + b) / 11.6e12 * 1e6
        t = perop.get(i, float("nan"))
        print(f"{i:3d} {o.tag:14s} {f'{cv.in_channels}->{cv.out_channels} k{k}s{s}d{d} @{Ho}x{Wo}':28s} {mode + ('/ws' if ws else ''):9s} {tiles:6d} {a / 1e6:7.1f} {b / 1e6:7.1f} "
              f"{l2us:6.1f} {hbm / 6.5e12 * 1e6:6.1f} {flop / 1.45e15 * 1e6:6.1f} {t:6.1f}")
        tot["a"] += a; tot["b"] += b; tot["t"] += 0 if t != t else t; tot["flop"] += flop; tot["l2"] += l2us
    print(f"total A {tot['a'] / 1e9:.2f} GB  B {tot['b'] / 1e9:.2f} GB  L2-bound {tot['l2']:.0f} us  tensor-bound {tot['flop'] / 1.45e15 * 1e6:.0f} us  measured {tot['t']:.0f} us")

main()
