"""Ground-truth device timeline of ONE graph-replayed step (CUPTI through torch.profiler): every kernel with start, duration, stream, and the
critical path statistics (busy time per stream, gaps).  Usage: python tools/kernel_trace.py [--cfg s_psp] [--batch 16] [--post]
Writes the per-kernel list to stdout (one step), sorted by start time."""
import argparse, os, sys, json, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from multiyolov5_b200.models.yolo import Model
from multiyolov5_b200.utils.general import non_max_suppression, seg_argmax
ap = argparse.ArgumentParser()
ap.add_argument("--cfg", default="s_psp")
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--post", action="store_true", help="include NMS + argmax")
args = ap.parse_args()
yml, cfg, sd = bench.make_weights(args.cfg)
model = Model(yml); model.load_state_dict(sd); model.cuda().eval().half()
xs = [torch.rand(args.batch, 3, 512, 1024, device="cuda").half() for _ in range(3)]
def step(i):
    (z, _), seg = model(xs[i % 3])
    if args.post:
        non_max_suppression(z, 0.25, 0.45, return_padded=True); seg_argmax(seg, (512, 1024))
for i in range(5):
    step(i)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(3):
        step(i)
        torch.cuda.synchronize()
f = tempfile.mktemp(suffix=".json")
prof.export_chrome_trace(f)
tr = json.load(open(f))
ks = [e for e in tr["traceEvents"] if e.get("cat") == "kernel"]
ks.sort(key=lambda e: e["ts"])
# split into the 3 steps by large gaps (synchronize between them)
steps, cur = [], [ks[0]]
for a, b in zip(ks, ks[1:]):
    if b["ts"] - (a["ts"] + a["dur"]) > 150:
        steps.append(cur); cur = []
    cur.append(b)
steps.append(cur)
st = steps[-1]
t0 = st[0]["ts"]
end = max(e["ts"] + e["dur"] for e in st)
print(f"# one step: {len(st)} kernels, wall {end - t0:.1f} us, sum of kernel durations {sum(e['dur'] for e in st):.1f} us")
busy = {}
for e in st:
    busy[e["args"].get("stream")] = busy.get(e["args"].get("stream"), 0) + e["dur"]
print("# busy us per stream:", {k: round(v, 1) for k, v in busy.items()})
# union coverage: time during which at least one kernel runs
iv = sorted((e["ts"], e["ts"] + e["dur"]) for e in st)
cov, cs, ce = 0.0, iv[0][0], iv[0][1]
for a, b in iv[1:]:
    if a > ce:
        cov += ce - cs; cs, ce = a, b
    else:
        ce = max(ce, b)
cov += ce - cs
print(f"# union of kernel intervals {cov:.1f} us  -> idle (no kernel running) {end - t0 - cov:.1f} us")
for e in st:
    name = e["name"].replace("myolo::", "").replace("void ", "")[:60]
    print(f"{e['ts'] - t0:9.1f} {e['dur']:8.1f} s{e['args'].get('stream')} grid{e['args'].get('grid')} {name}")
