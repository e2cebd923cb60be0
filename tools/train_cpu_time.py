"""Host time of Trainer.step (the enqueue cost, no synchronisation) next to the device time per step: is the train step host-bound?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools.bench_train import make_workload
wl = make_workload(1, 0, 4)
tr, NROT = wl["tr"], wl["NROT"]
def step(i):
    k = i % NROT
    return tr.step(wl["imgs"][k], wl["tg"][k], wl["segimgs"][k], wl["masks"][k])
for i in range(8):
    step(i)
torch.cuda.synchronize()
N = 40
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
t0 = time.perf_counter()
per = []
for i in range(N):
    a = time.perf_counter()
    step(i)
    per.append(time.perf_counter() - a)
t1 = time.perf_counter()
e1.record()
torch.cuda.synchronize()
t2 = time.perf_counter()
per.sort()
print(f"host enqueue per step: mean {1e3 * (t1 - t0) / N:.2f} ms, median {1e3 * per[N // 2]:.2f} ms, min {1e3 * per[0]:.2f} ms; "
      f"device per step {e0.elapsed_time(e1) / N:.2f} ms; host finished {1e3 * (t2 - t1):.1f} ms before the device")
