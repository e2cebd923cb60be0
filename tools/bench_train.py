"""Training-step measurement (BASELINE.json configs[3]): yolov5s_city_seg.yaml (PSP head), per GPU
4 det images + 4 seg images of 3x512x1024, 20 boxes per det image, seg labels randint(-1,19); one step = det forward/backward +
seg forward/backward + ONE flat-gradient all-reduce + SGD (reference train.py:363-401).  Not the headline metric (bench.py is);
this is the a13 row's number and its breakdown.

    python tools/bench_train.py [--steps 10 --warmup 3 --batch 4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/bench_train.py

Prints ONE JSON line on rank 0.  Timing: CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
H, W = 512, 1024
GFLOP_FWD_PER_IMG = 29.70          # SURVEY.md section 8d (Conv2d MACs x 2, s/PSP)


def make_workload(world, rank, B):
    """model, Trainer and NROT rotating synthetic batches of BASELINE.json configs[3]'s per-GPU slice"""
    from multiyolov5_b200 import synth
    from multiyolov5_b200.models.yolo import Model
    from multiyolov5_b200.train import Trainer, scale_hyp
    yml, tag = "yolov5s_city_seg.yaml", "s_psp"
    cfg = synth.load_cfg(yml)
    sd = synth.synth_state_dict(synth.load_manifest(tag), cfg, seed=1, gain=1.0)
    model = Model(yml)
    model.load_state_dict(sd)
    model.cuda().train()
    hyp = dict(lr0=0.0015, momentum=0.937, weight_decay=5e-4, box=0.05, cls=0.5, cls_pw=1.0, obj=1.0, obj_pw=1.0, anchor_t=4.0, fl_gamma=0.0)
    hyp = scale_hyp(hyp, nl=3, nc=cfg["nc"], imgsz=W, total_batch_size=B * world)      # data/hyp.scratch.yaml values
    tr = Trainer(model, hyp, batch_size=B, world_size=world, rank=rank if world > 1 else -1, init_scale=2.0 ** 10)
    gen = torch.Generator(device="cuda").manual_seed(77 + rank)
    NROT = 3
    imgs = [torch.rand((B, 3, H, W), device="cuda", generator=gen) for _ in range(NROT)]
    segimgs = [torch.rand((B, 3, H, W), device="cuda", generator=gen) for _ in range(NROT)]
    rs = np.random.RandomState(5 + rank)
    tg = []
    for _ in range(NROT):
        t = np.zeros((20 * B, 6), np.float32)
        t[:, 0] = np.repeat(np.arange(B), 20)
        t[:, 1] = rs.randint(0, 10, 20 * B)
        t[:, 2:4] = rs.uniform(0.1, 0.9, (20 * B, 2))
        t[:, 4:6] = rs.uniform(0.02, 0.22, (20 * B, 2))
        tg.append(torch.from_numpy(t).cuda())
    masks = [torch.randint(-1, 19, (B, H, W), device="cuda", generator=gen) for _ in range(NROT)]
    return dict(yml=yml, cfg=cfg, sd=sd, hyp=hyp, model=model, tr=tr, imgs=imgs, segimgs=segimgs, tg=tg, masks=masks, NROT=NROT)


def timed_steps(step, steps, world):
    """EXACTLY `steps` calls bracketed by barrier + synchronize, CUDA events on the launching stream, max over ranks (ms)"""
    import torch.distributed as dist
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _ar_path():
    from multiyolov5_b200 import parallel
    return parallel.LAST_ALLREDUCE_PATH


def train_record(world, rank, steps=10, warmup=3, B=4, min_seconds=1.0, with_reference_gpu=False):
    """the `train` sub-record of bench.py's JSON line (BASELINE.json configs[3]): step time, whole-job images/s, the all-reduce's own time,
    and a >= min_seconds sustained run next to the short one"""
    wl = make_workload(world, rank, B)
    tr, NROT = wl["tr"], wl["NROT"]

    def step(i):
        k = i % NROT
        return tr.step(wl["imgs"][k], wl["tg"][k], wl["segimgs"][k], wl["masks"][k])

    for i in range(max(warmup, 3)):
        step(i)
    ms = timed_steps(step, steps, world)
    n_long = max(steps, int(min_seconds * 1e3 / (ms / steps)) + 1)
    ms_long = timed_steps(step, n_long, world)
    # the exchange on its own: the flat fp32 gradient buffer through ONE all-reduce (what every optimiser step contains)
    ar_ms = 0.0
    if world > 1:
        def ar(i):
            tr.allreduce()
        ar(0)
        ar_ms = timed_steps(ar, 20, world) / 20
        tr.flat.grad.zero_()
    flops = 3.0 * GFLOP_FWD_PER_IMG * 1e9 * 2 * B * world
    rec = {"workload": f"{wl['yml']} train step (det fwd+bwd, seg fwd+bwd, ONE flat-gradient all-reduce, fused SGD; reference train.py:363-401), per GPU "
                       f"{B} det + {B} seg images 3x{H}x{W}, 20 boxes/img, global batch {2 * B * world}",
           "n_gpus": world, "steps": steps, "ms_per_step": ms / steps, "images_per_s": 2 * B * world * steps / (ms * 1e-3),
           "sustained": {"steps": n_long, "seconds": ms_long * 1e-3, "ms_per_step": ms_long / n_long,
                         "images_per_s": 2 * B * world * n_long / (ms_long * 1e-3)},
           "allreduce_ms": ar_ms, "allreduce_mbytes": tr.flat.n * 4 / 1e6, "allreduce_share_of_step": ar_ms / (ms / steps) if world > 1 else 0.0,
           "collective": ("all-reduce (sum) of the flat fp32 gradient buffer, averaging folded into the optimiser; route: " + _ar_path()) if world > 1 else "none (1 GPU)",
           "conv_tflops_algorithmic": flops / (ms / steps * 1e-3) / 1e12, "dtype": "f16 storage / f32 accumulate, fp32 master weights",
           "loss_scale": float(tr.scale)}
    if with_reference_gpu and rank == 0:
        rec["reference_gpu"] = reference_gpu_train(wl, B)
    del wl, tr
    torch.cuda.empty_cache()
    return rec


def reference_gpu_train(wl, B, steps=6):
    """torch autocast + cuDNN step of the same graph on this GPU (reference train.py:363-401): the reference's own GPU path as a baseline"""
    try:
        from oracle.gpu_pipeline import TorchAutocastTrainStep
        from multiyolov5_b200.models.yolo import Model
        shell = Model(wl["yml"]).cuda()
        ref = TorchAutocastTrainStep(wl["cfg"], wl["sd"], shell, wl["hyp"], B)
        NROT = wl["NROT"]

        def step(i):
            k = i % NROT
            ref.step(wl["imgs"][k], wl["tg"][k], wl["segimgs"][k], wl["masks"][k])
        for i in range(3):
            step(i)
        ms = timed_steps(step, steps, 1)
        return {"ms_per_step": ms / steps, "images_per_s": 2 * B * steps / (ms * 1e-3), "kind": "torch autocast(fp16) + cuDNN (cudnn.benchmark) + "
                "GradScaler + torch SGD on the restated train graph, 1 GPU", "steps": steps}
    except Exception as e:   # a baseline must never take the bench down
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=4, help="det images per GPU (= seg images per GPU)")
    ap.add_argument("--kernel-table", default=None, help="write a per-kernel device-time table (torch.profiler / CUPTI, 2 steps) to this file")
    ap.add_argument("--per-step", action="store_true", help="print synchronised wall time of every step to stderr (diagnostic)")
    ap.add_argument("--breakdown", action="store_true", help="also time the phases of one step separately (extra synchronisation)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    wl = make_workload(world, rank, args.batch)
    yml, model, tr, NROT, imgs, segimgs, tg, masks = wl["yml"], wl["model"], wl["tr"], wl["NROT"], wl["imgs"], wl["segimgs"], wl["tg"], wl["masks"]
    B = args.batch

    def step(i):
        k = i % NROT
        return tr.step(imgs[k], tg[k], segimgs[k], masks[k])

    if args.per_step:
        import time
        for i in range(args.steps):
            torch.cuda.synchronize(); t0 = time.time()
            tr.backward_det(imgs[i % NROT], tg[i % NROT]); torch.cuda.synchronize(); t1 = time.time()
            tr.backward_seg(segimgs[i % NROT], masks[i % NROT]); torch.cuda.synchronize(); t2 = time.time()
            tr.optimizer_step(); torch.cuda.synchronize(); t3 = time.time()
            print("step %2d: det %.1f ms  seg %.1f ms  opt %.1f ms" % (i, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3), file=sys.stderr)
    for i in range(max(args.warmup, 3)):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        items, segloss = step(i)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    breakdown = None
    if args.breakdown:
        def ev():
            e = torch.cuda.Event(enable_timing=True); e.record(); return e
        acc = np.zeros(7)
        reps = 3
        for r in range(reps):
            k = r % NROT
            t0 = ev(); pred = model(imgs[k])
            t1 = ev(); loss, _ = tr.compute_loss(pred[0], tg[k]); loss = loss * tr.detgain * tr.scale
            t2 = ev(); loss.backward()
            t3 = ev(); pred = model(segimgs[k])
            t4 = ev(); sl = tr.compute_seg_loss(pred[1], masks[k]) * B * tr.seggain * tr.scale
            t5 = ev(); sl.backward()
            t6 = ev(); tr.optimizer_step()
            t7 = ev()
            torch.cuda.synchronize()
            acc += np.array([a.elapsed_time(b) for a, b in zip((t0, t1, t2, t3, t4, t5, t6), (t1, t2, t3, t4, t5, t6, t7))])
        names = ["det_forward", "det_loss", "det_loss_bwd+net_bwd", "seg_forward", "seg_loss", "seg_loss_bwd+net_bwd", "allreduce+sgd"]
        breakdown = {n: round(v / reps, 3) for n, v in zip(names, acc)}
    if args.kernel_table and rank == 0:
        import time
        from torch.profiler import profile, ProfilerActivity
        torch.cuda.synchronize()
        w0 = time.time()
        for i in range(2):
            step(i)
        torch.cuda.synchronize()
        wall = (time.time() - w0) / 2
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            for i in range(2):
                step(i)
            torch.cuda.synchronize()
        with open(args.kernel_table, "w") as f:
            f.write("wall ms per step (no profiler): %.2f\n" % (wall * 1e3))
            f.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=70))
    if rank == 0:
        n_img = 2 * B * world * args.steps
        flops = 3.0 * GFLOP_FWD_PER_IMG * 1e9 * 2 * B * world           # per step, all ranks (fwd + dgrad + wgrad)
        line = {"metric": "train images/sec @1024x512 (det pass + seg pass + allreduce + SGD)", "value": n_img / (ms * 1e-3),
                "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "dtype": "f16 storage / f32 accumulate, fp32 master weights", "data": "synthetic",
                "config": {"workload": f"{yml} train step, per GPU {B} det + {B} seg images 3x{H}x{W}, 20 boxes/img", "global_batch": 2 * B * world,
                           "parallelism": f"dp{world} (one flat-gradient all-reduce of {tr.flat.n * 4 / 1e6:.1f} MB per step)"},
                "conv_tflops_algorithmic": flops / (ms / args.steps * 1e-3) / 1e12, "loss_scale": float(tr.scale),
                "last_losses": {"det": [float(v) for v in items], "seg": float(segloss)}, "breakdown_ms": breakdown}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
