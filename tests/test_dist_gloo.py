"""CPU, world_size 2 over gloo: the N>1 host logic of bench.py / the replica runner (sharding of a global batch across ranks,
barrier + max-over-ranks timing reduction, whole-job images/s aggregation).  Inference has no data-path collective; training has
one: the flat-gradient all-reduce, checked here against the reference's DDP arithmetic (det loss x world_size, gradients averaged)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from multiyolov5_b200.parallel import aggregate_throughput, shard_range
    lo, hi = shard_range(37, world, rank)                         # ragged global batch
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([hi - lo]))
    ms_local = 10.0 * (rank + 1)                                  # rank 1 is the slow one
    val = aggregate_throughput(images_local=hi - lo, ms_local=ms_local)
    dist.barrier()
    if rank == 0:
        ret["sizes"] = [int(s) for s in sizes]
        ret["value"] = val
    dist.destroy_process_group()


def test_two_rank_replicas_gloo():
    mgr = mp.Manager(); ret = mgr.dict(); port = _free_port()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert sum(ret["sizes"]) == 37 and max(ret["sizes"]) - min(ret["sizes"]) <= 1
    assert abs(ret["value"] - 37 / 0.020) < 1e-6                 # total images / max-over-ranks time


def _grad_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from multiyolov5_b200.parallel import allreduce_flat_grads
    # a tiny "model": flat parameter vector w, per-rank det / seg losses  L_det_r = a_r . w ,  L_seg_r = b_r . w
    n = 1001
    g = torch.Generator().manual_seed(100 + rank)
    a, b = torch.randn(n, generator=g), torch.randn(n, generator=g)
    scale = 1024.0
    flat = torch.zeros(n)
    flat += a * world * scale            # det loss is multiplied by world_size before backward (train.py:367-368)
    flat += b * scale                    # seg pass accumulates into the same buffer (train.py:392)
    w = allreduce_flat_grads(flat)
    inv = 1.0 / (scale * w)              # what Trainer.optimizer_step hands to myolo_sgd_step
    if rank == 0:
        ret["grad"] = (flat * inv).tolist()
        ret["world"] = w
    dist.destroy_process_group()


def test_flat_gradient_allreduce_matches_ddp_arithmetic():
    mgr = mp.Manager(); ret = mgr.dict(); port = _free_port()
    mp.spawn(_grad_worker, args=(2, port, ret), nprocs=2, join=True)
    n = 1001
    ab = []
    for r in range(2):
        g = torch.Generator().manual_seed(100 + r)
        ab.append((torch.randn(n, generator=g), torch.randn(n, generator=g)))
    # DDP: gradient = mean over ranks of d(per-rank loss)/dw, per-rank loss = world*L_det + L_seg
    want = sum(2 * a + b for a, b in ab) / 2
    assert ret["world"] == 2
    assert torch.allclose(torch.tensor(ret["grad"]), want, rtol=1e-5, atol=1e-5)
