"""CPU, world_size 2 over gloo: the N>1 host logic of bench.py / the replica runner (sharding of a global batch across ranks,
barrier + max-over-ranks timing reduction, whole-job images/s aggregation).  Inference has no data-path collective."""
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
