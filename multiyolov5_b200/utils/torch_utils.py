"""Step-glue helpers the reference's train.py touches around the hot path (reference utils/torch_utils.py)."""
import math
from copy import deepcopy

import torch


def is_parallel(model):
    return type(model) in (torch.nn.parallel.DataParallel, torch.nn.parallel.DistributedDataParallel)


class ModelEMA:
    """reference utils/torch_utils.py:270-304: exponential moving average of everything in the state_dict (parameters AND BN
    buffers), decay ramped by the update count.  The average is ONE multi-tensor lerp over all floating-point entries."""

    def __init__(self, model, decay=0.9999, updates=0):
        self.ema = deepcopy(model.module if is_parallel(model) else model).eval()   # Model.__getstate__ drops the compiled plans
        self.updates = updates
        self.decay = lambda x: decay * (1 - math.exp(-x / 2000))
        for p in self.ema.parameters():
            p.requires_grad_(False)

    def update(self, model):
        with torch.no_grad():
            self.updates += 1
            d = self.decay(self.updates)
            msd = (model.module if is_parallel(model) else model).state_dict()
            mine, theirs = [], []
            for k, v in self.ema.state_dict().items():
                if v.dtype.is_floating_point:
                    mine.append(v)
                    theirs.append(msd[k].detach())
            torch._foreach_mul_(mine, d)
            torch._foreach_add_(mine, theirs, alpha=1.0 - d)
        if hasattr(self.ema, "invalidate_weights"):
            self.ema.invalidate_weights()      # the EMA copy's packed fp16 weights are stale now

    def update_attr(self, model, include=(), exclude=("process_group", "reducer")):
        for k, v in model.__dict__.items():
            if (len(include) and k not in include) or k.startswith("_") or k in exclude:
                continue
            setattr(self.ema, k, v)
