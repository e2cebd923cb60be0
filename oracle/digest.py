"""TEST INFRASTRUCTURE - compact digests shared by oracle/make_golden.py (which writes them from the reference) and the tests (which
recompute them from the oracle / the CUDA path): fixed random cotangents for the train-mode probe loss and a per-tensor gradient digest."""
import numpy as np
import torch


def train_probe_tensors(shapes_raw, shapes_seg, seed=11):
    """the fixed random cotangents R_i, S_k of the scalar L = sum_i <x_i, R_i> + sum_k <seg_k, S_k> (same generator in the tests)"""
    rs = np.random.RandomState(seed)
    return ([torch.from_numpy(rs.normal(0, 1, s).astype(np.float32)) for s in shapes_raw],
            [torch.from_numpy(rs.normal(0, 1, s).astype(np.float32) * 0.05) for s in shapes_seg])


def grad_digest(g: torch.Tensor):
    """[L2 norm, sum, 32 leading values, 32 strided values] of one gradient tensor"""
    f = g.detach().reshape(-1).double()
    idx = torch.linspace(0, f.numel() - 1, 32).long()
    lead = torch.zeros(32, dtype=torch.float64)
    lead[:min(32, f.numel())] = f[:32]
    return torch.cat([f.norm().view(1), f.sum().view(1), lead, f[idx]]).numpy()
