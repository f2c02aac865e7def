"""Cohort sharding over the GPUs of one box: patients are independent,
so a shard never talks to another one on the data path.  The only cross-rank traffic is the
reduction of timings / counters for reporting and the final gather of fitted hyper-parameters,
both tiny; they use torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def lpt_assign(sizes, world):
    """Longest-processing-time-first: returns shard index per patient, load ~ n^3."""
    sizes = np.asarray(sizes, dtype=np.float64)
    order = np.argsort(-sizes, kind="stable")
    load = np.zeros(world)
    out = np.empty(len(sizes), dtype=np.int64)
    for k in order:
        tgt = int(np.argmin(load))
        load[tgt] += sizes[k] ** 3
        out[k] = tgt
    return out


def reduce_report(dist, device, ms_local, evals_local):
    """(max ms over ranks, total evaluations over ranks); dist may be None for world 1."""
    import torch
    t = torch.tensor([ms_local, -float(evals_local)], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        mx = t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = t.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        return float(mx[0]), -float(sm[1])
    return float(t[0]), float(evals_local)


def gather_theta(dist, device, theta_local, owner_index, n_total):
    """Final gather of fitted hyper-parameters: every rank contributes the rows it owns
    (owner_index = global patient indices of theta_local's rows); returns (n_total, P)."""
    import torch
    P = theta_local.shape[1]
    full = torch.zeros((n_total, P), dtype=torch.float64, device=device)
    if len(owner_index):
        full[torch.as_tensor(owner_index, device=device)] = torch.as_tensor(theta_local, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(full, op=dist.ReduceOp.SUM)  # disjoint rows: a sum is a gather
    return full.cpu().numpy()
