"""Multi-GPU driver of the path: scene pairs are independent (the reference iterates them serially,
registration_node.py:587-588), so they shard across ranks with NO data-path collective; the only exchange is one
all-gather of the per-pair 4x4 transforms (+ a small stats row) after the local loop.

One process per GPU (torchrun); NCCL over NVLink on the device, gloo in the CPU tests.  The payload is
P x (16 + 4) float64 = 160 B per pair: latency-bound, so it is a plain ``all_gather_into_tensor`` -- there is no
compute step to fuse with."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous block partition (keeps the scans of one map on one rank): returns (start, stop, per_rank)."""
    per = (n_items + world - 1) // world
    start = min(rank * per, n_items)
    return start, min(start + per, n_items), per


@dataclass
class GatheredResults:
    T: np.ndarray        # (P, 4, 4) float64, in global pair order
    stats: np.ndarray    # (P, 4) float64: fitness, rmse, n_corr, best_hyp


def register_pairs(pairs: Sequence, solve_fn: Optional[Callable] = None, *, device: Optional[torch.device] = None,
                   group=None, **register_kwargs) -> GatheredResults:
    """Solve this rank's shard of ``pairs`` and all-gather every rank's transforms.

    pairs[i] = (source_pcd, target_pcd, src_feats, tgt_feats); every rank passes the same global list (or any sequence
    that can be indexed by global pair id -- only the local shard is touched).  ``solve_fn(pair) -> (T 4x4, fitness, rmse,
    n_corr, best_hyp)`` defaults to ``vfm_registration_b200.register``."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = len(pairs)
    start, stop, per = shard_range(n, rank, world)
    if solve_fn is None:
        from . import api

        def solve_fn(pair):
            r = api.register(*pair, **register_kwargs)
            return r.T, r.fitness, r.rmse, len(r.corr), r.best_hyp
    local = np.zeros((per, 20), dtype=np.float64)
    for k, p in enumerate(range(start, stop)):
        t, fit, rmse, n_corr, best = solve_fn(pairs[p])
        local[k, :16] = np.asarray(t, dtype=np.float64).reshape(16)
        local[k, 16:] = (fit, rmse, n_corr, best)
    if world == 1:
        allr = local
    else:
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        send = torch.from_numpy(local).to(device)
        recv = torch.empty((world * per, 20), dtype=torch.float64, device=device)
        dist.all_gather_into_tensor(recv, send, group=group)
        allr = recv.cpu().numpy()
    allr = allr[:n] if world == 1 else np.concatenate([allr[r * per:r * per + max(0, min(per, n - r * per))] for r in range(world)])
    return GatheredResults(T=allr[:, :16].reshape(-1, 4, 4).copy(), stats=allr[:, 16:].copy())
