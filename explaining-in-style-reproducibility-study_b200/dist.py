"""Multi-GPU plumbing for AttFind: one process per GPU, latents sharded, ONE collective at the end.

The reference's AttFind is single-GPU (``batch_size == 1`` enforced, NB:284-285).  Here rank r sweeps the
contiguous latent shard ``shard_range(N, r, world)`` (no data-path collective: every coord-eval is
independent given minima/maxima, which every rank derives from all N latents), and the per-coordinate effect
scores are all-gathered once (NCCL over NVLink on the GPU box, gloo in the CPU tests).  74 MB at config 3.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from .attfind import shard_range


def init_from_env(backend: str = None) -> tuple:
    """Initialise torch.distributed from the torchrun environment; returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def gather_effects(local_effects: torch.Tensor, n_total: int, world_size: int) -> torch.Tensor:
    """All-gather the per-rank shards ``[n_r, 2, S, 2]`` into ``[N, 2, S, 2]`` (same order as the latents).

    Shards may differ by one row when N % world != 0; they are padded to the largest shard for the collective.
    """
    if world_size == 1:
        return local_effects
    sizes = [shard_range(n_total, r, world_size) for r in range(world_size)]
    n_max = max(hi - lo for lo, hi in sizes)
    tail = tuple(local_effects.shape[1:])
    send = local_effects
    if send.shape[0] != n_max:
        send = torch.zeros((n_max,) + tail, device=local_effects.device, dtype=local_effects.dtype)
        send[: local_effects.shape[0]] = local_effects
    recv = torch.empty((world_size * n_max,) + tail, device=local_effects.device, dtype=local_effects.dtype)
    dist.all_gather_into_tensor(recv, send.contiguous())
    if all(hi - lo == n_max for lo, hi in sizes):
        return recv
    parts = [recv[r * n_max: r * n_max + (hi - lo)] for r, (lo, hi) in enumerate(sizes)]
    return torch.cat(parts, dim=0)
