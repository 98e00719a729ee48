"""Multi-GPU plumbing: one process per GPU (torchrun), batch sharding, and the single
collective of the path -- the all-reduce of calibration min/max statistics.

Inference itself has no collective: per-tensor scales are static after calibration, so
disjoint batch shards are independent.
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np
import torch
import torch.distributed as dist


def world() -> tuple[int, int]:
    """(rank, world_size) of the default process group, (0, 1) if not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def init_from_env(backend: Optional[str] = None) -> tuple[int, int, int]:
    """Initialise torch.distributed from torchrun's environment (RANK / LOCAL_RANK /
    WORLD_SIZE / MASTER_*). Returns (rank, local_rank, world_size); no-op for a single process."""
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if ws > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        import datetime
        # a rank that dies or diverges must fail the job in minutes, not after the 10-minute default watchdog
        tmo = datetime.timedelta(seconds=int(os.environ.get("NQ_DIST_TIMEOUT_S", "240")))
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend=backend, rank=rank, world_size=ws, device_id=torch.device("cuda", local), timeout=tmo)
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=ws, timeout=tmo)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    return rank, local, ws


def shard_bounds(n: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous, balanced [lo, hi) slice of n items for `rank` (first n % ws ranks get one more)."""
    base, extra = divmod(n, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(arrays: list, rank: int, world_size: int) -> list:
    """Slice every input along axis 0 for this rank (images are the independent unit)."""
    out = []
    for a in arrays:
        lo, hi = shard_bounds(a.shape[0], rank, world_size)
        out.append(a[lo:hi])
    return out


def allreduce_minmax(mm: torch.Tensor, group=None) -> torch.Tensor:
    """mm[n, 2] = (min, max) per value on this rank -> global (min, max) on every rank (`group=False`: no exchange,
    this rank calibrates on its own -- for work that only one rank of a job performs).

    One all-reduce(MAX) over the packed vector [max_0.., -min_0..]; min/max are exact under
    any reduction order, so all ranks end up with bit-identical statistics (and therefore
    bit-identical quantization parameters) regardless of how the calibration batch was split.
    """
    if group is False:                                   # explicitly local calibration inside a multi-rank job
        return mm
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return mm
    packed = torch.cat([mm[:, 1], -mm[:, 0]]).contiguous()
    dist.all_reduce(packed, op=dist.ReduceOp.MAX, group=group)
    n = mm.shape[0]
    return torch.stack([-packed[n:], packed[:n]], dim=1)


def gather_outputs(local: np.ndarray, group=None) -> Optional[np.ndarray]:
    """Concatenate per-rank outputs along axis 0 on rank 0 (host side; outputs are tiny)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, local, group=group)
    return np.concatenate(parts, axis=0) if dist.get_rank(group) == 0 else None
