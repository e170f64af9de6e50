"""Multi-GPU plumbing: one process per GPU (torchrun), pure data parallelism over utterances.

The eval forward has no exchange step (utterances are independent), so the only
collectives are the barrier and the max-over-ranks of the device time used for reporting."""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def env() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (1-process defaults)."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init(backend: str = "nccl", device=None) -> bool:
    """Initialise the default process group when WORLD_SIZE > 1.  Returns True if distributed."""
    rank, _, world = env()
    if world <= 1:
        return False
    if not dist.is_initialized():
        kw = {}
        if backend == "nccl" and device is not None:
            kw["device_id"] = device
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return True


def shard_seed(base_seed: int) -> int:
    """Every rank draws its own synthetic utterances (weak scaling: per-GPU batch is fixed)."""
    return base_seed + env()[0]


def global_batch(per_gpu_batch: int) -> int:
    return per_gpu_batch * env()[2]


def max_over_ranks(value: float, device="cpu") -> float:
    """Max of a host scalar over all ranks (device-timed milliseconds -> whole-job time)."""
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def remaining_spans(done, n: int):
    """Complement of the (possibly unordered) half-open spans `done` inside [0, n): what the trainer still has
    to all-reduce after the backward sent the per-layer spans (trainer.FlatAdamTrainer.allreduce_grads)."""
    pos, rest = 0, []
    for lo, hi in sorted(done):
        if lo > pos:
            rest.append((pos, lo))
        pos = max(pos, hi)
    if pos < n:
        rest.append((pos, n))
    return rest
