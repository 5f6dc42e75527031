"""Data-parallel serving: one process per GPU, requests sharded statically,
weights broadcast once.

The reference is single-process (no distributed code at
all).  Requests never interact - each owns its page lists, pools are per model
instance - so the path shards as independent units: request ``i`` goes to rank
``i mod world`` and every rank runs the unchanged scheduler on its own replica.
The only collective is the start-up ``broadcast`` of the packed weights from
rank 0 (NCCL over NVLink/NVSwitch on GPUs, gloo in CPU tests) plus a final
reduction of timings for the report; there is no per-step communication, hence
nothing to fuse a collective into.
"""

from __future__ import annotations

import os

import torch
import torch.distributed as dist

from .synthetic import named_tensors, synthetic_qwen3


def init_distributed(device_type: str | None = None) -> tuple[int, int, torch.device]:
    """(rank, world, device) from the torchrun environment; world == 1 without it."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    use_cuda = (device_type or ("cuda" if torch.cuda.is_available() else "cpu")) == "cuda"
    device = torch.device("cuda", local) if use_cuda else torch.device("cpu")
    if use_cuda:
        torch.cuda.set_device(device)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl" if use_cuda else "gloo", rank=rank, world_size=world)
    return rank, world, device


def shard(items: list, rank: int, world: int) -> list:
    """Static round-robin: request i -> rank i mod world."""
    return [item for index, item in enumerate(items) if index % world == rank]


def shard_indices(count: int, rank: int, world: int) -> list[int]:
    return list(range(rank, count, world))


def broadcast_model(model_ns, src: int = 0) -> int:
    """Broadcast every tensor of the ``mlx_model`` namespace from ``src`` in a
    fixed order; returns the number of bytes sent.  uint32 words travel as
    int32 views (same bits)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return 0
    sent = 0
    for _, tensor in named_tensors(model_ns):
        wire = tensor.view(torch.int32) if tensor.dtype == torch.uint32 else tensor
        dist.broadcast(wire, src=src)
        sent += wire.numel() * wire.element_size()
    return sent


def replicated_model(name_or_dims, seed: int, rank: int, device, **overrides):
    """Rank 0 draws the seeded weights and broadcasts them; the other ranks only
    allocate.  Every rank ends up with identical bytes on its own GPU."""
    if rank == 0:
        model_ns = synthetic_qwen3(name_or_dims, seed=seed, device=device, **overrides)
    else:
        model_ns = synthetic_qwen3(name_or_dims, seed=seed, device=device, empty=True, **overrides)
    nbytes = broadcast_model(model_ns, src=0)
    return model_ns, nbytes


def max_over_ranks(value: float, device) -> float:
    """Timings are reported as the maximum over ranks."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def sum_over_ranks(value: float, device) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t[0])


def barrier(device) -> None:
    if dist.is_initialized() and dist.get_world_size() > 1:
        if device.type == "cuda":
            dist.barrier(device_ids=[device.index])
        else:
            dist.barrier()
