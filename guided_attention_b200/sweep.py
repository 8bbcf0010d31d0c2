"""Multi-GPU seed sweep (BASELINE config 5): seeds are independent units (reference run.py:97 loops them sequentially on
one GPU), so they shard one process per GPU with NO collective on the hot path; the only communication is one gather of
the final latents at the end (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

import os
from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist


def dist_env():
    """(rank, world_size, local_rank) from the torchrun environment; (0, 1, 0) when not launched by torchrun."""
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0)))


def init_distributed(backend: Optional[str] = None):
    rank, world, local_rank = dist_env()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        n_dev = torch.cuda.device_count() if torch.cuda.is_available() else 0
        if backend is None:
            # one process per GPU -> NCCL over NVLink.  More ranks than GPUs (several processes sharing a device, e.g.
            # the 2-rank equality test on a 1-GPU box) cannot form an NCCL communicator: the final gather then goes
            # through gloo on host copies (it is 32 KB per image and off the hot path either way).
            backend = "nccl" if n_dev >= world else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kw["device_id"] = torch.device("cuda", local_rank)
        elif n_dev > 0:
            torch.cuda.set_device(local_rank % n_dev)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local_rank


def local_device(local_rank: int) -> torch.device:
    """The CUDA device of this rank: its own GPU, or a shared one when there are more ranks than GPUs."""
    n_dev = torch.cuda.device_count()
    return torch.device("cuda", local_rank % n_dev) if n_dev > 0 else torch.device("cpu")


def shard_seeds(seeds: Sequence[int], rank: int, world: int) -> List[int]:
    """Static round-robin: seed index i goes to rank i % world."""
    return [s for i, s in enumerate(seeds) if i % world == rank]


def gather_results(local: torch.Tensor, n_total: int, rank: int, world: int) -> Optional[torch.Tensor]:
    """All ranks pass (n_local, ...) results for `shard_seeds` order; every rank gets (n_total, ...) back in seed order.
    One all_gather on equal-sized, zero-padded chunks."""
    if world == 1:
        return local
    home = local.device
    if dist.get_backend() == "gloo" and local.is_cuda:
        local = local.cpu()
    per = (n_total + world - 1) // world
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    chunks = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(chunks, pad)
    out = torch.empty((n_total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for r in range(world):
        idx = list(range(r, n_total, world))
        out[idx] = chunks[r][: len(idx)]
    return out.to(home)


def _agree_on_template(local_template: Optional[torch.Tensor], rank: int, world: int, device=None):
    """(shape, dtype) of one result, agreed across ranks: the lowest rank that has a result describes it to everybody.
    Every rank takes part (no early return before a collective); returns None on every rank when nobody has a result."""
    if world == 1:
        return None if local_template is None else (tuple(local_template.shape), local_template.dtype)
    mine = None if local_template is None else (tuple(local_template.shape), str(local_template.dtype).split(".")[-1])
    everyone = [None] * world
    dist.all_gather_object(everyone, mine)
    desc = next((d for d in everyone if d is not None), None)
    return None if desc is None else (tuple(desc[0]), getattr(torch, desc[1]))


def run_seed_sweep(generate: Callable[[int], torch.Tensor], seeds: Sequence[int], rank: int, world: int,
                   gather: bool = True, device=None):
    """`generate(seed) -> (C, H, W) tensor`.  Returns (all results in seed order or None, this rank's results, failures).
    A failing seed is reported and skipped (zeros in its slot); the other seeds continue.  A rank whose shard is empty
    or whose seeds all failed still joins the final gather with a zero chunk, so per-seed failures never turn into a
    job-wide hang; when no rank has any result every rank returns (None, [], failures)."""
    mine = shard_seeds(seeds, rank, world)
    results, failures = [], []
    for s in mine:
        try:
            results.append(generate(s))
        except Exception as e:  # per-seed failure isolation
            failures.append((s, repr(e)))
            results.append(None)
    template = next((r for r in results if r is not None), None)
    desc = _agree_on_template(template, rank, world) if gather else (
        None if template is None else (tuple(template.shape), template.dtype))
    if desc is None:
        return None, [], failures
    shape, dtype = desc
    dev = template.device if template is not None else (device if device is not None else (
        torch.device("cuda", torch.cuda.current_device()) if dist.is_initialized() and dist.get_backend() == "nccl"
        else torch.device("cpu")))
    zero = torch.zeros(shape, dtype=dtype, device=dev)
    local = torch.stack([r if r is not None else zero for r in results]) if results else zero.new_zeros((0,) + shape)
    full = gather_results(local, len(seeds), rank, world) if gather else None
    return full, local, failures
