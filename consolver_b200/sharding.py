"""Multi-GPU plumbing of the sampling path: one process per GPU (torchrun), each rank owns a contiguous shard
of the prompt/seed list and runs its own scheduler on its own stream — the path has NO data-path collective
(every preview is independent; coefficients are per-sample scalars).

Partitioning rules restated from the reference:
  * SD   (gen_ppo.py:349-357): floor(len/world) prompts per rank, the LAST rank also takes the remainder;
         per-rank batches of `batch_size`, generator seed = seed + batch_idx (gen_ppo.py:253-260);
  * FLUX (edit_ppo/generate_ours.py:176-177): chunks of ceil(len/world), trailing ranks may be empty.
The only exchange is the result gather (counts / timing / checksums), done once per job through
torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import math
import os
from typing import Iterator, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds_sd(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    per = n_items // world
    start = rank * per
    end = start + per if rank != world - 1 else n_items
    return start, end


def shard_bounds_flux(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    chunk = math.ceil(n_items / world) if n_items else 0
    start = min(rank * chunk, n_items)
    return start, min(start + chunk, n_items)


def batches(items: Sequence, batch_size: int, seed: int) -> Iterator[Tuple[int, Sequence, int]]:
    """(batch_idx, items, generator seed) — seed + batch_idx restarts on every rank exactly as the reference's
    per-device loop does (gen_ppo.py:253-260)."""
    total = len(items) // batch_size + (1 if len(items) % batch_size else 0)
    for b in range(total):
        yield b, items[b * batch_size:(b + 1) * batch_size], seed + b


def init_from_env(backend: str = None) -> Tuple[int, int, int]:
    """(rank, world, local_rank) from the torchrun environment; initialises the process group when world > 1."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kw)
    return rank, world, local


def bind_to_gpu_numa(local_rank: int) -> Optional[int]:
    """Pin this process (and therefore its pinned host allocations, first-touch) to the CPUs NVML reports as local
    to GPU `local_rank`.  With 8 ranks sharing two sockets this keeps every rank's host<->device copies on the
    PCIe root of its own socket.  Returns the number of CPUs bound, or None when NVML / affinity is unavailable."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:  # noqa: BLE001
        return None


def gather_job_stats(n_done: int, elapsed_s: float, checksum: float, device=None) -> dict:
    """max-over-ranks time, total units and a checksum-of-checksums (the reference's only multi-GPU self-check is
    a per-rank parameter checksum print, train_ppo.py:452-455).  Safe to call with world == 1."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(total=n_done, max_elapsed_s=elapsed_s, checksum=checksum, world=1, per_rank=[n_done])
    t = torch.tensor([float(n_done), elapsed_s, checksum], dtype=torch.float64, device=device)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    per = [int(o[0].item()) for o in out]
    return dict(total=sum(per), max_elapsed_s=max(o[1].item() for o in out), checksum=sum(o[2].item() for o in out),
                world=dist.get_world_size(), per_rank=per)
