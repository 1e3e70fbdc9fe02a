"""Multi-GPU plumbing: one process per GPU, problems sharded by index, no data-path collective.

The only exchange is the best-of-batch selection: an all-gather of one
(merit, global index) pair per rank, a local argmin, and a broadcast of the winner's decision vector
from the owning rank.  ``torch.distributed`` (NCCL on GPUs, gloo in CPU tests) is the plumbing.
"""
from __future__ import annotations

import os
from typing import Tuple

import numpy as np


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of problem indices owned by ``rank``: [lo, hi).  Remainder goes to the low ranks."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def env_rank_world() -> Tuple[int, int, int]:
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init_process_group(backend: str = "nccl"):
    import torch.distributed as dist
    rank, local_rank, world = env_rank_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kw = {}
        if backend == "nccl":
            import torch
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def bind_to_gpu_numa(device_index: int) -> int:
    """Pin this process to the CPU cores NVML reports as local to the GPU, so that pinned host buffers are first-touched
    on the GPU's own NUMA node (the host-buffer path moves 1.5 GB per step per GPU: with 8 ranks the host side is the
    bound).  Returns the number of CPUs in the new affinity mask, 0 if nothing was changed."""
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            n_cpu = os.cpu_count() or 1
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
            cpus = [64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1 and 64 * w + b < n_cpu]
            allowed = os.sched_getaffinity(0)
            cpus = [c for c in cpus if c in allowed]
            if cpus and len(cpus) < len(allowed):
                os.sched_setaffinity(0, cpus)
                return len(cpus)
        finally:
            pynvml.nvmlShutdown()
    except Exception:
        pass
    return 0


def merit(f, viol, feas_tol: float = 1e-4, penalty: float = 1e3):
    """Scalar used to rank candidates: the objective, plus a large penalty on constraint violation beyond tol."""
    import torch
    return f + penalty * torch.clamp(viol - feas_tol, min=0.0)


def select_best(local_merit, w_soa, first_index: int, n_w: int):
    """local_merit: (P_local,) tensor; w_soa: (n_w, ld) tensor of the local shard.
    Returns (best_merit: float, best_global_index: int, w_best: (n_w,) tensor on every rank)."""
    import torch
    import torch.distributed as dist
    dev = local_merit.device
    if local_merit.numel() > 0:
        val, idx = torch.min(local_merit, dim=0)
        pair = torch.stack([val.double(), (idx + first_index).double()])
    else:
        pair = torch.tensor([float("inf"), -1.0], dtype=torch.float64, device=dev)
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    if world == 1:
        best_idx = int(pair[1].item())
        return float(pair[0].item()), best_idx, w_soa[:n_w, best_idx - first_index].clone()
    gathered = [torch.empty_like(pair) for _ in range(world)]
    dist.all_gather(gathered, pair)                                  # world x 16 B over NVLink: latency only
    table = torch.stack(gathered).cpu().numpy()
    owner = int(np.lexsort((table[:, 1], table[:, 0]))[0])           # ties -> lowest global index
    best_val, best_idx = float(table[owner, 0]), int(table[owner, 1])
    w_best = torch.empty(n_w, dtype=w_soa.dtype, device=dev)
    if dist.get_rank() == owner:
        w_best.copy_(w_soa[:n_w, best_idx - first_index])
    dist.broadcast(w_best, src=owner)                                # <= 727 floats
    return best_val, best_idx, w_best
