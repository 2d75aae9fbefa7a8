"""Multi-GPU plumbing: one process per GPU, codeword batches shard embarrassingly (no data-path
collective); the only exchange is a sum of BLER counters / a max of timings per measurement point
(plot_BLER_vs_SNR.m:23-27 describes the reference's manual version:
independent seeds per instance, results aggregated afterwards)."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def env_world():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init(backend: str | None = None):
    """Initialise torch.distributed from the torchrun environment (no-op for a single process)."""
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def bind_to_gpu_numa_node(gpu_index: int) -> bool:
    """Pin this process to the CPUs NVML reports as local to `gpu_index`, so that the pinned host staging
    buffers it allocates afterwards (first touch) sit on the GPU's own NUMA node: with one process per GPU the
    host<->device copies of all ranks then do not funnel through one socket's memory controllers."""
    try:
        import pynvml
        pynvml.nvmlInit()
        dev = pynvml.nvmlDeviceGetHandleByIndex(int(gpu_index))
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(dev, (n_cpu + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus:
            os.sched_setaffinity(0, cpus)
            return True
    except Exception:
        pass
    return False


def shard_range(total: int, rank: int, world: int):
    """Contiguous [lo, hi) slice of `total` units for `rank`; sizes differ by at most one."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def rank_seed(seed: int, rank: int) -> int:
    """Distinct, reproducible random stream per rank (plot_BLER_vs_SNR.m:23-27)."""
    return (int(seed) * 0x9E3779B97F4A7C15 + int(rank) * 0xBF58476D1CE4E5B9 + 1) & 0xFFFFFFFFFFFFFFFF


def _device_for_collective():
    if dist.is_initialized() and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def sum_counters(counters):
    """All-reduce(sum) of a short int64 vector, e.g. [blocks, block_errors, bit_errors, iterations]."""
    t = torch.as_tensor(counters, dtype=torch.int64).clone()
    if dist.is_initialized() and dist.get_world_size() > 1:
        t = t.to(_device_for_collective())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        t = t.cpu()
    return t


def max_over_ranks(value: float) -> float:
    t = torch.tensor([float(value)], dtype=torch.float64)
    if dist.is_initialized() and dist.get_world_size() > 1:
        t = t.to(_device_for_collective())
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t = t.cpu()
    return float(t.item())


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
