"""Data-parallel plumbing of the hot path (one process per GPU, torch.distributed).

The path shards by batch sample and has no data-path collective: every rank
processes its own frames.  The only exchange is the scalar loss that the trainer logs
(reference: dro_sfm/utils/reduce.py:10-30, average=True) -- one all-reduce of one float per step --
plus, in the benchmark, the max-over-ranks of the device timings.
"""
import os

import torch
import torch.distributed as dist


def env_rank():
    """(rank, world_size, local_rank) from the torchrun environment; (0, 1, 0) when absent."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_seed(base_seed, rank, cfg_id=0):
    """Seed of the synthetic shard of `rank` (1234 + 1000*rank + cfg_id)."""
    return int(base_seed) + 1000 * int(rank) + int(cfg_id)


def init(backend, device=None):
    """Initialise the default process group from the torchrun environment if WORLD_SIZE > 1."""
    rank, world, _ = env_rank()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        kwargs = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, **kwargs)
    return rank, world


def average_loss(loss):
    """In-place mean of the per-rank loss over all ranks (no-op for a single process)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(loss, op=dist.ReduceOp.SUM)
        loss /= dist.get_world_size()
    return loss


def max_over_ranks(values, device="cpu"):
    """Element-wise maximum of a list of floats over all ranks (device timings are reported this way)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def whole_job_rate(units_per_rank, world, seconds):
    """Aggregate throughput of a weak-scaling run: all ranks' units over the slowest rank's time."""
    return units_per_rank * world / seconds


def bind_host_to_gpu(local_rank):
    """Best effort: run this process on the CPUs NVML reports as local to its GPU, so that the pinned staging buffers
    it allocates afterwards (first touch) live on the GPU's NUMA node and the per-step H2D copies of the ranks do not
    all cross the inter-socket link.  Returns the CPU set applied, or None (no NVML, no affinity API, empty mask)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(int(local_rank))
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (os.cpu_count() + 63) // 64 + 16)
        local = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = local & allowed
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:
        return None
