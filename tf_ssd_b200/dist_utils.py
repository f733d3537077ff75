"""Data-parallel plumbing for the hot path.

The reference has no distribution at all.  The path shards over the
batch axis only: inference shards are independent (no data-path collective);
training needs exactly one exchange, the mean of the gradients.  One process
per GPU, ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests) is the
only transport.
"""

from __future__ import annotations

import os
from typing import Iterable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def env_rank() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (1-process defaults)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init_from_env(backend: str = "nccl") -> Tuple[int, int, int]:
    """Initialise the default process group when WORLD_SIZE > 1 (rendezvous from MASTER_ADDR/PORT)."""
    rank, local_rank, world = env_rank()
    if world > 1 and not dist.is_initialized():
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of a global batch for ``rank``: sizes differ by at most one and
    the concatenation over ranks is the global batch in order (what "8-GPU step == 1-GPU step on the
    concatenated batch" needs)."""
    base, extra = divmod(n_items, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def max_over_ranks(value: float, device: torch.device) -> float:
    """Max of a host scalar over all ranks (multi-GPU timings are the slowest rank's)."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def world_size() -> int:
    return dist.get_world_size() if dist.is_initialized() else 1


class _Done(object):
    def wait(self) -> None:
        return None


def bind_to_gpu_numa(local_rank: int) -> bool:
    """Pin the calling process to the CPUs NVML reports as closest to GPU ``local_rank`` so that pinned staging
    buffers are first-touched on that GPU's NUMA node (8 ranks streaming images from one node's memory are
    host-bandwidth bound).  Best effort: returns False when NVML or the affinity call is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = int(vis.split(",")[local_rank]) if vis else local_rank
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = [64 * i + b for i, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1]
        if not cpus:
            return False
        os.sched_setaffinity(0, cpus)
        return True
    except Exception:  # noqa: BLE001
        return False


def barrier() -> None:
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


class GradBuckets(object):
    """Flat fp32 gradient buckets for the one training exchange: ``mean`` over ranks.

    Gradients are views into a few large contiguous buffers (``bucket_bytes`` each, 8-25 MB:
    sized for launch latency and overlap on NVSwitch, not for link count), so the all-reduce
    runs on whole buckets and a fused optimizer can consume them in place."""

    def __init__(self, shapes: Sequence[Tuple[str, Sequence[int]]], device: torch.device, bucket_bytes: int = 16 << 20):
        self.views = {}
        self.bucket_of = {}                               # variable name -> index of its bucket
        self.buckets: List[torch.Tensor] = []
        cur: List[Tuple[str, Sequence[int], int]] = []
        cur_elems = 0
        limit = max(1, bucket_bytes // 4)

        def flush():
            nonlocal cur, cur_elems
            if not cur:
                return
            buf = torch.zeros(cur_elems, dtype=torch.float32, device=device)
            off = 0
            for name, shape, n in cur:
                self.views[name] = buf[off:off + n].view(*shape)
                self.bucket_of[name] = len(self.buckets)
                off += (n + 3) // 4 * 4                    # every view starts on a 16-byte boundary
            self.buckets.append(buf)
            cur, cur_elems = [], 0

        for name, shape in shapes:
            n = 1
            for d in shape:
                n *= int(d)
            n_pad = (n + 3) // 4 * 4                       # keep every view 16-byte aligned
            if cur and cur_elems + n_pad > limit:
                flush()
            cur.append((name, tuple(shape), n))
            cur_elems += n_pad
        flush()

    def zero_(self) -> None:
        for b in self.buckets:
            b.zero_()

    def allreduce_sum_async(self, k: int):
        """Starts the SUM all-reduce of bucket ``k`` (NCCL runs it on its own stream behind everything enqueued on the
        current stream so far) and returns a handle whose ``wait()`` makes the current stream wait for it.  The 1 / world
        of the mean is the consumer's business (the fused optimizer folds it into its gradient scale)."""
        if not (dist.is_initialized() and dist.get_world_size() > 1):
            return _Done()
        return dist.all_reduce(self.buckets[k], op=dist.ReduceOp.SUM, async_op=True)

    def allreduce_mean_(self) -> None:
        """Sum over ranks then divide by the world size (Keras' batch mean over the global batch when
        every rank holds an equal shard); asynchronous launches, one wait at the end."""
        if not (dist.is_initialized() and dist.get_world_size() > 1):
            return
        world = dist.get_world_size()
        works = [dist.all_reduce(b, op=dist.ReduceOp.SUM, async_op=True) for b in self.buckets]
        for w in works:
            w.wait()
        for b in self.buckets:
            b.mul_(1.0 / world)
