"""Data-parallel plumbing (new in this build; the reference is single-device).

The train step shards by batch only (every loss term is a batch mean of per-image sums,
vae/trainer.py:13,18,127-128,161): each rank runs the step on its contiguous shard with gradients
pre-scaled by 1/(per_gpu_batch * world_size), and the only exchange is a SUM all-reduce of the
flat fp32 gradient arena, issued per backward segment so it overlaps the remaining backward
work.  torch.distributed (NCCL on GPUs, gloo in the CPU tests) carries the collective.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialises torch.distributed from torchrun's environment.  Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    return rank, world, local_rank


def shard_range(global_batch: int, world: int, rank: int):
    """Contiguous shard [start, stop) of the global batch owned by `rank` (equal shards required)."""
    if global_batch % world:
        raise ValueError(f"global batch {global_batch} is not divisible by world size {world}")
    per = global_batch // world
    return rank * per, (rank + 1) * per


class BucketReducer:
    """SUM all-reduce of "buckets" of one flat tensor, asynchronously.  A bucket = one backward segment = a few contiguous ranges
    (the arena keeps the Keras variable order, so the two encoders' layers of one segment are two ranges)."""

    def __init__(self, flat: torch.Tensor, segments, group=None):
        self.flat = flat
        self.segments = []
        for seg in segments:
            ranges = [seg] if seg and isinstance(seg[0], int) else list(seg)
            self.segments.append([(int(o), int(c)) for o, c in ranges])
        self.group = group
        self._work = []
        self.enabled = dist.is_initialized() and dist.get_world_size(group) > 1

    def reduce(self, seg: int):
        """Issues the all-reduce of bucket `seg` behind the work queued so far on the current stream; returns its work handles."""
        if not self.enabled:
            return []
        works = [dist.all_reduce(self.flat[off:off + cnt], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
                 for off, cnt in self.segments[seg]]
        self._work.extend(works)
        return works

    def wait_all(self):
        for w in self._work:
            w.wait()
        self._work = []


class NvlsArenas:
    """Symmetric (NVLS multicast) parameter and gradient arenas for the fused data-parallel optimizer (sv_nvls_adam_segment).

    torch.distributed._symmetric_memory provides the plumbing only: identical allocations on every rank, the multicast address that
    maps all of them (NVLink 5 / NVSwitch), and a device-side cross-rank barrier.  The reduction, the Adam update and the broadcast
    are libsplitvae's kernel.  `available()` is False when the group has no multicast support (then the NCCL bucket path is used)."""

    def __init__(self, group=None):
        import torch.distributed._symmetric_memory as symm
        self.symm = symm
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.tensors, self.handles = [], []

    @staticmethod
    def available(group=None):
        if os.environ.get("SV_NO_NVLS", "0") == "1" or not dist.is_initialized() or dist.get_world_size(group) < 2:
            return False
        try:
            import torch.distributed._symmetric_memory as symm
            t = symm.empty(1024, dtype=torch.float32, device=torch.device("cuda", torch.cuda.current_device()))
            h = symm.rendezvous(t, group if group is not None else dist.group.WORLD)
            return bool(h.has_multicast_support) and int(h.multicast_ptr) != 0
        except Exception:
            return False

    def alloc(self, n_floats):
        """Zeroed fp32 arena in symmetric memory (every rank must call this in the same order with the same size)."""
        t = self.symm.empty(int(n_floats), dtype=torch.float32, device=torch.device("cuda", torch.cuda.current_device()))
        t.zero_()
        h = self.symm.rendezvous(t, self.group)
        if not h.has_multicast_support or int(h.multicast_ptr) == 0:
            raise RuntimeError("symmetric memory without multicast support")
        self.tensors.append(t)
        self.handles.append(h)
        return t

    def multicast_ptr(self, tensor):
        i = next(k for k, t in enumerate(self.tensors) if t.data_ptr() == tensor.data_ptr())
        h = self.handles[i]
        return int(h.multicast_ptr) + int(getattr(h, "offset", 0) or 0)

    def barrier(self, channel=0):
        """Cross-rank barrier on the current stream (a small kernel over the signal pads; graph-capturable)."""
        self.handles[0].barrier(channel=channel)


def mean_scalars(values: torch.Tensor, group=None) -> torch.Tensor:
    """Average of per-rank loss scalars = the scalar of the global batch (equal shards)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        values = values.clone()
        dist.all_reduce(values, op=dist.ReduceOp.SUM, group=group)
        values /= dist.get_world_size(group)
    return values
