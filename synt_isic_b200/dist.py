"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink on the
GPU box, gloo in the CPU tests).  The reference has no distributed code at all (SURVEY.md
section 5); the hot path shards naturally because its units are independent:

* sampling    -- unit = one image (class, seed): ``partition`` assigns whole batches to ranks,
                 no data-path collective, one ``gather_images`` of the uint8 results at the end;
* classifier  -- unit = one evaluation (frame / coalition / intervention): ``sharded_eval``
                 splits the batch contiguously, every rank evaluates its slice with replicated
                 weights, ONE ``all_gather`` of the [n,7] logits.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None):
    """Initialises the default process group from torchrun's env (RANK/WORLD_SIZE/MASTER_*).
    Returns (rank, world, local_rank); a no-op (0, 1, 0) when not launched under torchrun."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 0, 1, int(os.environ.get("LOCAL_RANK", "0"))
    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous split of ``n`` units: the first ``n % world`` ranks get one extra unit."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def partition(units, rank: int, world: int):
    """Round-robin assignment of work units (e.g. (class, batch) pairs) to ranks."""
    return [u for i, u in enumerate(units) if i % world == rank]


def _world(group):
    """``group=None`` means "this rank alone" (no collective is issued); pass ``dist.group.WORLD`` (or a
    sub-group) to shard across ranks.  Every rank of the group must then make the same call."""
    if group is None or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def gather_rows(local: torch.Tensor | None, n: int, group=None, device=None) -> torch.Tensor:
    """The one exchange of the classifier-evaluation paths: every rank holds the rows [lo, hi) = ``shard_bounds(n, rank,
    world)`` of an [n, k] result; ONE ``all_gather`` (rows padded to ceil(n / world)) gives every rank the whole table."""
    rank, world = _world(group)
    if world == 1:
        return local
    lo, hi = shard_bounds(n, rank, world)
    if device is None:
        device = local.device
    if n >= world:
        k = local.shape[1]                                  # no rank slice is empty: every rank knows the row width
    else:                                                   # fewer rows than ranks: agree on the width (one tiny all_reduce)
        kk = torch.tensor([local.shape[1] if local is not None else 0], device=device)
        dist.all_reduce(kk, op=dist.ReduceOp.MAX, group=group)
        k = int(kk.item())
    per = (n + world - 1) // world
    pad = torch.zeros(per, k, dtype=torch.float32, device=device)
    if local is not None and hi > lo:
        pad[: hi - lo] = local.float()
    gathered = torch.empty(world * per, k, dtype=torch.float32, device=device)
    dist.all_gather_into_tensor(gathered, pad, group=group)
    if n == world * per:
        return gathered
    parts = []
    for r in range(world):
        a, b = shard_bounds(n, r, world)
        parts.append(gathered[r * per: r * per + (b - a)])
    return torch.cat(parts)


def sharded_eval(fn, items: torch.Tensor, group=None, chunk: int = 256) -> torch.Tensor:
    """``fn(items)`` row-wise (fn maps [n, ...] -> [n, k]) with the rows split across ranks and a
    single all_gather of the results.  With one rank it is just a chunked call.  ``items`` may live in host memory: every
    rank then touches (and uploads) only its own rows."""
    rank, world = _world(group)
    n = items.shape[0]
    lo, hi = shard_bounds(n, rank, world) if world > 1 else (0, n)
    outs = [fn(items[i:min(i + chunk, hi)]) for i in range(lo, hi, chunk)]
    if world == 1:
        return torch.cat(outs) if len(outs) > 1 else outs[0]
    local = torch.cat(outs) if outs else None
    dev = local.device if local is not None else items.device
    return gather_rows(local, n, group, dev)


def gather_images(local_u8: torch.Tensor, counts, group=None, dst: int = 0):
    """Collects each rank's uint8 images [n_r,128,128,3] on rank ``dst`` (one gather at the end of
    sampling).  ``counts[r]`` = images produced by rank r.  Returns the concatenation on dst, else None."""
    rank, world = _world(group)
    if world == 1:
        return local_u8
    per = max(counts)
    pad = torch.zeros((per,) + tuple(local_u8.shape[1:]), dtype=local_u8.dtype, device=local_u8.device)
    pad[: local_u8.shape[0]] = local_u8
    # a GATHER to `dst` (only dst receives the images), not an all_gather; `dst` is a rank of `group`
    dst_global = dist.get_global_rank(group, dst) if group is not None and group is not dist.group.WORLD else dst
    if rank != dst:
        dist.gather(pad, None, dst=dst_global, group=group)
        return None
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.gather(pad, bufs, dst=dst_global, group=group)
    return torch.cat([bufs[r][: counts[r]] for r in range(world)])


def max_over_ranks(value: float, device, group=None) -> float:
    rank, world = _world(group)
    if world == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
