"""Multi-GPU plumbing of the placement path: one process per GPU, queries sharded, index replicated.

The path shards by query with NO data-path collective (queries are independent in the reference:
core/src/use_cases/place_sequences/mod.rs:123-126).  What is left for
``torch.distributed`` is bookkeeping: which contiguous slice of the batch a rank places, and the
concatenation of the per-rank result arrays on rank 0 in input order (the reference's shared output
file).  Works with the ``nccl`` backend on GPUs and with ``gloo`` on CPU (tests).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from .engine import RESULT_DTYPES, BatchResult


def shard_bounds(n_queries: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced slice [lo, hi) of a batch of ``n_queries`` for ``rank`` of ``world``."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n_queries, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(bases: np.ndarray, offsets: np.ndarray, rank: int, world: int):
    """The (bases, offsets) sub-batch of ``rank``; offsets are rebased to start at 0."""
    lo, hi = shard_bounds(len(offsets) - 1, rank, world)
    off = offsets[lo:hi + 1]
    return bases[int(off[0]):int(off[-1])], (off - off[0]).astype(np.uint64), lo, hi


def gather_results(local: BatchResult, n_total: int, group=None, dst: int = 0) -> Optional[BatchResult]:
    """Concatenate the per-rank result arrays on ``dst`` in rank (= input) order.  Returns the full
    ``BatchResult`` on ``dst`` and ``None`` elsewhere.  Device-agnostic: tensors travel on the
    backend's device (CUDA for nccl, CPU for gloo)."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    sizes = [shard_bounds(n_total, r, world) for r in range(world)]
    out = BatchResult(n_total) if rank == dst else None
    max_len = max(hi - lo for lo, hi in sizes)
    for name, dt in RESULT_DTYPES:
        arr = getattr(local, name)
        # unsigned 64/32-bit integers travel as their signed bit patterns
        view = arr.view({np.dtype(np.uint64): np.int64, np.dtype(np.uint32): np.int32}.get(arr.dtype, arr.dtype))
        # ragged slices: pad to the largest slice, all_gather (same call for nccl and gloo), trim
        pad = np.zeros(max_len, dtype=view.dtype)
        pad[: len(view)] = view
        t = torch.from_numpy(pad).to(dev)
        bufs = [torch.empty(max_len, dtype=t.dtype, device=dev) for _ in range(world)]
        dist.all_gather(bufs, t, group=group)
        if rank == dst:
            full = torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, sizes)]).cpu().numpy().view(arr.dtype)
            getattr(out, name)[:] = full
    return out
