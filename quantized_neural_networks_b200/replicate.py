"""Input replication for multi-GPU jobs: every rank needs the WHOLE layer input (X, X~ replicated, neurons / channels
sharded -- BASELINE.json north_star), but the host link is the slow part (one PCIe x16 per GPU against 900 GB/s of NVLink).

So each rank copies only its 1/world slice of the leading axis host -> device, and the slices are exchanged with ONE
all-gather over NCCL / NVLink: PCIe bytes per rank drop by `world`, the NVLink step costs a few ms.
With a gloo group (the CPU tests) the same code runs on host tensors, which checks the slicing / padding logic without GPUs.

torch is plumbing here (buffers + torch.distributed); the compute stays in libgpfq.
"""
from __future__ import annotations

import numpy as np


def shard_range(n: int, rank: int, world: int):
    """Contiguous, balanced slice of n independent units for `rank` of `world`."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def replicate_leading_axis(host: np.ndarray, rank: int, world: int, device=None, group=None):
    """All ranks hold the same `host` array (float32, C-contiguous).  Returns a tensor on `device` (CUDA with NCCL, CPU with
    gloo) holding the whole array, of which only rows [rank * chunk, (rank + 1) * chunk) crossed this rank's host link.
    The all-gather needs equal chunks, so the leading axis is padded to world * chunk rows internally."""
    import torch
    import torch.distributed as dist
    host = np.ascontiguousarray(host)
    n = host.shape[0]
    if world == 1:
        t = torch.from_numpy(host)
        return t.to(device, non_blocking=True) if device is not None else t
    chunk = -(-n // world)
    lo, hi = min(rank * chunk, n), min((rank + 1) * chunk, n)
    full = torch.empty((world * chunk,) + host.shape[1:], dtype=torch.from_numpy(host[:0]).dtype, device=device)
    mine = full[rank * chunk:(rank + 1) * chunk]          # in place: this rank's slot of the gathered tensor
    if hi > lo:
        mine[:hi - lo].copy_(torch.from_numpy(host[lo:hi]), non_blocking=True)
    if hi - lo < chunk:
        mine[hi - lo:].zero_()
    dist.all_gather_into_tensor(full, mine, group=group)
    return full[:n]


def h2d_bytes_per_rank(shape, itemsize: int, world: int) -> int:
    """Bytes of one array that cross ONE rank's host link under `replicate_leading_axis`."""
    n = shape[0]
    chunk = -(-n // world)
    return int(chunk * int(np.prod(shape[1:], dtype=np.int64)) * itemsize)
