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


def sample_split_gram(engine, wX: np.ndarray, qX, rank: int, world: int, device=None, group=None):
    """Sample-split Gram stage (BASELINE.json north_star): `wX`, `qX` are the (N0, m) feature-major
    layer inputs every rank holds on the HOST (qX None / wX itself for the first layer).  Rank r moves only the samples
    `shard_range(m, r, world)` to its GPU, contracts them with `engine.gram_matrices` (libgpfq: tcgen05 / DMMA Gram kernels)
    and the (N0, N0) fp64 partial matrices are summed with one all-reduce per matrix (NCCL over NVLink; gloo in the CPU
    tests).  Returns (G1, G2) tensors holding the whole-sample Grams on every rank, G1 is G2 when qX is wX.

    Against input replication this moves 16 N0^2 bytes over NVLink instead of 8 N0 m, never holds more than m / world
    samples per GPU (the only way a layer with m x N0 beyond one GPU's HBM fits), and divides the Gram stage by `world`.
    The fp64 sums are associated differently from a one-GPU Gram (partial sums per rank), a 1e-16-relative effect."""
    import torch
    import torch.distributed as dist
    same = qX is None or qX is wX
    m = wX.shape[1]
    lo, hi = shard_range(m, rank, world)
    if hi == lo:                       # fewer samples than ranks: contribute zeros
        lo, hi = 0, 0

    N0 = wX.shape[0]
    if hi > lo:
        # strided views of this rank's sample range: the library moves them with one pitched H2D copy per matrix
        G1, G2 = engine.gram_matrices(wX[:, lo:hi], None if same else qX[:, lo:hi], device_out=True)
        if not isinstance(G2, torch.Tensor):       # an engine returning NumPy (tests)
            G2 = torch.from_numpy(np.ascontiguousarray(G2))
            G1 = G2 if same else torch.from_numpy(np.ascontiguousarray(G1))
    else:
        G2 = torch.zeros((N0, N0), dtype=torch.float64, device=device)
        G1 = G2 if same else torch.zeros((N0, N0), dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(G2, group=group)
        if not same:
            dist.all_reduce(G1, group=group)
    return (G2 if same else G1), G2


def prefer_sample_split(N0: int, m: int, world: int, hbm_bytes: float = 150e9) -> bool:
    """Sample split vs input replication for a Dense layer of a `world`-rank job: split when the all-reduce of the Grams
    (two (N0, N0) fp64 matrices) moves fewer bytes than the all-gather of the inputs (two (N0, m) fp32 matrices), i.e.
    m > 2 N0 -- the Gram stage then also shrinks by `world` -- or when the replicated inputs would not fit one GPU."""
    if world <= 1:
        return False
    return (8.0 * N0 * m > hbm_bytes) or (m > 2 * N0)


def image_split_conv_gram(engine, act: np.ndarray, actq, kernel_size, strides, padding, rate, rank: int, world: int,
                          group=None):
    """Image-split Gram stage of a conv layer: `act`, `actq` are the (n_img, H, W, C) layer inputs every rank holds on the
    HOST (actq None / act itself for the first layer).  Rank r hands only the images `shard_range(n_img, r, world)` to its
    GPU (`engine.conv_gram_nhwc`: the per-channel kk x kk Grams are sums over images) and the (C, 2, kk, kk) fp64 partial
    matrices are summed with ONE all-reduce (83 KB for 64 channels).  Returns the whole-batch Grams of ALL channels on
    every rank: 1 / world of the activations per host link and per GPU, no replication, no other collective."""
    import torch
    import torch.distributed as dist
    same = actq is None or actq is act
    n_img = act.shape[0]
    lo, hi = shard_range(n_img, rank, world)
    kk = int(kernel_size[0]) * int(kernel_size[1])
    if hi > lo:
        gram = engine.conv_gram_nhwc(act[lo:hi], None if same else actq[lo:hi], kernel_size, strides, padding, rate)
        if not isinstance(gram, torch.Tensor):       # an engine returning NumPy (tests)
            gram = torch.from_numpy(np.ascontiguousarray(gram))
    else:                                            # fewer images than ranks: contribute zeros
        dev = None if getattr(engine, "device", None) is None else torch.device("cuda", engine.device)
        gram = torch.zeros((act.shape[3], 2, kk, kk), dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(gram, group=group)
    return gram
