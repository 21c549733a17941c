#!/usr/bin/env python
"""bench.py -- GPFQ hot path on B200: quantized weights/s of a full-network pass (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's NumPy path on the host cores

Workload (config.workload): BASELINE.json configs[1], the CIFAR10 CNN of train_cifar10_cnn.py -- six 3x3 'same'
Conv2D layers handed over as per-channel patch matrices (9 x n_patches, m_img = 5008 images) and two Dense layers
(2048->128, 128->10, m = 5008), 4-bit alphabet (K = 16), synthetic activations, random-init weights.
One "step" = one pass of the hot path over all eight layers.

  value   whole-job weights/s with every input resident in HBM when the timed region starts (CUDA events on the
          launching stream, max over ranks).  Inputs (25.7 GB) are far larger than L2, so no flush is needed.
  e2e     the same pass through the reference-facing C ABI with HOST (pinned) buffers, copies inside the timed region.
          Conv layers go through gpfq_conv_layer_nhwc (the coarser override point of INTEGRATION.md: the layer's
          (n_img, H, W, C) activations are handed over, patches are extracted on the device -- 9x fewer PCIe bytes
          than per-channel patch matrices), Dense layers through gpfq_dense_layer; every Q comes back to the host.
  roofline  the dominant kernel (conv_gram9_tma_kernel, HBM-bound, 74 % of the step): algorithmic bytes / CUDA-event time
          of that stage.  roofline_tensor_stage: the Dense Gram stage on tcgen05 (int8 slices), against 2 x measured bf16.
  from_activations  the same device-resident pass with every conv layer handed over as its NHWC activation tensor (no
          patch matrices: the 9 x 9 Grams are 13 displacement sums of the activations, conv_corr.cu) -- value, ms/step and
          the roofline of conv_corr9_tma_kernel.
  cpu_baseline  the oracle's NumPy restatement of the reference walk (kind "port": the reference is Python and does
          not travel to the GPU box), one process per host core exactly like the reference's ProcessPoolExecutor, on a
          bounded sample of every layer, extrapolated linearly in the number of neurons/filters.
Multi-GPU: conv channels and Dense neurons shard over ranks (no data-path collective); total work is fixed => "strong".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_IMG = 5008
BITS, CSCALAR = 4, 4
# (name, kind, C or N0, F or N1, H)   -- SURVEY.md App. B, config 2
CIFAR_LAYERS = [
    ("conv0", "conv", 3, 32, 32), ("conv2", "conv", 32, 32, 32), ("conv6", "conv", 32, 64, 16),
    ("conv8", "conv", 64, 64, 16), ("conv12", "conv", 64, 128, 8), ("conv14", "conv", 128, 128, 8),
    ("dense19", "dense", 2048, 128, 0), ("dense22", "dense", 128, 10, 0),
]
MNIST_LAYERS = [("dense1", "dense", 784, 500, 0), ("dense3", "dense", 500, 300, 0), ("dense5", "dense", 300, 10, 0)]


def shard_range(n, rank, world):
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def layer_weights(layer):
    name, kind, a, b, H = layer
    return 9 * a * b if kind == "conv" else a * b


def make_alphabet(W_abs_median, bits=BITS, c=CSCALAR):
    return c * float(W_abs_median) * np.linspace(-1, 1, int(round(2 ** bits)))


# ----------------------------------------------------------------------------------------------------------------
# synthetic data (device side, torch is only the buffer/generator)
# ----------------------------------------------------------------------------------------------------------------
def build_device_inputs(layers, n_img, rank, world, dev):
    """Per layer: dict with device tensors.  conv: NHWC activations act/actq (what the reference's host code collects,
    quantized_network.py:468) and, derived from them, Xp/Xqp = per-channel (9, n) patch matrices of this rank's channels
    (what _build_patch_array hands to the workers, :789-797)."""
    import torch
    out = []
    for li, (name, kind, a, b, H) in enumerate(layers):
        g = torch.Generator(device=dev).manual_seed(1000 + li)
        if kind == "conv":
            C, F = a, b
            W = (torch.rand((3, 3, C, F), device=dev, generator=g) * 2 - 1) * float(np.sqrt(6.0 / (9 * C)))
            A = make_alphabet(torch.median(W.abs().flatten()))
            lo, hi = shard_range(C, rank, world)
            n = n_img * H * H
            first = li == 0
            shape = (n_img, H, H, C)
            if first:   # image-like: uniform[0,1) with half the pixels zero; X == Xq
                act = torch.rand(shape, device=dev, generator=g) * (torch.rand(shape, device=dev, generator=g) < 0.5)
                actq = None
            else:       # hidden: X = relu(Z), Xq = relu(Z + 0.05 N)
                Z = torch.randn(shape, device=dev, generator=g)
                act = torch.relu(Z)
                actq = torch.relu(Z + 0.05 * torch.randn(shape, device=dev, generator=g))
                del Z
            Xp, Xqp = [], []
            cb = max(1, min(hi - lo, int(1.5e9 // (36 * n))))
            for c0 in range(lo, hi, cb):
                c1 = min(hi, c0 + cb)
                mats = []
                for t in ((act,) if first else (act, actq)):
                    tc = t[..., c0:c1].permute(3, 0, 1, 2).reshape(-1, 1, H, H)
                    p = torch.nn.functional.unfold(tc, 3, padding=1)                          # (cb*n_img, 9, H*H)
                    p = p.reshape(c1 - c0, n_img, 9, H * H).permute(0, 2, 1, 3).reshape(c1 - c0, 9, n).contiguous()
                    mats.append(p)
                    del tc
                Xp += list(mats[0])
                Xqp += list(mats[0] if first else mats[1])
                del p, mats
            out.append(dict(name=name, kind=kind, W=W, A=A, Xp=Xp, Xqp=None if first else Xqp, c0=lo, n_ch=hi - lo,
                            n=n, C=C, F=F, first=first, act=act, actq=actq))
        else:
            N0, N1 = a, b
            m = n_img if n_img else 25000
            W = (torch.rand((N0, N1), device=dev, generator=g) * 2 - 1) * float(np.sqrt(6.0 / N0))
            A = make_alphabet(torch.median(W.abs().flatten()))
            first = (li == 0)
            if first:
                X = torch.rand((N0, m), device=dev, generator=g) * (torch.rand((N0, m), device=dev, generator=g) < 0.5)
                Xq = None
            else:
                Z = torch.randn((N0, m), device=dev, generator=g)
                X = torch.relu(Z)
                Xq = torch.relu(Z + 0.05 * torch.randn((N0, m), device=dev, generator=g))
                del Z
            lo, hi = shard_range(N1, rank, world)
            out.append(dict(name=name, kind=kind, W=W, A=A, X=X, Xq=Xq, j0=lo, j1=hi, N0=N0, N1=N1, m=m, first=first))
        torch.cuda.empty_cache()
    return out


def run_pass_device(eng, data, outs, sync=False):
    for d, o in zip(data, outs):
        if d["kind"] == "conv":
            eng.conv_channels(d["Xp"], d["Xqp"], d["W"], d["A"], c0=d["c0"], n_channels=d["n_ch"], out=o, sync=sync)
        else:
            eng.dense_layer(d["X"], d["Xq"], d["W"], d["A"], j0=d["j0"], j1=d["j1"], out=o, sync=sync)


def run_pass_device_nhwc(eng, data, outs, sync=False, rank=0, world=1):
    """The same pass with the conv layers handed over as NHWC activation tensors (the coarser override point of
    INTEGRATION.md: `_quantize_conv2D_layer_parallel_jit` before `_build_patch_array`), still device-resident.
    world > 1: conv layers split over IMAGES -- every rank contracts its n_img / world images of all channels, one NCCL
    all-reduce sums the per-channel 9 x 9 Grams (C x 162 doubles), every rank then walks every channel."""
    import torch.distributed as dist
    for d, o in zip(data, outs):
        if d["kind"] == "conv" and world > 1:
            lo, hi = shard_range(d["act"].shape[0], rank, world)
            gram = eng.conv_gram_nhwc(d["act"][lo:hi], None if d["actq"] is None else d["actq"][lo:hi], (3, 3))
            dist.all_reduce(gram)
            eng.conv_layer_from_gram(gram, d["W"], d["A"], out=o, sync=sync)
        elif d["kind"] == "conv":
            eng.conv_layer_nhwc(d["act"], d["actq"], d["W"], d["A"], c0=d["c0"], n_channels=d["n_ch"], out=o, sync=sync)
        else:
            eng.dense_layer(d["X"], d["Xq"], d["W"], d["A"], j0=d["j0"], j1=d["j1"], out=o, sync=sync)


def to_host_pinned(data):
    """Pinned host copies of what the reference's host code holds: NHWC activations per conv layer, (N0, m) matrices per
    Dense layer, the kernels."""
    import torch

    def pin(t):
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t)
        return h.numpy()

    host, nbytes = [], 0
    for d in data:
        if d["kind"] == "conv":
            act = pin(d["act"])
            actq = None if d["actq"] is None else pin(d["actq"])
            nbytes += act.nbytes + (0 if actq is None else actq.nbytes)
            h = dict(d, act=act, actq=actq, W=d["W"].cpu().numpy(), Xp=None, Xqp=None)
        else:
            X = pin(d["X"])
            Xq = None if d["Xq"] is None else pin(d["Xq"])
            nbytes += X.nbytes + (0 if Xq is None else Xq.nbytes)
            h = dict(d, X=X, Xq=Xq, W=d["W"].cpu().numpy())
        nbytes += h["W"].nbytes
        host.append(h)
    return host, nbytes


def run_pass_host(eng, host, keep=None, rank=0, world=1, dev=None):
    """One pass from HOST buffers.  world == 1: host pointers straight into the C ABI (the library overlaps its chunked
    H2D copies with the Gram kernels).  world > 1: every rank needs the whole layer input, so each copies 1/world of the
    leading axis over its own PCIe link: conv layers are split over images and Dense layers with m > 2 N0 over samples
    (the Gram matrices are sums over samples: one NCCL all-reduce of them, replicate.py), the remaining Dense layers are
    completed by ONE all-gather over NVLink; the walks run on the rank's shard of channels / neurons.
    Returns (d2h bytes, h2d bytes that crossed this rank's host link)."""
    import torch
    from quantized_neural_networks_b200.replicate import (h2d_bytes_per_rank, image_split_conv_gram, prefer_sample_split,
                                                            replicate_leading_axis, sample_split_gram)
    d2h = h2d = 0
    for d in host:
        conv = d["kind"] == "conv"
        a, aq = (d["act"], d["actq"]) if conv else (d["X"], d["Xq"])
        if world == 1:
            Q = (eng.conv_layer_nhwc(a, aq, d["W"], d["A"], c0=d["c0"], n_channels=d["n_ch"]) if conv
                 else eng.dense_layer(a, aq, d["W"], d["A"], j0=d["j0"], j1=d["j1"]))
            h2d += a.nbytes + (0 if aq is None else aq.nbytes) + d["W"].nbytes
        elif conv:
            # image split: this rank's n_img / world images of every channel over its own PCIe link, one all-reduce of the
            # per-channel Grams, every channel walked on every rank (no replication, no Q exchange)
            gram = image_split_conv_gram(eng, a, aq, (3, 3), (1, 1), "SAME", (1, 1), rank, world)
            Q = eng.conv_layer_from_gram(gram, d["W"], d["A"])[:, :, d["c0"]:d["c0"] + d["n_ch"]]
            lo, hi = shard_range(a.shape[0], rank, world)
            h2d += (hi - lo) * int(np.prod(a.shape[1:])) * 4 * (1 if aq is None else 2) + d["W"].nbytes
        elif prefer_sample_split(d["N0"], d["m"], world):
            # Dense, m > 2 N0: this rank's m / world samples, all-reduce of the (N0, N0) Grams, walk of this rank's neurons
            G1, G2 = sample_split_gram(eng, a, aq, rank, world, device=dev)
            Q = eng.dense_layer_from_gram(G1, G2, d["W"], d["A"], j0=d["j0"], j1=d["j1"])[:, d["j0"]:d["j1"]]
            lo, hi = shard_range(a.shape[1], rank, world)
            h2d += d["N0"] * (hi - lo) * 4 * (1 if aq is None else 2) + d["W"].nbytes
        else:
            ad = replicate_leading_axis(a, rank, world, dev)
            aqd = None if aq is None else replicate_leading_axis(aq, rank, world, dev)
            Wd = torch.from_numpy(d["W"]).to(dev, non_blocking=True)
            h2d += h2d_bytes_per_rank(a.shape, 4, world) * 2 + d["W"].nbytes
            Qd = eng.dense_layer(ad, aqd, Wd, d["A"], j0=d["j0"], j1=d["j1"])
            Q = Qd[:, d["j0"]:d["j1"]].cpu().numpy()
            del ad, aqd, Qd
        d2h += (9 * d["n_ch"] * d["F"] if conv else d["N0"] * (d["j1"] - d["j0"])) * 8
        if keep is not None:
            keep.append(Q)
    return d2h, h2d


def host_channel_patches(act, c):
    """(9, n) patch matrix of channel c of an NHWC host array, 3x3 'same' stride 1 -- what _build_patch_array yields."""
    n_img, H = act.shape[0], act.shape[1]
    p = np.zeros((n_img, H + 2, H + 2), np.float32)
    p[:, 1:-1, 1:-1] = act[..., c]
    cols = np.empty((9, n_img * H * H), np.float32)
    for r in range(3):
        for cc in range(3):
            cols[r * 3 + cc] = p[:, r:r + H, cc:cc + H].reshape(-1)
    return cols


# ----------------------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------
# CPU baseline: the oracle's NumPy walk in a fork pool, one process per core (the reference's own fan-out)
# ----------------------------------------------------------------------------------------------------------------
_CPU = {}


def _cpu_neuron(args):
    key, j = args
    from oracle import gpfq_oracle as O
    d = _CPU[key]
    return O.quantize_neuron(d["W"][:, j], d["X"], d["Xq"], d["A"])


def cpu_sample_pass(host_layers, cores, budget_per_layer=2.5):
    """Times a bounded sample of every layer with `cores` worker processes; returns (extrapolated full-net seconds,
    description).  Neurons / filters are independent and equal-cost, so the extrapolation is linear in their count."""
    import concurrent.futures as cf
    import multiprocessing as mp
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    try:
        from threadpoolctl import threadpool_limits
        limiter = threadpool_limits(limits=1)
    except Exception:
        limiter = None
    _CPU.clear()
    jobs = []
    for li, d in enumerate(host_layers):
        if d["kind"] == "conv":
            if d.get("Xp"):
                X = d["Xp"][0]
                Xq = X if d["Xqp"] is None else d["Xqp"][0]
            else:
                X = host_channel_patches(d["act"], d["c0"])
                Xq = X if d["actq"] is None else host_channel_patches(d["actq"], d["c0"])
            Wc = np.ascontiguousarray(d["W"][:, :, d["c0"], :].reshape(9, d["F"]))
            units_total = d["n_ch"] * d["F"]
            take = min(d["F"], cores)
            weights_per_unit = 9
        else:
            X, Xq = d["X"], (d["X"] if d["Xq"] is None else d["Xq"])
            Wc = d["W"]
            units_total = d["j1"] - d["j0"]
            take = min(units_total, 2 * cores if d["N0"] >= 1024 else 4 * cores)
            weights_per_unit = d["N0"]
        _CPU[li] = dict(W=Wc, X=X, Xq=Xq, A=np.asarray(d["A"], dtype=np.float64))
        jobs.append((li, d["name"], take, units_total, weights_per_unit))
    total_s, desc, sampled_w = 0.0, [], 0
    ctx = mp.get_context("fork")
    with cf.ProcessPoolExecutor(max_workers=cores, mp_context=ctx) as ex:
        list(ex.map(_cpu_neuron, [(jobs[-1][0], 0)] * cores))  # start the workers before timing
        for li, name, take, units_total, wpu in jobs:
            t0 = time.perf_counter()
            list(ex.map(_cpu_neuron, [(li, j) for j in range(take)]))
            dt = time.perf_counter() - t0
            total_s += dt * units_total / take
            sampled_w += take * wpu
            desc.append(f"{name}:{take}/{units_total}")
    if limiter is not None:
        limiter.restore_original_limits()
    return total_s, sampled_w, "units timed per layer " + " ".join(desc)


# ----------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gpfq", choices=["gpfq", "reference"])
    ap.add_argument("--workload", default="cifar10_cnn", choices=["cifar10_cnn", "mnist_mlp"])
    ap.add_argument("--n-img", type=int, default=0, help="override the image/sample count (debug only)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-activations-leg", action="store_true", help="skip the from_activations leg (profiling the value leg)")
    ap.add_argument("--profile-activations-leg", action="store_true",
                    help="profiling only: run nothing but warm-up + timed passes of the from_activations leg")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    layers = CIFAR_LAYERS if args.workload == "cifar10_cnn" else MNIST_LAYERS
    n_img = args.n_img or (N_IMG if args.workload == "cifar10_cnn" else 25000)
    total_weights = sum(layer_weights(l) for l in layers)
    cores = len(os.sched_getaffinity(0))
    config = {"workload": f"{args.workload}: " + ("CIFAR10 CNN 6x Conv2D 3x3 (per-channel patch matrices 9 x n_patches) + Dense 2048->128->10"
                                                    if args.workload == "cifar10_cnn" else "MNIST MLP 784-500-300-10"),
              "samples": n_img, "alphabet": f"bits={BITS} (K=16), alphabet_scalar={CSCALAR}" if args.workload == "cifar10_cnn"
              else "ternary", "weights_per_step": total_weights, "sharding": f"conv channels / dense neurons over {world} rank(s)",
              "l2": ("inputs (25.7 GB of patch matrices per pass) exceed L2; no flush needed" if args.workload == "cifar10_cnn"
                     else "inputs (240 MB per pass) exceed the 126 MB L2; no flush needed")}

    import torch
    if args.impl == "reference":
        # The reference's own CPU implementation of the path: Python/NumPy, does not travel -> the oracle port, run
        # exactly as BASELINE.md section 3 prescribes.  Rank 0 alone works.
        if rank != 0:
            return
        dev = torch.device("cuda", local_rank) if torch.cuda.is_available() else torch.device("cpu")
        data = build_inputs_for_cpu(layers, n_img)
        vals = []
        for it in range(args.warmup + args.steps):
            sec, sampled_w, desc = cpu_sample_pass(data, cores, 2.0)
            if it >= args.warmup:
                vals.append(total_weights / sec)
            if it == 0 and args.warmup + args.steps > 2 and sec > 0:
                pass
        v = float(np.mean(vals))
        line = {"impl": "reference", "metric": "quantized weights/s, full-network GPFQ pass", "value": v, "unit": "weights/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_weights / v * 1e3,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": v, "unit": "weights/s", "cores": cores, "kind": "port",
                                 "sample": desc + "; extrapolated linearly in neurons/filters; HDF5 I/O excluded"},
                "e2e": {"value": v, "unit": "weights/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback for the GPFQ hot path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from quantized_neural_networks_b200 import get_engine
    eng = get_engine(local_rank)

    data = build_device_inputs(layers, n_img, rank, world, dev)
    outs = []
    for d in data:
        if d["kind"] == "conv":
            outs.append(torch.zeros((1, 3, 3, d["C"], d["F"]), dtype=torch.float64, device=dev))
        else:
            outs.append(torch.zeros((1, d["N0"], d["N1"]), dtype=torch.float64, device=dev))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.profile_activations_leg:
        for _ in range(max(args.warmup, 3) + args.steps):
            run_pass_device_nhwc(eng, data, outs)
        barrier()
        return
    for _ in range(max(args.warmup, 3)):
        run_pass_device(eng, data, outs)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        e0.record()
        for _ in range(args.steps):
            run_pass_device(eng, data, outs)
        e1.record()
        barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = total_weights / (ms_per_step * 1e-3)

    # per-call stage times of the timed region (most recent calls first)
    ncall = min(len(layers) * args.steps, 120)
    per_layer = {}
    launches_per_step = 0
    for back in range(ncall):
        st = eng.query_stats(back)
        name = layers[(len(layers) - 1 - back) % len(layers)][0]
        per_layer.setdefault(name, []).append(st)
    conv_bytes = conv_ms = 0.0
    layer_report = {}
    for name, sts in per_layer.items():
        kind = dict((l[0], l[1]) for l in layers)[name]
        launches_per_step += sts[0]["kernel_launches"]
        msl = float(np.mean([s["ms_total"] for s in sts]))
        layer_report[name] = {"ms": round(msl, 4), "weights_per_s": round(sts[0]["weights"] / (msl * 1e-3)) if msl > 0 else None,
                              "method": {1: "stream", 2: "gram", 3: "stream_fast"}.get(sts[0]["method"]),
                              "gram_kernel": {1: "dmma", 2: "i8_tcgen05", 3: "block_diagonal"}.get(sts[0].get("gram_kernel"))}
        if kind == "conv":
            conv_bytes += sum(s["bytes_algorithmic"] for s in sts)
            conv_ms += sum(s["ms_gram"] for s in sts)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = conv_bytes / (conv_ms * 1e-3) / 1e9 if conv_ms > 0 else None
    # DRAM bytes ncu measured for three launches of this kernel (conv8, conv12, conv14; profiles/r1c_conv_gram9_tma.md)
    # against their algorithmic bytes: 10.352e9 vs 10.339e9 -- every byte is read exactly once
    roofline = {"kernel": "conv_gram9_tma_kernel (3x3 per-channel patch Grams: UBLKCP/mbarrier ring, DMMA corners + DFMA edge; "
                          "74 % of the step, profiles/r1g_bench_launches.md)", "bound": "hbm",
                "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": (achieved / hbm_peak) if achieved else None,
                "traffic": 10.352e9, "traffic_note": "dram read+write of the conv8+conv12+conv14 launches (ncu --set full); "
                                                     "algorithmic bytes of the same launches: 10.339e9",
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (copy)" if peaks else "fallback (B200_PROFILING.md)",
                "algorithmic_bytes": "72 B per patch column per channel (36 when X == Xq); the stage time is CUDA events around "
                                     "the Gram launches of every conv call of the timed region"}
    # the one tensor-core stage of the pass: the Dense Gram (dense19, 2048 x 2048 over m = 5008) as int8 slices on tcgen05
    tensor = None
    for name, sts in per_layer.items():
        kind = dict((l[0], l[1]) for l in layers)[name]
        if kind == "dense" and sts[0].get("gram_kernel") == 2 and sts[0]["ms_gram"] > 0:
            d = [x for x in data if x["name"] == name][0]
            N0, m = d["N0"], d["m"]
            tiles = sum((ti >> 1) + 1 for ti in range(-(-N0 // 128)))
            ops = 15 * tiles * 128 * 256 * 2 * (-(-m // 128) * 128) * (1 if d["first"] else 2)
            ms_g = float(np.mean([x["ms_gram"] for x in sts]))
            peak_i8 = 2.0 * float(peaks.get("bf16_tflops", 1590.0))
            tensor = {"kernel": "gram_i8_kernel (tcgen05.mma kind::i8 + TMA + TMEM; 15 int8 slice pairs)", "layer": name, "bound": "tensor",
                      "achieved": ops / (ms_g * 1e-3) / 1e12, "peak": peak_i8, "unit": "TOP/s (int8)",
                      "frac": ops / (ms_g * 1e-3) / 1e12 / peak_i8,
                      "peak_source": "2 x MEASURED_PEAKS.json bf16_tflops (int8 issues at twice the bf16 rate; no int8 peak is measured)",
                      "fp64_equivalent_tflops": (1 if d["first"] else 2) * m * N0 * (N0 + 1) / (ms_g * 1e-3) / 1e12,
                      "note": "stage time includes the slicing and exponent kernels; a 0.5 ms stage is mostly fill/drain"}

    # ---- the same pass from the layers' NHWC activations (device-resident): no patch matrices anywhere -----------------
    nhwc = None
    if args.workload == "cifar10_cnn" and not args.no_activations_leg:
        outs2 = [torch.zeros_like(o) for o in outs]
        for _ in range(max(args.warmup, 3)):
            run_pass_device_nhwc(eng, data, outs2, rank=rank, world=world)
        barrier()
        def _agree(a, b):   # world > 1: `outs` holds this rank's channels / neurons only, the image-split pass every channel
            m = (a != 0) if world > 1 else torch.ones_like(a, dtype=torch.bool)
            return float((a == b)[m].double().mean()) if bool(m.any()) else 1.0
        agree = min(_agree(a, b) for a, b in zip(outs, outs2))
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        f0.record()
        for _ in range(args.steps):
            run_pass_device_nhwc(eng, data, outs2, rank=rank, world=world)
        f1.record()
        barrier()
        ms2 = f0.elapsed_time(f1)
        if world > 1:
            t = torch.tensor([ms2], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms2 = float(t.item())
        ms2 /= args.steps
        cb = cm = cf = 0.0
        kinds = {}
        for back in range(min(len(layers) * args.steps, 120)):
            st = eng.query_stats(back)
            lname, lkind = layers[(len(layers) - 1 - back) % len(layers)][:2]
            if lkind == "conv" and st.get("gram_kernel") in (4, 5):
                cb += st["bytes_algorithmic"]
                cf += st["flops_algorithmic"]
                cm += st["ms_gram"]
            if lkind == "conv":
                kinds[lname] = {0: "patch form (shared-memory planes)", 4: "correlation form", 5: "correlation form, packed images"}.get(
                    st.get("gram_kernel"), str(st.get("gram_kernel")))
        nhwc = {"value": total_weights / (ms2 * 1e-3), "unit": "weights/s", "ms_per_step": ms2,
                "inputs": "NHWC activations of every conv layer resident in HBM (2.9 GB per pass, larger than L2) instead of "
                          "per-channel patch matrices (25.7 GB)",
                "agreement_with_patch_matrix_pass": agree, "conv_gram_form": kinds,
                "roofline": None if cm <= 0 else {
                    "kernel": "conv_corr9_tma_kernel (13 displacement sums per Gram; TMA boxes -> per-warp mbarrier ring -> "
                              "fp64 register window, DFMA)", "bound": "hbm", "achieved": cb / (cm * 1e-3) / 1e9,
                    "peak": hbm_peak, "unit": "GB/s", "frac": cb / (cm * 1e-3) / 1e9 / hbm_peak,
                    "algorithmic_bytes": "4 B per pixel and channel per tensor (the activations, read once)",
                    "dfma_pipe_frac": cf / 2 / (cm * 1e-3) / 17.05e12,
                    "dfma_pipe_note": "13 DFMA per pixel, channel and Gram against the measured 17.05e12 DFMA/s "
                                      "(profiles/fp64_pipes_r1.txt); the stage time includes image packing, row launches and assembly"}}
        del outs2

    # ---- e2e: host (pinned) buffers through the C ABI, copies inside the timed region ---------------------------
    e2e = None
    if not args.no_e2e:
        host, _ = to_host_pinned(data)
        kept = []
        run_pass_host(eng, host, kept, rank, world, dev)  # warm-up (allocates the staging workspaces)
        barrier()
        for d, o, Qh in zip(data, outs, kept):   # both entry points must agree (north-star gate: >= 99.99 % of entries;
            Qd = o[0].cpu().numpy()               # the correlation form re-associates fp64 sums, so not always bit for bit)
            if d["kind"] == "conv":
                ref_blk = Qd[:, :, d["c0"]:d["c0"] + d["n_ch"]]
                got = Qh[:, :, d["c0"]:d["c0"] + d["n_ch"]] if world == 1 else Qh
            else:
                ref_blk = Qd[:, d["j0"]:d["j1"]]
                got = Qh[:, d["j0"]:d["j1"]] if world == 1 else Qh
            assert ref_blk.size == 0 or float(np.mean(ref_blk == got)) >= 0.9999, d["name"]
        del kept
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            d2h_bytes, h2d_bytes = run_pass_host(eng, host, None, rank, world, dev)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": total_weights * args.e2e_steps / dt, "unit": "weights/s", "h2d_bytes_per_step": int(h2d_bytes),
               "d2h_bytes_per_step": int(d2h_bytes), "steps": args.e2e_steps, "ms_per_step": dt / args.e2e_steps * 1e3,
               "timing": "host wall clock around synchronous C-ABI calls (copies + kernels), max over ranks",
               "api": "gpfq_conv_layer_nhwc + gpfq_dense_layer from pinned host buffers" +
                      ("" if world == 1 else f"; per rank 1/{world} of every input over PCIe: conv layers split over images and Dense "
                                             "layers with m > 2 N0 over samples (one NCCL all-reduce of the Gram matrices each), other "
                                             "Dense layers replicated by one all-gather over NVLink (h2d/d2h bytes are per rank)")}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        host_cpu = host if e2e is not None else to_host_pinned(data)[0]
        sec, sampled_w, desc = cpu_sample_pass(host_cpu, cores)
        cpu = {"value": total_weights / sec, "unit": "weights/s", "cores": cores, "kind": "port",
               "sample": desc + "; extrapolated linearly in neurons/filters; HDF5 I/O excluded",
               "extrapolated_full_pass_s": sec}

    if rank == 0:
        line = {"metric": "quantized weights/s, full-network GPFQ pass", "value": value, "unit": "weights/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": int(launches_per_step * args.steps),
                "roofline": roofline, "roofline_tensor_stage": tensor, "from_activations": nhwc, "cpu_baseline": cpu,
                "layers": layer_report}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def build_inputs_for_cpu(layers, n_img):
    """Host-only synthetic inputs for the reference arm: one channel of patches per conv layer is enough for the sample."""
    rng = np.random.default_rng(0)
    out = []
    for li, (name, kind, a, b, H) in enumerate(layers):
        if kind == "conv":
            C, F = a, b
            W = (rng.uniform(-1, 1, (3, 3, C, F)) * np.sqrt(6.0 / (9 * C))).astype(np.float32)
            A = make_alphabet(np.median(np.abs(W)))
            n = n_img * H * H
            first = li == 0
            shape = (n_img, H, H)
            if first:
                act = (rng.random(shape, dtype=np.float32) * (rng.random(shape, dtype=np.float32) < 0.5)).astype(np.float32)
                acts = [act]
            else:
                Z = rng.standard_normal(shape, dtype=np.float32)
                acts = [np.maximum(Z, 0), np.maximum(Z + 0.05 * rng.standard_normal(shape, dtype=np.float32), 0)]
            mats = []
            for t in acts:
                p = np.zeros((n_img, H + 2, H + 2), np.float32)
                p[:, 1:-1, 1:-1] = t
                cols = np.empty((9, n), np.float32)
                for r in range(3):
                    for c in range(3):
                        cols[r * 3 + c] = p[:, r:r + H, c:c + H].reshape(-1)
                mats.append(cols)
            out.append(dict(name=name, kind=kind, W=W, A=A, Xp=[mats[0]], Xqp=None if first else [mats[1]], c0=0, n_ch=C,
                            n=n, C=C, F=F))
        else:
            N0, N1 = a, b
            m = n_img
            W = (rng.uniform(-1, 1, (N0, N1)) * np.sqrt(6.0 / N0)).astype(np.float32)
            A = make_alphabet(np.median(np.abs(W)))
            Z = rng.standard_normal((N0, m), dtype=np.float32)
            X = np.maximum(Z, 0)
            Xq = np.maximum(Z + 0.05 * rng.standard_normal((N0, m), dtype=np.float32), 0)
            out.append(dict(name=name, kind=kind, W=W, A=A, X=X, Xq=Xq, j0=0, j1=N1, N0=N0, N1=N1, m=m))
    return out


if __name__ == "__main__":
    main()
