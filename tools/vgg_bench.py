"""VGG16-shaped full-network GPFQ pass (BASELINE.json configs[2]: quantize_pretrained_imagenet.py, ternary), device-resident
synthetic inputs, random-init weights -- per-layer times and roofline fractions.

    python tools/vgg_bench.py [--n-img 1504] [--reps 2] [--gpus-emulated 1]

Conv layers go through gpfq_conv_layer_nhwc (activations (n_img, H, W, C): at 1504 images the per-channel patch matrices of
block 1 would be 348 GB, the activations are 19 GB), Dense layers through gpfq_dense_layer.  One JSON line per layer + a total.
Roofline per layer: conv -> the fp64 pipe (168 slots per patch column and channel, 18.55e12 slots/s measured) and the HBM bytes
actually needed (8 B per column and channel); Dense -> fp64 pipe for the sweep (N0^2 N1 MACs), int8 tensor pipe for the Gram.
`--shard r/w` runs rank r's share of every layer (channels / neurons) to emulate one rank of a w-GPU job on one GPU.
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CONV = [(3, 64, 224), (64, 64, 224), (64, 128, 112), (128, 128, 112), (128, 256, 56), (256, 256, 56), (256, 256, 56),
        (256, 512, 28), (512, 512, 28), (512, 512, 28), (512, 512, 14), (512, 512, 14), (512, 512, 14)]
DENSE = [(25088, 4096), (4096, 4096), (4096, 1000)]
FP64_SLOTS = 18.55e12   # DMMA.8x8x4 (profiles/fp64_pipes_r1.txt)
DFMA_SLOTS = 17.05e12   # DFMA, same pipe


def shard_range(n, rank, world):
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-img", type=int, default=1504)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--shard", default="0/1")
    ap.add_argument("--net", default="vgg16", choices=["vgg16", "cifar"], help="layer shapes: VGG16 (default) or the CIFAR10 CNN of config 2 (use --n-img 5008)")
    ap.add_argument("--conv-split", default="images", choices=["images", "channels"],
                    help="with --shard r/w: conv layers split over images (Gram-only call on n_img / w images of every channel + "
                         "the walks of every channel; the all-reduce of C x 162 doubles is not emulated) or over channels")
    ap.add_argument("--skip-conv", action="store_true")
    ap.add_argument("--skip-dense", action="store_true")
    ap.add_argument("--corr-rows", type=int, default=0, help="gpfq_set_option corr_rows (0 auto, 4, 6, 8)")
    ap.add_argument("--layers", default="", help="comma-separated conv layer indices to run (default: all)")
    ap.add_argument("--conv-kernel", type=int, default=0, help="gpfq_set_option conv_kernel: 0 correlation form, 3 planes kernel")
    args = ap.parse_args()
    rank, world = (int(v) for v in args.shard.split("/"))
    global CONV, DENSE
    if args.net == "cifar":
        CONV = [(3, 32, 32), (32, 32, 32), (32, 64, 16), (64, 64, 16), (64, 128, 8), (128, 128, 8)]
        DENSE = [(2048, 128), (128, 10)]
    import torch
    from quantized_neural_networks_b200 import get_engine
    eng = get_engine(0)
    eng.set_option("conv_kernel", args.conv_kernel)
    eng.set_option("corr_rows", args.corr_rows)
    dev = torch.device("cuda", 0)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0)) * 1e9
    i8_peak = 2.0 * float(peaks.get("bf16_tflops", 1590.0)) * 1e12
    alph = lambda W: 3 * float(torch.median(W.abs().flatten())) * np.linspace(-1, 1, 3)
    total_ms, total_w = 0.0, 0
    if not args.skip_conv:
        only = {int(v) for v in args.layers.split(',') if v != ''}
        for li, (C, F, H) in enumerate(CONV):
            if only and li not in only:
                continue
            g = torch.Generator(device=dev).manual_seed(100 + li)
            shape = (args.n_img, H, H, C)
            if li == 0:
                act = torch.rand(shape, device=dev, generator=g)
                actq = None
            else:
                act = torch.empty(shape, device=dev)
                actq = torch.empty(shape, device=dev)
                step = max(1, args.n_img // 8)
                for i0 in range(0, args.n_img, step):   # chunked: the temporaries of a 19 GB randn would double the footprint
                    z = torch.randn((min(step, args.n_img - i0), H, H, C), device=dev, generator=g)
                    act[i0:i0 + step] = torch.relu(z)
                    actq[i0:i0 + step] = torch.relu(z + 0.05 * torch.randn(z.shape, device=dev, generator=g))
                    del z
            W = (torch.rand((3, 3, C, F), device=dev, generator=g) * 2 - 1) * float(np.sqrt(6.0 / (9 * C)))
            A = alph(W)
            lo, hi = shard_range(C, rank, world)
            if hi == lo and not (world > 1 and args.conv_split == "images"):   # fewer channels than ranks: nothing to do
                del act, actq, W
                continue
            out = torch.zeros((1, 3, 3, C, F), dtype=torch.float64, device=dev)
            best = None
            img_split = world > 1 and args.conv_split == "images"
            ilo, ihi = shard_range(args.n_img, rank, world)
            for _ in range(args.reps):
                if img_split:
                    gram = eng.conv_gram_nhwc(act[ilo:ihi], None if actq is None else actq[ilo:ihi], (3, 3))
                    st = dict(eng.last_stats)
                    eng.conv_layer_from_gram(gram, W, A, out=out, sync=True)
                    st["ms_total"] += eng.last_stats["ms_total"]
                else:
                    eng.conv_layer_nhwc(act, actq, W, A, c0=lo, n_channels=hi - lo, out=out, sync=True)
                    st = dict(eng.last_stats)
                if best is None or st["ms_total"] < best["ms_total"]:
                    best = st
            if img_split:
                lo, hi = 0, C
            colch = (ihi - ilo if img_split else args.n_img) * H * H * (hi - lo)
            corr = best["gram_kernel"] in (4, 5)
            # patch form: 168 (84 when X == Xq) fp64-pipe slots per patch column and channel at the DMMA rate; correlation
            # form: 26 (13) DFMAs per pixel and channel at the measured DFMA rate
            slots = colch * ((13 if li == 0 else 26) if corr else (84 if li == 0 else 168))
            pipe = DFMA_SLOTS if corr else FP64_SLOTS
            bytes_alg = colch * (4 if li == 0 else 8)
            ms = best["ms_total"]
            print(json.dumps({"layer": f"conv{li}", "C": C, "F": F, "H": H, "channels": [lo, hi],
                              "images": [ilo, ihi] if img_split else [0, args.n_img], "weights": 9 * (hi - lo) * F,
                              "ms": round(ms, 3), "ms_gram": round(best["ms_gram"], 3),
                              "form": ("correlation, packed images" if best["gram_kernel"] == 5 else "correlation (13 DFMA / pixel / Gram)") if corr else "patch (126 MACs / column)",
                              "fp64_pipe_frac": round(slots / pipe / (best["ms_gram"] * 1e-3), 3),
                              "hbm_frac": round(bytes_alg / hbm / (best["ms_gram"] * 1e-3), 3),
                              "weights_per_s": round(9 * (hi - lo) * F / (ms * 1e-3))}), flush=True)
            total_ms += ms
            total_w += 9 * (hi - lo) * F
            del act, actq, W, out
            torch.cuda.empty_cache()
            eng.trim()
    m = args.n_img
    for li, (N0, N1) in enumerate([] if args.skip_dense else DENSE):
        g = torch.Generator(device=dev).manual_seed(200 + li)
        Z = torch.randn((N0, m), device=dev, generator=g)
        X = torch.relu(Z)
        Xq = torch.relu(Z + 0.05 * torch.randn((N0, m), device=dev, generator=g))
        del Z
        W = (torch.rand((N0, N1), device=dev, generator=g) * 2 - 1) * float(np.sqrt(6.0 / (N0 + N1)))
        A = alph(W)
        lo, hi = shard_range(N1, rank, world)
        out = torch.zeros((1, N0, N1), dtype=torch.float64, device=dev)
        best = None
        for _ in range(args.reps):
            eng.dense_layer(X, Xq, W, A, j0=lo, j1=hi, out=out, sync=True)
            st = dict(eng.last_stats)
            if best is None or st["ms_total"] < best["ms_total"]:
                best = st
        ms = best["ms_total"]
        nj = hi - lo
        line = {"layer": f"fc{li + 1}", "N0": N0, "N1": N1, "m": m, "neurons": [lo, hi], "weights": N0 * nj, "ms": round(ms, 3),
                "method": {1: "stream", 2: "gram", 3: "stream_fast"}[best["method"]], "ms_gram": round(best["ms_gram"], 3),
                "ms_sweep": round(best["ms_sweep"], 3), "ms_stream": round(best["ms_stream"], 3),
                "weights_per_s": round(N0 * nj / (ms * 1e-3))}
        if best["method"] == 2:
            macs = 3 * m * N0 * nj if best["gram_kernel"] == 3 else N0 * N0 * nj   # residual (low-rank) vs Gram-row outer level
            line["sweep_form"] = "carried residuals: 3 m N0 N1 MACs" if best["gram_kernel"] == 3 else "Gram rows: N0^2 N1 MACs"
            line["sweep_fp64_pipe_frac"] = round(macs / FP64_SLOTS / (best["ms_sweep"] * 1e-3), 3)
            if best["gram_kernel"] == 2:
                tiles = sum((ti >> 1) + 1 for ti in range(-(-N0 // 128)))
                ops = 15 * tiles * 128 * 256 * 2 * (-(-m // 128) * 128) * 2
                line["gram_int8_frac_of_2x_bf16"] = round(ops / i8_peak / (best["ms_gram"] * 1e-3), 3)
        print(json.dumps(line), flush=True)
        total_ms += ms
        total_w += N0 * nj
        del X, Xq, W, out
        torch.cuda.empty_cache()
        eng.trim()
    print(json.dumps({"total_ms": round(total_ms, 2), "weights": total_w, "weights_per_s": round(total_w / (total_ms * 1e-3)),
                      "n_img": args.n_img, "shard": args.shard}))


if __name__ == "__main__":
    main()
