"""Per-shape timing of the Dense path (device-resident inputs, CUDA-event stage times from the library).

    python tools/dense_bench.py [--shapes N0xN1xm,...] [--methods gram_i8,gram_dmma,stream_fast,auto] [--reps 3]

Prints one JSON line per (shape, method): total / Gram / sweep / stream ms, weights/s, and for the Gram stage the
fp64-equivalent TFLOP/s (2*m*N0*(N0+1) flops per Gram pair over the lower triangle) plus, for the int8 tcgen05 kernel,
the int8 TOP/s actually issued (slice pairs x 128x256 tiles).
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="2048x128x5008,784x500x25000,1024x1024x5000,4096x4096x25000")
    ap.add_argument("--methods", default="gram_i8,gram_dmma,stream_fast,auto")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--first", action="store_true", help="X == Xq (first layer)")
    ap.add_argument("--bits", type=float, default=np.log2(3))
    ap.add_argument("--opt", default="", help="comma-separated key=value engine options (gpfq_set_option)")
    args = ap.parse_args()
    import torch
    from quantized_neural_networks_b200 import get_engine
    eng = get_engine(0)
    dev = torch.device("cuda", 0)
    for kv in [v for v in args.opt.split(",") if v]:
        eng.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    for shp in args.shapes.split(","):
        N0, N1, m = (int(v) for v in shp.split("x"))
        g = torch.Generator(device=dev).manual_seed(N0 + N1 + m)
        Z = torch.randn((N0, m), device=dev, generator=g)
        X = torch.relu(Z)
        Xq = None if args.first else torch.relu(Z + 0.05 * torch.randn((N0, m), device=dev, generator=g))
        del Z
        W = (torch.rand((N0, N1), device=dev, generator=g) * 2 - 1) * float(np.sqrt(6.0 / (N0 + N1)))
        A = 2 * float(torch.median(W.abs().flatten())) * np.linspace(-1, 1, int(round(2 ** args.bits)))
        out = torch.zeros((1, N0, N1), dtype=torch.float64, device=dev)
        ref = None
        for meth in args.methods.split(","):
            method, variant = {"gram_i8": ("gram", 2), "gram_dmma": ("gram", 1), "stream_fast": ("stream_fast", 0),
                               "stream": ("stream", 0), "auto": ("auto", 0)}[meth]
            if meth == "stream_fast" and 3.0 * m * N0 * N1 > 4e14:
                continue
            eng.set_option("gram_kernel", variant)
            best = None
            for _ in range(args.reps + 1):
                eng.dense_layer(X, Xq, W, A, method=method, out=out, sync=True)
                st = dict(eng.last_stats)
                if best is None or st["ms_total"] < best["ms_total"]:
                    best = st
            eng.set_option("gram_kernel", 0)
            Q = out[0].clone()
            agree = None
            if ref is None:
                ref = Q
            else:
                agree = float((Q == ref).double().mean())
            pairs = 2 if Xq is not None else 1
            gram_flops = pairs * m * N0 * (N0 + 1)
            line = {"shape": [N0, N1, m], "requested": meth, "method": {1: "stream", 2: "gram", 3: "stream_fast"}[best["method"]],
                    "gram_kernel": {0: None, 1: "dmma", 2: "i8_tcgen05", 3: "block_diagonal_dmma (residual outer level)"}[best["gram_kernel"]],
                    "ms_total": round(best["ms_total"], 4), "ms_gram": round(best["ms_gram"], 4), "ms_sweep": round(best["ms_sweep"], 4),
                    "ms_stream": round(best["ms_stream"], 4), "launches": best["kernel_launches"],
                    "weights_per_s": round(N0 * N1 / (best["ms_total"] * 1e-3)),
                    "gram_fp64_equiv_tflops": round(gram_flops / (best["ms_gram"] * 1e-3) / 1e12, 2) if best["ms_gram"] > 0 else None,
                    "agreement_with_first": agree}
            print(json.dumps(line), flush=True)
        del X, Xq, W, out
        torch.cuda.empty_cache()
        eng.trim()


if __name__ == "__main__":
    main()
