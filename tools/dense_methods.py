"""Stream-vs-Gram by measurement (BASELINE north_star: "benchmarked per layer shape, the faster one picked by measurement").

    python tools/dense_methods.py [--out profiles/dense_methods_r2.md] [--reps 3]

Every Dense shape of BASELINE configs 1-4 (device-resident synthetic inputs), every method the library has:
  stream_fast  the literal residual walk on exact products (dense_stream.cu)
  gram_i8      Gram stage on tcgen05 (int8 slices) + blocked sweep   [sweep_outer = 1: Gram rows]
  gram_dmma    Gram stage on the fp64 DMMA pipe + blocked sweep      [sweep_outer = 1]
  residual_i8  carried-residual sweep, contractions on tcgen05 (slgemm_i8.cu)   [sweep_outer = 2]
  residual_f64 carried-residual sweep, contractions on the fp64 DMMA pipe       [sweep_outer = 2, sweep_i8 = 2]
  auto         what gpfq_dense_layer picks (choose_dense_method + dense_uses_lowrank)
Best-of-`reps` CUDA-event time of the whole call.  tests/test_gpu_parity.py::test_auto_picks_a_measured_best_method asserts that
auto stays within 10 % (+ 50 us) of the best measured method on the same list.
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

# (config, N0, N1, m, first layer, bits, alphabet_scalar)
SHAPES = [("1 MNIST", 784, 500, 25000, True, np.log2(3), 3), ("1 MNIST", 500, 300, 25000, False, np.log2(3), 3),
          ("1 MNIST", 300, 10, 25000, False, np.log2(3), 3), ("2 CIFAR", 2048, 128, 5008, False, 4, 4),
          ("2 CIFAR", 128, 10, 5008, False, 4, 4), ("3 VGG16", 25088, 4096, 1504, False, np.log2(3), 3),
          ("3 VGG16", 4096, 4096, 1504, False, np.log2(3), 3), ("3 VGG16", 4096, 1000, 1504, False, np.log2(3), 3),
          ("4 sweep", 1024, 1024, 5000, False, np.log2(3), 3), ("4 sweep", 1024, 1024, 25000, False, np.log2(3), 3),
          ("4 sweep", 4096, 4096, 5000, False, np.log2(3), 3), ("4 sweep", 4096, 4096, 25000, False, np.log2(3), 3),
          ("4 sweep", 16384, 16384, 5000, False, np.log2(3), 3)]
METHODS = {"stream_fast": ("stream_fast", {}), "gram_i8": ("gram", {"gram_kernel": 2, "sweep_outer": 1}),
           "gram_dmma": ("gram", {"gram_kernel": 1, "sweep_outer": 1}), "residual_i8": ("gram", {"sweep_outer": 2}),
           "residual_f64": ("gram", {"sweep_outer": 2, "sweep_i8": 2}), "auto": ("auto", {})}


def inputs(N0, N1, m, first, bits, c, dev):
    import torch
    g = torch.Generator(device=dev).manual_seed(N0 + N1 + m)
    X = torch.empty((N0, m), device=dev)
    Xq = None if first else torch.empty((N0, m), device=dev)
    step = max(1, (1 << 27) // m)
    for t0 in range(0, N0, step):
        n = min(step, N0 - t0)
        if first:
            X[t0:t0 + n] = torch.rand((n, m), device=dev, generator=g) * (torch.rand((n, m), device=dev, generator=g) < 0.5)
        else:
            z = torch.randn((n, m), device=dev, generator=g)
            X[t0:t0 + n] = torch.relu(z)
            Xq[t0:t0 + n] = torch.relu(z + 0.05 * torch.randn((n, m), device=dev, generator=g))
    W = (torch.rand((N0, N1), device=dev, generator=g) * 2 - 1) * float(np.sqrt(6.0 / (N0 + N1)))
    A = c * float(torch.median(W.abs().flatten())) * np.linspace(-1, 1, int(round(2 ** bits)))
    return X, Xq, W, A


def skip(name, N0, N1, m):
    if name == "stream_fast" and 3.0 * m * N0 * N1 > 2e14:
        return True              # minutes on the fp64 pipe
    if name in ("gram_i8", "gram_dmma") and 16.0 * N0 * N0 > 40e9:
        return True
    if name == "gram_dmma" and float(m) * N0 * N0 > 3e13:
        return True
    if name == "residual_f64" and 3.0 * m * N0 * N1 > 4e13:
        return True
    return False


def measure(eng, shape, reps, methods=METHODS):
    import torch
    cfg, N0, N1, m, first, bits, c = shape
    dev = torch.device("cuda", eng.device)
    X, Xq, W, A = inputs(N0, N1, m, first, bits, c, dev)
    out = torch.zeros((1, N0, N1), dtype=torch.float64, device=dev)
    res, ref = {}, None
    for name, (method, opts) in methods.items():
        if skip(name, N0, N1, m):
            continue
        for k, v in opts.items():
            eng.set_option(k, v)
        try:
            best = None
            for _ in range(reps + 1):
                eng.dense_layer(X, Xq, W, A, method=method, out=out, sync=True)
                st = dict(eng.last_stats)
                if best is None or st["ms_total"] < best["ms_total"]:
                    best = st
        finally:
            for k in opts:
                eng.set_option(k, 0)
        Q = out[0].clone()
        if ref is None:
            ref = Q
        res[name] = {"ms": best["ms_total"], "ms_gram": best["ms_gram"], "ms_sweep": best["ms_sweep"], "ms_stream": best["ms_stream"],
                     "picked": {1: "stream", 2: "gram", 3: "stream_fast"}[best["method"]] + (
                         "" if best["method"] != 2 else {1: "/dmma", 2: "/i8", 3: "/residual"}.get(best["gram_kernel"], "")) + (
                         "+i8" if best.get("reserved", 0) & 1 else ""),
                     "agreement": float((Q == ref).double().mean())}
    del X, Xq, W, out
    torch.cuda.empty_cache()
    eng.trim()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    from quantized_neural_networks_b200 import get_engine
    eng = get_engine(0)
    names = list(METHODS)
    lines = ["# Dense methods by measurement (round 2)", "",
             "`python tools/dense_methods.py` on one B200: best-of-%d CUDA-event time of the whole `gpfq_dense_layer` call in ms, device-resident "
             "synthetic inputs (SURVEY.md 8d).  `auto` = what the library picks; **bold** = fastest measured.  Every method agrees with the "
             "first one on >= 99.99 %% of the entries (last column: the minimum)." % args.reps, "",
             "| config | (N0, N1, m) | " + " | ".join(names) + " | auto picked | auto / best | min agreement |", "|---|---|" + "---|" * (len(names) + 3)]
    for shape in SHAPES:
        res = measure(eng, shape, args.reps)
        print(json.dumps({"shape": shape[1:4], **res}), flush=True)
        best = min(v["ms"] for k, v in res.items() if k != "auto")
        cells = []
        for n in names:
            if n not in res:
                cells.append("--")
            else:
                v = res[n]["ms"]
                cells.append(f"**{v:.3f}**" if v == best and n != "auto" else f"{v:.3f}")
        lines.append(f"| {shape[0]} | ({shape[1]}, {shape[2]}, {shape[3]}){' X==Xq' if shape[4] else ''} | " + " | ".join(cells) +
                     f" | {res['auto']['picked']} | {res['auto']['ms'] / best:.2f} | {min(v['agreement'] for v in res.values()):.6f} |")
    text = "\n".join(lines) + "\n"
    if args.out:
        open(args.out, "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
