"""Quick per-shape timing probe (not the bench): prints library-reported CUDA-event times per stage."""
import os
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quantized_neural_networks_b200 import get_engine  # noqa: E402

HBM = 6545.3e9


def alphabet(W, bits, c):
    med = float(torch.median(W.abs().flatten()))
    return c * med * np.linspace(-1, 1, int(round(2 ** bits)))


def dense(eng, N0, N1, m, bits, c, same, methods=("stream", "stream_fast", "gram"), reps=2):
    g = torch.Generator(device="cuda").manual_seed(0)
    Z = torch.randn((N0, m), device="cuda", generator=g)
    X = torch.relu(Z)
    Xq = X if same else torch.relu(Z + 0.05 * torch.randn((N0, m), device="cuda", generator=g))
    W = (torch.rand((N0, N1), device="cuda", generator=g) * 2 - 1) * float(np.sqrt(6 / (N0 + N1)))
    A = alphabet(W, bits, c)
    res = {}
    for meth in methods:
        for r in range(reps):
            try:
                Q = eng.dense_layer(X, None if same else Xq, W, A, method=meth)
            except Exception as e:
                print(f"dense ({N0},{N1},{m}) {meth}: FAILED {e}")
                break
            st = dict(eng.last_stats)
        else:
            res[meth] = Q
            w = N0 * N1
            print(f"dense ({N0},{N1},{m}) same={same} K={len(A)} {meth:6s}: total {st['ms_total']:9.3f} ms  gram {st['ms_gram']:8.3f}"
                  f"  sweep {st['ms_sweep']:8.3f}  stream {st['ms_stream']:8.3f}  launches {st['kernel_launches']:5d}"
                  f"  {w / st['ms_total'] * 1e3:.3e} w/s  flops_alg {st['flops_algorithmic']:.3e}"
                  f" -> {st['flops_algorithmic'] / max(st['ms_gram'] if meth == 'gram' else st['ms_stream'], 1e-6) / 1e9:.1f} TF/s")
    names = list(res)
    for other in names[1:]:
        print(f"    {names[0]} vs {other} agreement: {float((res[names[0]] == res[other]).double().mean()):.6f}")
    del X, Xq, Z, W
    torch.cuda.empty_cache()


def conv(eng, n_img, H, C, F, bits, c, same, reps=2, nhwc=True):
    g = torch.Generator(device="cuda").manual_seed(1)
    act = torch.relu(torch.randn((n_img, H, H, C), device="cuda", generator=g))
    actq = act if same else torch.relu(act + 0.05 * torch.randn(act.shape, device="cuda", generator=g))
    W = (torch.rand((3, 3, C, F), device="cuda", generator=g) * 2 - 1) * float(np.sqrt(6 / (9 * C)))
    A = alphabet(W, bits, c)
    n = n_img * H * H
    for r in range(reps):
        Q = eng.conv_layer_nhwc(act, None if same else actq, W, A)
        st = dict(eng.last_stats)
    by = st["bytes_algorithmic"]
    print(f"conv nhwc ({n_img}x{H}x{H}x{C} -> {F}) same={same}: total {st['ms_total']:9.3f} ms gram(+im2col) {st['ms_gram']:9.3f}"
          f" sweep {st['ms_sweep']:7.3f} launches {st['kernel_launches']}  alg bytes {by:.3e} -> {by / st['ms_gram'] / 1e6:.1f} GB/s")
    # patch-matrix API with device-resident patches (a few channels at a time to bound memory)
    cb = min(C, max(1, int(6e9 // (72 * n))))
    pat = torch.nn.functional.unfold(act[..., :cb].permute(3, 0, 1, 2).reshape(cb * n_img, 1, H, H), 3, padding=1)
    Xp = pat.reshape(cb, n_img, 9, H * H).permute(0, 2, 1, 3).reshape(cb, 9, n).contiguous()
    if same:
        Xqp = None
    else:
        patq = torch.nn.functional.unfold(actq[..., :cb].permute(3, 0, 1, 2).reshape(cb * n_img, 1, H, H), 3, padding=1)
        Xqp = patq.reshape(cb, n_img, 9, H * H).permute(0, 2, 1, 3).reshape(cb, 9, n).contiguous()
    Wc = W[:, :, :cb, :].contiguous()
    for r in range(reps + 1):
        Q2 = eng.conv_channels(list(Xp), None if same else list(Xqp), Wc, A)
        st = dict(eng.last_stats)
    by = st["bytes_algorithmic"]
    print(f"conv patches ({cb} ch x 9 x {n}) same={same}: total {st['ms_total']:9.3f} ms gram {st['ms_gram']:9.3f}"
          f" -> {by / st['ms_gram'] / 1e6:.1f} GB/s = {by / st['ms_gram'] / 1e6 / 6545.3:.3f} of HBM peak; agreement with nhwc"
          f" {float((Q2 == Q[:, :, :cb, :]).double().mean()):.6f}")
    del act, actq, pat, Xp, Xqp
    torch.cuda.empty_cache()


if __name__ == "__main__":
    print(subprocess.run("nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv; free -g; nproc",
                         shell=True, capture_output=True, text=True).stdout)
    eng = get_engine(0)
    t0 = time.time()
    if len(sys.argv) > 1 and sys.argv[1] == "dense":
        # python tools/gpu_probe.py dense N0 N1 m bits c same method[,method...] [reps]
        N0, N1, m = (int(v) for v in sys.argv[2:5])
        bits, c, same = float(sys.argv[5]), float(sys.argv[6]), sys.argv[7] in ("1", "true", "same")
        dense(eng, N0, N1, m, bits, c, same, methods=tuple(sys.argv[8].split(",")), reps=int(sys.argv[9]) if len(sys.argv) > 9 else 2)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "sweep":   # BASELINE configs 1-4 Dense shapes
        shapes = [(784, 500, 25000, True), (500, 300, 25000, False), (300, 10, 25000, False), (2048, 128, 5008, False),
                  (128, 10, 5008, False), (4096, 4096, 1504, False), (4096, 1000, 1504, False), (25088, 4096, 1504, False),
                  (1024, 1024, 5000, False), (1024, 1024, 25000, False), (1024, 1024, 100000, False),
                  (4096, 4096, 5000, False), (4096, 4096, 25000, False), (4096, 4096, 100000, False)]
        for N0, N1, m, same in shapes:
            meths = ("stream_fast", "gram") if N0 * N0 * 16 < 40e9 and not (N0 > 8192) else ("stream_fast",)
            dense(eng, N0, N1, m, np.log2(3), 2, same, methods=meths, reps=1)
        sys.exit(0)
    dense(eng, 784, 500, 25000, np.log2(3), 2, True)
    dense(eng, 500, 300, 25000, np.log2(3), 2, False)
    dense(eng, 300, 10, 25000, np.log2(3), 2, False)
    dense(eng, 2048, 128, 5008, 4, 4, False)
    dense(eng, 128, 10, 5008, 4, 4, False)
    dense(eng, 4096, 4096, 1504, np.log2(3), 2, False)
    dense(eng, 1024, 1024, 25000, np.log2(3), 2, False)
    dense(eng, 4096, 4096, 25000, np.log2(3), 2, False, methods=("gram",), reps=1)
    conv(eng, 5008, 32, 3, 32, 4, 4, True)
    conv(eng, 5008, 32, 32, 32, 4, 4, False)
    conv(eng, 5008, 16, 64, 64, 4, 4, False)
    conv(eng, 5008, 8, 128, 128, 4, 4, False)
    print(f"probe wall {time.time() - t0:.1f} s")
