"""GPU: parity with the LITERAL C oracle at the sizes bench.py measures (BASELINE.json configs 2, 3, 4).

The literal walk (oracle/gpfq_oracle.c: quantized_network.py:91-121, :185-233) costs N0 x m per neuron, so the oracle runs
on a seeded SUBSET of neurons / filters -- neurons are independent (quantized_network.py:553-556), so a subset at full
(N0, m) exercises everything the full layer does: every K chunk of the tcgen05 Gram, every carried-residual range of the
low-rank sweep, every band of the correlation-form conv kernel.  Bars as in test_gpu_parity.py.
"""
import numpy as np
import pytest

from oracle import c_oracle, gpfq_oracle as O

pytestmark = pytest.mark.gpu
AGREE = 0.9999
RESID_RTOL = 1e-6


def _hidden(N0, m, seed, device="cuda"):
    """X = relu(Z), Xq = relu(Z + 0.05 N) as float32 CUDA tensors (torch is only the generator / buffer)."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    X = torch.empty((N0, m), device=device)
    Xq = torch.empty((N0, m), device=device)
    step = max(1, (1 << 27) // m)
    for t0 in range(0, N0, step):
        n = min(step, N0 - t0)
        z = torch.randn((n, m), device=device, generator=g)
        X[t0:t0 + n] = torch.relu(z)
        Xq[t0:t0 + n] = torch.relu(z + 0.05 * torch.randn((n, m), device=device, generator=g))
    return X, Xq


def _check_subset(Q, W, X, Xq, A, cols):
    """Q: (N0, N1) result of the CUDA path; literal oracle on columns `cols` at full size."""
    cols = list(cols)
    Ws = np.ascontiguousarray(W[:, cols])
    Qref = c_oracle.quantize_layer(Ws, X, Xq, A)
    Qs = np.asarray(Q)[:, cols]
    agree = O.agreement(Qs, Qref)
    assert agree >= AGREE, f"agreement {agree} on {len(cols)} neurons"
    if agree < 1.0:
        r, rref = O.relative_residual(Ws, Qs, X, Xq), O.relative_residual(Ws, Qref, X, Xq)
        assert abs(r - rref) <= RESID_RTOL * max(rref, 1e-30), (r, rref)
    return agree


def _dense_case(engine, N0, N1, m, n_check, seed, bits, c, variants):
    import torch
    Xd, Xqd = _hidden(N0, m, seed)
    rng = np.random.default_rng(seed)
    W = (rng.uniform(-1, 1, (N0, N1)) * np.sqrt(6.0 / (N0 + N1))).astype(np.float32)
    Wd = torch.from_numpy(W).cuda()
    A = O.layer_alphabet(W, c, O.unit_alphabet(bits))
    X, Xq = Xd.cpu().numpy(), Xqd.cpu().numpy()
    cols = rng.choice(N1, size=min(n_check, N1), replace=False)
    results = {}
    for name, opts, method in variants:
        for k, v in opts.items():
            engine.set_option(k, v)
        try:
            Q = engine.dense_layer(Xd, Xqd, Wd, A, method=method).cpu().numpy()
            st = dict(engine.last_stats)
        finally:
            for k in opts:
                engine.set_option(k, 0)
        results[name] = (Q, st)
    # the oracle runs once; every variant is compared with it
    Ws = np.ascontiguousarray(W[:, cols])
    Qref = c_oracle.quantize_layer(Ws, X, Xq, A)
    for name, (Q, st) in results.items():
        agree = O.agreement(Q[:, cols], Qref)
        assert agree >= AGREE, (name, agree, st)
        if agree < 1.0:
            r, rref = O.relative_residual(Ws, Q[:, cols], X, Xq), O.relative_residual(Ws, Qref, X, Xq)
            assert abs(r - rref) <= RESID_RTOL * max(rref, 1e-30), (name, r, rref)
    engine.trim()
    return results


def test_cifar_dense19_full_size(engine):
    """CIFAR10 CNN layer 19: (2048, 128), m = 5008, 4-bit (bench --workload cifar10_cnn)."""
    res = _dense_case(engine, 2048, 128, 5008, 16, 19, 4, 4,
                      [("auto", {}, "auto"), ("gram_i8", {"gram_kernel": 2}, "gram"), ("gram_dmma", {"gram_kernel": 1}, "gram"),
                       ("stream_fast", {}, "stream_fast")])
    assert res["gram_i8"][1]["gram_kernel"] == 2 and res["gram_dmma"][1]["gram_kernel"] == 1


def test_vgg_fc1_full_size_lowrank_sweep(engine):
    """VGG16 fc1: N0 = 25088, m = 1504, 2048 of the 4096 neurons (the two-stream split of the carried-residual sweep needs
    >= 2048), ternary; 49 carried-residual ranges.  Both stream arrangements and the int8 / fp64 contraction paths."""
    res = _dense_case(engine, 25088, 2048, 1504, 8, 31, np.log2(3), 3,
                      [("auto", {}, "auto"), ("one_chain", {"sweep_outer": 3}, "gram"),
                       ("fp64_contractions", {"sweep_i8": 2}, "gram")])
    assert res["auto"][1]["gram_kernel"] == 3, res["auto"][1]          # the residual (low-rank) form was chosen
    assert np.array_equal(res["auto"][0], res["one_chain"][0])


def test_vgg_fc2_full_size(engine):
    """VGG16 fc2: (4096, 4096), m = 1504, ternary."""
    _dense_case(engine, 4096, 4096, 1504, 16, 32, np.log2(3), 3, [("auto", {}, "auto"), ("gram_rows", {"sweep_outer": 1}, "gram")])


def test_config4_4096_25000(engine):
    """Synthetic sweep point N0 = N1 = 4096, m = 25000: tcgen05 Gram (196 K blocks per tile) + Gram-row sweep."""
    res = _dense_case(engine, 4096, 4096, 25000, 4, 44, np.log2(3), 3, [("auto", {}, "auto")])
    assert res["auto"][1]["gram_kernel"] == 2


def test_config4_16384_100000(engine):
    """Synthetic sweep corner N0 = 16384, m = 100000 (782 K blocks: four K chunks of the int8 accumulators), 256 neurons."""
    res = _dense_case(engine, 16384, 256, 100000, 4, 45, np.log2(3), 3, [("auto", {}, "auto")])
    assert res["auto"][1]["gram_kernel"] == 2


def _plane_patches(plane):
    n_img, H, Wd = plane.shape
    p = np.zeros((n_img, H + 2, Wd + 2), np.float32)
    p[:, 1:-1, 1:-1] = plane
    cols = np.empty((9, n_img * H * Wd), np.float32)
    for r in range(3):
        for cc in range(3):
            cols[r * 3 + cc] = p[:, r:r + H, cc:cc + Wd].reshape(-1)
    return cols


@pytest.mark.parametrize("n_img,H,C,F,c_check", [(5008, 32, 32, 32, 9),      # CIFAR conv2 (bench --workload cifar10_cnn)
                                                 (96, 224, 64, 64, 33),      # VGG conv1 geometry (RB = 8 bands), fewer images
                                                 (1504, 14, 512, 512, 77)])  # VGG conv10-12
def test_conv_full_size_all_paths(engine, n_img, H, C, F, c_check):
    """One channel of a full-size conv layer against the literal oracle walk over its (9, n_patches) patch matrices, through
    every conv entry point: NHWC activations (correlation form and the shared-memory planes kernel), per-channel patch
    matrices, and the image-split pair (Gram of two image halves summed, then the walks)."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(H * C)
    z = torch.randn((n_img, H, H, C), device="cuda", generator=g)
    act = torch.relu(z)
    actq = torch.relu(z + 0.05 * torch.randn((n_img, H, H, C), device="cuda", generator=g))
    del z
    rng = np.random.default_rng(H + C)
    W = (rng.uniform(-1, 1, (3, 3, C, F)) * np.sqrt(6.0 / (9 * C))).astype(np.float32)
    Wd = torch.from_numpy(W).cuda()
    A = O.layer_alphabet(W, 3, O.unit_alphabet(np.log2(3)))
    Xp = _plane_patches(act[..., c_check].cpu().numpy())
    Xqp = _plane_patches(actq[..., c_check].cpu().numpy())
    f_check = list(range(min(F, 16)))
    Wc = np.ascontiguousarray(W[:, :, c_check, :].reshape(9, F)[:, f_check])
    Qref = c_oracle.quantize_layer(Wc, Xp, Xqp, A)

    def compare(Q, name):
        Qs = np.asarray(Q)[:, :, c_check, :].reshape(9, F)[:, f_check]
        agree = O.agreement(Qs, Qref)
        assert agree >= AGREE, (name, agree)
        if agree < 1.0:
            r, rref = O.relative_residual(Wc, Qs, Xp, Xqp), O.relative_residual(Wc, Qref, Xp, Xqp)
            assert abs(r - rref) <= RESID_RTOL * max(rref, 1e-30), (name, r, rref)

    c0 = (c_check // 8) * 8
    Q = engine.conv_layer_nhwc(act, actq, Wd, A, c0=c0, n_channels=min(32, C - c0)).cpu().numpy()
    form = engine.last_stats["gram_kernel"]
    compare(Q, f"nhwc (gram_kernel {form})")
    engine.set_option("conv_kernel", 3)          # the planes kernel (patch form from shared memory)
    try:
        compare(engine.conv_layer_nhwc(act, actq, Wd, A, c0=c0, n_channels=8).cpu().numpy(), "nhwc planes")
    finally:
        engine.set_option("conv_kernel", 0)
    half = n_img // 2
    gram = engine.conv_gram_nhwc(act[:half], actq[:half], (3, 3), c0=c0, n_channels=8) + \
        engine.conv_gram_nhwc(act[half:], actq[half:], (3, 3), c0=c0, n_channels=8)
    compare(engine.conv_layer_from_gram(gram, Wd, A, c0=c0, n_channels=8).cpu().numpy(), "image split")
    Xpd, Xqpd = torch.from_numpy(Xp).cuda(), torch.from_numpy(Xqp).cuda()
    compare(engine.conv_channels([Xpd], [Xqpd], Wd, A, c0=c_check, n_channels=1).cpu().numpy(), "patch matrices")
    engine.trim()
