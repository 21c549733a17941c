"""Bring-up check of the int8-slice tcgen05 Gram kernel (gram_i8.cu) against an fp64 NumPy Gram; run under `timeout`.

    timeout 300 python tools/i8_check.py [quick]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quantized_neural_networks_b200 import get_engine  # noqa: E402


def check(eng, N0, m, kind, d=0):
    rng = np.random.default_rng(N0 * 7 + m)
    if kind == "hidden":
        Z = rng.standard_normal((N0, m))
        X = np.maximum(Z, 0).astype(np.float32)
        Xq = np.maximum(Z + 0.05 * rng.standard_normal((N0, m)), 0).astype(np.float32)
    elif kind == "signed":
        X = rng.standard_normal((N0, m)).astype(np.float32)
        Xq = (X + 0.05 * rng.standard_normal((N0, m))).astype(np.float32)
    elif kind == "pixels":
        X = (rng.integers(0, 256, (N0, m)) * (rng.random((N0, m)) < 0.5)).astype(np.float32)
        X[:3] = 0
        Xq = None
    elif kind == "wide":
        X = (rng.standard_normal((N0, m)) * 1e-4).astype(np.float32)
        X[:, ::97] = 50.0
        Xq = (X * (1 + 1e-3 * rng.standard_normal((N0, m)))).astype(np.float32)
    A = X.astype(np.float64)
    B = A if Xq is None else Xq.astype(np.float64)
    R2 = B @ B.T
    R1 = B @ A.T
    S2 = np.abs(B) @ np.abs(B).T
    S1 = np.abs(B) @ np.abs(A).T
    eng.set_option("gram_kernel", 2)
    eng.set_option("i8_pairs_d", d)
    t0 = time.perf_counter()
    G1, G2 = eng.gram_matrices(X, Xq)
    dt = time.perf_counter() - t0
    eng.set_option("gram_kernel", 1)
    H1, H2 = eng.gram_matrices(X, Xq)
    eng.set_option("gram_kernel", 0)
    tri = np.tril_indices(N0)
    out = []
    for G, H, R, Sabs in ((G2, H2, R2, S2),) + (() if Xq is None else ((G1, H1, R1, S1),)):
        e_i8 = np.max(np.abs(G[tri] - R[tri]) / np.maximum(Sabs[tri], 1e-300))
        e_dm = np.max(np.abs(H[tri] - R[tri]) / np.maximum(Sabs[tri], 1e-300))
        out.append((e_i8, e_dm))
    exact = kind == "pixels" and np.array_equal(G2[tri], R2[tri])
    print(f"N0={N0:6d} m={m:7d} {kind:7s} D={d:2d}  err_i8={[f'{a:.2e}' for a, _ in out]}  err_dmma={[f'{b:.2e}' for _, b in out]}"
          f"  exact={exact}  host_s={dt:.3f}", flush=True)
    return max(a for a, _ in out)


def main():
    eng = get_engine(0)
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    cases = [(256, 1024, "hidden"), (96, 400, "hidden"), (300, 3001, "hidden"), (784, 5000, "pixels"), (1000, 129, "signed"),
             (520, 2048, "wide"), (2048, 5008, "hidden")]
    if not quick:
        cases += [(256, 30000, "hidden"), (1500, 9000, "signed"), (4096, 4096, "hidden")]
    worst = 0.0
    for N0, m, kind in cases:
        worst = max(worst, check(eng, N0, m, kind))
    check(eng, 300, 3001, "hidden", d=6)
    check(eng, 300, 3001, "hidden", d=10)
    print("worst", worst)
    assert worst < 2e-10, worst


if __name__ == "__main__":
    main()
