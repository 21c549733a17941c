"""Probe: durations of slgemm_i8_kernel for a few shapes (run under `ncu --metrics gpu__time_duration.sum`)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quantized_neural_networks_b200 import get_engine
eng = get_engine(0)
rng = np.random.default_rng(0)
for (M, N, K, D, tb) in [(2048, 1504, 512, 7, False), (2048, 1504, 512, 6, False), (2048, 1504, 512, 7, True), (2048, 512, 1536, 6, False),
                         (2048, 1536, 512, 7, False), (2048, 1504, 1536, 6, False), (4096, 512, 1536, 6, False)]:
    A = rng.standard_normal((M, K))
    B = rng.standard_normal((K, N) if tb else (N, K)).astype(np.float32)
    for _ in range(2):
        C = eng.debug_slgemm(A, B, D=D, transposed_b=tb)
    ref = A @ (B.astype(np.float64) if tb else B.astype(np.float64).T)
    print(M, N, K, D, tb, float(np.max(np.abs(C - ref)) / np.max(np.abs(ref))), flush=True)
