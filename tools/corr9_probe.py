"""Probe: run the correlation-form conv path on a list of small shapes, each in its own process (CUDA errors are sticky)."""
import subprocess
import sys

CASES = [(4, 7, 7, 40), (4, 7, 7, 32), (4, 7, 7, 64), (4, 14, 14, 40), (4, 14, 14, 64), (40, 7, 7, 64), (4, 7, 12, 64), (4, 12, 7, 64),
         (4, 6, 5, 32), (4, 16, 16, 32)]
CHILD = r'''
import sys, numpy as np, torch
sys.path.insert(0, ".")
from quantized_neural_networks_b200 import get_engine
n, H, W, C = (int(v) for v in sys.argv[1:5])
eng = get_engine(0)
rng = np.random.default_rng(0)
act = torch.from_numpy(np.maximum(rng.standard_normal((n, H, W, C)), 0).astype(np.float32)).cuda()
Wt = torch.from_numpy((rng.uniform(-1, 1, (3, 3, C, 2)) * 0.3).astype(np.float32)).cuda()
A = np.linspace(-1, 1, 3) * 0.2
Q = eng.conv_layer_nhwc(act, None, Wt, A)
torch.cuda.synchronize()
print("ok", eng.last_stats["gram_kernel"])
'''
for case in CASES:
    r = subprocess.run([sys.executable, "-c", CHILD, *map(str, case)], capture_output=True, text=True)
    print(case, (r.stdout.strip() or r.stderr.strip()[-300:]).replace("\n", " | "), flush=True)
