"""CPU: the C-ABI library builds, loads, and exports every symbol include/gpfq.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "gpfq.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gpfq_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for s in ("gpfq_create", "gpfq_destroy", "gpfq_last_error", "gpfq_dense_layer", "gpfq_conv_channels",
              "gpfq_conv_layer_nhwc", "gpfq_msq", "gpfq_bit_round", "gpfq_gram_matrices", "gpfq_set_stream"):
        assert s in syms


def test_library_builds_and_exports_every_declared_symbol():
    from quantized_neural_networks_b200 import _lib
    path = _lib.build()
    handle = ctypes.CDLL(path)
    for s in declared_symbols():
        assert hasattr(handle, s), f"{s} declared in include/gpfq.h but not exported by libgpfq.so"
    assert set(_lib.EXPORTS) == set(declared_symbols())
    lib = _lib.lib()
    assert lib.gpfq_version() == 100


def test_cubin_is_sm100a_only():
    import subprocess
    from quantized_neural_networks_b200 import _lib
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.build()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_gpu():
    """Without a GPU the engine must refuse to exist (GPFQ_ERR_UNSUPPORTED), never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from quantized_neural_networks_b200 import GpfqEngine, GpfqError
    with pytest.raises(GpfqError):
        GpfqEngine(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "quantized_neural_networks_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower().replace("no oracle", ""), f"{f} mentions the oracle"
