"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference module.

The reference (`/root/reference/scripts/quantized_network.py`) imports TensorFlow and
h5py at module scope (`:21-25`, `:31`); neither is installed here.  The arithmetic of
the hot path (`_bit_round_parallel :40`, `_quantize_weight_parallel :59`,
`_quantize_neuron_parallel :91`, `_quantize_filter2D_parallel_jit :185`) only needs
NumPy + SciPy.  This shim registers empty stand-in modules for the TensorFlow names
and an in-memory `h5py.File` (a dict of ndarrays behind a context manager), then
imports the reference file from where it lies.  Nothing of the reference is copied.

`/root/reference` exists only in the build container, never on the GPU box: this
module is used by `oracle/make_golden.py` (fixture generation) and by the CPU tests
that pin the oracle restatement (they skip when the reference is absent).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("GPFQ_REFERENCE_ROOT", "/root/reference")
_SCRIPTS = os.path.join(REFERENCE_ROOT, "scripts")

# name -> {dataset -> ndarray}; the reference passes file *names* to its workers.
MEMORY_FILES: dict[str, dict] = {}


class _MemFile:
    """Stand-in for h5py.File: 'w' creates/clears a named dict, 'r' opens it."""

    def __init__(self, name, mode="r"):
        key = os.path.basename(str(name))
        if mode == "w":
            MEMORY_FILES[key] = {}
        self._d = MEMORY_FILES[key]

    def __enter__(self):
        return self._d

    def __exit__(self, *exc):
        return False


def available() -> bool:
    return os.path.isfile(os.path.join(_SCRIPTS, "quantized_network.py"))


def load():
    """Import and return the reference `quantized_network` module (unmodified)."""
    if "quantized_network" in sys.modules and getattr(
        sys.modules["quantized_network"], "__gpfq_shim__", False
    ):
        return sys.modules["quantized_network"]
    if not available():
        raise FileNotFoundError(f"reference not present under {REFERENCE_ROOT}")

    def mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    class Sequence:  # base class only
        pass

    def _absent(*a, **k):
        raise RuntimeError("TensorFlow is not installed; host-side collection cannot run")

    saved = {k: sys.modules.get(k) for k in (
        "tensorflow", "tensorflow.keras", "tensorflow.keras.utils", "tensorflow.keras.backend",
        "tensorflow.keras.models", "tensorflow.image", "h5py")}
    tf = mod("tensorflow", convert_to_tensor=_absent)
    keras = mod("tensorflow.keras")
    tf.keras = keras
    keras.utils = mod("tensorflow.keras.utils", Sequence=Sequence)
    keras.backend = mod("tensorflow.keras.backend", function=_absent)
    keras.models = mod("tensorflow.keras.models", Model=_absent, clone_model=_absent)
    tf.image = mod("tensorflow.image", extract_patches=_absent)
    mod("h5py", File=_MemFile)

    sys.path.insert(0, _SCRIPTS)
    try:
        ref = importlib.import_module("quantized_network")
    finally:
        sys.path.remove(_SCRIPTS)
        # leave the stubs only inside the imported module's namespace
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    ref.__gpfq_shim__ = True
    return ref


def ref_quantize_neuron(w, X, Xq, alphabet):
    """Run the reference's `_quantize_neuron_parallel` (:91-121) on in-memory data.

    X, Xq: float32 (N0, m) feature-major, exactly the `wX`/`qX` datasets of `:469-492`.
    """
    ref = load()
    name = f"layer_mem_{os.getpid()}_{id(X)}.h5"
    MEMORY_FILES[name] = {"wX": X, "qX": Xq}
    try:
        return ref._quantize_neuron_parallel(w, name, alphabet)
    finally:
        MEMORY_FILES.pop(name, None)


def ref_quantize_filter(chan_filter, Xp, Xqp, alphabet, channel_idx=0):
    """Run the reference's `_quantize_filter2D_parallel_jit` (:185-233) on in-memory patches."""
    ref = load()
    name = f"channel_mem_{os.getpid()}_{id(Xp)}.h5"
    MEMORY_FILES[name] = {f"wX_channel{channel_idx}": Xp, f"qX_channel{channel_idx}": Xqp}
    try:
        return ref._quantize_filter2D_parallel_jit(chan_filter, channel_idx, name, alphabet)
    finally:
        MEMORY_FILES.pop(name, None)
