"""quantized_neural_networks_b200 -- B200-native GPFQ (greedy path-following quantization) hot path.

Host-side mirror of elybrand/quantized_neural_networks' `scripts/quantized_network.py` API
(`QuantizedNeuralNetwork`, `QuantizedCNN`, the module-level worker functions) over libgpfq.so,
a C-ABI library of hand-written sm_100a CUDA kernels (include/gpfq.h)."""
from ._lib import GpfqError, build  # noqa: F401
from .engine import GpfqEngine, get_engine  # noqa: F401
from .grid import QuantizedCNNGrid, pack_levels, unpack_levels  # noqa: F401
from .quantized_network import (  # noqa: F401
    QuantizedCNN,
    QuantizedNeuralNetwork,
    _bit_round_parallel,
    _quantize_filter2D_parallel_jit,
    _quantize_neuron_parallel,
)

__all__ = ["GpfqEngine", "GpfqError", "QuantizedCNN", "QuantizedCNNGrid", "QuantizedNeuralNetwork", "build", "get_engine",
           "pack_levels", "unpack_levels"]
