"""TEST INFRASTRUCTURE ONLY -- ctypes loader for oracle/libgpfq_oracle.so (the C restatement)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libgpfq_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "gpfq_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libgpfq_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.gpfq_oracle_layer.restype = ctypes.c_int
        _lib.gpfq_oracle_layer.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_long,
            ctypes.c_void_p, ctypes.c_long, ctypes.c_long, ctypes.c_long,
            ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_long, ctypes.c_int]
    return _lib


def quantize_layer(W, X, Xq, alphabet, j0=0, j1=None, nthreads=None) -> np.ndarray:
    """Literal walk of neurons j0..j1-1 of a (N0, N1) layer; returns Q (N0, N1) fp64 (other columns 0)."""
    X = np.ascontiguousarray(X, dtype=np.float32)
    Xq = X if Xq is None else np.ascontiguousarray(Xq, dtype=np.float32)
    W = np.ascontiguousarray(W, dtype=np.float32)
    A = np.ascontiguousarray(alphabet, dtype=np.float64)
    N0, m = X.shape
    assert W.shape[0] == N0 and Xq.shape == X.shape
    N1 = W.shape[1]
    j1 = N1 if j1 is None else j1
    Q = np.zeros((N0, N1))
    nthreads = nthreads or len(os.sched_getaffinity(0))
    rc = lib().gpfq_oracle_layer(X.ctypes.data, Xq.ctypes.data, N0, m, W.ctypes.data, N1, j0, j1,
                                 A.ctypes.data, len(A), Q.ctypes.data, N1, nthreads)
    if rc:
        raise RuntimeError(f"gpfq_oracle_layer failed rc={rc}")
    return Q


def quantize_conv_layer(W, patches, alphabet, nthreads=None) -> np.ndarray:
    """W (kh, kw, C, F); patches(c) -> (Xp, Xqp) each (kh*kw, n).  Channel loop of :844-860."""
    kh, kw, C, F = W.shape
    Q = np.zeros(W.shape)
    for c in range(C):
        Xp, Xqp = patches(c)
        Wc = np.ascontiguousarray(W[:, :, c, :].reshape(kh * kw, F))
        Q[:, :, c, :] = quantize_layer(Wc, Xp, Xqp, alphabet, nthreads=nthreads).reshape(kh, kw, F)
    return Q
