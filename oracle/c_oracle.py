"""ctypes wrapper of oracle/fa_oracle.c (the C/OpenMP restatement of the reference's CPU loops).
TEST INFRASTRUCTURE ONLY -- see the header of fa_oracle.c."""
import ctypes
import os

import numpy as np

_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "build", "libfa_oracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise ImportError(f"{_SO} missing: run `make -C oracle` (or __graft_entry__.build())")
        _lib = ctypes.CDLL(_SO)
        fp = ctypes.POINTER(ctypes.c_float)
        lg = ctypes.c_long
        _lib.fa_oracle_threads.restype = ctypes.c_int
        _lib.fa_oracle_set_threads.argtypes = [ctypes.c_int]
        _lib.fa_oracle_dense_fwd.argtypes = [fp] * 6 + [lg] * 4
        _lib.fa_oracle_circulant_fwd.argtypes = [fp] * 6 + [lg] * 5
        _lib.fa_oracle_windowed_fwd.argtypes = [fp] * 6 + [ctypes.c_int, ctypes.POINTER(lg)] + [lg] * 6
    return _lib


def threads() -> int:
    return lib().fa_oracle_threads()


def set_threads(n: int) -> int:
    """Use ``n`` OpenMP threads (torchrun exports OMP_NUM_THREADS=1); returns the new count."""
    lib().fa_oracle_set_threads(int(n))
    return threads()


def _f(x):
    x = np.asfortranarray(x, dtype=np.float32)
    return x, x.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def dense_fa(q, k, v):
    """dense_fa (src/dense.jl:1-102) on Float32 Julia-shaped arrays -> (y, l, m)."""
    q, pq = _f(q); k, pk = _f(k); v, pv = _f(v)
    d, B, dv = q.shape[-2], q.shape[-1], v.shape[-2]
    N = int(np.prod(q.shape[:-2]))
    y = np.zeros(q.shape[:-2] + (dv, B), np.float32, order="F")
    l = np.zeros((N, 1, B), np.float32, order="F")
    m = np.zeros((N, 1, B), np.float32, order="F")
    fp = ctypes.POINTER(ctypes.c_float)
    lib().fa_oracle_dense_fwd(pq, pk, pv, y.ctypes.data_as(fp), l.ctypes.data_as(fp), m.ctypes.data_as(fp), N, d, dv, B)
    return y, l, m


def circulant_fa(Q, K, V, W):
    """circulant_fa! (src/circulant.jl:9-118) -> (O, l, m)."""
    Q, pq = _f(Q); K, pk = _f(K); V, pv = _f(V)
    N, d, B = Q.shape
    dv = V.shape[1]
    O = np.zeros((N, dv, B), np.float32, order="F")
    l = np.zeros((N, 1, B), np.float32, order="F")
    m = np.zeros((N, 1, B), np.float32, order="F")
    fp = ctypes.POINTER(ctypes.c_float)
    lib().fa_oracle_circulant_fwd(pq, pk, pv, O.ctypes.data_as(fp), l.ctypes.data_as(fp), m.ctypes.data_as(fp), N, d, dv, B, W)
    return O, l, m


def windowed_fa(q, k, v, W, stride=None, pad=None):
    """windowed_fa (src/windowed.jl:3-23) -> (y, l, m)."""
    stride = W if stride is None else stride
    pad = (W - 1) // 2 if pad is None else pad
    q, pq = _f(q); k, pk = _f(k); v, pv = _f(v)
    spatial, d, B, dv = q.shape[:-2], q.shape[-2], q.shape[-1], v.shape[-2]
    o = [(s + 2 * pad - W) // stride + 1 for s in spatial]
    L, WD = int(np.prod(o)), W ** len(spatial)
    y = np.zeros(spatial + (dv, B), np.float32, order="F")
    l = np.zeros((WD, 1, L, B), np.float32, order="F")
    m = np.zeros((WD, 1, L, B), np.float32, order="F")
    fp = ctypes.POINTER(ctypes.c_float)
    dims = (ctypes.c_long * len(spatial))(*spatial)
    lib().fa_oracle_windowed_fwd(pq, pk, pv, y.ctypes.data_as(fp), l.ctypes.data_as(fp), m.ctypes.data_as(fp),
                                 len(spatial), dims, d, dv, B, W, stride, pad)
    return y, l, m
