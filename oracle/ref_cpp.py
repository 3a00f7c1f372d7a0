"""ctypes wrapper of oracle/_ref/libfa_ref_cpp.so -- TEST INFRASTRUCTURE ONLY.

The library is the reference's own ``src_cpp/FlashAttention.cpp`` (unmodified, compiled where it
lies under /root/reference by ``oracle/ref_build/Makefile`` against a stand-in Eigen header; Eigen,
MKL and icpc are absent from this image).  It is the executable pin of the numpy restatement
``oracle/fa_oracle.py`` for dense forward / backward and for 1-D block attention with exact cover:

  OneDNaive           src_cpp/FlashAttention.cpp:15-47     <-> fa_oracle.dense_dpa / block_dpa(1-D)
  OneDFast            :49-100                              <-> fa_oracle.dense_fa  / block_fa(1-D)
  OneDParallelCPU     :102-159                             <-> fa_oracle.dense_fa
  OneDNaiveBack       :161-192                             <-> fa_oracle.dense_backward
  OneDFastBack        :194-252                             <-> fa_oracle.dense_fa_backward_blocked
  OneDParallelCPUBack :253-317  (1 thread; it races, B-6)  <-> fa_oracle.dense_fa_backward_blocked

Only ``tests/`` may import this module.  All matrices are (N, d) Float64 in column-major order,
i.e. one batch slice of the Julia (N, d, B) arrays.
"""
import ctypes
import os

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_libs = {}


def available(checked=False) -> bool:
    return os.path.exists(os.path.join(_DIR, "libfa_ref_cpp_checked.so" if checked else "libfa_ref_cpp.so"))


def lib(checked=False):
    name = "libfa_ref_cpp_checked.so" if checked else "libfa_ref_cpp.so"
    if name not in _libs:
        path = os.path.join(_DIR, name)
        if not os.path.exists(path):
            raise ImportError(f"{path} missing: run `make -C oracle/ref_build` where /root/reference exists")
        L = ctypes.CDLL(path)
        dp, lg, db, it = ctypes.POINTER(ctypes.c_double), ctypes.c_long, ctypes.c_double, ctypes.c_int
        L.fa_ref_OneDNaive.argtypes = [dp] * 4 + [lg, lg, lg, db]
        L.fa_ref_OneDFast.argtypes = [dp] * 4 + [lg, lg, lg, lg, db]
        L.fa_ref_OneDParallelCPU.argtypes = [dp] * 4 + [lg, lg, lg, lg, db, it]
        L.fa_ref_OneDNaiveBack.argtypes = [dp] * 8 + [lg, lg, db]
        L.fa_ref_OneDFastBack.argtypes = [dp] * 10 + [lg, lg, lg, db]
        L.fa_ref_OneDParallelCPUBack.argtypes = [dp] * 10 + [lg, lg, lg, db, it]
        for f in ("OneDNaive", "OneDFast", "OneDParallelCPU", "OneDNaiveBack", "OneDFastBack", "OneDParallelCPUBack"):
            getattr(L, "fa_ref_" + f).restype = None
        _libs[name] = L
    return _libs[name]


def _in(x):
    x = np.asfortranarray(x, dtype=np.float64)
    return x, x.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _out(shape):
    x = np.zeros(shape, np.float64, order="F")
    return x, x.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def one_d_naive(Q, K, V, wsize=0, lam=1.0, checked=False):
    Q, pq = _in(Q); K, pk = _in(K); V, pv = _in(V)
    N, d = Q.shape
    O, po = _out((N, d))
    lib(checked).fa_ref_OneDNaive(pq, pk, pv, po, N, d, wsize, lam)
    return O


def one_d_fast(Q, K, V, cache, wsize=0, lam=1.0, parallel=False, threads=0, checked=False):
    Q, pq = _in(Q); K, pk = _in(K); V, pv = _in(V)
    N, d = Q.shape
    O, po = _out((N, d))
    if parallel:
        lib(checked).fa_ref_OneDParallelCPU(pq, pk, pv, po, N, d, cache, wsize, lam, threads)
    else:
        lib(checked).fa_ref_OneDFast(pq, pk, pv, po, N, d, cache, wsize, lam)
    return O


def one_d_naive_back(Q, K, V, P, dO, lam=1.0, checked=False):
    Q, pq = _in(Q); K, pk = _in(K); V, pv = _in(V); P, pp = _in(P); dO, pg = _in(dO)
    N, d = Q.shape
    (dQ, a), (dK, b), (dV, c) = _out((N, d)), _out((N, d)), _out((N, d))
    lib(checked).fa_ref_OneDNaiveBack(pq, pk, pv, pp, pg, a, b, c, N, d, lam)
    return dQ, dK, dV


def one_d_fast_back(Q, K, V, O, dO, l, m, cache, lam=1.0, parallel=False, checked=False):
    Q, pq = _in(Q); K, pk = _in(K); V, pv = _in(V); O, po = _in(O); dO, pg = _in(dO)
    l, pl = _in(np.ravel(l)); m, pm = _in(np.ravel(m))
    N, d = Q.shape
    (dQ, a), (dK, b), (dV, c) = _out((N, d)), _out((N, d)), _out((N, d))
    if parallel:
        lib(checked).fa_ref_OneDParallelCPUBack(pq, pk, pv, po, pg, pl, pm, a, b, c, N, d, cache, lam, 1)
    else:
        lib(checked).fa_ref_OneDFastBack(pq, pk, pv, po, pg, pl, pm, a, b, c, N, d, cache, lam)
    return dQ, dK, dV
