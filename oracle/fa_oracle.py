"""CPU oracle for the FlashAttention.jl hot path -- TEST INFRASTRUCTURE ONLY.

This module is a numpy restatement of the reference's CPU algorithms.  It is the
*checker*: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  The product path
(``flashattention.jl_b200/``) never imports anything under ``oracle/``.

PARITY PINNING STATUS (round 2): dense forward / backward and 1-D block attention are PINNED to the reference's own
code -- src_cpp/FlashAttention.cpp (OneDNaive, OneDFast, OneDParallelCPU, OneDNaiveBack, OneDFastBack,
OneDParallelCPUBack) compiled unmodified into oracle/_ref/libfa_ref_cpp.so (oracle/ref_build/) and compared with
this module at 1e-12, live and through the frozen vectors tests/golden/ref_cpp_*.npz (tests/test_ref_pin.py).
**Parity unpinned** for what has no runnable reference code here (no Julia runtime; the reference ships no golden
vectors): NNlib unfold/fold with padding or overlap (``window``/``unwindow``), the circulant
storage order and ``fused_softmax``.  Those are pinned against *independent* implementations (tests/test_oracle.py):
  * ``dense_dpa``            vs ``torch.nn.functional.scaled_dot_product_attention``
                             (the analogue of test/test.jl:19 vs NNlib.dot_product_attention)
  * ``window``/``unwindow``  vs ``torch.nn.functional.unfold``/``fold`` (1-D, 2-D) and a
                             brute-force per-element definition (1-D, 2-D, 3-D)
  * ``*_fa``                 vs ``*_dpa``  (test/test.jl:20, bench/compare.jl:20,47,74)
  * ``cartesian_circulant``  vs the closed-form key set  mod(j-1-p+t, N)+1
  * every backward           vs central finite differences in float64
and the outputs on seeded inputs are frozen in ``tests/golden/`` by ``tests/golden/make_golden.py``.  A maintainer
with Julia can export vectors from the ORIGINAL package (julia/FlashAttention/bench/export_golden.jl) for
tests/test_julia_golden.py.

Array convention: arrays have the *Julia shapes* of the reference, e.g. ``(N, d, B)`` or
``(X, Y, d, B)``, and are handled in Fortran (column-major) order so that linear
memory matches what a Julia ``Array`` / ``CuArray`` hands to ``ccall``.  All indices in
the docstrings are the reference's 1-based indices; the code is 0-based.

All ``file:line`` citations are relative to ``/root/reference``.
"""
from __future__ import annotations

import math
from concurrent.futures import ThreadPoolExecutor

import numpy as np

__all__ = [
    "cld", "fa_blocks", "dense_dpa", "dense_fa", "dense_fa_inplace", "dense_backward",
    "dense_fa_backward_blocked", "window_counts", "window_index", "window", "unwindow",
    "windowed_dpa", "windowed_fa", "block_fa", "block_dpa", "windowed_backward",
    "circshift_index", "cartesian_circulant", "circulant_keys", "circulant_dpa",
    "circulant_fa", "circulant_backward", "circulant_backward_given", "circulant2d_backward_given", "fused_softmax",
]


def cld(a: int, b: int) -> int:
    """Julia ``cld`` (ceiling division) for positive ints."""
    return -(-a // b)


def _F(x):
    return np.asfortranarray(x)


def _flatten3(x):
    """``reshape(q, :, d, batchsize)`` (src/dense.jl:6-8): flatten spatial dims, s1 fastest."""
    d, B = x.shape[-2], x.shape[-1]
    return np.reshape(_F(x), (-1, d, B), order="F")


# --------------------------------------------------------------------------------------
# dense
# --------------------------------------------------------------------------------------
def fa_blocks(N: int, d: int, M: int = 32_000):
    """Row/column block lengths of ``dense_fa!`` (src/dense.jl:28-36)."""
    Bc = min(max(cld(M, d), 1), N)
    Br = min(max(min(d, cld(M, d)), 1), N)
    return Br, Bc


def dense_dpa(q, k, v):
    """Naive attention, returns ``(y, P)`` (src/naive/dense.jl:8-35).

    ``P = softmax(tau * Q K^T, dims=2)`` of shape ``(N, N, B)``, ``tau = 1/sqrt(d)``.
    """
    Q, K, V = _flatten3(q), _flatten3(k), _flatten3(v)
    N, d, B = Q.shape
    dv = V.shape[1]
    tau = Q.dtype.type(1) / Q.dtype.type(math.sqrt(d))
    # batched_mul!(P, Q, K^T, tau)  (src/naive/dense.jl:14)
    S = np.einsum("ikb,jkb->ijb", Q, K) * tau
    S = S - S.max(axis=1, keepdims=True)           # NNlib.softmax!(dims=2) (:15)
    P = np.exp(S)
    P = P / P.sum(axis=1, keepdims=True)
    O = np.einsum("ijb,jcb->icb", P, V)            # batched_mul!(O, P, V)  (:16)
    y = np.reshape(_F(O), q.shape[:-2] + (dv, B), order="F")
    return y, _F(P.astype(Q.dtype))


def dense_fa_inplace(O, l, m, Q, K, V, M: int = 32_000, threads: int = 1):
    """Blocked online-softmax attention, ``dense_fa!`` (src/dense.jl:21-102).

    Same loop nest, block sizes, update rule (normalise every step, :89) and
    task decomposition (one task per (batch, row-block), :45) as the reference;
    ``threads`` plays the role of ``Threads.nthreads()``.
    """
    N, d, B = Q.shape
    T = Q.dtype.type
    Br, Bc = fa_blocks(N, d, M)
    Tr, Tc = cld(N, Br), cld(N, Bc)
    tau = T(1) / T(math.sqrt(d))                                   # :43

    def task(bi):
        b, i = bi
        r0, r1 = i * Br, min(N, (i + 1) * Br)
        Qi = Q[r0:r1, :, b]
        Oi = np.zeros((r1 - r0, V.shape[1]), dtype=Q.dtype)        # fill! :58
        li = np.zeros((r1 - r0, 1), dtype=Q.dtype)                 # :59
        mi = np.full((r1 - r0, 1), -np.inf, dtype=Q.dtype)         # :60
        for j in range(Tc):                                        # :70
            c0, c1 = j * Bc, min(N, (j + 1) * Bc)
            Kj, Vj = K[c0:c1, :, b], V[c0:c1, :, b]
            Pij = (Qi @ Kj.T) * tau                                # :77
            mij = Pij.max(axis=1, keepdims=True)                   # :78
            Pij = np.exp(Pij - mij)                                # :79
            lij = Pij.sum(axis=1, keepdims=True)                   # :80
            mi_new = np.maximum(mi, mij)                           # :82
            ei = np.exp(mi - mi_new)                               # :83
            eij = np.exp(mij - mi_new)                             # :84
            li_new = ei * li + eij * lij                           # :85
            Oi_new = Pij @ Vj                                      # :88
            Oi = (li * ei * Oi + eij * Oi_new) / li_new            # :89
            li, mi = li_new, mi_new                                # :90-91
        O[r0:r1, :, b] = Oi
        l[r0:r1, :, b] = li
        m[r0:r1, :, b] = mi

    tasks = [(b, i) for i in range(Tr) for b in range(B)]
    if threads > 1:
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(task, tasks))
    else:
        for t in tasks:
            task(t)
    return O, l, m


def dense_fa(q, k, v, M: int = 32_000, threads: int = 1):
    """``dense_fa(q,k,v) -> (y, l, m)`` (src/dense.jl:1-19).

    ``l_i = sum_j exp(s_ij - m_i)``, ``m_i = max_j s_ij``; ``l, m :: (N, 1, B)``.
    Unlike the reference ``O`` is allocated with ``dv`` columns so
    ``dv != d`` works.
    """
    Q, K, V = _flatten3(q), _flatten3(k), _flatten3(v)
    N, d, B = Q.shape
    dv = V.shape[1]
    O = np.zeros((N, dv, B), dtype=Q.dtype, order="F")
    l = np.zeros((N, 1, B), dtype=Q.dtype, order="F")
    m = np.zeros((N, 1, B), dtype=Q.dtype, order="F")
    dense_fa_inplace(O, l, m, Q, K, V, M=M, threads=threads)
    y = np.reshape(O, q.shape[:-2] + (dv, B), order="F")
    return y, l, m


def merge_partials(Oa, la, ma, Ob, lb, mb):
    """Online-softmax merge of two partial results over disjoint key sets: the update rule of
    src/dense.jl:82-91 (m = max, l rescaled by exp(m_old - m), O kept normalised)."""
    mn = np.maximum(ma, mb)
    wa = np.where(np.isneginf(ma), 0.0, la * np.exp(ma - mn))
    wb = np.where(np.isneginf(mb), 0.0, lb * np.exp(mb - mn))
    ln = wa + wb
    return (Oa * wa + Ob * wb) / ln, ln, mn


def ring_dense_fa(q_shards, k_shards, v_shards):
    """Ring schedule of SURVEY 8e on a list of per-rank token shards: at step s rank r holds the
    K/V block of rank (r - s) mod G, runs dense_fa of its queries against it and merges.  Returns
    per-rank (O, l, m); equal to dense_fa on the concatenated sequence."""
    G = len(q_shards)
    out = []
    for r in range(G):
        acc = None
        for s in range(G):
            src = (r - s) % G
            part = dense_fa(q_shards[r], k_shards[src], v_shards[src])
            acc = part if acc is None else merge_partials(*acc, *part)
        out.append(acc)
    return out


def dense_backward(Q, K, V, dO):
    """Naive backward, the runnable statement ``OneDNaiveBack``
    (src_cpp/FlashAttention.cpp:161-175).  Returns ``(dQ, dK, dV)``."""
    Q, K, V, dO = _flatten3(Q), _flatten3(K), _flatten3(V), _flatten3(dO)
    N, d, B = Q.shape
    tau = Q.dtype.type(1) / Q.dtype.type(math.sqrt(d))
    S = np.einsum("ikb,jkb->ijb", Q, K) * tau
    P = np.exp(S - S.max(axis=1, keepdims=True))
    P = P / P.sum(axis=1, keepdims=True)
    dV = np.einsum("ijb,icb->jcb", P, dO)                 # dV = P^T dO        (:169)
    dP = np.einsum("icb,jcb->ijb", dO, V)                 # dP = dO V^T        (:170)
    Bv = (P * dP).sum(axis=1, keepdims=True)              # rowsum(P o dP)     (:171)
    dS = P * (dP - Bv)                                    # (:172-173)
    dQ = np.einsum("ijb,jkb->ikb", dS, K) * tau           # (:174)
    dK = np.einsum("ijb,ikb->jkb", dS, Q) * tau           # (:175)
    return _F(dQ), _F(dK), _F(dV)


def dense_fa_backward_blocked(Q, K, V, O, dO, l, m, M: int = 32_000):
    """Flash backward with P recomputed from the saved ``(l, m)``: ``OneDFastBack``
    (src_cpp/FlashAttention.cpp:194-252) -- the working statement of the broken
    ``dense_fa_backward`` (src/dense.jl:104-167).  Block sizes follow src/dense.jl:118-119."""
    N, d, B = Q.shape
    T = Q.dtype.type
    Br, Bc = fa_blocks(N, d, M)
    tau = T(1) / T(math.sqrt(d))
    dQ, dK, dV = (np.zeros_like(_F(x)) for x in (Q, K, V))
    for b in range(B):
        for i in range(cld(N, Br)):
            r = slice(i * Br, min(N, (i + 1) * Br))
            Di = (dO[r, :, b] * O[r, :, b]).sum(axis=1, keepdims=True)       # cpp:243
            for j in range(cld(N, Bc)):
                c = slice(j * Bc, min(N, (j + 1) * Bc))
                S = (Q[r, :, b] @ K[c, :, b].T) * tau                        # cpp:238
                P = np.exp(S - m[r, :, b]) / l[r, :, b]                      # cpp:239-240
                dV[c, :, b] += P.T @ dO[r, :, b]                             # cpp:241
                dP = dO[r, :, b] @ V[c, :, b].T                              # cpp:242
                dS = P * (dP - Di)                                           # cpp:244
                dQ[r, :, b] += tau * (dS @ K[c, :, b])                       # cpp:245
                dK[c, :, b] += tau * (dS.T @ Q[r, :, b])                     # cpp:246
    return dQ, dK, dV


# --------------------------------------------------------------------------------------
# window / unwindow  (NNlib.unfold / NNlib.fold as called at src/utils.jl:40,52)
# --------------------------------------------------------------------------------------
def _win_kws(W, stride, pad):
    stride = W if stride is None else stride          # default stride=windowsize (src/utils.jl:36)
    pad = (W - 1) // 2 if pad is None else pad        # default pad=(W-1)/2      (src/utils.jl:36)
    return int(stride), int(pad)


def window_counts(spatial, W, stride=None, pad=None):
    """Windows per spatial dim ``o_k = (s_k + 2 pad - W) div stride + 1``."""
    stride, pad = _win_kws(W, stride, pad)
    return tuple((s + 2 * pad - W) // stride + 1 for s in spatial)


def window_index(spatial, W, stride=None, pad=None):
    """Index set of ``window``: int64 array ``(W^D, L)``, Fortran order.

    Entry ``[kappa, w]`` is the 0-based *linear* (s1-fastest) spatial index read by slot
    ``kappa`` of window ``w``, or ``-1`` where the slot falls in the zero padding.
    Slot and window linear indices are first-dim-fastest.
    """
    stride, pad = _win_kws(W, stride, pad)
    D = len(spatial)
    o = window_counts(spatial, W, stride, pad)
    L = int(np.prod(o))
    WD = W ** D
    # coordinates per dim, shape (WD, L)
    kap = np.unravel_index(np.arange(WD), (W,) * D, order="F")
    win = np.unravel_index(np.arange(L), o, order="F")
    lin = np.zeros((WD, L), dtype=np.int64)
    ok = np.ones((WD, L), dtype=bool)
    mult = 1
    for kdim in range(D):
        pos = win[kdim][None, :] * stride - pad + kap[kdim][:, None]   # (w-1)*stride - pad + kappa
        ok &= (pos >= 0) & (pos < spatial[kdim])
        lin += pos * mult
        mult *= spatial[kdim]
    lin[~ok] = -1
    return _F(lin)


def window(x, W, stride=None, pad=None):
    """``window(x, W; stride, pad)`` (src/utils.jl:36-44): ``(s.., d, B) -> (W^D, d, L, B)``;
    zero outside the array (im2col, cross-correlation order)."""
    x = _F(x)
    spatial, d, B = x.shape[:-2], x.shape[-2], x.shape[-1]
    idx = window_index(spatial, W, stride, pad)            # (WD, L)
    X = np.reshape(x, (-1, d, B), order="F")
    Xp = np.concatenate([X, np.zeros((1, d, B), dtype=x.dtype)], axis=0)   # row -1 == zero pad
    out = Xp[idx]                                          # (WD, L, d, B)
    return _F(np.transpose(out, (0, 2, 1, 3)))             # (WD, d, L, B)


def unwindow(Xw, outputsize, W, stride=None, pad=None):
    """``unwindow(X, outputsize, W; stride, pad)`` (src/utils.jl:46-54): col2im scatter-add
    ``(W^D, d, L, B) -> outputsize = (s.., d, B)``; padded slots are dropped."""
    spatial, d, B = tuple(outputsize[:-2]), outputsize[-2], outputsize[-1]
    idx = window_index(spatial, W, stride, pad)            # (WD, L)
    Ntok = int(np.prod(spatial))
    out = np.zeros((Ntok + 1, d, B), dtype=Xw.dtype)
    src = np.transpose(np.asarray(Xw), (0, 2, 1, 3))       # (WD, L, d, B)
    flat = np.where(idx < 0, Ntok, idx).reshape(-1)
    np.add.at(out, flat, src.reshape(-1, d, B))
    return np.reshape(_F(out[:Ntok]), tuple(spatial) + (d, B), order="F")


def _windowed(attn, q, k, v, W, stride, pad, **kw):
    qw, kw_, vw = (window(t, W, stride, pad) for t in (q, k, v))          # src/windowed.jl:4-6
    WD, d, L, B = qw.shape
    dv = vw.shape[1]
    r3 = lambda t: np.reshape(t, (t.shape[0], t.shape[1], -1), order="F")
    res = attn(r3(qw), r3(kw_), r3(vw), **kw)                             # :8-11
    yw = np.reshape(_F(res[0]), (WD, dv, L, B), order="F")                # :13
    szy = tuple(q.shape[:-2]) + (dv, q.shape[-1])                         # :14
    ones = np.ones(szy, dtype=q.dtype)
    divisor = unwindow(window(ones, W, stride, pad), szy, W, stride, pad)  # :16-17
    with np.errstate(divide="ignore", invalid="ignore"):
        y = unwindow(yw, szy, W, stride, pad) / divisor                   # :19  (0/0 -> NaN)
    return _F(y), res[1:], (WD, L, B)


def windowed_fa(q, k, v, W, stride=None, pad=None, threads: int = 1):
    """``windowed_fa(q,k,v,W; stride, pad) -> (y, l, m)`` (src/windowed.jl:3-23);
    ``l, m :: (W^D, 1, L, B)`` (:20-21)."""
    y, (lw, mw), (WD, L, B) = _windowed(dense_fa, q, k, v, W, stride, pad, threads=threads)
    l = np.reshape(_F(lw), (WD, 1, L, B), order="F")
    m = np.reshape(_F(mw), (WD, 1, L, B), order="F")
    return y, l, m


def windowed_dpa(q, k, v, W, stride=None, pad=None):
    """``windowed_dpa -> (y, P)`` with ``P :: (W^D, W^D, L, B)`` (src/naive/windowed.jl:3-23)."""
    y, (Pw,), (WD, L, B) = _windowed(dense_dpa, q, k, v, W, stride, pad)
    return y, np.reshape(_F(Pw), (WD, WD, L, B), order="F")


def block_fa(q, k, v, W, pad=0, threads: int = 1):
    """``block_fa`` forces ``stride=W`` and defaults ``pad=0`` (src/windowed.jl:1)."""
    return windowed_fa(q, k, v, W, stride=W, pad=pad, threads=threads)


def block_dpa(q, k, v, W):
    """``block_dpa`` forwards no kwargs, so inherits ``pad=(W-1)/2`` (src/naive/windowed.jl:1)."""
    return windowed_dpa(q, k, v, W)


def windowed_backward(q, k, v, dy, W, stride=None, pad=None):
    """Backward of ``windowed_fa`` (no reference code exists):
    ``dYw = window(dY ./ count)``; per-window dense backward; ``dq = unwindow(dQw)`` etc."""
    q, k, v, dy = _F(q), _F(k), _F(v), _F(dy)
    szy = dy.shape
    ones = np.ones(szy, dtype=q.dtype)
    cnt = unwindow(window(ones, W, stride, pad), szy, W, stride, pad)
    with np.errstate(divide="ignore", invalid="ignore"):
        g = np.where(cnt > 0, dy / cnt, 0).astype(q.dtype)
    qw, kw_, vw, gw = (window(t, W, stride, pad) for t in (q, k, v, g))
    WD, d, L, B = qw.shape
    r3 = lambda t: np.reshape(t, (t.shape[0], t.shape[1], -1), order="F")
    dQw, dKw, dVw = dense_backward(r3(qw), r3(kw_), r3(vw), r3(gw))
    r4 = lambda t, c: np.reshape(_F(t), (WD, c, L, B), order="F")
    dq = unwindow(r4(dQw, d), q.shape, W, stride, pad)
    dk = unwindow(r4(dKw, d), k.shape, W, stride, pad)
    dv = unwindow(r4(dVw, vw.shape[1]), v.shape, W, stride, pad)
    return dq, dk, dv


# --------------------------------------------------------------------------------------
# one volume over several ranks: slab split on window boundaries (no reference code --
# windows of src/utils.jl:36-44 with stride >= W never share a token, so whole window planes can be
# handed to different ranks and nothing is exchanged)
# --------------------------------------------------------------------------------------
def windowed_slab_plan(spatial, W, stride=None, pad=None, rank=0, nranks=1):
    """``(plane_lo, plane_hi, win_lo, win_hi, pad_lo)`` of rank ``rank``, by brute force from the window
    index set: the window planes of the slowest dim are dealt out in contiguous balanced ranges; a token
    plane belongs to the window plane that reads it, an unread one to the nearest window plane in front
    of it (plane 0 side: the first)."""
    stride, pad = _win_kws(W, stride, pad)
    assert stride >= W > pad
    S = spatial[-1]
    nw = window_counts(spatial, W, stride, pad)[-1]
    base, rem = divmod(nw, nranks)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    if hi <= lo:
        return (0, 0, lo, hi, 0)
    owner = np.full(S, -1)
    for w in range(nw):
        for t in range(W):
            z = w * stride - pad + t
            if 0 <= z < S:
                owner[z] = w
    for z in range(S):                       # unread planes: owner of the nearest read plane in front
        if owner[z] < 0:
            owner[z] = owner[z - 1] if z > 0 and owner[z - 1] >= 0 else 0
    mine = np.nonzero((owner >= lo) & (owner < hi))[0]
    plane_lo, plane_hi = int(mine[0]), int(mine[-1]) + 1
    assert len(mine) == plane_hi - plane_lo
    return (plane_lo, plane_hi, lo, hi, plane_lo - (lo * stride - pad))


def _slab_embed(x, W, stride, pad, pad_lo, nwin):
    """Slab ``(s.., planes, d, B)`` -> explicitly zero-padded volume on which ``window(.; pad=0)`` yields the slab's windows."""
    x = _F(x)
    spatial = x.shape[:-2]
    D = len(spatial)
    want = (nwin - 1) * stride + W                      # planes the slab's windows span
    widths = [(pad, pad)] * (D - 1) + [(pad_lo, max(0, want - pad_lo - spatial[-1]))] + [(0, 0), (0, 0)]
    xp = np.pad(x, widths)
    sl = [slice(None)] * x.ndim
    sl[D - 1] = slice(0, want)                          # planes behind the last window are not read
    return _F(xp[tuple(sl)]), widths


def windowed_fa_slab(q, k, v, W, stride, pad, pad_lo, nwin, threads: int = 1):
    """``windowed_fa`` restricted to one slab: the slab's planes of ``y`` and its windows' ``l, m``."""
    stride, pad = _win_kws(W, stride, pad)
    spatial = q.shape[:-2]
    D = len(spatial)
    (qp, widths), (kp, _), (vp, _) = (_slab_embed(t, W, stride, pad, pad_lo, nwin) for t in (q, k, v))
    yp, l, m = windowed_fa(qp, kp, vp, W, stride=stride, pad=0, threads=threads)
    # crop the explicit padding; planes behind the last window (cut off above) are unread: 0/0 = NaN
    sl = [slice(None)] * yp.ndim
    for ax in range(D - 1):
        sl[ax] = slice(pad, pad + spatial[ax])
    have = min(spatial[-1], yp.shape[D - 1] - pad_lo)
    sl[D - 1] = slice(pad_lo, pad_lo + have)
    y = yp[tuple(sl)]
    if have < spatial[-1]:
        tail = list(y.shape); tail[D - 1] = spatial[-1] - have
        y = np.concatenate([y, np.full(tail, np.nan, dtype=y.dtype)], axis=D - 1)
    return _F(y), l, m


# --------------------------------------------------------------------------------------
# circulant (1-D)
# --------------------------------------------------------------------------------------
def circshift_index(m: int, s: int, M: int) -> int:
    """``circshift_index(m, s, M) = mod(m - 1 - s, M) + 1`` (src/utils.jl:4), 1-based."""
    return (m - 1 - s) % M + 1


def cartesian_circulant(n: int, N: int, M: int):
    """``cartesian_circulant(n, N, M) -> (i, j)`` (src/utils.jl:6-17), all 1-based:
    nz-index ``n`` of the banded circulant CSC matrix -> (row = key, col = query)."""
    p = (M - 1) // 2                                   # :8
    j = cld(n, M)                                      # :9
    m = (n - 1) % M + 1                                # :10
    if j <= p:                                         # :11
        m = circshift_index(m, j - p - 1, M)           # :12
    elif j > N - p:
        m = circshift_index(m, p - N + j, M)           # :13
    i = ((m - 1) + (j - 1) - p) % N + 1                # :15
    return i, j


def circulant_keys(N: int, W: int):
    """Vectorised ``first(cartesian_circulant((j-1)W + w, N, W))`` for all ``j, w``:
    int64 ``(W, N)`` Fortran array of 0-based key indices, in the reference's storage order
    (the order of ``P[ww, ii, bb]`` in src/naive/circulant.jl:23)."""
    p = (W - 1) // 2
    j = np.arange(1, N + 1)[None, :]
    m = np.arange(1, W + 1)[:, None] + np.zeros_like(j)
    m = np.where(j <= p, (m - 1 - (j - p - 1)) % W + 1,
                 np.where(j > N - p, (m - 1 - (p - N + j)) % W + 1, m))
    i = ((m - 1) + (j - 1) - p) % N + 1
    return _F((i - 1).astype(np.int64))


def circulant_dpa(Q, K, V, W: int):
    """``circulant_dpa!`` (src/naive/circulant.jl:8-36): returns ``(O, P)`` with the dense
    ``P :: (W, N, B)`` nz-values in ``cartesian_circulant`` order (softmax over dim 1, :27)."""
    Q, K, V = _F(Q), _F(K), _F(V)
    N, d, B = Q.shape
    tau = Q.dtype.type(1) / Q.dtype.type(math.sqrt(d))
    keys = circulant_keys(N, W)                                   # (W, N)
    Kg = K[keys]                                                  # (W, N, d, B)
    S = np.einsum("ikb,wikb->wib", Q, Kg) * tau                   # :23
    S = S - S.max(axis=0, keepdims=True)
    P = np.exp(S)
    P = P / P.sum(axis=0, keepdims=True)                          # :27
    O = np.einsum("wib,wicb->icb", P, V[keys])                    # :28-34  (P^T * V)
    return _F(O), _F(P.astype(Q.dtype))


def circulant_fa(Q, K, V, W: int, M: int = 32_000, threads: int = 1):
    """``circulant_fa!(O,l,m,Q,K,V,W) -> (O, l, m)`` (src/circulant.jl:9-118): online softmax
    over window blocks of ``Bw`` keys (:24,:61), rows in blocks of ``Br`` (:25,:35).  The
    scalar loops (:68-79, :90-102) are vectorised per row-block; index set is identical."""
    Q, K, V = _F(Q), _F(K), _F(V)
    N, d, B = Q.shape
    T = Q.dtype.type
    Bw = min(max(cld(M, d), 1), W)                                # :24
    Br = min(max(min(d, cld(M, d)), 1), N)                        # :25
    Tr, Tw = cld(N, Br), cld(W, Bw)
    tau = T(1) / T(math.sqrt(d))                                  # :33
    keys = circulant_keys(N, W)                                   # (W, N)
    O = np.zeros((N, V.shape[1], B), dtype=Q.dtype, order="F")
    l = np.zeros((N, 1, B), dtype=Q.dtype, order="F")
    m = np.zeros((N, 1, B), dtype=Q.dtype, order="F")

    def task(bi):
        b, i = bi
        r0, r1 = i * Br, min(N, (i + 1) * Br)
        Oi = np.zeros((r1 - r0, V.shape[1]), dtype=Q.dtype)       # :50
        li = np.zeros((r1 - r0, 1), dtype=Q.dtype)                # :51
        mi = np.full((r1 - r0, 1), -np.inf, dtype=Q.dtype)        # :52
        for w in range(Tw):                                       # :61
            w0, w1 = w * Bw, min(W, (w + 1) * Bw)
            kk = keys[w0:w1, r0:r1].T                             # (rows, Bw) key index per (ii, ww)
            Piw = np.einsum("ik,iwk->iw", Q[r0:r1, :, b], K[:, :, b][kk]) * tau   # :68-79
            miw = Piw.max(axis=1, keepdims=True)                  # :80
            Piw = np.exp(Piw - miw)                               # :81
            liw = Piw.sum(axis=1, keepdims=True)                  # :82
            mi_new = np.maximum(mi, miw)                          # :84
            ei = np.exp(mi - mi_new)                              # :85
            eiw = np.exp(miw - mi_new)                            # :86
            li_new = ei * li + eiw * liw                          # :87
            t = np.einsum("iw,iwc->ic", Piw, V[:, :, b][kk])      # :90-99
            Oi = (li * ei * Oi + eiw * t) / li_new                # :101
            li, mi = li_new, mi_new                               # :106-107
        O[r0:r1, :, b], l[r0:r1, :, b], m[r0:r1, :, b] = Oi, li, mi

    tasks = [(b, i) for i in range(Tr) for b in range(B)]
    if threads > 1:
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(task, tasks))
    else:
        for t in tasks:
            task(t)
    return O, l, m


def circulant_backward(Q, K, V, dO, W: int):
    """Backward of circulant attention: A.5.1 restricted to the band;
    ``dK_j, dV_j`` scatter-add from the W queries whose window holds ``j``."""
    Q, K, V, dO = _F(Q), _F(K), _F(V), _F(dO)
    N, d, B = Q.shape
    tau = Q.dtype.type(1) / Q.dtype.type(math.sqrt(d))
    keys = circulant_keys(N, W)                                   # (W, N)
    S = np.einsum("ikb,wikb->wib", Q, K[keys]) * tau
    P = np.exp(S - S.max(axis=0, keepdims=True))
    P = P / P.sum(axis=0, keepdims=True)
    dP = np.einsum("icb,wicb->wib", dO, V[keys])
    dS = P * (dP - (P * dP).sum(axis=0, keepdims=True))
    dQ = np.einsum("wib,wikb->ikb", dS, K[keys]) * tau
    dK = np.zeros_like(K)
    dV = np.zeros_like(V)
    flat = keys.reshape(-1)
    np.add.at(dK, flat, (dS[:, :, None, :] * Q[None, :, :, :]).reshape(-1, d, B) * tau)
    np.add.at(dV, flat, (P[:, :, None, :] * dO[None, :, :, :]).reshape(-1, V.shape[1], B))
    return _F(dQ), _F(dK), _F(dV)


def circulant2d_keys(X: int, Y: int, W: int):
    """Keys of the 2-D circulant attention (the reference's todo, README.md:38-41,53): direct product of the
    1-D key set mod(j-1-p+t, N)+1 (src/utils.jl:6-17).  0-based (W*W, X*Y): row t*W+s, column y*X+x."""
    p = (W - 1) // 2
    x = np.arange(X)[None, None, None, :]
    y = np.arange(Y)[None, None, :, None]
    s = np.arange(W)[None, :, None, None]
    t = np.arange(W)[:, None, None, None]
    xx = np.mod(x - p + s, X)
    yy = np.mod(y - p + t, Y)
    return (yy * X + xx).reshape(W * W, X * Y)


def circulant2d_fa(Q, K, V, W: int):
    """2-D circulant attention on (X, Y, d, B) arrays -> (O (X,Y,dv,B), l, m (X*Y,1,B))."""
    Q, K, V = _F(Q), _F(K), _F(V)
    X, Y, d, B = Q.shape
    dv = V.shape[2]
    if W > X or W > Y:
        raise ValueError("W must not exceed the spatial extents")
    q, k, v = (t.reshape(X * Y, t.shape[2], B, order="F") for t in (Q, K, V))
    keys = circulant2d_keys(X, Y, W)
    tau = Q.dtype.type(1) / Q.dtype.type(math.sqrt(d))
    S = np.einsum("ikb,wikb->wib", q, k[keys]) * tau
    m = S.max(axis=0, keepdims=True)
    P = np.exp(S - m)
    l = P.sum(axis=0, keepdims=True)
    O = np.einsum("wib,wicb->icb", P / l, v[keys])
    return _F(O.reshape(X, Y, dv, B, order="F")), _F(l.reshape(X * Y, 1, B)), _F(m.reshape(X * Y, 1, B))


def circulant2d_backward(Q, K, V, dO, W: int):
    """Backward of circulant2d_fa (A.5.1 restricted to the periodic neighbourhood)."""
    Q, K, V, dO = _F(Q), _F(K), _F(V), _F(dO)
    X, Y, d, B = Q.shape
    dv = V.shape[2]
    q, k, v, g = (t.reshape(X * Y, t.shape[2], B, order="F") for t in (Q, K, V, dO))
    keys = circulant2d_keys(X, Y, W)
    tau = Q.dtype.type(1) / Q.dtype.type(math.sqrt(d))
    S = np.einsum("ikb,wikb->wib", q, k[keys]) * tau
    P = np.exp(S - S.max(axis=0, keepdims=True))
    P = P / P.sum(axis=0, keepdims=True)
    dP = np.einsum("icb,wicb->wib", g, v[keys])
    dS = P * (dP - (P * dP).sum(axis=0, keepdims=True))
    dq = np.einsum("wib,wikb->ikb", dS, k[keys]) * tau
    dk = np.zeros_like(k)
    dvv = np.zeros_like(v)
    flat = keys.reshape(-1)
    np.add.at(dk, flat, (dS[:, :, None, :] * q[None, :, :, :]).reshape(-1, d, B) * tau)
    np.add.at(dvv, flat, (P[:, :, None, :] * g[None, :, :, :]).reshape(-1, dv, B))
    rs = lambda t, c: _F(t.reshape(X, Y, c, B, order="F"))
    return rs(dq, d), rs(dk, d), rs(dvv, dv)


def circulant_backward_given(Q, K, V, O, dO, l, m, W: int):
    """Band-restricted ``OneDFastBack`` (src_cpp/FlashAttention.cpp:238-246) on the
    SAVED forward results: ``P = exp(tau q.k - m) / l`` and ``D = rowsum(dO o O)`` use the
    ``(O, l, m)`` the caller passes, as the flash backward does, instead of recomputing them."""
    Q, K, V, O, dO = _F(Q), _F(K), _F(V), _F(O), _F(dO)
    N, d, B = Q.shape
    tau = Q.dtype.type(1) / Q.dtype.type(math.sqrt(d))
    keys = circulant_keys(N, W)                                   # (W, N)
    S = np.einsum("ikb,wikb->wib", Q, K[keys]) * tau
    P = np.exp(S - np.asarray(m).reshape(1, N, B)) / np.asarray(l).reshape(1, N, B)      # cpp:239-240
    dP = np.einsum("icb,wicb->wib", dO, V[keys])                                          # cpp:242
    Di = (dO * O).sum(axis=1)[None, :, :]                                                 # cpp:243
    dS = P * (dP - Di)                                                                    # cpp:244
    dQ = np.einsum("wib,wikb->ikb", dS, K[keys]) * tau
    dK = np.zeros_like(K)
    dV = np.zeros_like(V)
    flat = keys.reshape(-1)
    np.add.at(dK, flat, (dS[:, :, None, :] * Q[None, :, :, :]).reshape(-1, d, B) * tau)
    np.add.at(dV, flat, (P[:, :, None, :] * dO[None, :, :, :]).reshape(-1, V.shape[1], B))
    return _F(dQ), _F(dK), _F(dV)


def circulant2d_backward_given(Q, K, V, O, dO, l, m, W: int):
    """2-D periodic neighbourhood form of :func:`circulant_backward_given`: the flash backward
    (src_cpp/FlashAttention.cpp:238-246) restricted to the W x W neighbourhood, on the SAVED (O, l, m)."""
    Q, K, V, O, dO = _F(Q), _F(K), _F(V), _F(O), _F(dO)
    X, Y, d, B = Q.shape
    dv = V.shape[2]
    q, k, v, o, g = (t.reshape(X * Y, t.shape[2], B, order="F") for t in (Q, K, V, O, dO))
    keys = circulant2d_keys(X, Y, W)
    tau = Q.dtype.type(1) / Q.dtype.type(math.sqrt(d))
    S = np.einsum("ikb,wikb->wib", q, k[keys]) * tau
    P = np.exp(S - np.asarray(m).reshape(1, X * Y, B)) / np.asarray(l).reshape(1, X * Y, B)
    dP = np.einsum("icb,wicb->wib", g, v[keys])
    dS = P * (dP - (g * o).sum(axis=1)[None, :, :])
    dq = np.einsum("wib,wikb->ikb", dS, k[keys]) * tau
    dk = np.zeros_like(k)
    dvv = np.zeros_like(v)
    flat = keys.reshape(-1)
    np.add.at(dk, flat, (dS[:, :, None, :] * q[None, :, :, :]).reshape(-1, d, B) * tau)
    np.add.at(dvv, flat, (P[:, :, None, :] * g[None, :, :, :]).reshape(-1, dv, B))
    rs = lambda t, c: _F(t.reshape(X, Y, c, B, order="F"))
    return rs(dq, d), rs(dk, d), rs(dvv, dv)


# --------------------------------------------------------------------------------------
# softmax
# --------------------------------------------------------------------------------------
def fused_softmax(S, dims: int = 1):
    """``fused_softmax(S; dims)`` (src/fused_softmax.jl:1-39): safe softmax along Julia dim 1
    (columns) or 2 (rows) of ``(M, N[, B])``; any other ``dims`` is an assertion error (:12)."""
    assert dims in (1, 2), "only softmax in dims 1 or 2 supported"
    S = _F(S)
    ax = dims - 1
    e = np.exp(S - S.max(axis=ax, keepdims=True))
    return _F(e / e.sum(axis=ax, keepdims=True))
