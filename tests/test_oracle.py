"""CPU tests that pin the oracle (oracle/fa_oracle.py) against independent implementations.

The reference ships no golden vectors and cannot be run here, so every
function of the restatement is checked against something that does not share its code:
torch's SDPA / unfold / fold, brute-force per-element definitions, closed-form index
sets and float64 finite differences.  Mirrors test/test.jl:5-21 and the ``@test O1 ~ O2``
checks of bench/compare.jl:20,47,74.
"""
import itertools
import math

import numpy as np
import pytest
import torch

from oracle import fa_oracle as fo

RTOL = math.sqrt(np.finfo(np.float64).eps)  # Julia isapprox default rtol for Float64


def randn(shape, seed, dtype=np.float64):
    return np.asfortranarray(np.random.default_rng(seed).standard_normal(shape).astype(dtype))


def approx(a, b, rtol=RTOL):
    """Julia ``isapprox``: norm(a-b) <= rtol*max(norm(a), norm(b)), NaNs must coincide."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    na, nb = np.isnan(a), np.isnan(b)
    if not np.array_equal(na, nb):
        return False
    a, b = np.where(na, 0, a), np.where(nb, 0, b)
    return np.linalg.norm(a - b) <= rtol * max(np.linalg.norm(a), np.linalg.norm(b))


# ---------------------------------------------------------------- dense
def test_dense_dpa_vs_torch_sdpa():
    # test/test.jl:6-19 shapes (Nq=Nkv=30, dqk=12, dv=6, bs=2), rand inputs
    rng = np.random.default_rng(0)
    q, k, v = (np.asfortranarray(rng.random(s)) for s in ((30, 12, 2), (30, 12, 2), (30, 6, 2)))
    y1, P1 = fo.dense_dpa(q, k, v)
    tq, tk, tv = (torch.from_numpy(np.ascontiguousarray(t.transpose(2, 0, 1))) for t in (q, k, v))
    y0 = torch.nn.functional.scaled_dot_product_attention(tq, tk, tv).numpy().transpose(1, 2, 0)
    assert approx(y1, y0)
    assert np.allclose(P1.sum(axis=1), 1.0)


@pytest.mark.parametrize("shape,dv", [((30, 12, 2), 6), ((30, 12, 2), 12), ((1024, 64, 1), 64),
                                      ((7, 5, 3, 8, 2), 8), ((600, 64, 1), 64)])
def test_dense_fa_vs_dpa(shape, dv):
    # test/test.jl:20 (dv != d allowed here, SURVEY B-3) and bench/compare.jl:20
    q, k = randn(shape, 0), randn(shape, 1)
    v = randn(shape[:-2] + (dv, shape[-1]), 2)
    y1, P = fo.dense_dpa(q, k, v)
    y2, l, m = fo.dense_fa(q, k, v)
    assert y2.shape == shape[:-2] + (dv, shape[-1])
    assert approx(y2, y1)
    N, d = int(np.prod(shape[:-2])), shape[-2]
    S = np.einsum("ikb,jkb->ijb", q.reshape((N, d, -1), order="F"),
                  k.reshape((N, d, -1), order="F")) / math.sqrt(d)
    assert l.shape == m.shape == (N, 1, shape[-1])
    assert approx(m[:, 0], S.max(axis=1))
    assert approx(l[:, 0], np.exp(S - S.max(axis=1, keepdims=True)).sum(axis=1))


def test_dense_fa_blocks_and_threads():
    assert fo.fa_blocks(1024, 64) == (64, 500)
    assert fo.fa_blocks(8192, 128) == (128, 250)
    assert fo.fa_blocks(8192, 32) == (32, 1000)
    assert fo.fa_blocks(30, 12) == (12, 30)
    q, k, v = randn((700, 32, 3), 0), randn((700, 32, 3), 1), randn((700, 32, 3), 2)
    a = fo.dense_fa(q, k, v, threads=1)
    b = fo.dense_fa(q, k, v, threads=4)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def _fd_grad(f, xs, gout, eps=1e-6):
    """Central finite differences of sum(f(xs) * gout) w.r.t. each array in xs."""
    grads = []
    for t in range(len(xs)):
        g = np.zeros_like(xs[t])
        it = np.nditer(xs[t], flags=["multi_index"])
        for _ in it:
            idx = it.multi_index
            old = xs[t][idx]
            xs[t][idx] = old + eps
            fp = np.nansum(f(*xs) * gout)
            xs[t][idx] = old - eps
            fm = np.nansum(f(*xs) * gout)
            xs[t][idx] = old
            g[idx] = (fp - fm) / (2 * eps)
        grads.append(g)
    return grads


def test_dense_backward_fd_and_blocked():
    q, k, v, g = randn((9, 4, 2), 0), randn((9, 4, 2), 1), randn((9, 3, 2), 2), randn((9, 3, 2), 3)
    dq, dk, dv = fo.dense_backward(q, k, v, g)
    fq, fk, fv = _fd_grad(lambda a, b, c: fo.dense_dpa(a, b, c)[0], [q, k, v], g)
    for a, b in ((dq, fq), (dk, fk), (dv, fv)):
        assert np.abs(a - b).max() < 1e-7
    # blocked flash backward (cpp:194-252) == naive backward (cpp:161-175)
    q, k, v, g = (randn((300, 64, 2), s) for s in range(4))
    y, l, m = fo.dense_fa(q, k, v)
    b1 = fo.dense_fa_backward_blocked(q, k, v, y, g, l, m, M=4000)
    b0 = fo.dense_backward(q, k, v, g)
    for a, b in zip(b1, b0):
        assert approx(a, b)


# ---------------------------------------------------------------- window / unwindow
def _window_bruteforce(x, W, stride, pad):
    spatial, d, B = x.shape[:-2], x.shape[-2], x.shape[-1]
    D = len(spatial)
    o = [(s + 2 * pad - W) // stride + 1 for s in spatial]
    out = np.zeros((W ** D, d, int(np.prod(o)), B), dtype=x.dtype)
    for wl, w in enumerate(itertools.product(*[range(n) for n in reversed(o)])):
        w = w[::-1]                                      # first dim fastest
        for kl, kap in enumerate(itertools.product(*[range(W)] * D)):
            kap = kap[::-1]
            pos = tuple(w[i] * stride - pad + kap[i] for i in range(D))
            if all(0 <= pos[i] < spatial[i] for i in range(D)):
                out[kl, :, wl, :] = x[pos]
    return out


@pytest.mark.parametrize("spatial,W,stride,pad", [
    ((20,), 5, 2, 2), ((22,), 5, 5, 0), ((64,), 7, 7, 3), ((16,), 4, 3, 1),
    ((9, 8), 3, 2, 1), ((10, 10), 7, 7, 3), ((6, 7), 4, 1, 2),
    ((5, 6, 4), 3, 2, 1), ((6, 6, 6), 5, 5, 2), ((4, 5, 6), 2, 1, 0)])
def test_window_vs_bruteforce(spatial, W, stride, pad):
    x = randn(spatial + (3, 2), 5)
    got = fo.window(x, W, stride, pad)
    assert np.array_equal(got, _window_bruteforce(x, W, stride, pad))
    # unwindow is the exact adjoint of window: <window(x), Y> == <x, unwindow(Y)>
    Y = randn(got.shape, 6)
    lhs = (got * Y).sum()
    rhs = (x * fo.unwindow(Y, x.shape, W, stride, pad)).sum()
    assert abs(lhs - rhs) <= 1e-10 * max(1.0, abs(lhs))


@pytest.mark.parametrize("spatial,W,stride,pad", [
    ((20,), 5, 2, 2), ((64,), 7, 7, 3), ((4096,), 64, 16, 0),
    ((9, 8), 3, 2, 1), ((64, 64), 7, 7, 3), ((12, 10), 5, 1, 2)])
def test_window_vs_torch_unfold_fold(spatial, W, stride, pad):
    """NNlib.unfold/fold (not vendored) restated; torch's im2col/col2im is the independent pin."""
    d, B = 3, 2
    x = randn(spatial + (d, B), 7)
    sp2 = spatial if len(spatial) == 2 else (spatial[0], 1)
    # Julia (s1, s2, d, B) column-major == torch (B, d, s2, s1) row-major
    xt = torch.from_numpy(np.ascontiguousarray(x.reshape(sp2 + (d, B), order="F").transpose(3, 2, 1, 0)))
    ks = (W, W) if len(spatial) == 2 else (1, W)
    pd = (pad, pad) if len(spatial) == 2 else (0, pad)
    cols = torch.nn.functional.unfold(xt, ks, padding=pd, stride=stride)    # (B, d*W^D, L)
    got = fo.window(x, W, stride, pad)                                       # (W^D, d, L, B)
    WD, _, L, _ = got.shape
    want = cols.numpy().reshape(B, d, WD, L).transpose(2, 1, 3, 0)
    assert np.array_equal(got, want)
    Y = randn(got.shape, 8)
    Yt = torch.from_numpy(np.ascontiguousarray(Y.transpose(3, 1, 0, 2).reshape(B, d * WD, L)))
    folded = torch.nn.functional.fold(Yt, (sp2[1], sp2[0]), ks, padding=pd, stride=stride)
    want_f = folded.numpy().transpose(3, 2, 1, 0).reshape(spatial + (d, B), order="F")
    assert np.allclose(fo.unwindow(Y, x.shape, W, stride, pad), want_f, rtol=1e-12, atol=1e-12)


def test_window_coverage_facts():
    # SURVEY A.3 coverage table
    assert fo.window_counts((64,), 7) == (10,)
    assert fo.window_counts((64,), 5) == (13,)
    assert fo.window_counts((64,), 5, 5, 3) == (14,)
    assert fo.window_counts((4096,), 64, 16, 0) == (253,)
    idx = fo.window_index((64,), 5)                       # default stride 5 pad 2
    covered = np.zeros(64, int)
    np.add.at(covered, idx[idx >= 0], 1)
    assert (covered == 0).sum() == 1 and covered.max() == 1
    idx = fo.window_index((64,), 5, 1, 2)
    covered = np.zeros(64, int)
    np.add.at(covered, idx[idx >= 0], 1)
    assert covered.max() == 5 and covered.min() == 3


@pytest.mark.parametrize("spatial,W,kws", [
    ((64,), 16, dict(stride=16, pad=0)), ((64,), 16, dict(stride=4, pad=0)),
    ((20, 12), 7, {}), ((16, 16), 3, dict(stride=1, pad=1)),
    ((6, 7, 8), 3, {}), ((8, 8, 8), 5, dict(stride=5, pad=1))])
def test_windowed_fa_vs_dpa(spatial, W, kws):
    # bench/compare.jl:47
    q, k, v = (randn(spatial + (8, 2), s) for s in range(3))
    y1, P = fo.windowed_dpa(q, k, v, W, **kws)
    y2, l, m = fo.windowed_fa(q, k, v, W, **kws)
    assert approx(y2, y1)
    WD = W ** len(spatial)
    L = int(np.prod(fo.window_counts(spatial, W, kws.get("stride"), kws.get("pad"))))
    assert l.shape == m.shape == (WD, 1, L, 2) and P.shape == (WD, WD, L, 2)


def test_windowed_nan_for_uncovered_and_pad_tokens():
    # uncovered positions are 0/0 = NaN; zero-pad tokens take part in softmax
    q, k, v = (randn((22, 4, 1), s) for s in range(3))
    y, l, m = fo.windowed_fa(q, k, v, 5, stride=5, pad=0)
    assert np.isnan(y[20:]).all() and not np.isnan(y[:20]).any()
    y, l, m = fo.windowed_fa(q, k, v, 5)                  # pad=2: first window has 2 pad slots
    assert (m[:2, 0, 0, 0] == 0).all()                    # pad query rows: scores all exactly 0
    assert np.allclose(l[:2, 0, 0, 0], 5.0)
    assert (m[2:, 0, 0, 0] >= 0).all()                    # pad keys contribute score 0


def test_block_aliases():
    q, k, v = (randn((32, 4, 2), s) for s in range(3))
    assert np.array_equal(fo.block_fa(q, k, v, 8)[0], fo.windowed_fa(q, k, v, 8, stride=8, pad=0)[0])
    assert np.array_equal(fo.block_dpa(q, k, v, 8)[0], fo.windowed_dpa(q, k, v, 8)[0], equal_nan=True)


@pytest.mark.parametrize("spatial,W,stride,pad", [((20,), 5, 2, 2), ((22,), 5, 5, 0), ((6, 5), 3, 2, 1)])
def test_windowed_backward_fd(spatial, W, stride, pad):
    q, k, v = (randn(spatial + (3, 2), s) for s in range(3))
    g = randn(spatial + (3, 2), 3)
    dq, dk, dv = fo.windowed_backward(q, k, v, g, W, stride, pad)
    f = lambda a, b, c: fo.windowed_dpa(a, b, c, W, stride, pad)[0]
    fq, fk, fv = _fd_grad(f, [q, k, v], g)
    for a, b in ((dq, fq), (dk, fk), (dv, fv)):
        assert np.abs(a - b).max() < 1e-7


# ---------------------------------------------------------------- circulant
@pytest.mark.parametrize("N,W", [(8, 3), (16, 5), (16, 7), (32, 8), (16, 4), (4096, 255), (64, 63), (9, 9)])
def test_cartesian_circulant_index_set(N, W):
    p = (W - 1) // 2
    keys = fo.circulant_keys(N, W)
    for j in (list(range(1, min(N, 2 * W) + 1)) + list(range(max(1, N - 2 * W), N + 1))):
        scalar = [fo.cartesian_circulant((j - 1) * W + w, N, W) for w in range(1, W + 1)]
        assert all(jj == j for _, jj in scalar)
        ks = [i for i, _ in scalar]
        assert ks == list(keys[:, j - 1] + 1)                       # vectorised == scalar, in order
        assert sorted(ks) == sorted(((j - 1 - p + t) % N) + 1 for t in range(W))
        if W % 2 == 1:
            assert ks == sorted(ks)                                 # valid CSC for odd W
    if (N, W) == (8, 3):
        assert list(keys[:, 0] + 1) == [1, 2, 8] and list(keys[:, 7] + 1) == [1, 7, 8]


@pytest.mark.parametrize("N,d,B,W", [(64, 8, 2, 9), (256, 32, 1, 129), (128, 16, 2, 16), (40, 4, 1, 40),
                                     (700, 64, 1, 65)])
def test_circulant_fa_vs_dpa(N, d, B, W):
    # bench/compare.jl:74
    Q, K, V = (randn((N, d, B), s) for s in range(3))
    O1, P = fo.circulant_dpa(Q, K, V, W)
    O2, l, m = fo.circulant_fa(Q, K, V, W, M=max(4 * d, 1000))       # several window blocks
    assert approx(O2, O1)
    assert P.shape == (W, N, B) and np.allclose(P.sum(axis=0), 1)
    # banded == dense attention with everything outside the band masked
    p = (W - 1) // 2
    S = np.einsum("ikb,jkb->ijb", Q, K) / math.sqrt(d)
    i, j = np.arange(N)[:, None], np.arange(N)[None, :]
    band = ((j - (i - p)) % N) < W
    S = np.where(band[:, :, None], S, -np.inf)
    Pd = np.exp(S - S.max(axis=1, keepdims=True))
    assert approx(l[:, 0], Pd.sum(axis=1)) and approx(m[:, 0], S.max(axis=1))
    Od = np.einsum("ijb,jcb->icb", Pd / Pd.sum(axis=1, keepdims=True), V)
    assert approx(O1, Od)


def test_circulant_full_window_is_dense():
    Q, K, V = (randn((33, 8, 2), s) for s in range(3))
    assert approx(fo.circulant_fa(Q, K, V, 33)[0], fo.dense_dpa(Q, K, V)[0])


def test_circulant_backward_fd():
    Q, K, V, g = (randn((12, 3, 2), s) for s in range(4))
    for W in (5, 4):
        dq, dk, dv = fo.circulant_backward(Q, K, V, g, W)
        fq, fk, fv = _fd_grad(lambda a, b, c: fo.circulant_dpa(a, b, c, W)[0], [Q, K, V], g)
        for a, b in ((dq, fq), (dk, fk), (dv, fv)):
            assert np.abs(a - b).max() < 1e-7


# ---------------------------------------------------------------- softmax
def test_fused_softmax():
    S = randn((7, 9, 3), 0)
    t = torch.from_numpy(np.ascontiguousarray(S))
    assert np.allclose(fo.fused_softmax(S, 1), torch.softmax(t, 0).numpy())
    assert np.allclose(fo.fused_softmax(S, 2), torch.softmax(t, 1).numpy())
    assert np.allclose(fo.fused_softmax(S[:, :, 0], 2).sum(axis=1), 1)
    with pytest.raises(AssertionError):
        fo.fused_softmax(S, 3)


def test_backward_given_forward_results_equals_recomputed():
    """The flash backwards take the saved (O, l, m) (src_cpp/FlashAttention.cpp:194-252); with the
    exact forward results they must reproduce the naive backwards that recompute everything."""
    rng = np.random.default_rng(11)
    Q, K, V, G = (np.asfortranarray(rng.standard_normal((48, 6, 2))) for _ in range(4))
    O, l, m = fo.circulant_fa(Q, K, V, 9)
    a = fo.circulant_backward_given(Q, K, V, O, G, l, m, 9)
    b = fo.circulant_backward(Q, K, V, G, 9)
    for x, y in zip(a, b):
        assert np.abs(x - y).max() < 1e-12
    O, l, m = fo.dense_fa(Q, K, V)
    a = fo.dense_fa_backward_blocked(Q, K, V, O, G, l, m)
    b = fo.dense_backward(Q, K, V, G)
    for x, y in zip(a, b):
        assert np.abs(x - y).max() < 1e-12


def test_circulant2d_oracle_definition():
    """2-D circulant attention (reference todo, README.md:38-41,53): with W == X == Y every key is used once,
    so it must equal dense attention; its backward against central finite differences."""
    rng = np.random.default_rng(5)
    Q, K, V = (np.asfortranarray(rng.standard_normal((5, 5, 4, 2))) for _ in range(3))
    O, l, m = fo.circulant2d_fa(Q, K, V, 5)
    y, l2, m2 = fo.dense_fa(Q, K, V)
    assert np.abs(O - y).max() < 1e-12 and np.abs(l.reshape(-1) - l2.reshape(-1)).max() < 1e-12
    keys = fo.circulant2d_keys(6, 7, 3)
    assert keys.shape == (9, 42) and keys[4, 0] == 0 and sorted(keys[:, 0]) == sorted([0, 1, 5, 6, 7, 11, 36, 37, 41])
    Q, K, V, G = (np.asfortranarray(rng.standard_normal((6, 7, 3, 1))) for _ in range(4))
    grads = fo.circulant2d_backward(Q, K, V, G, 3)
    f = lambda q, k, v: (fo.circulant2d_fa(q, k, v, 3)[0] * G).sum()
    eps = 1e-6
    for which, gr in enumerate(grads):
        for idx in ((2, 3, 1, 0), (0, 0, 0, 0), (5, 6, 2, 0)):
            args_p, args_m = [Q.copy(), K.copy(), V.copy()], [Q.copy(), K.copy(), V.copy()]
            args_p[which][idx] += eps
            args_m[which][idx] -= eps
            assert abs((f(*args_p) - f(*args_m)) / (2 * eps) - gr[idx]) < 1e-7


def test_circulant2d_backward_given_equals_recomputing_form():
    """the (O, l, m)-taking form of the 2-D circulant backward equals the recomputing one when handed exact stats"""
    rng = np.random.default_rng(5)
    Q, K, V, G = (np.asfortranarray(rng.standard_normal((8, 6, 4, 2))) for _ in range(4))
    O, l, m = fo.circulant2d_fa(Q, K, V, 3)
    for a, b in zip(fo.circulant2d_backward_given(Q, K, V, O, G, l, m, 3), fo.circulant2d_backward(Q, K, V, G, 3)):
        assert np.abs(a - b).max() < 1e-12
