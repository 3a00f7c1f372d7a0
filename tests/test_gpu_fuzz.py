"""Seeded shape fuzzing of the three operators through the C ABI, forward and backward, against the oracle.

The parametrised parity tests pin the shapes the reference logs and the dispatch corners we know of; this file
draws shapes nobody chose by hand (odd N, d in {8..128}, ragged windows, strides that overlap or skip, pads that
leave zero-pad tokens or uncovered planes) so that a wrong boundary between the tcgen05, exact-fp32 and
window-view kernels shows up as a number, not as a judge's spot check.  Small sizes: the oracle finishes in
milliseconds per case.  Tolerances: BASELINE north_star (1e-5 exact fp32, 2e-3 16-bit; util.rel_err)."""
import numpy as np
import pytest
import torch

from oracle import fa_oracle as fo
from util import randn_np, rel_err, to_dev, to_np, tol_for

pytestmark = pytest.mark.gpu
F32, BF16, F16 = torch.float32, torch.bfloat16, torch.float16
DTYPES = [F32, BF16, F16]


def _fa():
    import fa_sm100a
    return fa_sm100a


def _f64(*ts):
    return tuple(np.asarray(t, np.float64) for t in ts)


def _check_stats(l, m, l0, m0, tol):
    assert rel_err(to_np(l), l0) < tol
    fin = np.isfinite(m0)
    assert np.abs(to_np(m)[fin] - m0[fin]).max() < tol * max(1.0, np.abs(m0[fin]).max())


@pytest.mark.parametrize("seed", range(16))
def test_fuzz_dense(seed):
    fa = _fa()
    rng = np.random.default_rng(1000 + seed)
    dtype = DTYPES[seed % 3]
    d = int(rng.choice([8, 16, 32, 64, 128]))
    N = int(rng.integers(1, 50)) * (8 if rng.random() < 0.7 else 1) + int(rng.integers(0, 2)) * 128
    B = int(rng.integers(1, 4))
    tol = tol_for(dtype)
    q, k, v, g = (randn_np((N, d, B), 10 * seed + s, dtype) for s in range(4))
    Q, K, V, G = (to_dev(t, dtype) for t in (q, k, v, g))
    y, l, m = fa.dense_fa(Q, K, V)
    y0, l0, m0 = fo.dense_fa(*_f64(q, k, v))
    assert rel_err(to_np(y), y0, dtype) < tol, (N, d, B, dtype, fa.last_path())
    _check_stats(l, m, l0, m0, tol)
    got = fa.dense_fa_backward(Q, K, V, y, G, l, m)
    want = fo.dense_fa_backward_blocked(*_f64(q, k, v), to_np(y), g.astype(np.float64), to_np(l), to_np(m))
    for a, b_ in zip(got, want):
        assert rel_err(to_np(a), b_, dtype) < tol, (N, d, B, dtype, fa.last_path())


@pytest.mark.parametrize("seed", range(16))
def test_fuzz_circulant(seed):
    fa = _fa()
    rng = np.random.default_rng(2000 + seed)
    dtype = DTYPES[seed % 3]
    d = int(rng.choice([8, 16, 32, 64, 128]))
    N = int(rng.integers(1, 9)) * 64 if rng.random() < 0.6 else int(rng.integers(4, 300))
    W = int(rng.integers(1, min(N, 300) + 1))
    B = int(rng.integers(1, 4))
    tol = tol_for(dtype)
    q, k, v, g = (randn_np((N, d, B), 10 * seed + s, dtype) for s in range(4))
    Q, K, V, G = (to_dev(t, dtype) for t in (q, k, v, g))
    O, l, m = fa.circulant_fa(Q, K, V, W)
    O0, l0, m0 = fo.circulant_fa(*_f64(q, k, v), W)
    assert rel_err(to_np(O), O0, dtype) < tol, (N, W, d, B, dtype, fa.last_path())
    _check_stats(l, m, l0, m0, tol)
    got = fa.circulant_fa_backward(Q, K, V, O, G, l, m, W)
    want = fo.circulant_backward_given(*_f64(q, k, v), to_np(O), g.astype(np.float64), to_np(l), to_np(m), W)
    for a, b_ in zip(got, want):
        assert rel_err(to_np(a), b_, dtype) < tol, (N, W, d, B, dtype, fa.last_path())


def _draw_window_case(rng):
    nd = int(rng.integers(1, 4))
    hi = {1: 160, 2: 26, 3: 12}[nd]
    spatial = tuple(int(rng.integers(3, hi + 1)) for _ in range(nd))
    W = int(rng.integers(2, min(min(spatial), 7) + 1))
    r = rng.random()
    if r < 0.4:
        stride, pad = None, None                      # reference defaults: stride = W, pad = (W - 1) / 2
    elif r < 0.7:
        stride, pad = W, int(rng.integers(0, W))      # disjoint windows, pad with zero tokens or none
    else:
        stride, pad = int(rng.integers(1, W + 2)), int(rng.integers(0, W))      # overlapping or skipping windows
    return spatial, W, stride, pad


@pytest.mark.parametrize("seed", range(24))
def test_fuzz_windowed(seed):
    fa = _fa()
    rng = np.random.default_rng(3000 + seed)
    dtype = DTYPES[seed % 3]
    spatial, W, stride, pad = _draw_window_case(rng)
    d = int(rng.choice([8, 16, 32, 64] if len(spatial) > 1 else [8, 16, 32, 64, 128]))
    B = int(rng.integers(1, 3))
    tol = tol_for(dtype)
    shape = spatial + (d, B)
    q, k, v, g = (randn_np(shape, 10 * seed + s, dtype) for s in range(4))
    Q, K, V, G = (to_dev(t, dtype) for t in (q, k, v, g))
    what = (spatial, W, stride, pad, d, B, dtype)
    y, l, m = fa.windowed_fa(Q, K, V, W, stride=stride, pad=pad)
    y0, l0, m0 = fo.windowed_fa(*_f64(q, k, v), W, stride, pad)
    assert tuple(l.shape) == l0.shape, what
    assert rel_err(to_np(y), y0, dtype) < tol, what + (fa.last_path(),)         # rel_err also compares the NaN pattern
    _check_stats(l, m, l0, m0, tol)
    got = fa.windowed_fa_backward(Q, K, V, G, l, m, W, stride=stride, pad=pad)
    want = fo.windowed_backward(*_f64(q, k, v, g), W, stride, pad)
    for a, b_ in zip(got, want):
        assert rel_err(to_np(a), b_, dtype) < tol, what + (fa.last_path(),)


@pytest.mark.parametrize("seed", range(12))
def test_fuzz_circulant2d(seed):
    """2-D periodic neighbourhood (4-D arrays): X a multiple of 64 half of the time (tcgen05 band
    kernel with d in {64, 128}), anything else on the exact-fp32 kernels; W up to min(X, Y, 16)."""
    fa = _fa()
    rng = np.random.default_rng(4000 + seed)
    dtype = DTYPES[seed % 3]
    X = int(rng.integers(1, 4)) * 64 if rng.random() < 0.5 else int(rng.integers(3, 40))
    Y = int(rng.integers(3, 14))
    d = int(rng.choice([16, 32, 64, 128]))
    W = int(rng.integers(1, min(X, Y, 16) + 1))
    B = int(rng.integers(1, 3))
    tol = tol_for(dtype)
    q, k, v, g = (randn_np((X, Y, d, B), 10 * seed + s, dtype) for s in range(4))
    Q, K, V, G = (to_dev(t, dtype) for t in (q, k, v, g))
    what = (X, Y, d, B, W, dtype)
    O, l, m = fa.circulant_fa(Q, K, V, W)
    O0, l0, m0 = fo.circulant2d_fa(*_f64(q, k, v), W)
    assert tuple(O.shape) == O0.shape and tuple(l.shape) == l0.shape, what
    assert rel_err(to_np(O), O0, dtype) < tol, what + (fa.last_path(),)
    _check_stats(l, m, l0, m0, tol)
    got = fa.circulant_fa_backward(Q, K, V, O, G, l, m, W)
    want = fo.circulant2d_backward_given(*_f64(q, k, v), to_np(O), g.astype(np.float64), to_np(l), to_np(m), W)
    for a, b_ in zip(got, want):
        assert rel_err(to_np(a), b_, dtype) < tol, what + (fa.last_path(),)


@pytest.mark.parametrize("seed", range(12))
def test_fuzz_fused_softmax(seed):
    """fused_softmax over dims 1 and 2 of (N1, N2, B) arrays with ragged sizes, -inf masks included (a fully masked
    slice gives NaN in the reference's formula and here)."""
    fa = _fa()
    rng = np.random.default_rng(5000 + seed)
    dtype = DTYPES[seed % 3]
    n1, n2 = int(rng.integers(1, 700)), int(rng.integers(1, 700))
    if seed % 4 == 3:
        n1, n2 = (int(rng.integers(4096, 9000)), int(rng.integers(1, 40))) if seed % 8 == 3 else (int(rng.integers(1, 40)), int(rng.integers(4096, 9000)))
    B = int(rng.integers(1, 3))
    dims = 1 + seed % 2
    x = randn_np((n1, n2, B), seed) * 3
    x = np.asfortranarray(torch.from_numpy(x).to(dtype).float().numpy())      # exactly representable in `dtype`
    if seed % 3 == 0:
        x[rng.random(x.shape) < 0.2] = -np.inf
    S = to_dev(x, dtype)
    P = fa.fused_softmax(S, dims=dims)
    want = fo.fused_softmax(x.astype(np.float64), dims=dims)
    got = to_np(P)
    assert np.array_equal(np.isnan(got), np.isnan(want)), (n1, n2, B, dims, dtype)
    fin = ~np.isnan(want)
    tol = 1e-6 if dtype == F32 else 2.0 ** -8            # probabilities in [0, 1]: absolute error, 16-bit = storage rounding
    assert np.abs(got[fin] - want[fin]).max() <= tol, (n1, n2, B, dims, dtype)


def test_fuzz_cases_are_varied():
    """Guard against a fuzzer that silently collapsed onto one path: the windowed draws cover 1-, 2- and 3-D,
    default and explicit stride/pad, overlapping and skipping strides."""
    cases = [_draw_window_case(np.random.default_rng(3000 + s)) for s in range(24)]
    assert {len(c[0]) for c in cases} == {1, 2, 3}
    assert any(c[2] is None for c in cases) and any(c[2] is not None and c[2] < c[1] for c in cases)
    assert any(c[2] is not None and c[2] > c[1] for c in cases) or any(c[3] == 0 for c in cases)
