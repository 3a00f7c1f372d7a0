"""The C/OpenMP restatement (oracle/fa_oracle.c, the timed CPU baseline) against the numpy oracle."""
import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import fa_oracle as fo


def randn(shape, seed):
    return np.asfortranarray(np.random.default_rng(seed).standard_normal(shape).astype(np.float32))


def close(a, b, tol=2e-5):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    return np.nanmax(np.abs(a - b)) <= tol * max(1.0, np.nanmax(np.abs(b)))


@pytest.mark.parametrize("shape,dv", [((30, 12, 2), 6), ((1024, 64, 2), 64), ((600, 128, 1), 128), ((7, 5, 8, 2), 8)])
def test_c_dense(shape, dv):
    q, k = randn(shape, 0), randn(shape, 1)
    v = randn(shape[:-2] + (dv, shape[-1]), 2)
    for a, b in zip(co.dense_fa(q, k, v), fo.dense_fa(q.astype(np.float64), k.astype(np.float64), v.astype(np.float64))):
        assert a.shape == b.shape and close(a, b)


@pytest.mark.parametrize("N,d,B,W", [(64, 8, 2, 9), (128, 16, 2, 16), (300, 64, 1, 65), (40, 4, 1, 40)])
def test_c_circulant(N, d, B, W):
    Q, K, V = (randn((N, d, B), s) for s in range(3))
    for a, b in zip(co.circulant_fa(Q, K, V, W), fo.circulant_fa(*(t.astype(np.float64) for t in (Q, K, V)), W)):
        assert close(a, b)


@pytest.mark.parametrize("spatial,W,kws", [((64,), 16, dict(stride=4, pad=0)), ((22,), 5, dict(stride=5, pad=0)),
                                           ((20, 12), 7, {}), ((6, 7, 8), 3, {})])
def test_c_windowed(spatial, W, kws):
    q, k, v = (randn(spatial + (8, 2), s) for s in range(3))
    for a, b in zip(co.windowed_fa(q, k, v, W, **kws), fo.windowed_fa(*(t.astype(np.float64) for t in (q, k, v)), W, **kws)):
        assert a.shape == b.shape and close(a, b)
