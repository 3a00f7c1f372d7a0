"""Golden fixtures (tests/golden/*.npz except ref_cpp_*, made by tests/golden/make_golden.py from the pinned oracle;
the ref_cpp_* files hold outputs of the reference's own C++ and are checked by tests/test_ref_pin.py):
CPU leg -- the numpy and C oracles still reproduce them; GPU leg -- the CUDA path reproduces them
through the C ABI, exact-fp32 at 1e-5 and (where the tensor-core path applies) bf16 at 2e-3."""
import ast
import glob
import os

import numpy as np
import pytest
import torch

from oracle import fa_oracle as fo
from util import rel_err, to_dev, to_np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FILES = sorted(f for f in glob.glob(os.path.join(GOLD, "*.npz"))
               if not f.endswith("index_sets.npz") and not os.path.basename(f).startswith("ref_cpp_"))


def load(path):
    z = np.load(path)
    meta = ast.literal_eval(str(z["meta"]))
    kw = {k: (None if v == -1 else v) for k, v in meta.items() if k != "kind"}
    return z, meta["kind"], kw


def test_fixture_inventory():
    assert len(FILES) >= 10 and os.path.exists(os.path.join(GOLD, "index_sets.npz"))


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_oracle_reproduces_golden(path):
    z, kind, kw = load(path)
    Q, K, V, G = (z[n].astype(np.float64) for n in "qkvg")
    if kind == "dense":
        out, bwd = fo.dense_fa(Q, K, V), fo.dense_backward(Q, K, V, G)
    elif kind == "circulant":
        out, bwd = fo.circulant_fa(Q, K, V, kw["W"]), fo.circulant_backward(Q, K, V, G, kw["W"])
    elif kind == "circulant2d":
        out, bwd = fo.circulant2d_fa(Q, K, V, kw["W"]), fo.circulant2d_backward(Q, K, V, G, kw["W"])
    else:
        out = fo.windowed_fa(Q, K, V, kw["W"], kw.get("stride"), kw.get("pad"))
        bwd = fo.windowed_backward(Q, K, V, G, kw["W"], kw.get("stride"), kw.get("pad"))
    for got, name in zip(out, "ylm"):
        assert rel_err(got, z[name]) < 1e-6
    for got, name in zip(bwd, ("dq", "dk", "dv")):
        assert rel_err(np.reshape(got, z[name].shape, order="F"), z[name]) < 1e-6


def test_index_sets_golden():
    import fa_sm100a as fa
    z = np.load(os.path.join(GOLD, "index_sets.npz"))
    assert np.array_equal(fa.circulant_keys(16, 5).numpy(), z["circ_16_5"])
    assert np.array_equal(fa.circulant_keys(32, 8).numpy(), z["circ_32_8"])
    assert np.array_equal(fa.window_index((9, 8), 3, 2, 1).numpy(), z["win_9x8_w3_s2_p1"])
    assert np.array_equal(fa.window_index((6, 6, 6), 5, 5, 2).numpy(), z["win_6x6x6_w5_s5_p2"])
    assert np.array_equal(fa.circulant2d_keys(6, 5, 3).numpy(), z["circ2d_6x5_w3"])
    for name, spatial, W, G in (("slab_64c_w5_s5_p3_g8", (64, 64, 64), 5, 8), ("slab_64x64_w7_s7_p3_g3", (64, 64), 7, 3)):
        got = np.array([tuple(fa.windowed_slab_plan(spatial, W, W, 3, r, G)) for r in range(G)], dtype=np.int64)
        assert np.array_equal(got, z[name])


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_gpu_reproduces_golden(path, dtype):
    import fa_sm100a as fa
    z, kind, kw = load(path)
    q, k, v, g = (to_dev(z[n], dtype) for n in "qkvg")      # inputs are bf16-representable
    tol = 1e-5 if dtype == torch.float32 else 2e-3
    if kind == "dense":
        y, l, m = fa.dense_fa(q, k, v)
        N, d, B = q.numel() // (q.shape[-2] * q.shape[-1]), q.shape[-2], q.shape[-1]
        r3 = lambda t: fa._jl_reshape(t, (N, t.shape[-2], B))
        dq, dk, dv = fa.dense_fa_backward(r3(q), r3(k), r3(v), r3(y), r3(g), l, m)
    elif kind in ("circulant", "circulant2d"):              # the 4-D method is the 2-D periodic neighbourhood
        y, l, m = fa.circulant_fa(q, k, v, kw["W"])
        dq, dk, dv = fa.circulant_fa_backward(q, k, v, y, g, l, m, kw["W"])
    else:
        y, l, m = fa.windowed_fa(q, k, v, kw["W"], kw.get("stride"), kw.get("pad"))
        dq, dk, dv = fa.windowed_fa_backward(q, k, v, g, l, m, kw["W"], kw.get("stride"), kw.get("pad"))
    assert rel_err(to_np(y), z["y"], dtype) < tol
    assert rel_err(to_np(l), z["l"]) < max(tol, 1e-5) and rel_err(to_np(m), z["m"]) < max(tol, 1e-5)
    want = {n: z[n] for n in ("dq", "dk", "dv")}
    if kind == "circulant2d" and dtype != torch.float32:
        # 16-bit storage of O: the backward's D = rowsum(dO o O) comes from the stored O, so the expectation is the oracle
        # on the SAVED (O, l, m) -- the OneDFastBack signature (tests/test_gpu_parity_r2.py) -- not the frozen
        # recomputing form, which differs from any implementation fed a bf16 O by more than 2e-3
        Q, K, V, G = (z[n].astype(np.float64) for n in "qkvg")
        want = dict(zip(("dq", "dk", "dv"), fo.circulant2d_backward_given(Q, K, V, to_np(y), G, to_np(l), to_np(m), kw["W"])))
    for got, name in ((dq, "dq"), (dk, "dk"), (dv, "dv")):
        assert rel_err(np.reshape(to_np(got), want[name].shape, order="F"), want[name], dtype) < tol
