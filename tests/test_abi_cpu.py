"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol
include/fa_sm100a.h declares, validates arguments like the reference's Julia signatures, refuses
to compute without a device (no CPU fallback), and its integer index sets are bit-exact."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import fa_oracle as fo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def fa():
    import fa_sm100a
    return fa_sm100a


def test_header_symbols_exported(fa):
    hdr = open(os.path.join(ROOT, "include", "fa_sm100a.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(fa_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 23
    assert declared == set(fa.EXPORTED_SYMBOLS)
    raw = ctypes.CDLL(os.path.join(ROOT, "flashattention.jl_b200", "lib", "libfa_sm100a.so"))
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in include/fa_sm100a.h but not exported"
    assert fa.lib.fa_version() == 1


def test_library_is_sm100a_only():
    import subprocess
    so = os.path.join(ROOT, "flashattention.jl_b200", "lib", "libfa_sm100a.so")
    out = subprocess.run(["cuobjdump", "--list-elf", so], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


@pytest.mark.parametrize("N,W", [(8, 3), (16, 5), (16, 7), (32, 8), (16, 4), (64, 63), (9, 9), (4096, 255)])
def test_circulant_index_bit_exact(fa, N, W):
    keys = fa.circulant_keys(N, W).numpy()
    assert np.array_equal(keys, fo.circulant_keys(N, W))
    for n in (1, W, W + 1, N * W // 2, N * W):
        assert fa.cartesian_circulant(n, N, W) == fo.cartesian_circulant(n, N, W)


@pytest.mark.parametrize("spatial,W,stride,pad", [
    ((20,), 5, 2, 2), ((22,), 5, 5, 0), ((64,), 7, None, None), ((9, 8), 3, 2, 1), ((64, 64), 7, None, None),
    ((5, 6, 4), 3, 2, 1), ((6, 6, 6), 5, 5, 2), ((64, 64, 64), 5, 5, 3)])
def test_window_index_and_count_bit_exact(fa, spatial, W, stride, pad):
    assert fa.window_counts(spatial, W, stride, pad) == fo.window_counts(spatial, W, stride, pad)
    if np.prod(spatial) <= 4096:
        idx = fo.window_index(spatial, W, stride, pad)
        assert np.array_equal(fa.window_index(spatial, W, stride, pad).numpy(), idx)
        cnt = np.zeros(int(np.prod(spatial)), np.int64)
        np.add.at(cnt, idx[idx >= 0], 1)
        assert np.array_equal(fa.window_count(spatial, W, stride, pad).numpy().reshape(-1, order="F"), cnt)


def test_argument_validation(fa):
    L = fa.lib
    buf = (ctypes.c_int64 * 16)()
    assert L.fa_circulant_index(4, 5, buf) == 1                       # W > N
    assert b"W <= N" in L.fa_last_error_string()
    dims = (ctypes.c_int64 * 1)(4)
    assert L.fa_window_index(1, dims, 9, 9, 0, buf, None) == 1        # window larger than padded extent
    assert L.fa_window_index(4, dims, 3, 3, 0, buf, None) == 1        # ndim > 3
    assert L.fa_softmax(None, None, 4, 4, 1, 3, 0, None) == 1         # dims not in (1,2): src/fused_softmax.jl:12
    assert L.fa_dense_fwd(None, None, None, None, None, None, 8, 8, 8, 1, 0, 0, None) == 1   # NULL pointers
    assert L.fa_dense_fwd(None, None, None, None, None, None, 8, 8, 8, 1, 7, 0, None) == 1   # bad dtype


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(fa):
    q = fa.jl_randn((16, 8, 2), 0, device="cpu")
    for call in (lambda: fa.dense_fa(q, q, q), lambda: fa.circulant_fa(q, q, q, 5), lambda: fa.windowed_fa(q, q, q, 4)):
        with pytest.raises(fa.FaError, match="no CUDA device"):
            call()


def test_julia_layout_helpers(fa):
    x = np.asfortranarray(np.arange(24, dtype=np.float32).reshape((2, 3, 4), order="F"))
    t = fa.jl_array(x)
    assert tuple(t.shape) == (2, 3, 4) and t.stride() == (1, 2, 6) and fa.is_jl_contiguous(t)
    assert t.permute(2, 1, 0).reshape(-1).tolist() == list(range(24))            # linear memory == Julia order
    r = fa._jl_reshape(t, (6, 4))
    assert r.stride() == (1, 6) and r[5, 3] == 23
    e = fa.jl_empty((5, 7, 3), torch.bfloat16, "cpu")
    assert e.stride() == (1, 5, 35)
    a, b = fa.jl_randn((4, 3, 2), 1, device="cpu"), fa.jl_randn((4, 3, 2), 1, device="cpu")
    assert torch.equal(a, b)


def test_circulant_sparse_matrices_match_reference_construction():
    """circulant(N, M) / circulant(V) / batch_circulant (src/utils.jl:19-34): host-side index logic, no GPU.
    Dense form against the definition: column j has V[w, j] at row first(cartesian_circulant((j-1)M + w))."""
    import numpy as np
    import torch
    import fa_sm100a as fa
    from oracle import fa_oracle as fo
    N, M, B = 12, 5, 3
    rng = np.random.default_rng(0)
    V = rng.standard_normal((M, N, B))
    dense = fa.circulant(torch.from_numpy(V[:, :, 0])).to_dense().numpy()
    want = np.zeros((N, N))
    for j in range(1, N + 1):
        for w in range(1, M + 1):
            i, jj = fo.cartesian_circulant((j - 1) * M + w, N, M)
            assert jj == j
            want[i - 1, j - 1] = V[w - 1, j - 1, 0]
    assert np.array_equal(dense, want)
    ones = fa.circulant(N, M).to_dense().numpy()
    assert np.array_equal(ones, (want != 0).astype(np.float64)) and ones.sum() == N * M
    bd = fa.batch_circulant(torch.from_numpy(V)).to_dense().numpy()
    for b in range(B):
        blk = bd[b * N:(b + 1) * N, b * N:(b + 1) * N]
        assert np.array_equal(blk, fa.circulant(torch.from_numpy(V[:, :, b])).to_dense().numpy())
    assert np.count_nonzero(bd) == B * N * M


def test_circulant2d_index_bit_exact():
    import numpy as np
    import fa_sm100a as fa
    from oracle import fa_oracle as fo
    for X, Y, W in ((6, 7, 3), (8, 8, 8), (5, 9, 4), (16, 3, 1)):
        got = fa.circulant2d_keys(X, Y, W).numpy()
        assert np.array_equal(got, fo.circulant2d_keys(X, Y, W))
    import pytest
    with pytest.raises(fa.FaError):
        fa.circulant2d_keys(4, 8, 5)          # W > X would duplicate keys


def test_workspace_queries_are_pure_host_functions(fa):
    """Workspace sizes are computed on the host (no device needed) and ordered as documented in include/fa_sm100a.h:
    the _ex query of the 2-D neighbourhood backward is never below the basic one, larger exactly where the tcgen05
    backward applies (16-bit, d = dv in {64, 128}, X % 64 == 0), and equal to it when FORCE_SIMT is set."""
    L = fa.lib
    base = L.fa_workspace_bytes_circulant2d_bwd(128, 16, 2)
    assert base >= 128 * 16 * 2 * 4
    import torch
    bf, f32 = fa._DTYPES[torch.bfloat16], fa._DTYPES[torch.float32]
    assert L.fa_workspace_bytes_circulant2d_bwd_ex(128, 16, 64, 64, 2, 7, bf, 0) > base          # tcgen05 path
    assert L.fa_workspace_bytes_circulant2d_bwd_ex(128, 16, 64, 64, 2, 7, bf, fa.FA_FLAG_FORCE_SIMT) == base
    assert L.fa_workspace_bytes_circulant2d_bwd_ex(128, 16, 64, 64, 2, 7, f32, 0) == base        # Float32: exact kernels
    assert L.fa_workspace_bytes_circulant2d_bwd_ex(100, 16, 64, 64, 2, 7, bf, 0) == L.fa_workspace_bytes_circulant2d_bwd(100, 16, 2)
    assert L.fa_workspace_bytes_circulant2d_bwd_ex(128, 16, 32, 32, 2, 7, bf, 0) == base         # d = 32: exact kernels
    # slab backward workspace: the delta buffer of the slab's own windows
    dims = fa._i64arr((64, 64, 10))
    assert L.fa_workspace_bytes_windowed_slab_bwd(3, dims, 64, 64, 1, 5, 5, 3, 0, 2) >= 125 * 14 * 14 * 2 * 4
    assert L.fa_workspace_bytes_windowed_slab_bwd(3, dims, 64, 64, 1, 5, 4, 3, 0, 2) == 0        # overlapping windows: rejected
