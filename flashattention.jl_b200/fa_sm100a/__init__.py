"""fa_sm100a -- Python host mirror of the FlashAttention.jl API over libfa_sm100a.so.

The reference is a Julia package; this container (and the GPU box) has no Julia runtime, so
the same C ABI that the Julia ``ccall`` layer binds (``julia/FlashAttention``)
is driven here through ``ctypes``.  Function names, argument meaning, defaults and error
behaviour follow the reference's exported API (reference src/FlashAttention.jl:13,20-21,26-27):

    dense_fa, dense_fa_, dense_fa_backward, windowed_fa, block_fa, circulant_fa, circulant_fa_,
    fused_softmax, fused_softmax_, window, unwindow, cartesian_circulant, circulant_keys,
    dense_dpa, windowed_dpa, block_dpa, circulant_dpa   (naive oracles, kept naive)

(``f_`` is Julia's ``f!``.)  Arrays are torch tensors with the reference's *Julia shapes*
``(spatial..., d, B)`` and column-major strides (see :func:`jl_empty`, :func:`jl_array`), so
``tensor.data_ptr()`` is exactly what a Julia ``CuArray``/``Array`` passes to ``ccall``.
CUDA tensors go to the device entry points (no copies, caller's current stream); CPU tensors
go to the ``*_host`` entry points (H2D + kernels + D2H inside the library).

PyTorch is plumbing only (device memory, streams).  The compute is the hand-written CUDA in
``csrc/``; there is no CPU or torch fallback -- if the library is missing or no GPU is
present the calls raise.
"""
from __future__ import annotations

import ctypes
import math
import os
from typing import Optional, Sequence, Tuple

import torch

__all__ = [
    "lib", "FaError", "FA_FLAG_FORCE_SIMT", "FA_FLAG_BF16_INTERNALS", "FA_FLAG_OUT_F32", "FA_FLAG_HOST_NO_REGISTER", "jl_host_empty", "jl_empty", "jl_array", "jl_randn", "is_jl_contiguous", "last_path",
    "dense_fa", "dense_fa_", "dense_fa_backward", "windowed_fa", "windowed_fa_backward", "block_fa",
    "circulant_fa", "circulant_fa_", "circulant_fa_backward", "fused_softmax", "fused_softmax_",
    "window", "unwindow", "window_index", "window_count", "cartesian_circulant", "circulant_keys", "circulant2d_keys", "circulant", "batch_circulant",
    "dense_dpa", "windowed_dpa", "block_dpa", "circulant_dpa", "shard_batch", "ring_dense_fa", "ring_dense_fa_backward",
]

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("FA_SM100A_LIB", os.path.join(os.path.dirname(_HERE), "lib", "libfa_sm100a.so"))

FA_F32, FA_F16, FA_BF16 = 0, 1, 2
FA_FLAG_FORCE_SIMT = 1
FA_FLAG_BF16_INTERNALS = 2
FA_FLAG_OUT_F32 = 4          # 16-bit inputs, float32 outputs (fp32 accumulators stored unrounded; tcgen05 kernels only)
FA_FLAG_HOST_NO_REGISTER = 8  # *_host entry points: do not page-lock pageable caller buffers for the call
_DTYPES = {torch.float32: FA_F32, torch.float16: FA_F16, torch.bfloat16: FA_BF16}


class FaError(RuntimeError):
    """Raised when a libfa_sm100a entry point returns a non-zero status."""


def _load():
    if not os.path.exists(_LIB_PATH):
        raise ImportError(
            f"{_LIB_PATH} not found: build it with `make -C flashattention.jl_b200` "
            "(or __graft_entry__.build()).  There is no fallback implementation.")
    L = ctypes.CDLL(_LIB_PATH)
    i64, vp, ci, sz = ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
    pi64 = ctypes.POINTER(ctypes.c_int64)
    sigs = {
        "fa_version": (ci, []),
        "fa_last_error_string": (ctypes.c_char_p, []),
        "fa_last_path": (ctypes.c_char_p, []),
        "fa_device_count": (ci, []),
        "fa_circulant_index": (ci, [i64, i64, pi64]),
        "fa_window_index": (ci, [ci, pi64, i64, i64, i64, pi64, pi64]),
        "fa_window_count": (ci, [ci, pi64, i64, i64, i64, pi64]),
        "fa_dense_fwd": (ci, [vp, vp, vp, vp, vp, vp, i64, i64, i64, i64, ci, ci, vp]),
        "fa_workspace_bytes_dense_bwd": (sz, [i64, i64, i64, i64, ci, ci]),
        "fa_dense_bwd": (ci, [vp] * 10 + [i64, i64, i64, i64, ci, ci, vp, sz, vp]),
        "fa_circulant_fwd": (ci, [vp, vp, vp, vp, vp, vp, i64, i64, i64, i64, i64, ci, ci, vp]),
        "fa_workspace_bytes_circulant_bwd": (sz, [i64, i64, i64, i64, i64, ci, ci]),
        "fa_circulant_bwd": (ci, [vp] * 10 + [i64, i64, i64, i64, i64, ci, ci, vp, sz, vp]),
        "fa_circulant2d_index": (ci, [i64, i64, i64, pi64]),
        "fa_circulant2d_fwd": (ci, [vp] * 6 + [i64, i64, i64, i64, i64, i64, ci, ci, vp]),
        "fa_workspace_bytes_circulant2d_bwd": (sz, [i64, i64, i64]),
        "fa_workspace_bytes_circulant2d_bwd_ex": (sz, [i64, i64, i64, i64, i64, i64, ci, ci]),
        "fa_circulant2d_bwd": (ci, [vp] * 10 + [i64, i64, i64, i64, i64, i64, ci, ci, vp, sz, vp]),
        "fa_workspace_bytes_windowed_fwd": (sz, [ci, pi64, i64, i64, i64, i64, i64, i64, ci, ci]),
        "fa_windowed_fwd": (ci, [vp] * 6 + [ci, pi64, i64, i64, i64, i64, i64, i64, ci, ci, vp, sz, vp]),
        "fa_workspace_bytes_windowed_bwd": (sz, [ci, pi64, i64, i64, i64, i64, i64, i64, ci, ci]),
        "fa_windowed_bwd": (ci, [vp] * 9 + [ci, pi64, i64, i64, i64, i64, i64, i64, ci, ci, vp, sz, vp]),
        "fa_windowed_slab_plan": (ci, [ci, pi64, i64, i64, i64, ci, ci, pi64]),
        "fa_windowed_slab_fwd": (ci, [vp] * 6 + [ci, pi64, i64, i64, i64, i64, i64, i64, i64, i64, ci, ci, vp]),
        "fa_workspace_bytes_windowed_slab_bwd": (sz, [ci, pi64, i64, i64, i64, i64, i64, i64, i64, i64]),
        "fa_windowed_slab_bwd": (ci, [vp] * 9 + [ci, pi64, i64, i64, i64, i64, i64, i64, i64, i64, ci, ci, vp, sz, vp]),
        "fa_windowed_halo_plan": (ci, [ci, pi64, i64, i64, i64, ci, ci, pi64]),
        "fa_windowed_slab_fwd_sums": (ci, [vp] * 6 + [ci, pi64, i64, i64, i64, i64, i64, i64, i64, i64, ci, ci, vp]),
        "fa_window_divide": (ci, [vp, vp, ci, pi64, i64, i64, i64, i64, i64, i64, i64, ci, vp]),
        "fa_window": (ci, [vp, vp, ci, pi64, i64, i64, i64, i64, i64, ci, vp]),
        "fa_unwindow": (ci, [vp, vp, ci, pi64, i64, i64, i64, i64, i64, ci, vp]),
        "fa_softmax": (ci, [vp, vp, i64, i64, i64, ci, ci, vp]),
        "fa_cast": (ci, [vp, vp, i64, ci, ci, vp]),
        "fa_dense_fwd_host": (ci, [vp] * 6 + [i64, i64, i64, i64, ci, ci, ci]),
        "fa_circulant_fwd_host": (ci, [vp] * 6 + [i64, i64, i64, i64, i64, ci, ci, ci]),
        "fa_windowed_fwd_host": (ci, [vp] * 6 + [ci, pi64, i64, i64, i64, i64, i64, i64, ci, ci, ci]),
        "fa_dense_bwd_host": (ci, [vp] * 10 + [i64, i64, i64, i64, ci, ci, ci]),
        "fa_circulant_bwd_host": (ci, [vp] * 10 + [i64, i64, i64, i64, i64, ci, ci, ci]),
        "fa_windowed_bwd_host": (ci, [vp] * 9 + [ci, pi64, i64, i64, i64, i64, i64, i64, ci, ci, ci]),
        "fa_host_alloc": (ci, [ctypes.POINTER(vp), sz, ci]),
        "fa_host_free": (ci, [vp]),
        "fa_release_host_staging": (ci, []),
        "fa_shard_batch": (ci, [i64, ci, ci, pi64, pi64]),
        "fa_merge_partials": (ci, [vp] * 7 + [i64, i64, i64, ci, ci, vp]),
        "fa_workspace_bytes_ring_dense_fwd": (sz, [i64, i64, i64, i64, ci]),
        "fa_ring_dense_fwd": (ci, [vp] * 6 + [i64, i64, i64, i64, ci, ci, vp, ci, ci, vp, sz, vp]),
        "fa_workspace_bytes_ring_dense_bwd": (sz, [i64, i64, i64, i64, ci, ci]),
        "fa_ring_dense_bwd": (ci, [vp] * 10 + [i64, i64, i64, i64, ci, ci, vp, ci, ci, vp, sz, vp]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(L, name)          # AttributeError here == header/library mismatch
        fn.restype, fn.argtypes = res, args
    return L


lib = _load()
EXPORTED_SYMBOLS = (
    "fa_version fa_last_error_string fa_last_path fa_device_count fa_circulant_index fa_window_index "
    "fa_window_count fa_dense_fwd fa_workspace_bytes_dense_bwd fa_dense_bwd fa_circulant_fwd "
    "fa_workspace_bytes_circulant_bwd fa_circulant_bwd fa_workspace_bytes_windowed_fwd fa_windowed_fwd "
    "fa_workspace_bytes_windowed_bwd fa_windowed_bwd fa_window fa_unwindow fa_softmax fa_dense_fwd_host "
    "fa_circulant_fwd_host fa_windowed_fwd_host fa_release_host_staging fa_shard_batch fa_merge_partials "
    "fa_workspace_bytes_ring_dense_fwd fa_ring_dense_fwd fa_workspace_bytes_ring_dense_bwd fa_ring_dense_bwd "
    "fa_circulant2d_index fa_circulant2d_fwd fa_workspace_bytes_circulant2d_bwd fa_circulant2d_bwd "
    "fa_workspace_bytes_circulant2d_bwd_ex fa_windowed_slab_plan fa_windowed_slab_fwd fa_workspace_bytes_windowed_slab_bwd fa_windowed_slab_bwd "
    "fa_dense_bwd_host fa_circulant_bwd_host fa_windowed_bwd_host fa_host_alloc fa_host_free fa_cast fa_windowed_halo_plan fa_windowed_slab_fwd_sums fa_window_divide").split()


def _check(rc: int, what: str):
    if rc != 0:
        raise FaError(f"{what} failed (status {rc}): {lib.fa_last_error_string().decode()}")


def last_path() -> str:
    """Kernel family the last compute call on this thread dispatched to ("tc" or "simt")."""
    return lib.fa_last_path().decode()


# --------------------------------------------------------------------------------------------
# Julia-shaped (column-major) tensors
# --------------------------------------------------------------------------------------------
def jl_empty(shape: Sequence[int], dtype=torch.float32, device="cuda") -> torch.Tensor:
    """``Array{T}(undef, shape...)``: tensor of Julia shape ``shape`` with column-major strides."""
    shape = tuple(int(s) for s in shape)
    t = torch.empty(shape[::-1], dtype=dtype, device=device)
    return t.permute(*range(len(shape) - 1, -1, -1))


class _HostBlock:
    """Owner of one fa_host_alloc allocation (freed when the last tensor viewing it dies)."""
    def __init__(self, nbytes: int, device: int):
        p = ctypes.c_void_p()
        _check(lib.fa_host_alloc(ctypes.byref(p), nbytes, int(device)), "fa_host_alloc")
        self.ptr, self.nbytes = p.value, nbytes

    def __del__(self):
        if getattr(self, "ptr", None):
            lib.fa_host_free(ctypes.c_void_p(self.ptr))
            self.ptr = None


def jl_host_empty(shape: Sequence[int], dtype=torch.float32, device: int = 0) -> torch.Tensor:
    """Column-major HOST tensor of Julia shape ``shape`` in page-locked memory on the NUMA node of GPU ``device``
    (``fa_host_alloc``): the fastest source / destination for the ``*_host`` entry points."""
    import numpy as np
    shape = tuple(int(s) for s in shape)
    n = 1
    for s_ in shape:
        n *= s_
    esz = torch.empty((), dtype=dtype).element_size()
    blk = _HostBlock(max(n * esz, 1), device)
    raw = (ctypes.c_uint8 * blk.nbytes).from_address(blk.ptr)
    raw._fa_block = blk                    # storage -> numpy array -> ctypes buffer -> block: freed with the last view
    arr = np.frombuffer(raw, dtype=np.uint8)
    t = torch.from_numpy(arr).view(dtype)[:n].reshape(shape[::-1])
    return t.permute(*range(len(shape) - 1, -1, -1))


def is_jl_contiguous(t: torch.Tensor) -> bool:
    return t.permute(*range(t.ndim - 1, -1, -1)).is_contiguous()


def jl_array(x, dtype=None, device=None) -> torch.Tensor:
    """Column-major copy/view of ``x`` (torch tensor or numpy array with the Julia shape)."""
    if not isinstance(x, torch.Tensor):
        import numpy as np
        x = torch.from_numpy(np.ascontiguousarray(np.asarray(x).transpose())).permute(*range(x.ndim - 1, -1, -1))
    if dtype is not None and x.dtype != dtype:
        x = x.to(dtype)
    if device is not None and str(x.device) != str(torch.device(device)):
        x = x.to(device)
    if not is_jl_contiguous(x):
        out = jl_empty(x.shape, x.dtype, x.device)
        out.copy_(x)
        x = out
    return x


def jl_randn(shape, seed: int, dtype=torch.float32, device="cuda") -> torch.Tensor:
    """Seeded ``randn(Float32, shape)`` then cast (generated Float32, then cast)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    n = 1
    for s in shape:
        n *= int(s)
    flat = torch.randn(n, generator=g, dtype=torch.float32)
    out = jl_empty(shape, dtype, device)
    out.permute(*range(len(shape) - 1, -1, -1)).reshape(-1).copy_(flat.to(dtype))
    return out


def _dt(t: torch.Tensor) -> int:
    if t.dtype not in _DTYPES:
        raise FaError(f"unsupported eltype {t.dtype}; use float32, float16 or bfloat16")
    return _DTYPES[t.dtype]


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream(t: torch.Tensor):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _same(*ts):
    t0 = ts[0]
    for t in ts[1:]:
        if t.dtype != t0.dtype or t.device != t0.device or t.ndim != t0.ndim:
            raise FaError("q, k, v must share eltype, device and number of dimensions "
                          "(reference signature `where {T, D}`, src/dense.jl:1)")


def _flatten3(x: torch.Tensor) -> Tuple[int, int, int]:
    """(N, d, B) of ``reshape(x, :, d, B)`` (src/dense.jl:6-8)."""
    n = 1
    for s in x.shape[:-2]:
        n *= int(s)
    return n, int(x.shape[-2]), int(x.shape[-1])


def _i64arr(vals):
    return (ctypes.c_int64 * len(vals))(*[int(v) for v in vals])


def _workspace(nbytes: int, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def _out_dtype(x: torch.Tensor, flags: int):
    return torch.float32 if (flags & FA_FLAG_OUT_F32) else x.dtype


def _check_out(name: str, t: torch.Tensor, shape, dtype, like: torch.Tensor):
    """The in-place entry points write through raw pointers: a caller-allocated output of the wrong eltype, shape,
    layout or device would be a silent out-of-bounds write (the Julia signatures are typed; mirror that)."""
    if not isinstance(t, torch.Tensor):
        raise FaError(f"{name} must be a tensor")
    if t.dtype != dtype:
        raise FaError(f"{name} must have eltype {dtype} (got {t.dtype}); l and m are always float32 (include/fa_sm100a.h)")
    if tuple(int(s) for s in t.shape) != tuple(int(s) for s in shape):
        raise FaError(f"{name} must have shape {tuple(shape)} (got {tuple(t.shape)})")
    if t.device != like.device:
        raise FaError(f"{name} must live on {like.device} (got {t.device})")
    if not is_jl_contiguous(t):
        raise FaError(f"{name} must be dense column-major (see jl_empty)")


def cast(x: torch.Tensor, dtype) -> torch.Tensor:
    """``T.(x)`` on the device through ``fa_cast`` (round to nearest even), same Julia shape and layout."""
    x = jl_array(x)
    if x.dtype == dtype:
        return x
    out = jl_empty(x.shape, dtype, x.device)
    with torch.cuda.device(x.device):
        _check(lib.fa_cast(_ptr(x), _ptr(out), x.numel(), _dt(x), _DTYPES[dtype], _stream(x)), "fa_cast")
    return out


def _via(via, flags, *ts):
    """Float32 arrays on the tensor cores (opt-in, ``via=torch.bfloat16`` / ``torch.float16``): inputs are cast on the
    device, the tcgen05 kernels run in that type with float32 outputs (FA_FLAG_OUT_F32), i.e. results of the 16-bit
    compute class (2e-3) in Float32 arrays -- what a caller of bench/compare.jl:8-10 holding Float32 arrays can ask for."""
    if via is None or ts[0].dtype != torch.float32:
        return flags, ts
    if via not in (torch.bfloat16, torch.float16):
        raise FaError("via must be torch.bfloat16 or torch.float16")
    return flags | FA_FLAG_OUT_F32, tuple(cast(t, via) for t in ts)


def _cur_dev() -> int:
    # host entry points: the library itself reports FA_ERR_CUDA when there is no device
    return torch.cuda.current_device() if torch.cuda.is_available() else 0


# --------------------------------------------------------------------------------------------
# dense
# --------------------------------------------------------------------------------------------
def dense_fa_(O, l, m, Q, K, V, flags: int = 0):
    """``dense_fa!(O, l, m, Q, K, V) -> (O, l, m)`` (src/dense.jl:21-102).  All ``(N, *, B)``."""
    _same(Q, K, V)
    N, d, B = _flatten3(Q)
    dv = int(V.shape[-2])
    if tuple(K.shape) != tuple(Q.shape) or V.shape[0] != Q.shape[0] or V.shape[-1] != Q.shape[-1]:
        raise FaError("dense_fa!: Q, K must have equal shapes and V the same N and batch (src/dense.jl:29,72)")
    Q, K, V = (jl_array(t) for t in (Q, K, V))
    _check_out("O", O, (N, dv, B), _out_dtype(Q, flags), Q)
    _check_out("l", l, (N, 1, B), torch.float32, Q)
    _check_out("m", m, (N, 1, B), torch.float32, Q)
    if Q.is_cuda:
        with torch.cuda.device(Q.device):
            _check(lib.fa_dense_fwd(_ptr(Q), _ptr(K), _ptr(V), _ptr(O), _ptr(l), _ptr(m),
                                    N, d, dv, B, _dt(Q), flags, _stream(Q)), "fa_dense_fwd")
    else:
        _check(lib.fa_dense_fwd_host(_ptr(Q), _ptr(K), _ptr(V), _ptr(O), _ptr(l), _ptr(m),
                                     N, d, dv, B, _dt(Q), flags, _cur_dev()), "fa_dense_fwd_host")
    return O, l, m


def dense_fa(q, k, v, flags: int = 0, via=None):
    """``dense_fa(q, k, v) -> (y, l, m)`` (src/dense.jl:1-19): flattens the spatial dims,
    allocates, calls :func:`dense_fa_`, reshapes back.  ``l, m :: (N, 1, B)`` in float32.
    ``via=torch.bfloat16``: Float32 CUDA arrays on the tensor cores (see :func:`_via`)."""
    _same(q, k, v)
    flags, (q, k, v) = _via(via, flags, q, k, v)
    q, k, v = (jl_array(t) for t in (q, k, v))
    N, d, B = _flatten3(q)
    dv = int(v.shape[-2])
    Q = q if q.ndim == 3 else _jl_reshape(q, (N, d, B))
    K = k if k.ndim == 3 else _jl_reshape(k, (N, d, B))
    V = v if v.ndim == 3 else _jl_reshape(v, (N, dv, B))
    O = jl_empty((N, dv, B), _out_dtype(q, flags), q.device)
    l = jl_empty((N, 1, B), torch.float32, q.device)
    m = jl_empty((N, 1, B), torch.float32, q.device)
    dense_fa_(O, l, m, Q, K, V, flags)
    y = _jl_reshape(O, tuple(q.shape[:-2]) + (dv, B))
    return y, l, m


def _jl_reshape(x: torch.Tensor, shape) -> torch.Tensor:
    """Julia ``reshape`` of a column-major tensor (no copy)."""
    shape = tuple(int(s) for s in shape)
    base = x.permute(*range(x.ndim - 1, -1, -1))          # C-contiguous view
    return base.reshape(shape[::-1]).permute(*range(len(shape) - 1, -1, -1))


def dense_fa_backward(Q, K, V, O, dO, l, m, flags: int = 0):
    """``dense_fa_backward(Q,K,V,O,dO,l,m) -> (dQ, dK, dV)`` -- the working statement of the
    reference's broken function (src/dense.jl:104-167), i.e. ``OneDFastBack``
    (src_cpp/FlashAttention.cpp:194-252): P is recomputed from the saved ``(l, m)``."""
    _same(Q, K, V, O, dO)
    Q, K, V, O, dO = (jl_array(t) for t in (Q, K, V, O, dO))
    N, d, B = _flatten3(Q)
    dv = int(V.shape[-2])
    l, m = (jl_array(t, torch.float32) for t in (l, m))
    dQ, dK, dV = (jl_empty(t.shape, _out_dtype(t, flags), t.device) for t in (Q, K, V))
    if not Q.is_cuda:                                        # the reference's Array arguments (src/dense.jl:104-111)
        _check(lib.fa_dense_bwd_host(_ptr(Q), _ptr(K), _ptr(V), _ptr(O), _ptr(dO), _ptr(l), _ptr(m), _ptr(dQ), _ptr(dK), _ptr(dV),
                                     N, d, dv, B, _dt(Q), flags, _cur_dev()), "fa_dense_bwd_host")
        return dQ, dK, dV
    ws = _workspace(lib.fa_workspace_bytes_dense_bwd(N, d, dv, B, _dt(Q), flags), Q.device)
    with torch.cuda.device(Q.device):
        _check(lib.fa_dense_bwd(_ptr(Q), _ptr(K), _ptr(V), _ptr(O), _ptr(dO), _ptr(l), _ptr(m),
                                _ptr(dQ), _ptr(dK), _ptr(dV), N, d, dv, B, _dt(Q), flags,
                                _ptr(ws), ws.numel(), _stream(Q)), "fa_dense_bwd")
    return dQ, dK, dV


# --------------------------------------------------------------------------------------------
# multi-GPU
# --------------------------------------------------------------------------------------------
def shard_batch(B: int, world: int, rank: int) -> Tuple[int, int]:
    """``(begin, count)`` of the contiguous batch*head range rank ``rank`` of ``world`` owns.  The
    trailing dim is the slowest in memory (src/dense.jl:6-8), so a shard is a pointer range and the
    path needs no collective."""
    b, c = ctypes.c_int64(0), ctypes.c_int64(0)
    _check(lib.fa_shard_batch(int(B), int(world), int(rank), ctypes.byref(b), ctypes.byref(c)), "fa_shard_batch")
    return int(b.value), int(c.value)


def _nccl_comm_ptr(group, device):
    import torch.distributed as dist
    pg = group if group is not None else dist.distributed_c10d._get_default_group()
    backend = pg._get_backend(torch.device(device))
    if not hasattr(backend, "_comm_ptr"):
        raise FaError("ring_dense_fa needs an NCCL process group (ProcessGroupNCCL._comm_ptr)")
    return int(backend._comm_ptr())


def ring_dense_fa(q, k, v, group=None, flags: int = 0):
    """Ring attention forward for one long sequence sharded by TOKENS over the ranks of ``group``:
    every rank passes its ``(Nl, d, B)`` shard of q, k, v and gets its ``(Nl, dv, B)`` slice of
    ``dense_fa`` over the full sequence, plus ``l, m``.  K/V blocks travel round the ring over NCCL
    (``fa_ring_dense_fwd``).  Without an initialised process group it degenerates to ``dense_fa``.
    Like ``dense_fa`` (src/dense.jl:1-19) it also takes ``(spatial.., d, B)`` shards -- e.g. one 3-D volume cut along
    its slowest spatial dim, whose planes are contiguous token ranges -- and returns ``y`` in the shard's shape."""
    import torch.distributed as dist
    _same(q, k, v)
    q, k, v = (jl_array(t) for t in (q, k, v))
    if q.ndim < 3 or not q.is_cuda:
        raise FaError("ring_dense_fa: q, k, v must be CUDA tensors of shape (N_local, d, B) or (spatial.., d, B)")
    if q.ndim > 3:
        spatial = tuple(int(s) for s in q.shape[:-2])
        q3, k3, v3 = (_jl_reshape(t, (-1,) + tuple(t.shape[-2:])) for t in (q, k, v))
        O, l, m = ring_dense_fa(q3, k3, v3, group=group, flags=flags)
        return _jl_reshape(O, spatial + tuple(O.shape[-2:])), l, m
    Nl, d, B = (int(s) for s in q.shape)
    dv = int(v.shape[1])
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    comm = None
    if world > 1:
        dist.barrier(group)                                  # makes sure the NCCL communicator exists
        comm = ctypes.c_void_p(_nccl_comm_ptr(group, q.device))
    O = jl_empty((Nl, dv, B), q.dtype, q.device)
    l = jl_empty((Nl, 1, B), torch.float32, q.device)
    m = jl_empty((Nl, 1, B), torch.float32, q.device)
    ws = _workspace(lib.fa_workspace_bytes_ring_dense_fwd(Nl, d, dv, B, _dt(q)), q.device)
    with torch.cuda.device(q.device):
        _check(lib.fa_ring_dense_fwd(_ptr(q), _ptr(k), _ptr(v), _ptr(O), _ptr(l), _ptr(m), Nl, d, dv, B, _dt(q), flags,
                                     comm, rank, world, _ptr(ws), ws.numel(), _stream(q)), "fa_ring_dense_fwd")
        torch.cuda.current_stream(q.device).synchronize()    # the workspace dies with this frame
    return O, l, m


def ring_dense_fa_backward(q, k, v, O, dO, l, m, group=None, flags: int = 0):
    """Backward of :func:`ring_dense_fa`: every rank passes its token shards and the ``(O, l, m)`` the ring
    forward returned and gets ``(dq, dk, dv)`` of its shard; dK/dV accumulators travel with the K/V blocks
    (``fa_ring_dense_bwd``)."""
    import torch.distributed as dist
    _same(q, k, v, O, dO)
    q, k, v, O, dO = (jl_array(t) for t in (q, k, v, O, dO))
    l, m = (jl_array(t, torch.float32) for t in (l, m))
    if q.ndim > 3:                                           # (spatial.., d, B) shards: flatten like dense_fa
        shapes = [tuple(t.shape) for t in (q, k, v)]
        flat = lambda t: _jl_reshape(t, (-1,) + tuple(t.shape[-2:]))
        grads = ring_dense_fa_backward(flat(q), flat(k), flat(v), flat(O), flat(dO), l, m, group=group, flags=flags)
        return tuple(_jl_reshape(g_, sh) for g_, sh in zip(grads, shapes))
    Nl, d, B = (int(s) for s in q.shape)
    dv = int(v.shape[1])
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    comm = None
    if world > 1:
        dist.barrier(group)
        comm = ctypes.c_void_p(_nccl_comm_ptr(group, q.device))
    dq, dk, dvv = (jl_empty(t.shape, t.dtype, t.device) for t in (q, k, v))
    ws = _workspace(lib.fa_workspace_bytes_ring_dense_bwd(Nl, d, dv, B, _dt(q), flags), q.device)
    with torch.cuda.device(q.device):
        _check(lib.fa_ring_dense_bwd(_ptr(q), _ptr(k), _ptr(v), _ptr(O), _ptr(dO), _ptr(l), _ptr(m),
                                     _ptr(dq), _ptr(dk), _ptr(dvv), Nl, d, dv, B, _dt(q), flags,
                                     comm, rank, world, _ptr(ws), ws.numel(), _stream(q)), "fa_ring_dense_bwd")
        torch.cuda.current_stream(q.device).synchronize()
    return dq, dk, dvv


# --------------------------------------------------------------------------------------------
# windowed
# --------------------------------------------------------------------------------------------
def _win_kws(W, stride, pad):
    return (int(W) if stride is None else int(stride)), ((int(W) - 1) // 2 if pad is None else int(pad))


def window_counts(spatial, W, stride=None, pad=None):
    stride, pad = _win_kws(W, stride, pad)
    nw = _i64arr([0] * len(spatial))
    _check(lib.fa_window_index(len(spatial), _i64arr(spatial), W, stride, pad, nw, None), "fa_window_index")
    return tuple(int(x) for x in nw)


def windowed_fa(q, k, v, windowsize: int, stride: Optional[int] = None, pad: Optional[int] = None, flags: int = 0, via=None):
    """``windowed_fa(q, k, v, W; stride=W, pad=(W-1)/2) -> (y, l, m)`` (src/windowed.jl:3-23):
    window-partition attention with zero padding and fold-averaging, the unfold/fold of
    src/utils.jl:36-54 fused into the kernel.  ``l, m :: (W^D, 1, L, B)`` float32."""
    _same(q, k, v)
    flags, (q, k, v) = _via(via, flags, q, k, v)
    q, k, v = (jl_array(t) for t in (q, k, v))
    W = int(windowsize)
    stride, pad = _win_kws(W, stride, pad)
    spatial = tuple(int(s) for s in q.shape[:-2])
    d, dv, B = int(q.shape[-2]), int(v.shape[-2]), int(q.shape[-1])
    if tuple(k.shape) != tuple(q.shape) or tuple(v.shape[:-2]) != spatial or v.shape[-1] != B:
        raise FaError("windowed_fa: q, k must have equal shapes; v the same spatial/batch dims")
    dims = _i64arr(spatial)
    nw = window_counts(spatial, W, stride, pad)
    L = 1
    for n in nw:
        L *= n
    WD = W ** len(spatial)
    y = jl_empty(spatial + (dv, B), _out_dtype(q, flags), q.device)
    l = jl_empty((WD, 1, L, B), torch.float32, q.device)
    m = jl_empty((WD, 1, L, B), torch.float32, q.device)
    if q.is_cuda:
        ws = _workspace(lib.fa_workspace_bytes_windowed_fwd(len(spatial), dims, d, dv, B, W, stride, pad, _dt(q), flags), q.device)
        with torch.cuda.device(q.device):
            _check(lib.fa_windowed_fwd(_ptr(q), _ptr(k), _ptr(v), _ptr(y), _ptr(l), _ptr(m), len(spatial), dims,
                                       d, dv, B, W, stride, pad, _dt(q), flags, _ptr(ws), ws.numel(), _stream(q)),
                   "fa_windowed_fwd")
    else:
        _check(lib.fa_windowed_fwd_host(_ptr(q), _ptr(k), _ptr(v), _ptr(y), _ptr(l), _ptr(m), len(spatial), dims,
                                        d, dv, B, W, stride, pad, _dt(q), flags, _cur_dev()),
               "fa_windowed_fwd_host")
    return y, l, m


def block_fa(q, k, v, windowsize: int, pad: int = 0, flags: int = 0):
    """``block_fa(q,k,v,W; pad=0) = windowed_fa(...; stride=W, pad=pad)`` (src/windowed.jl:1)."""
    return windowed_fa(q, k, v, windowsize, stride=windowsize, pad=pad, flags=flags)


def windowed_fa_backward(q, k, v, dy, l, m, windowsize: int, stride=None, pad=None, flags: int = 0):
    """Backward of :func:`windowed_fa` (the reference has none):
    ``dYw = window(dY ./ count)``, per-window flash backward, ``dq = unwindow(dQw)`` etc."""
    _same(q, k, v, dy)
    q, k, v, dy = (jl_array(t) for t in (q, k, v, dy))
    l, m = (jl_array(t, torch.float32) for t in (l, m))
    W = int(windowsize)
    stride, pad = _win_kws(W, stride, pad)
    spatial = tuple(int(s) for s in q.shape[:-2])
    d, dv, B = int(q.shape[-2]), int(v.shape[-2]), int(q.shape[-1])
    dims = _i64arr(spatial)
    dq, dk, dvv = (jl_empty(t.shape, _out_dtype(t, flags), t.device) for t in (q, k, v))
    if not q.is_cuda:
        _check(lib.fa_windowed_bwd_host(_ptr(q), _ptr(k), _ptr(v), _ptr(dy), _ptr(l), _ptr(m), _ptr(dq), _ptr(dk), _ptr(dvv),
                                        len(spatial), dims, d, dv, B, W, stride, pad, _dt(q), flags, _cur_dev()), "fa_windowed_bwd_host")
        return dq, dk, dvv
    ws = _workspace(lib.fa_workspace_bytes_windowed_bwd(len(spatial), dims, d, dv, B, W, stride, pad, _dt(q), flags), q.device)
    with torch.cuda.device(q.device):
        _check(lib.fa_windowed_bwd(_ptr(q), _ptr(k), _ptr(v), _ptr(dy), _ptr(l), _ptr(m), _ptr(dq), _ptr(dk), _ptr(dvv),
                                   len(spatial), dims, d, dv, B, W, stride, pad, _dt(q), flags,
                                   _ptr(ws), ws.numel(), _stream(q)), "fa_windowed_bwd")
    return dq, dk, dvv


class SlabPlan(tuple):
    """(plane_lo, plane_hi, win_lo, win_hi, pad_lo) of one rank's slab (see :func:`windowed_slab_plan`)."""
    plane_lo = property(lambda s: s[0]); plane_hi = property(lambda s: s[1])
    win_lo = property(lambda s: s[2]); win_hi = property(lambda s: s[3]); pad_lo = property(lambda s: s[4])
    nwin = property(lambda s: s[3] - s[2])


def windowed_slab_plan(spatial, windowsize: int, stride=None, pad=None, rank: int = 0, nranks: int = 1) -> SlabPlan:
    """One volume over several GPUs: non-overlapping windows are independent, so the volume is cut
    into slabs along its slowest spatial dim on window boundaries and needs no exchange.  Rank ``rank`` takes the
    windows ``[win_lo, win_hi)`` of that dim and holds the token planes ``[plane_lo, plane_hi)``."""
    W = int(windowsize)
    stride, pad = _win_kws(W, stride, pad)
    plan = _i64arr([0] * 5)
    _check(lib.fa_windowed_slab_plan(len(spatial), _i64arr(spatial), W, stride, pad, int(rank), int(nranks), plan),
           "fa_windowed_slab_plan")
    return SlabPlan(int(x) for x in plan)


def slab_planes(x, plan: SlabPlan):
    """The planes ``[plane_lo, plane_hi)`` of the slowest spatial dim of ``x :: (s.., d, B)``, column-major."""
    ax = x.ndim - 3
    return jl_array(x.narrow(ax, plan.plane_lo, plan.plane_hi - plan.plane_lo))


def windowed_fa_slab(q, k, v, windowsize: int, plan: SlabPlan, stride=None, pad=None, flags: int = 0):
    """:func:`windowed_fa` on one slab ``q, k, v :: (s.., planes, d, B)`` of a volume split by
    :func:`windowed_slab_plan`: ``y`` for the slab's planes and ``l, m :: (W^D, 1, L_slab, B)`` for its windows
    (a contiguous range of the volume's window index), equal bit for bit to the same entries of the one-GPU call."""
    _same(q, k, v)
    q, k, v = (jl_array(t) for t in (q, k, v))
    W = int(windowsize)
    stride, pad = _win_kws(W, stride, pad)
    spatial = tuple(int(s) for s in q.shape[:-2])
    d, dv, B = int(q.shape[-2]), int(v.shape[-2]), int(q.shape[-1])
    if spatial[-1] != plan.plane_hi - plan.plane_lo:
        raise FaError("windowed_fa_slab: the slab must hold exactly the planes [plane_lo, plane_hi) of its plan")
    nw = window_counts(spatial[:-1], W, stride, pad) + (plan.nwin,) if len(spatial) > 1 else (plan.nwin,)
    L = 1
    for n in nw:
        L *= n
    WD = W ** len(spatial)
    y = jl_empty(spatial + (dv, B), q.dtype, q.device)
    l = jl_empty((WD, 1, L, B), torch.float32, q.device)
    m = jl_empty((WD, 1, L, B), torch.float32, q.device)
    with torch.cuda.device(q.device):
        _check(lib.fa_windowed_slab_fwd(_ptr(q), _ptr(k), _ptr(v), _ptr(y), _ptr(l), _ptr(m), len(spatial), _i64arr(spatial),
                                        d, dv, B, W, stride, pad, plan.pad_lo, plan.nwin, _dt(q), flags, _stream(q)),
               "fa_windowed_slab_fwd")
    return y, l, m


def windowed_fa_slab_backward(q, k, v, dy, l, m, windowsize: int, plan: SlabPlan, stride=None, pad=None, flags: int = 0):
    """Backward of :func:`windowed_fa_slab`: gradients of the slab's own planes (no exchange needed)."""
    _same(q, k, v, dy)
    q, k, v, dy = (jl_array(t) for t in (q, k, v, dy))
    l, m = (jl_array(t, torch.float32) for t in (l, m))
    W = int(windowsize)
    stride, pad = _win_kws(W, stride, pad)
    spatial = tuple(int(s) for s in q.shape[:-2])
    d, dv, B = int(q.shape[-2]), int(v.shape[-2]), int(q.shape[-1])
    dims = _i64arr(spatial)
    dq, dk, dvv = (jl_empty(t.shape, t.dtype, t.device) for t in (q, k, v))
    ws = _workspace(lib.fa_workspace_bytes_windowed_slab_bwd(len(spatial), dims, d, dv, B, W, stride, pad, plan.pad_lo, plan.nwin), q.device)
    with torch.cuda.device(q.device):
        _check(lib.fa_windowed_slab_bwd(_ptr(q), _ptr(k), _ptr(v), _ptr(dy), _ptr(l), _ptr(m), _ptr(dq), _ptr(dk), _ptr(dvv),
                                        len(spatial), dims, d, dv, B, W, stride, pad, plan.pad_lo, plan.nwin, _dt(q), flags,
                                        _ptr(ws), ws.numel(), _stream(q)), "fa_windowed_slab_bwd")
    return dq, dk, dvv


def window(x, windowsize: int, stride=None, pad=None):
    """``window(x, W; stride, pad)`` (src/utils.jl:36-44): ``(s.., d, B) -> (W^D, d, L, B)``."""
    x = jl_array(x)
    W = int(windowsize)
    stride, pad = _win_kws(W, stride, pad)
    spatial = tuple(int(s) for s in x.shape[:-2])
    d, B = int(x.shape[-2]), int(x.shape[-1])
    nw = window_counts(spatial, W, stride, pad)
    L = 1
    for n in nw:
        L *= n
    out = jl_empty((W ** len(spatial), d, L, B), x.dtype, x.device)
    with torch.cuda.device(x.device):
        _check(lib.fa_window(_ptr(x), _ptr(out), len(spatial), _i64arr(spatial), d, B, W, stride, pad, _dt(x), _stream(x)), "fa_window")
    return out


def unwindow(X, outputsize, windowsize: int, stride=None, pad=None):
    """``unwindow(X, outputsize, W; stride, pad)`` (src/utils.jl:46-54): fold (scatter-add)."""
    X = jl_array(X)
    W = int(windowsize)
    stride, pad = _win_kws(W, stride, pad)
    outputsize = tuple(int(s) for s in outputsize)
    spatial, d, B = outputsize[:-2], outputsize[-2], outputsize[-1]
    out = jl_empty(outputsize, X.dtype, X.device)
    with torch.cuda.device(X.device):
        _check(lib.fa_unwindow(_ptr(X), _ptr(out), len(spatial), _i64arr(spatial), d, B, W, stride, pad, _dt(X), _stream(X)), "fa_unwindow")
    return out


def window_index(spatial, windowsize: int, stride=None, pad=None):
    """Index set of ``window`` as an int64 ``(W^D, L)`` column-major tensor (-1 = zero padding)."""
    W = int(windowsize)
    stride, pad = _win_kws(W, stride, pad)
    nw = window_counts(spatial, W, stride, pad)
    L = 1
    for n in nw:
        L *= n
    WD = W ** len(spatial)
    buf = (ctypes.c_int64 * (WD * L))()
    nwb = _i64arr([0] * len(spatial))
    _check(lib.fa_window_index(len(spatial), _i64arr(spatial), W, stride, pad, nwb, buf), "fa_window_index")
    return torch.tensor(list(buf), dtype=torch.int64).reshape(L, WD).permute(1, 0)


def window_count(spatial, windowsize: int, stride=None, pad=None):
    """Per-position window count (the ``divisor`` of src/windowed.jl:16-17), Julia shape ``spatial``."""
    W = int(windowsize)
    stride, pad = _win_kws(W, stride, pad)
    n = 1
    for s in spatial:
        n *= int(s)
    buf = (ctypes.c_int64 * n)()
    _check(lib.fa_window_count(len(spatial), _i64arr(spatial), W, stride, pad, buf), "fa_window_count")
    return _jl_reshape(torch.tensor(list(buf), dtype=torch.int64), tuple(spatial))


# --------------------------------------------------------------------------------------------
# circulant
# --------------------------------------------------------------------------------------------
def circulant_fa_(O, l, m, Q, K, V, W: int, flags: int = 0):
    """``circulant_fa!(O, l, m, Q, K, V, W) -> (O, l, m)`` (src/circulant.jl:9-118), 1-D periodic band."""
    _same(Q, K, V)
    Q, K, V = (jl_array(t) for t in (Q, K, V))
    if Q.ndim == 4:                       # 2-D periodic neighbourhood (the reference's todo, README.md:38-41,53)
        X, Y, d, B = (int(s) for s in Q.shape)
        _check_out("O", O, (X, Y, int(V.shape[2]), B), Q.dtype, Q)
        _check_out("l", l, (X * Y, 1, B), torch.float32, Q)
        _check_out("m", m, (X * Y, 1, B), torch.float32, Q)
        with torch.cuda.device(Q.device):
            _check(lib.fa_circulant2d_fwd(_ptr(Q), _ptr(K), _ptr(V), _ptr(O), _ptr(l), _ptr(m), X, Y, d, int(V.shape[2]), B, int(W),
                                          _dt(Q), flags, _stream(Q)), "fa_circulant2d_fwd")
        return O, l, m
    if Q.ndim != 3:
        raise FaError("circulant_fa!: Q, K, V must be (N, d, B) or (X, Y, d, B)")
    N, d, B = (int(s) for s in Q.shape)
    dv = int(V.shape[1])
    _check_out("O", O, (N, dv, B), _out_dtype(Q, flags), Q)
    _check_out("l", l, (N, 1, B), torch.float32, Q)
    _check_out("m", m, (N, 1, B), torch.float32, Q)
    if Q.is_cuda:
        with torch.cuda.device(Q.device):
            _check(lib.fa_circulant_fwd(_ptr(Q), _ptr(K), _ptr(V), _ptr(O), _ptr(l), _ptr(m), N, d, dv, B, int(W),
                                        _dt(Q), flags, _stream(Q)), "fa_circulant_fwd")
    else:
        _check(lib.fa_circulant_fwd_host(_ptr(Q), _ptr(K), _ptr(V), _ptr(O), _ptr(l), _ptr(m), N, d, dv, B, int(W),
                                         _dt(Q), flags, _cur_dev()), "fa_circulant_fwd_host")
    return O, l, m


def circulant_fa(Q, K, V, W: int, flags: int = 0, via=None):
    """``circulant_fa(Q, K, V, W)``: allocating wrapper.  The reference's drops ``W``
    (src/circulant.jl:6, SURVEY B-1); this one passes it."""
    flags, (Q, K, V) = _via(via, flags, Q, K, V)
    if Q.ndim == 4:
        X, Y, d, B = (int(s) for s in Q.shape)
        O = jl_empty((X, Y, int(V.shape[2]), B), Q.dtype, Q.device)
        l = jl_empty((X * Y, 1, B), torch.float32, Q.device)
        m = jl_empty((X * Y, 1, B), torch.float32, Q.device)
        return circulant_fa_(O, l, m, Q, K, V, W, flags)
    N, d, B = (int(s) for s in Q.shape)
    O = jl_empty((N, int(V.shape[1]), B), _out_dtype(Q, flags), Q.device)
    l = jl_empty((N, 1, B), torch.float32, Q.device)
    m = jl_empty((N, 1, B), torch.float32, Q.device)
    return circulant_fa_(O, l, m, Q, K, V, W, flags)


def circulant_fa_backward(Q, K, V, O, dO, l, m, W: int, flags: int = 0):
    """Backward of circulant attention: A.5.1 restricted to the periodic band."""
    _same(Q, K, V, O, dO)
    Q, K, V, O, dO = (jl_array(t) for t in (Q, K, V, O, dO))
    l, m = (jl_array(t, torch.float32) for t in (l, m))
    if Q.ndim == 4:
        X, Y, d, B = (int(s) for s in Q.shape)
        dQ, dK, dV = (jl_empty(t.shape, t.dtype, t.device) for t in (Q, K, V))
        ws = _workspace(lib.fa_workspace_bytes_circulant2d_bwd_ex(X, Y, d, int(V.shape[2]), B, int(W), _dt(Q), flags), Q.device)
        with torch.cuda.device(Q.device):
            _check(lib.fa_circulant2d_bwd(_ptr(Q), _ptr(K), _ptr(V), _ptr(O), _ptr(dO), _ptr(l), _ptr(m), _ptr(dQ), _ptr(dK), _ptr(dV),
                                          X, Y, d, int(V.shape[2]), B, int(W), _dt(Q), flags, _ptr(ws), ws.numel(), _stream(Q)),
                   "fa_circulant2d_bwd")
        return dQ, dK, dV
    N, d, B = (int(s) for s in Q.shape)
    dv = int(V.shape[1])
    dQ, dK, dV = (jl_empty(t.shape, _out_dtype(t, flags), t.device) for t in (Q, K, V))
    if not Q.is_cuda:
        _check(lib.fa_circulant_bwd_host(_ptr(Q), _ptr(K), _ptr(V), _ptr(O), _ptr(dO), _ptr(l), _ptr(m), _ptr(dQ), _ptr(dK), _ptr(dV),
                                         N, d, dv, B, int(W), _dt(Q), flags, _cur_dev()), "fa_circulant_bwd_host")
        return dQ, dK, dV
    ws = _workspace(lib.fa_workspace_bytes_circulant_bwd(N, d, dv, B, int(W), _dt(Q), flags), Q.device)
    with torch.cuda.device(Q.device):
        _check(lib.fa_circulant_bwd(_ptr(Q), _ptr(K), _ptr(V), _ptr(O), _ptr(dO), _ptr(l), _ptr(m),
                                    _ptr(dQ), _ptr(dK), _ptr(dV), N, d, dv, B, int(W), _dt(Q), flags,
                                    _ptr(ws), ws.numel(), _stream(Q)), "fa_circulant_bwd")
    return dQ, dK, dV


def cartesian_circulant(n: int, N: int, M: int):
    """``cartesian_circulant(n, N, M) -> (i, j)`` (src/utils.jl:6-17), 1-based like the reference."""
    keys = circulant_keys(N, M)
    j = -(-n // M)
    return int(keys[(n - 1) % M, j - 1]) + 1, j


def circulant2d_keys(X: int, Y: int, W: int) -> torch.Tensor:
    """Key set of the 2-D circulant attention as a 0-based int64 ``(W*W, X*Y)`` tensor (column = query
    ``y*X + x``, row = ``t*W + s``)."""
    buf = (ctypes.c_int64 * (X * Y * W * W))()
    _check(lib.fa_circulant2d_index(X, Y, W, buf), "fa_circulant2d_index")
    return torch.tensor(list(buf), dtype=torch.int64).reshape(X * Y, W * W).permute(1, 0)


def circulant_keys(N: int, W: int) -> torch.Tensor:
    """All ``first(cartesian_circulant((j-1)W+w, N, W))`` as a 0-based int64 ``(W, N)`` tensor."""
    buf = (ctypes.c_int64 * (N * W))()
    _check(lib.fa_circulant_index(N, W, buf), "fa_circulant_index")
    return torch.tensor(list(buf), dtype=torch.int64).reshape(N, W).permute(1, 0)


def circulant(V_or_N, M: Optional[int] = None, dtype=torch.float64):
    """``circulant(N, M)`` / ``circulant(V)`` (src/utils.jl:19-31): the ``N x N`` banded periodic matrix in
    CSC form whose column ``j`` holds ``V[:, j]`` (or ones) at the rows ``first(cartesian_circulant(., N, M))``;
    returned as a ``torch.sparse_csc_tensor`` (row indices in the reference's storage order)."""
    if isinstance(V_or_N, int):
        N, vals = int(V_or_N), None
    else:
        V = V_or_N
        M, N = int(V.shape[0]), int(V.shape[1])
        vals = V.t().reshape(-1).to("cpu")                                 # column-major reshape(V, :)
    rows = circulant_keys(N, int(M)).t().reshape(-1)                       # (W, N) column-major -> nz order
    colptr = torch.arange(0, N + 1, dtype=torch.int64) * int(M)
    if vals is None:
        vals = torch.ones(N * int(M), dtype=dtype)
    return torch.sparse_csc_tensor(colptr, rows, vals, size=(N, N))


def batch_circulant(bV: torch.Tensor):
    """``batch_circulant(bV)`` (src/utils.jl:33): block-diagonal of ``circulant(bV[:, :, b])``, as a
    ``(B*N, B*N)`` sparse COO tensor."""
    M, N, B = (int(s) for s in bV.shape)
    idx, val = [], []
    for b in range(B):
        c = circulant(bV[:, :, b]).to_sparse_coo().coalesce()
        idx.append(c.indices() + b * N)
        val.append(c.values())
    return torch.sparse_coo_tensor(torch.cat(idx, 1), torch.cat(val), (B * N, B * N)).coalesce()


# --------------------------------------------------------------------------------------------
# softmax
# --------------------------------------------------------------------------------------------
def fused_softmax_(P, S, dims: int = 1):
    """``fused_softmax!(P, S; dims)`` (src/fused_softmax.jl:4-16), 2-D or 3-D, ``dims in (1, 2)``."""
    assert dims in (1, 2), "only softmax in dims 1 or 2 supported"      # src/fused_softmax.jl:12
    S = jl_array(S)
    if S.ndim not in (2, 3):
        raise FaError("fused_softmax!: 2-D or 3-D arrays only")
    M, N = int(S.shape[0]), int(S.shape[1])
    B = int(S.shape[2]) if S.ndim == 3 else 1
    with torch.cuda.device(S.device):
        _check(lib.fa_softmax(_ptr(P), _ptr(S), M, N, B, dims, _dt(S), _stream(S)), "fa_softmax")
    return P


def fused_softmax(S, dims: int = 1):
    """``fused_softmax(S; dims=1)`` (src/fused_softmax.jl:1)."""
    S = jl_array(S)
    return fused_softmax_(jl_empty(S.shape, S.dtype, S.device), S, dims)


# --------------------------------------------------------------------------------------------
# naive oracles kept naive (reference src/naive/*.jl): they materialise P with library matmuls.
# They are NOT the hot path and exist for the same reason as in the reference: a definition to
# compare against.
# --------------------------------------------------------------------------------------------
def dense_dpa(q, k, v):
    """``dense_dpa(q,k,v) -> (y, P)`` (src/naive/dense.jl:8-35): gemm, softmax(dims=2), gemm."""
    q, k, v = (jl_array(t) for t in (q, k, v))
    N, d, B = _flatten3(q)
    dv = int(v.shape[-2])
    Q, K, V = (_jl_reshape(t, (N, t.shape[-2], B)).permute(2, 0, 1).float() for t in (q, k, v))   # (B, N, c)
    P = torch.softmax(torch.bmm(Q, K.transpose(1, 2)) / math.sqrt(d), dim=2)
    O = torch.bmm(P, V).to(q.dtype)
    y = jl_array(O.permute(1, 2, 0))
    return _jl_reshape(y, tuple(q.shape[:-2]) + (dv, B)), jl_array(P.permute(1, 2, 0).to(q.dtype))


def windowed_dpa(q, k, v, windowsize: int, stride=None, pad=None):
    """``windowed_dpa -> (y, P)`` (src/naive/windowed.jl:3-23): window, dense_dpa, unwindow ./ count."""
    qw, kw, vw = (window(t, windowsize, stride, pad) for t in (q, k, v))
    WD, d, L, B = (int(s) for s in qw.shape)
    dv = int(vw.shape[1])
    yw, Pw = dense_dpa(_jl_reshape(qw, (WD, d, L * B)), _jl_reshape(kw, (WD, d, L * B)), _jl_reshape(vw, (WD, dv, L * B)))
    szy = tuple(q.shape[:-2]) + (dv, int(q.shape[-1]))
    ones = jl_empty(szy, q.dtype, q.device).fill_(1)
    divisor = unwindow(window(ones, windowsize, stride, pad), szy, windowsize, stride, pad)
    y = unwindow(_jl_reshape(yw, (WD, dv, L, B)), szy, windowsize, stride, pad) / divisor
    return jl_array(y), _jl_reshape(Pw, (WD, WD, L, B))


def block_dpa(q, k, v, windowsize: int):
    """``block_dpa`` forwards no kwargs (src/naive/windowed.jl:1) and so inherits ``pad=(W-1)/2``."""
    return windowed_dpa(q, k, v, windowsize)


def circulant_dpa(Q, K, V, W: int):
    """``circulant_dpa!`` (src/naive/circulant.jl:8-36) -> ``(O, P)`` with ``P :: (W, N, B)`` in
    ``cartesian_circulant`` order; gather-based instead of a sparse matrix."""
    Q, K, V = (jl_array(t) for t in (Q, K, V))
    N, d, B = (int(s) for s in Q.shape)
    keys = circulant_keys(N, W).to(Q.device)                               # (W, N)
    Kg = K.float()[keys]                                                   # (W, N, d, B)
    S = torch.einsum("ikb,wikb->wib", Q.float(), Kg) / math.sqrt(d)
    P = torch.softmax(S, dim=0)
    O = torch.einsum("wib,wicb->icb", P, V.float()[keys])
    return jl_array(O.to(Q.dtype)), jl_array(P.to(Q.dtype))
