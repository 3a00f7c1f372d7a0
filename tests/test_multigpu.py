"""Multi-GPU rows of SURVEY 8e.

CPU (gloo, world_size 2):  the batch*head shard arithmetic (fa_shard_batch), and the ring schedule --
K/V token blocks passed rank -> rank+1 while each rank merges partial (O, l, m) with the update rule
of src/dense.jl:82-91 -- run with REAL send/recv between two processes, the oracle standing in for the
block kernel (the CUDA kernels cannot run on the CPU and there is no fallback).
GPU: fa_ring_dense_fwd on one rank (degenerate ring) and, when two GPUs are visible, on two ranks over
NCCL, against the oracle on the whole sequence.
One volume over several ranks (windowed, non-overlapping windows): the slab plan against its brute-force oracle,
slabs computed by two gloo processes assembling to the whole-volume oracle result, and on the GPU the slab
kernels reproducing the one-GPU call bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import fa_oracle as fo
from util import randn_np, rel_err, to_dev, to_np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_batch_partitions_exactly():
    import fa_sm100a as fa
    for B in (1, 7, 8, 64, 511, 512):
        for world in (1, 2, 3, 4, 8):
            spans = [fa.shard_batch(B, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == B
            for (b0, c0), (b1, _) in zip(spans, spans[1:]):
                assert b0 + c0 == b1                       # contiguous, in rank order
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(fa.FaError):
        fa.shard_batch(8, 2, 2)


def test_oracle_ring_equals_dense():
    q, k, v = (randn_np((96, 8, 2), s).astype(np.float64) for s in range(3))
    y0, l0, m0 = fo.dense_fa(q, k, v)
    for G in (1, 2, 3, 4):
        n = 96 // G
        sh = lambda t: [np.asfortranarray(t[r * n:(r + 1) * n]) for r in range(G)]
        parts = fo.ring_dense_fa(sh(q), sh(k), sh(v))
        y = np.concatenate([p[0] for p in parts]); l = np.concatenate([p[1] for p in parts]); m = np.concatenate([p[2] for p in parts])
        assert np.abs(y - y0).max() < 1e-12 and np.abs(m - m0).max() < 1e-12 and np.abs(l / l0 - 1).max() < 1e-12


def _gloo_ring_worker(rank, world, port, q, k, v, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = q.shape[0] // world
    ql, kl, vl = (np.asfortranarray(t[rank * n:(rank + 1) * n]) for t in (q, k, v))
    nxt, prv = (rank + 1) % world, (rank - 1) % world
    cur_k, cur_v, acc = kl, vl, None
    for s in range(world):
        reqs, rk, rv = [], None, None
        if s + 1 < world:                                  # same schedule as fa_ring_dense_fwd: post the exchange, then compute
            rk, rv = torch.empty(cur_k.shape[::-1], dtype=torch.float64), torch.empty(cur_v.shape[::-1], dtype=torch.float64)
            ops = [dist.P2POp(dist.isend, torch.from_numpy(np.ascontiguousarray(cur_k.T)), nxt),
                   dist.P2POp(dist.isend, torch.from_numpy(np.ascontiguousarray(cur_v.T)), nxt),
                   dist.P2POp(dist.irecv, rk, prv), dist.P2POp(dist.irecv, rv, prv)]
            reqs = dist.batch_isend_irecv(ops)
        part = fo.dense_fa(ql, cur_k, cur_v)
        acc = part if acc is None else fo.merge_partials(*acc, *part)
        for r in reqs:
            r.wait()
        if s + 1 < world:
            cur_k, cur_v = np.asfortranarray(rk.numpy().T), np.asfortranarray(rv.numpy().T)
    out[rank] = tuple(torch.from_numpy(np.ascontiguousarray(a)) for a in acc)
    dist.barrier()
    dist.destroy_process_group()


def test_ring_schedule_gloo_world2():
    q, k, v = (randn_np((64, 8, 2), s).astype(np.float64) for s in range(3))
    y0, l0, m0 = fo.dense_fa(q, k, v)
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_gloo_ring_worker, args=(world, port, q, k, v, out), nprocs=world, join=True)
    y = np.concatenate([out[r][0].numpy() for r in range(world)])
    l = np.concatenate([out[r][1].numpy() for r in range(world)])
    m = np.concatenate([out[r][2].numpy() for r in range(world)])
    assert np.abs(y - y0).max() < 1e-12 and np.abs(l / l0 - 1).max() < 1e-12 and np.abs(m - m0).max() < 1e-12


# ------------------------------------------------------------------------------------------- windowed slabs
SLAB_GEOS = [((64, 64, 64), 5, 5, 3), ((64, 64), 7, 7, 3), ((16,), 4, 4, 0), ((20, 9), 3, 4, 1), ((10,), 5, 5, 2),
             ((12, 13), 5, 5, 2), ((9, 31), 4, 6, 3), ((7, 8, 30), 3, 3, 1), ((5, 40), 2, 7, 1)]


def test_windowed_slab_plan_matches_oracle_and_tiles_the_volume():
    import fa_sm100a as fa
    for spatial, W, stride, pad in SLAB_GEOS:
        nw = fo.window_counts(spatial, W, stride, pad)[-1]
        for G in (1, 2, 3, 4, 8):
            plans = [fa.windowed_slab_plan(spatial, W, stride, pad, r, G) for r in range(G)]
            for r, pl in enumerate(plans):
                assert tuple(pl) == tuple(fo.windowed_slab_plan(spatial, W, stride, pad, r, G))
            live = [pl for pl in plans if pl.nwin > 0]
            assert live[0].plane_lo == 0 and live[-1].plane_hi == spatial[-1] and live[0].win_lo == 0 and live[-1].win_hi == nw
            for a, b in zip(live, live[1:]):
                assert a.plane_hi == b.plane_lo and a.win_hi == b.win_lo
            assert max(pl.nwin for pl in plans) - min(pl.nwin for pl in plans) <= 1
    with pytest.raises(fa.FaError):
        fa.windowed_slab_plan((16, 16), 5, 1, 2, 0, 2)      # overlapping windows: needs a halo reduce, not built


def _slab_of(t, pl):
    sl = [slice(None)] * t.ndim
    sl[t.ndim - 3] = slice(pl[0], pl[1])
    return np.asfortranarray(t[tuple(sl)])


def test_oracle_slabs_assemble_to_whole_volume():
    for spatial, W, stride, pad, G in (((12, 13), 5, 5, 2, 3), ((9, 31), 4, 6, 3, 2), ((6, 5, 14), 3, 3, 1, 3), ((10,), 5, 5, 2, 2)):
        q, k, v = (randn_np(spatial + (4, 2), s).astype(np.float64) for s in range(3))
        y0, l0, m0 = fo.windowed_fa(q, k, v, W, stride=stride, pad=pad)
        parts = []
        for r in range(G):
            pl = fo.windowed_slab_plan(spatial, W, stride, pad, r, G)
            if pl[3] > pl[2]:
                parts.append(fo.windowed_fa_slab(*(_slab_of(t, pl) for t in (q, k, v)), W, stride, pad, pl[4], pl[3] - pl[2]))
        y = np.concatenate([p[0] for p in parts], axis=len(spatial) - 1)
        l = np.concatenate([p[1] for p in parts], axis=2)
        m = np.concatenate([p[2] for p in parts], axis=2)
        assert np.array_equal(np.isnan(y), np.isnan(y0)) and np.nanmax(np.abs(y - y0)) == 0
        assert np.array_equal(l, l0) and np.array_equal(m, m0)


def _gloo_slab_worker(rank, world, port, q, k, v, geo, out):
    """Each rank works on its own slab only; nothing but the final gather crosses ranks."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(ROOT, "flashattention.jl_b200"))
    import fa_sm100a as fa
    spatial, W, stride, pad = geo
    pl = fa.windowed_slab_plan(spatial, W, stride, pad, rank, world)
    y, l, m = fo.windowed_fa_slab(*(_slab_of(t, pl) for t in (q, k, v)), W, stride, pad, pl.pad_lo, pl.nwin)
    gathered = [None] * world
    dist.all_gather_object(gathered, (tuple(pl), np.ascontiguousarray(y), np.ascontiguousarray(l), np.ascontiguousarray(m)))
    if rank == 0:
        out["res"] = gathered
    dist.barrier()
    dist.destroy_process_group()


def test_windowed_slabs_gloo_world2():
    geo = ((8, 6, 12), 3, 3, 1)
    spatial, W, stride, pad = geo
    q, k, v = (randn_np(spatial + (4, 1), s).astype(np.float64) for s in range(3))
    y0, l0, m0 = fo.windowed_fa(q, k, v, W, stride=stride, pad=pad)
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_gloo_slab_worker, args=(world, port, q, k, v, geo, out), nprocs=world, join=True)
    res = sorted(out["res"], key=lambda t: t[0][2])
    y = np.concatenate([r[1] for r in res], axis=2)
    l = np.concatenate([r[2] for r in res], axis=2)
    assert np.array_equal(np.isnan(y), np.isnan(y0)) and np.nanmax(np.abs(y - y0)) == 0 and np.array_equal(l, l0)


# ------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("dtype,N,d", [(torch.bfloat16, 512, 128), (torch.float32, 200, 32)])
def test_ring_single_rank_equals_dense(dtype, N, d):
    import fa_sm100a as fa
    q, k, v = (randn_np((N, d, 2), s, dtype) for s in range(3))
    y0, l0, m0 = fo.dense_fa(*(t.astype(np.float64) for t in (q, k, v)))
    O, l, m = fa.ring_dense_fa(*(to_dev(t, dtype) for t in (q, k, v)))
    tol = 1e-5 if dtype == torch.float32 else 2e-3
    assert rel_err(to_np(O), y0, dtype) < tol and rel_err(to_np(l), l0) < tol


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,N,d", [(torch.bfloat16, 512, 128), (torch.float32, 200, 32)])
def test_ring_backward_single_rank_equals_dense(dtype, N, d):
    import fa_sm100a as fa
    q, k, v, g = (randn_np((N, d, 2), s, dtype) for s in range(4))
    Q, K, V, G = (to_dev(t, dtype) for t in (q, k, v, g))
    O, l, m = fa.ring_dense_fa(Q, K, V)
    got = fa.ring_dense_fa_backward(Q, K, V, O, G, l, m)
    want = fo.dense_fa_backward_blocked(*(t.astype(np.float64) for t in (q, k, v)), to_np(O), g.astype(np.float64), to_np(l), to_np(m))
    tol = 1e-5 if dtype == torch.float32 else 2e-3
    for a, b in zip(got, want):
        assert rel_err(to_np(a), b, dtype) < tol


@pytest.mark.gpu
def test_ring_accepts_volume_shards():
    """A 3-D volume shard ``(X, Y, Z_local, d, B)`` (planes of the slowest spatial dim = contiguous tokens) goes through
    the ring entry points like ``dense_fa`` takes N-D inputs (src/dense.jl:1-19); one rank: equals ``dense_fa``."""
    import fa_sm100a as fa
    dtype = torch.bfloat16
    q, k, v, g = (to_dev(randn_np((8, 8, 8, 128, 2), s, dtype), dtype) for s in range(4))
    y0, l0, m0 = fa.dense_fa(q, k, v)
    y, l, m = fa.ring_dense_fa(q, k, v)
    assert tuple(y.shape) == (8, 8, 8, 128, 2) and rel_err(to_np(y), to_np(y0).astype(np.float64), dtype, want_rounded=True) < 2e-3
    grads = fa.ring_dense_fa_backward(q, k, v, y, g, l, m)
    want = fa.dense_fa_backward(*(fa._jl_reshape(t, (-1, 128, 2)) for t in (q, k, v, y, g)), l, m)
    for a, b in zip(grads, want):
        assert tuple(a.shape) == (8, 8, 8, 128, 2)
        assert rel_err(to_np(a).reshape((512, 128, 2), order="F"), to_np(b).astype(np.float64), dtype, want_rounded=True) < 2e-3


@pytest.mark.gpu
def test_merge_partials_kernel():
    import ctypes
    import fa_sm100a as fa
    N, dv, B = 300, 16, 3
    rng = np.random.default_rng(0)
    Oa, Ob = (np.asfortranarray(rng.standard_normal((N, dv, B)).astype(np.float32)) for _ in range(2))
    la, lb = (np.asfortranarray(rng.uniform(0.5, 50, (N, 1, B)).astype(np.float32)) for _ in range(2))
    ma, mb = (np.asfortranarray(rng.standard_normal((N, 1, B)).astype(np.float32) * 5) for _ in range(2))
    want = fo.merge_partials(*(t.astype(np.float64) for t in (Oa, la, ma, Ob, lb, mb)))
    d = [to_dev(t) for t in (Oa, la, ma, Ob, lb, mb)]
    out = fa.jl_empty((N, dv, B), torch.float32)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    rc = fa.lib.fa_merge_partials(p(d[0]), p(d[1]), p(d[2]), p(d[3]), p(d[4]), p(d[5]), p(out), N, dv, B, 0, 0, None)
    assert rc == 0, fa.lib.fa_last_error_string()
    torch.cuda.synchronize()
    assert rel_err(to_np(out), want[0]) < 1e-5 and rel_err(to_np(d[1]), want[1]) < 1e-5 and rel_err(to_np(d[2]), want[2]) < 1e-6


def _nccl_ring_worker(rank, world, port, q, k, v, dtype, out, g=None):
    sys.path.insert(0, os.path.join(ROOT, "flashattention.jl_b200"))
    import fa_sm100a as fa
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    n = q.shape[0] // world
    sh = lambda t: fa.jl_array(np.asfortranarray(t[rank * n:(rank + 1) * n]), dtype=dtype, device=f"cuda:{rank}")
    Q, K, V = sh(q), sh(k), sh(v)
    O, l, m = fa.ring_dense_fa(Q, K, V)
    grads = fa.ring_dense_fa_backward(Q, K, V, O, sh(g), l, m) if g is not None else ()
    out[rank] = (O.float().cpu(), l.cpu(), m.cpu()) + tuple(x.float().cpu() for x in grads)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_ring_two_ranks_nccl(dtype):
    N, d, B = (1024, 128, 2) if dtype == torch.bfloat16 else (256, 32, 2)
    q, k, v = (randn_np((N, d, B), s, dtype) for s in range(3))
    y0, l0, m0 = fo.dense_fa(*(t.astype(np.float64) for t in (q, k, v)))
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_nccl_ring_worker, args=(world, port, q, k, v, dtype, out), nprocs=world, join=True)
    y = np.concatenate([out[r][0].numpy() for r in range(world)])
    l = np.concatenate([out[r][1].numpy() for r in range(world)])
    m = np.concatenate([out[r][2].numpy() for r in range(world)])
    tol = 1e-5 if dtype == torch.float32 else 2e-3
    assert rel_err(y, y0, dtype) < tol
    assert rel_err(l * np.exp(m - m0), l0) < tol         # l is tied to its m


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("dtype,world", [(torch.bfloat16, 2), (torch.float32, 2)] + ([(torch.bfloat16, 4)] if torch.cuda.device_count() >= 4 else []))
def test_ring_backward_nccl(dtype, world):
    """dK/dV accumulators travelling with the K/V blocks over NCCL: every rank's (dq, dk, dv) shard against
    the oracle backward on the whole sequence (given the forward results the ranks computed)."""
    N, d, B = (1024, 128, 2) if dtype == torch.bfloat16 else (256, 32, 2)
    q, k, v, g = (randn_np((N, d, B), s, dtype) for s in range(4))
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_nccl_ring_worker, args=(world, port, q, k, v, dtype, out, g), nprocs=world, join=True)
    cat = lambda i: np.concatenate([out[r][i].numpy() for r in range(world)])
    O, l, m = cat(0), cat(1), cat(2)
    want = fo.dense_fa_backward_blocked(*(t.astype(np.float64) for t in (q, k, v)), O.astype(np.float64), g.astype(np.float64),
                                        l.astype(np.float64), m.astype(np.float64))
    tol = 1e-5 if dtype == torch.float32 else 2e-3
    for i, w in zip((3, 4, 5), want):
        assert rel_err(cat(i), w, dtype) < tol


@pytest.mark.gpu
@pytest.mark.parametrize("spatial,W,stride,pad,G,d,dtype", [
    ((64, 64, 64), 5, 5, 3, 8, 64, torch.bfloat16),        # BASELINE config 5, single volume over 8 ranks
    ((64, 64), 7, 7, 3, 3, 64, torch.bfloat16),            # config 2 geometry
    ((20, 9), 3, 4, 1, 3, 32, torch.float32),              # gaps between windows: uncovered planes stay NaN
    ((16,), 4, 4, 0, 2, 128, torch.float16),
    ((7, 8, 30), 3, 3, 1, 4, 16, torch.float32),
])
def test_windowed_slabs_reproduce_whole_volume_bitwise(spatial, W, stride, pad, G, d, dtype):
    """SURVEY 8(e), one volume over G ranks: every rank's slab (computed here one after the other on one GPU)
    holds exactly the planes / windows of the one-GPU call, forward and backward, bit for bit."""
    import fa_sm100a as fa
    B = 1
    q, k, v, g = (to_dev(randn_np(spatial + (d, B), s, dtype), dtype) for s in range(4))
    y0, l0, m0 = fa.windowed_fa(q, k, v, W, stride=stride, pad=pad)
    g0 = fa.windowed_fa_backward(q, k, v, g, l0, m0, W, stride=stride, pad=pad)
    ax = len(spatial) - 1
    ys, ls, ms, gs = [], [], [], ([], [], [])
    for r in range(G):
        pl = fa.windowed_slab_plan(spatial, W, stride, pad, r, G)
        if pl.nwin == 0:
            continue
        qs, ks, vs, dys = (fa.slab_planes(t, pl) for t in (q, k, v, g))
        y, l, m = fa.windowed_fa_slab(qs, ks, vs, W, pl, stride=stride, pad=pad)
        grads = fa.windowed_fa_slab_backward(qs, ks, vs, dys, l, m, W, pl, stride=stride, pad=pad)
        ys.append(y); ls.append(l); ms.append(m)
        for acc, t in zip(gs, grads):
            acc.append(t)
    eq = lambda a, b: torch.equal(torch.nan_to_num(a.float(), nan=12345.0), torch.nan_to_num(b.float(), nan=12345.0))
    assert eq(torch.cat(ys, dim=ax), y0) and eq(torch.cat(ls, dim=2), l0) and eq(torch.cat(ms, dim=2), m0)
    for acc, want in zip(gs, g0):
        assert eq(torch.cat(acc, dim=ax), want)


def test_windowed_slab_plan_random_geometries():
    """Randomised check of the slab plan (CPU only): for random extents / windows / strides / paddings / rank counts the C
    plan equals the brute-force oracle, the slabs tile the slowest dim, every window plane is owned exactly once, and each
    slab's own window geometry (pad_lo, nwin) reproduces the window starts of the whole volume."""
    import fa_sm100a as fa
    rng = np.random.default_rng(123)
    for _ in range(300):
        nd = int(rng.integers(1, 4))
        W = int(rng.integers(1, 9))
        stride = W + int(rng.integers(0, 4))
        pad = int(rng.integers(0, W))
        spatial = tuple(int(rng.integers(W, 40)) for _ in range(nd))
        G = int(rng.integers(1, 9))
        nw = fo.window_counts(spatial, W, stride, pad)[-1]
        plans = [fa.windowed_slab_plan(spatial, W, stride, pad, r, G) for r in range(G)]
        for r, pl in enumerate(plans):
            assert tuple(pl) == tuple(fo.windowed_slab_plan(spatial, W, stride, pad, r, G)), (spatial, W, stride, pad, G, r)
        live = [pl for pl in plans if pl.nwin > 0]
        assert sum(pl.nwin for pl in plans) == nw and live[0].plane_lo == 0 and live[-1].plane_hi == spatial[-1]
        for a, b in zip(live, live[1:]):
            assert a.plane_hi == b.plane_lo and a.win_hi == b.win_lo
        for pl in live:
            # window w of the slab starts at w * stride - pad_lo in slab coordinates = (win_lo + w) * stride - pad in the volume
            for w in range(pl.nwin):
                assert pl.plane_lo + w * stride - pl.pad_lo == (pl.win_lo + w) * stride - pad
            # every plane the slab's windows read and the volume holds is inside the slab
            first = pl.win_lo * stride - pad
            last = (pl.win_hi - 1) * stride - pad + W - 1
            assert pl.plane_lo <= max(first, 0) and min(last, spatial[-1] - 1) < pl.plane_hi


# ------------------------------------------------------------------------------------------------ halo splits
def test_windowed_halo_plan_tiles_volume_and_covers_windows():
    """fa_windowed_halo_plan (host, integer): owned planes tile the volume, window ranges tile the windows, and every
    window of a rank reads only planes [own_lo, ext_hi) with pad_lo zero planes in front -- against the definition."""
    import fa_sm100a as fa
    from fa_sm100a import halo
    rng = np.random.default_rng(0)
    for _ in range(300):
        S = int(rng.integers(8, 200)); W = int(rng.integers(1, 12)); stride = int(rng.integers(1, W + 3)); pad = int(rng.integers(0, W))
        G = int(rng.integers(1, 9))
        if S + 2 * pad < W:
            continue
        nw = (S + 2 * pad - W) // stride + 1
        plans = [halo.windowed_halo_plan((S,), W, stride, pad, r, G) for r in range(G)]
        assert plans[0].own_lo == 0 and plans[-1].own_hi == S and plans[0].win_lo == 0 and plans[-1].win_hi == nw
        for a, b in zip(plans, plans[1:]):
            assert a.own_hi == b.own_lo and a.win_hi == b.win_lo
        for pl in plans:
            assert pl.own_lo <= pl.own_hi <= pl.ext_hi <= S
            for w in range(pl.win_lo, pl.win_hi):
                lo, hi = w * stride - pad, w * stride - pad + W          # planes the window reads (outside [0, S): padding)
                assert max(lo, 0) >= pl.own_lo
                assert min(hi, S) <= pl.ext_hi
                assert lo - (pl.win_lo * stride - pad) == (w - pl.win_lo) * stride
            if pl.nwin:
                assert pl.pad_lo == pl.own_lo - (pl.win_lo * stride - pad) and pl.pad_lo >= 0


def test_local_comm_is_a_ring():
    from fa_sm100a.halo import LocalComm
    def fn(c):
        a, b = c.exchange(torch.tensor([c.rank * 10.0]), torch.tensor([c.rank * 10.0 + 1]), torch.zeros(1), torch.zeros(1))
        return float(a), float(b)
    out = LocalComm(4).run(fn)
    assert out == [(31.0, 10.0), (1.0, 20.0), (11.0, 30.0), (21.0, 0.0)]      # (prev's to_next, next's to_prev)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,N,d,W,G", [(torch.bfloat16, 1024, 64, 255, 4), (torch.bfloat16, 512, 64, 64, 2), (torch.float32, 240, 16, 9, 3),
                                           (torch.float16, 768, 128, 33, 3)])
def test_circulant_halo_split_equals_whole_sequence(dtype, N, d, W, G):
    """One periodic-band sequence sharded by tokens over G ranks (all ranks emulated in this process, LocalComm): every
    rank's (O, l, m) and (dq, dk, dv) against the unsharded call / the oracle; the halo exchange is the only coupling."""
    import fa_sm100a as fa
    from fa_sm100a import halo
    q, k, v, g = (randn_np((N, d, 2), s, dtype) for s in range(4))
    Nl = N // G
    sh = lambda t, r: to_dev(np.asfortranarray(t[r * Nl:(r + 1) * Nl]), dtype)

    def rank_fn(c):
        r = c.rank
        O, l, m, ctx = halo.circulant_fa_halo(sh(q, r), sh(k, r), sh(v, r), W, c)
        grads = halo.circulant_fa_halo_backward(ctx, sh(g, r), W, c)
        return tuple(to_np(t) for t in (O, l, m) + tuple(grads))
    outs = halo.LocalComm(G).run(rank_fn)
    cat = lambda i: np.concatenate([o[i] for o in outs])
    O0, l0, m0 = fo.circulant_fa(*(t.astype(np.float64) for t in (q, k, v)), W)
    tol = 1e-5 if dtype == torch.float32 else 2e-3
    assert rel_err(cat(0), O0, dtype) < tol and rel_err(cat(1), l0) < tol and np.abs(cat(2) - m0).max() < tol * max(1.0, np.abs(m0).max())
    want = fo.circulant_backward_given(*(t.astype(np.float64) for t in (q, k, v)), cat(0), g.astype(np.float64), cat(1), cat(2), W)
    for i, w in zip((3, 4, 5), want):
        assert rel_err(cat(i), w, dtype) < tol


@pytest.mark.gpu
@pytest.mark.parametrize("spatial,W,stride,pad,G,d,dtype", [
    ((64,), 16, 4, 0, 3, 64, torch.bfloat16),              # 1-D, heavy overlap
    ((4096,), 64, 16, 0, 4, 64, torch.bfloat16),           # logs/compare1.txt windowed geometry (stride 16)
    ((16, 24), 3, 1, 1, 4, 64, torch.float16),             # 2-D sliding windows, padding on both sides
    ((12, 10, 20), 5, 2, 2, 3, 64, torch.bfloat16),        # 3-D overlapping
    ((20, 30), 7, 7, 3, 2, 16, torch.float32),             # exact cover goes through the same path (no halo needed)
    ((9, 40), 4, 3, 0, 5, 8, torch.float32),               # uncovered tail planes -> NaN
])
def test_windowed_halo_split_equals_whole_volume(spatial, W, stride, pad, G, d, dtype):
    """One volume with OVERLAPPING windows over G ranks (emulated in-process): y of every rank's planes and (l, m) of its
    windows against windowed_fa on the whole volume (oracle), NaN pattern included."""
    import fa_sm100a as fa
    from fa_sm100a import halo
    q, k, v = (randn_np(spatial + (d, 2), s, dtype) for s in range(3))
    ax = len(spatial) - 1
    plans = [halo.windowed_halo_plan(spatial, W, stride, pad, r, G) for r in range(G)]

    def rank_fn(c):
        pl = plans[c.rank]
        cut = lambda t: to_dev(np.asfortranarray(np.take(t, range(pl.own_lo, pl.own_hi), axis=ax)), dtype)
        y, l, m = halo.windowed_fa_halo(cut(q), cut(k), cut(v), spatial, W, c, stride, pad)
        return to_np(y), to_np(l), to_np(m)
    outs = halo.LocalComm(G).run(rank_fn)
    y = np.concatenate([o[0] for o in outs], axis=ax)
    l = np.concatenate([o[1] for o in outs], axis=2)
    m = np.concatenate([o[2] for o in outs], axis=2)
    y0, l0, m0 = fo.windowed_fa(*(t.astype(np.float64) for t in (q, k, v)), W, stride, pad)
    tol = 1e-5 if dtype == torch.float32 else 2e-3
    assert y.shape == y0.shape and l.shape == l0.shape
    assert rel_err(y, y0, dtype) < tol and rel_err(l, l0) < max(tol, 1e-5) and rel_err(m, m0) < max(tol, 1e-5)


def _nccl_halo_worker(rank, world, port, q, k, v, g, W, out):
    sys.path.insert(0, os.path.join(ROOT, "flashattention.jl_b200"))
    import fa_sm100a as fa
    from fa_sm100a import halo
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    comm = halo.DistComm()
    dt = torch.bfloat16
    n = q.shape[0] // world
    sh = lambda t: fa.jl_array(np.asfortranarray(t[rank * n:(rank + 1) * n]), dtype=dt, device=f"cuda:{rank}")
    O, l, m, ctx = halo.circulant_fa_halo(sh(q), sh(k), sh(v), W, comm)
    grads = halo.circulant_fa_halo_backward(ctx, sh(g), W, comm)
    # windowed: a 1-D sequence with overlapping windows, planes = tokens
    pl = halo.windowed_halo_plan((q.shape[0],), 64, 16, 0, rank, world)
    cut = lambda t: fa.jl_array(np.asfortranarray(t[pl.own_lo:pl.own_hi]), dtype=dt, device=f"cuda:{rank}")
    yw, lw, mw = halo.windowed_fa_halo(cut(q), cut(k), cut(v), (q.shape[0],), 64, comm, 16, 0)
    out[rank] = tuple(t.float().cpu() for t in (O, l, m) + tuple(grads) + (yw, lw))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_halo_splits_two_ranks_nccl():
    """The two halo splits with REAL NCCL point-to-point exchanges between two GPUs."""
    N, d, W = 2048, 64, 255
    q, k, v, g = (randn_np((N, d, 2), s, torch.bfloat16) for s in range(4))
    world, port = 2, _free_port()
    out = mp.Manager().dict()
    mp.spawn(_nccl_halo_worker, args=(world, port, q, k, v, g, W, out), nprocs=world, join=True)
    cat = lambda i, ax=0: np.concatenate([out[r][i].numpy().astype(np.float64) for r in range(world)], axis=ax)
    O0, l0, m0 = fo.circulant_fa(*(t.astype(np.float64) for t in (q, k, v)), W)
    assert rel_err(cat(0), O0, torch.bfloat16) < 2e-3 and rel_err(cat(1), l0) < 2e-3
    want = fo.circulant_backward_given(*(t.astype(np.float64) for t in (q, k, v)), cat(0), g.astype(np.float64), cat(1), cat(2), W)
    for i, w in zip((3, 4, 5), want):
        assert rel_err(cat(i), w, torch.bfloat16) < 2e-3
    y0, lw0, _ = fo.windowed_fa(*(t.astype(np.float64) for t in (q, k, v)), 64, 16, 0)
    assert rel_err(cat(6), y0, torch.bfloat16) < 2e-3 and rel_err(cat(7, 2), lw0) < 2e-3
