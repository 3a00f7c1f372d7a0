"""Multi-GPU rows of SURVEY 8e.

CPU (gloo, world_size 2):  the batch*head shard arithmetic (fa_shard_batch), and the ring schedule --
K/V token blocks passed rank -> rank+1 while each rank merges partial (O, l, m) with the update rule
of src/dense.jl:82-91 -- run with REAL send/recv between two processes, the oracle standing in for the
block kernel (the CUDA kernels cannot run on the CPU and there is no fallback).
GPU: fa_ring_dense_fwd on one rank (degenerate ring) and, when two GPUs are visible, on two ranks over
NCCL, against the oracle on the whole sequence."""
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
