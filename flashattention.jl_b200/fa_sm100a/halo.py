"""One problem over several GPUs with a HALO exchange:

  * ``circulant_fa_halo`` / ``circulant_fa_halo_backward`` -- one long sequence sharded by tokens, periodic band of W
    keys: every rank needs ``p = (W-1)//2`` keys of its left and ``W-1-p`` of its right neighbour ("ring/halo of p tokens
    each side").  The shard is extended by the two halos (rounded up to 64 tokens so that the tcgen05 kernels apply) and
    the unchanged 1-D circulant kernels run on the extended arrays: real queries never see the wrap-around of the
    extended ring, the halo rows are dummy queries (q = 0, dO = 0: they contribute nothing to dK / dV).  Backward:
    the halo parts of dK, dV go back to the neighbours and are added there.
  * ``windowed_fa_halo`` -- one volume with OVERLAPPING windows split along its slowest spatial dim: halo of
    ``W - stride`` planes of q, k, v in, partial-y halo reduce out (``fa_windowed_halo_plan``,
    ``fa_windowed_slab_fwd_sums``, ``fa_window_divide``).  Forward.

The neighbour exchanges are ``torch.distributed`` point-to-point operations (NCCL over NVLink between GPUs); a ``Comm``
object abstracts them so that tests can also run all ranks of a split inside one process (``LocalComm``).
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

import torch

from . import (FA_FLAG_OUT_F32, FaError, _check, _dt, _i64arr, _ptr, _stream, _win_kws, circulant_fa, circulant_fa_backward,
               jl_array, jl_empty, lib, window_counts)

__all__ = ["DistComm", "LocalComm", "circulant_fa_halo", "circulant_fa_halo_backward", "windowed_halo_plan", "windowed_fa_halo"]


class DistComm:
    """Neighbour exchange over a torch.distributed process group (ranks form a ring)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def exchange(self, to_prev: Optional[torch.Tensor], to_next: Optional[torch.Tensor], from_prev_like, from_next_like):
        """Send ``to_prev`` / ``to_next`` to ranks r-1 / r+1 (ring) and return (from_prev, from_next) shaped like the given
        templates (None = nothing expected from that side)."""
        d = self.dist
        prev, nxt = (self.rank - 1) % self.world, (self.rank + 1) % self.world
        ops, rp, rn = [], None, None
        if to_prev is not None:
            ops.append(d.P2POp(d.isend, to_prev.contiguous(), prev, self.group))
        if to_next is not None:
            ops.append(d.P2POp(d.isend, to_next.contiguous(), nxt, self.group))
        # receives are posted in the order the PEER sends (to_prev first, to_next second): with two ranks prev == next and
        # messages between one pair of ranks match in posting order -- my next's `to_prev` is my `from_next`
        if from_next_like is not None:
            rn = torch.empty_like(from_next_like, memory_format=torch.contiguous_format)
            ops.append(d.P2POp(d.irecv, rn, nxt, self.group))
        if from_prev_like is not None:
            rp = torch.empty_like(from_prev_like, memory_format=torch.contiguous_format)
            ops.append(d.P2POp(d.irecv, rp, prev, self.group))
        for w in (d.batch_isend_irecv(ops) if ops else []):
            w.wait()
        return rp, rn


class LocalComm:
    """All ranks of a split inside ONE process (tests, single-GPU emulation): ``run(fn)`` calls ``fn(comm_r)`` for every
    rank in lock step -- each ``exchange`` is a rendezvous of all ranks, implemented with one thread per rank."""

    def __init__(self, world: int):
        import threading
        self.world = world
        self._barrier = threading.Barrier(world)
        self._box = [None] * world

    class _Rank:
        def __init__(self, parent, rank):
            self.p, self.rank, self.world = parent, rank, parent.world

        def exchange(self, to_prev, to_next, from_prev_like, from_next_like):
            p = self.p
            p._box[self.rank] = (to_prev, to_next)
            p._barrier.wait()
            prev, nxt = (self.rank - 1) % self.world, (self.rank + 1) % self.world
            rp = p._box[prev][1].clone() if from_prev_like is not None else None     # what prev sent to ITS next
            rn = p._box[nxt][0].clone() if from_next_like is not None else None      # what next sent to ITS prev
            p._barrier.wait()
            return rp, rn

    def run(self, fn):
        import threading
        out, err = [None] * self.world, []

        def work(r):
            try:
                out[r] = fn(LocalComm._Rank(self, r))
            except BaseException as e:      # noqa: BLE001 -- re-raised below; release the other ranks
                err.append(e)
                self._barrier.abort()
        ts = [threading.Thread(target=work, args=(r,)) for r in range(self.world)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        if err:
            raise err[0]
        return out


def _up64(x: int) -> int:
    return (x + 63) // 64 * 64


def _tok(x: torch.Tensor, lo: int, hi: int) -> torch.Tensor:
    return x[lo:hi]


# ---------------------------------------------------------------------------------------------- circulant
def _circ_halos(Nl: int, W: int):
    p = (W - 1) // 2                                        # src/utils.jl:8
    right = W - 1 - p
    hl, hr = (_up64(p) if p else 0), (_up64(right) if right else 0)
    if hl > Nl or hr > Nl:
        raise FaError(f"circulant_fa_halo: the band (W = {W}) reaches beyond the neighbouring shard ({Nl} tokens); use ring_dense_fa")
    return hl, hr


def circulant_fa_halo(q, k, v, W: int, comm, flags: int = 0):
    """``circulant_fa!`` (src/circulant.jl:9-118) on ONE sequence of ``world * Nl`` tokens sharded by tokens: rank r
    passes its ``(Nl, d, B)`` shards and gets ``(O, l, m)`` of its queries plus a context for the backward."""
    q, k, v = (jl_array(t) for t in (q, k, v))
    Nl, d, B = (int(s) for s in q.shape)
    if comm.world == 1:
        O, l, m = circulant_fa(q, k, v, W, flags)
        return O, l, m, None
    hl, hr = _circ_halos(Nl, W)
    # my last hl tokens are the left halo of rank r+1; my first hr tokens the right halo of rank r-1
    kv = torch.cat([k.permute(2, 1, 0), v.permute(2, 1, 0)], dim=1)            # (B, d + dv, Nl), token-contiguous rows
    from_prev, from_next = comm.exchange(kv[:, :, :hr] if hr else None, kv[:, :, Nl - hl:] if hl else None,
                                         kv[:, :, :hl] if hl else None, kv[:, :, :hr] if hr else None)
    parts = [t for t in (from_prev, kv, from_next) if t is not None]
    kv_ext = torch.cat(parts, dim=2)
    Ne = kv_ext.shape[2]
    K_ext = jl_array(kv_ext[:, :d].permute(2, 1, 0))
    V_ext = jl_array(kv_ext[:, d:].permute(2, 1, 0))
    Q_ext = jl_empty((Ne, d, B), q.dtype, q.device).zero_()
    Q_ext[hl:hl + Nl] = q
    O_ext, l_ext, m_ext = circulant_fa(Q_ext, K_ext, V_ext, W, flags)
    ctx = (Q_ext, K_ext, V_ext, O_ext, l_ext, m_ext, hl, hr, Nl)
    return jl_array(O_ext[hl:hl + Nl]), jl_array(l_ext[hl:hl + Nl]), jl_array(m_ext[hl:hl + Nl]), ctx


def circulant_fa_halo_backward(ctx, dO, W: int, comm, q=None, k=None, v=None, O=None, l=None, m=None, flags: int = 0):
    """Backward of :func:`circulant_fa_halo`: ``(dq, dk, dv)`` of the rank's shard.  The gradients of the halo keys are
    sent back to the neighbours that own them and added there (float32 sums)."""
    if ctx is None:                                          # one rank
        return circulant_fa_backward(q, k, v, O, dO, l, m, W, flags)
    Q_ext, K_ext, V_ext, O_ext, l_ext, m_ext, hl, hr, Nl = ctx
    dO = jl_array(dO)
    G_ext = jl_empty(tuple(O_ext.shape), dO.dtype, dO.device).zero_()
    G_ext[hl:hl + Nl] = dO
    use_f32 = Q_ext.dtype != torch.float32
    try:
        dQe, dKe, dVe = circulant_fa_backward(Q_ext, K_ext, V_ext, O_ext, G_ext, l_ext, m_ext, W, flags | (FA_FLAG_OUT_F32 if use_f32 else 0))
    except FaError:                                          # shape on the exact-fp32 kernels: gradients in the input type
        dQe, dKe, dVe = circulant_fa_backward(Q_ext, K_ext, V_ext, O_ext, G_ext, l_ext, m_ext, W, flags)
    g = torch.cat([dKe.float().permute(2, 1, 0), dVe.float().permute(2, 1, 0)], dim=1)       # (B, d + dv, Ne)
    d = int(dKe.shape[1])
    own = g[:, :, hl:hl + Nl].clone()
    # left-halo gradients belong to rank r-1's LAST hl tokens, right-halo gradients to rank r+1's FIRST hr tokens
    from_prev, from_next = comm.exchange(g[:, :, :hl] if hl else None, g[:, :, hl + Nl:] if hr else None,
                                         g[:, :, :hr] if hr else None, g[:, :, :hl] if hl else None)
    if from_prev is not None:
        own[:, :, :hr] += from_prev
    if from_next is not None:
        own[:, :, Nl - hl:] += from_next
    dt = Q_ext.dtype
    dq = jl_array(dQe[hl:hl + Nl].to(dt))
    dk = jl_array(own[:, :d].permute(2, 1, 0).to(dt))
    dv = jl_array(own[:, d:].permute(2, 1, 0).to(dt))
    return dq, dk, dv


# ---------------------------------------------------------------------------------------------- windowed, overlapping
class HaloPlan(tuple):
    """(own_lo, own_hi, ext_hi, win_lo, win_hi, pad_lo) -- see fa_windowed_halo_plan (include/fa_sm100a.h)."""
    own_lo = property(lambda s: s[0]); own_hi = property(lambda s: s[1]); ext_hi = property(lambda s: s[2])
    win_lo = property(lambda s: s[3]); win_hi = property(lambda s: s[4]); pad_lo = property(lambda s: s[5])
    nwin = property(lambda s: s[4] - s[3]); halo = property(lambda s: s[2] - s[1])


def windowed_halo_plan(spatial: Sequence[int], W: int, stride=None, pad=None, rank: int = 0, nranks: int = 1) -> HaloPlan:
    stride, pad = _win_kws(W, stride, pad)
    plan = _i64arr([0] * 6)
    _check(lib.fa_windowed_halo_plan(len(spatial), _i64arr(spatial), int(W), stride, pad, int(rank), int(nranks), plan), "fa_windowed_halo_plan")
    return HaloPlan(int(x) for x in plan)


def windowed_fa_halo(q, k, v, spatial: Sequence[int], W: int, comm, stride=None, pad=None, flags: int = 0):
    """``windowed_fa`` (src/windowed.jl:3-23) on ONE volume of extents ``spatial`` split along its slowest spatial dim,
    windows may overlap.  Rank r passes the planes ``[own_lo, own_hi)`` of q, k, v (``windowed_halo_plan``) and gets
    ``y`` for those planes and ``l, m :: (W^D, 1, L_r, B)`` for its windows (a contiguous range of the window index)."""
    stride, pad = _win_kws(W, stride, pad)
    q, k, v = (jl_array(t) for t in (q, k, v))
    spatial = tuple(int(s) for s in spatial)
    nd = len(spatial)
    plans = [windowed_halo_plan(spatial, W, stride, pad, r, comm.world) for r in range(comm.world)]
    me = plans[comm.rank]
    for r, pl in enumerate(plans):
        if pl.halo and (r + 1 >= comm.world or pl.ext_hi > plans[r + 1].own_hi):
            raise FaError("windowed_fa_halo: a window spans more than two slabs (slabs thinner than W - stride planes)")
    if tuple(int(s) for s in q.shape[:nd - 1]) != spatial[:-1] or int(q.shape[nd - 1]) != me.own_hi - me.own_lo:
        raise FaError("windowed_fa_halo: q, k, v must hold exactly the planes [own_lo, own_hi) of the plan")
    d, dv, B = int(q.shape[-2]), int(v.shape[-2]), int(q.shape[-1])
    ax = nd - 1
    # halo in: the first plans[r-1].halo planes of my slab go to rank r-1; I receive me.halo planes from rank r+1
    give = plans[comm.rank - 1].halo if comm.rank > 0 else 0
    qkv = torch.cat([t.movedim(ax, 0).reshape(t.shape[ax], -1) for t in (q, k, v)], dim=1)       # (planes, rest) rows
    like_next = qkv[:1].expand(me.halo, -1) if me.halo else None
    _, from_next = comm.exchange(qkv[:give] if give else None, None, None, like_next)
    ext = torch.cat([qkv, from_next], dim=0) if me.halo else qkv
    planes = ext.shape[0]

    def unpack(block, ch):
        shp = (planes,) + spatial[:-1] + (ch, B)
        return jl_array(block.reshape(shp).movedim(0, ax))
    nq = q.numel() // q.shape[ax]
    nv = v.numel() // v.shape[ax]
    Qe, Ke, Ve = unpack(ext[:, :nq], d), unpack(ext[:, nq:2 * nq], d), unpack(ext[:, 2 * nq:], dv)
    slab_dims = spatial[:-1] + (planes,)
    nw = (window_counts(spatial[:-1], W, stride, pad) if nd > 1 else ()) + (me.nwin,)
    L = 1
    for n in nw:
        L *= n
    WD = W ** nd
    acc = jl_empty(slab_dims + (dv, B), torch.float32, q.device)
    l = jl_empty((WD, 1, L, B), torch.float32, q.device)
    m = jl_empty((WD, 1, L, B), torch.float32, q.device)
    if me.nwin > 0:
        with torch.cuda.device(q.device):
            _check(lib.fa_windowed_slab_fwd_sums(_ptr(Qe), _ptr(Ke), _ptr(Ve), _ptr(acc), _ptr(l), _ptr(m), nd, _i64arr(slab_dims),
                                                 d, dv, B, W, stride, pad, me.pad_lo, me.nwin, _dt(q), flags, _stream(q)),
                   "fa_windowed_slab_fwd_sums")
    else:
        acc.zero_()
    # halo reduce out: my sums for the planes of rank r+1 go there; the sums rank r-1 computed for my first planes come in
    rows = acc.movedim(ax, 0).reshape(planes, -1)
    own_n = me.own_hi - me.own_lo
    like_prev = rows[:1].expand(give, -1) if give else None
    from_prev, _ = comm.exchange(None, rows[own_n:].contiguous() if me.halo else None, like_prev, None)
    own = rows[:own_n].clone()
    if give:
        own[:give] += from_prev
    own_acc = jl_array(own.reshape((own_n,) + spatial[:-1] + (dv, B)).movedim(0, ax))
    y = jl_empty(spatial[:-1] + (own_n, dv, B), q.dtype, q.device)
    if own_n > 0:
        with torch.cuda.device(q.device):
            _check(lib.fa_window_divide(_ptr(own_acc), _ptr(y), nd, _i64arr(spatial), W, stride, pad, me.own_lo, own_n, dv, B,
                                        _dt(q), _stream(q)), "fa_window_divide")
    return y, l, m
