"""Differentiable wrappers (the reference ships no rrules, README.md todo): the same pairs
of forward / backward entry points the Julia `rrules.jl` binds through ChainRulesCore, here as
``torch.autograd.Function`` so the gradients can be exercised end to end.  Inputs and outputs are
Julia-shaped column-major tensors as everywhere in :mod:`fa_sm100a`."""
import torch

from . import (circulant_fa, circulant_fa_backward, dense_fa, dense_fa_backward, jl_array, windowed_fa,
               windowed_fa_backward, _flatten3, _jl_reshape)


class _Dense(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v):
        y, l, m = dense_fa(q, k, v)
        ctx.save_for_backward(q, k, v, y, l, m)
        return y

    @staticmethod
    def backward(ctx, dy):
        q, k, v, y, l, m = ctx.saved_tensors
        N, d, B = _flatten3(q)
        dv = int(v.shape[-2])
        f = lambda t, c: _jl_reshape(jl_array(t), (N, c, B))
        dq, dk, dvv = dense_fa_backward(f(q, d), f(k, d), f(v, dv), f(y, dv), f(dy, dv), l, m)
        return _jl_reshape(dq, q.shape), _jl_reshape(dk, k.shape), _jl_reshape(dvv, v.shape)


class _Windowed(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, W, stride, pad):
        y, l, m = windowed_fa(q, k, v, W, stride, pad)
        ctx.save_for_backward(q, k, v, l, m)
        ctx.cfg = (W, stride, pad)
        return y

    @staticmethod
    def backward(ctx, dy):
        q, k, v, l, m = ctx.saved_tensors
        W, stride, pad = ctx.cfg
        dq, dk, dv = windowed_fa_backward(q, k, v, jl_array(dy), l, m, W, stride, pad)
        return dq, dk, dv, None, None, None


class _Circulant(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, W):
        O, l, m = circulant_fa(q, k, v, W)
        ctx.save_for_backward(q, k, v, O, l, m)
        ctx.W = W
        return O

    @staticmethod
    def backward(ctx, dO):
        q, k, v, O, l, m = ctx.saved_tensors
        dq, dk, dv = circulant_fa_backward(q, k, v, O, jl_array(dO), l, m, ctx.W)
        return dq, dk, dv, None


def dense_attention(q, k, v):
    """``first(dense_fa(q, k, v))`` with gradients (rrule of dense_fa)."""
    return _Dense.apply(q, k, v)


def windowed_attention(q, k, v, windowsize, stride=None, pad=None):
    """``first(windowed_fa(q, k, v, W; stride, pad))`` with gradients."""
    return _Windowed.apply(q, k, v, windowsize, stride, pad)


def circulant_attention(q, k, v, W):
    """``first(circulant_fa(Q, K, V, W))`` with gradients."""
    return _Circulant.apply(q, k, v, W)
