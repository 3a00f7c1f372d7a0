# ChainRulesCore rrules for the three flash ops (SURVEY 8f-1; the reference has none, README.md:40,53):
# Zygote / Flux users get the backward kernels of libfa_sm100a.  Only the first output (y) carries a
# cotangent; l and m are saved for the pullback.  (UNVERIFIED HERE: no Julia runtime in the build image;
# the same forward/backward pairs are exercised through fa_sm100a/autograd.py.)
using ChainRulesCore

function ChainRulesCore.rrule(::typeof(dense_fa), q::CuArray{T, D}, k::CuArray{T, D}, v::CuArray{T, D}) where {T, D}
    y, l, m = dense_fa(q, k, v)
    function dense_fa_pullback(Δ)
        dy = unthunk(Δ[1])
        d, dv, B = size(q, D-1), size(v, D-1), size(q, D)
        r(x, c) = reshape(x, :, c, B)
        dQ, dK, dV = dense_fa_backward(r(q, d), r(k, d), r(v, dv), r(y, dv), r(dy, dv), l, m)
        return NoTangent(), reshape(dQ, size(q)), reshape(dK, size(k)), reshape(dV, size(v))
    end
    return (y, l, m), dense_fa_pullback
end

function ChainRulesCore.rrule(::typeof(windowed_fa), q::CuArray, k::CuArray, v::CuArray, windowsize::Int;
                              stride=windowsize, pad=(windowsize-1)÷2)
    y, l, m = windowed_fa(q, k, v, windowsize; stride=stride, pad=pad)
    function windowed_fa_pullback(Δ)
        dq, dk, dv = windowed_fa_backward(q, k, v, unthunk(Δ[1]), l, m, windowsize; stride=stride, pad=pad)
        return NoTangent(), dq, dk, dv, NoTangent()
    end
    return (y, l, m), windowed_fa_pullback
end

function ChainRulesCore.rrule(::typeof(circulant_fa), Q::CuArray, K::CuArray, V::CuArray, W::Int)
    O, l, m = circulant_fa(Q, K, V, W)
    function circulant_fa_pullback(Δ)
        dQ, dK, dV = circulant_fa_backward(Q, K, V, O, unthunk(Δ[1]), l, m, W)
        return NoTangent(), dQ, dK, dV, NoTangent()
    end
    return (O, l, m), circulant_fa_pullback
end
