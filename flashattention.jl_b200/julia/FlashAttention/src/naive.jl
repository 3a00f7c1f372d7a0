# Naive oracles, retained naive (north_star): they materialise P and are only definitions to
# compare against -- same role and results as reference src/naive/{dense,windowed,circulant}.jl,
# with the reference's undefined-variable bugs (SURVEY Appendix B-2) not reproduced.
function dense_dpa!(O::AbstractArray{T, 3}, P::AbstractArray{T, 3}, Q::AbstractArray{T, 3},
                    K::AbstractArray{T, 3}, V::AbstractArray{T, 3}) where T
    batched_mul!(P, Q, batched_transpose(K), one(T) / T(sqrt(size(Q, 2))), zero(T))
    NNlib.softmax!(P, dims=2)
    batched_mul!(O, P, V)
    return O, P
end

function dense_dpa(q::AbstractArray{T, N}, k::AbstractArray{T, N}, v::AbstractArray{T, N}) where {T, N}
    dqk, dvo, bs = size(q, N-1), size(v, N-1), size(q, N)
    Q, K, V = reshape(q, :, dqk, bs), reshape(k, :, dqk, bs), reshape(v, :, dvo, bs)
    O = similar(Q, size(Q, 1), dvo, bs)
    P = similar(Q, size(Q, 1), size(Q, 1), bs)
    dense_dpa!(O, P, Q, K, V)
    return reshape(O, size(q)[1:N-2]..., dvo, bs), P
end

block_dpa(q, k, v, windowsize) = windowed_dpa(q, k, v, windowsize)

function windowed_dpa(q::AbstractArray{T, N}, k::AbstractArray{T, N}, v::AbstractArray{T, N}, windowsize; kws...) where {T, N}
    qw, kw, vw = window(q, windowsize; kws...), window(k, windowsize; kws...), window(v, windowsize; kws...)
    yw, Pw = dense_dpa(reshape(qw, size(qw, 1), size(qw, 2), :), reshape(kw, size(kw, 1), size(kw, 2), :),
                       reshape(vw, size(vw, 1), size(vw, 2), :))
    yw = reshape(yw, size(yw, 1), size(yw, 2), size(qw, 3), :)
    szy = (size(q)[1:N-2]..., size(v, N-1), size(q, N))
    divisor = unwindow(window(ones_like(v, szy), windowsize; kws...), szy, windowsize; kws...)
    y = unwindow(yw, szy, windowsize; kws...) ./ divisor
    return y, reshape(Pw, size(Pw, 1), size(Pw, 2), :, size(q, N))
end

function circulant_dpa(Q, K, V, W)
    N, d, batchsize = size(Q)
    return circulant_dpa!(similar(Q, N, size(V, 2), batchsize), similar(Q, W, N, batchsize), Q, K, V, W)
end

function circulant_dpa!(O::Array{T, 3}, P::Array{T, 3}, Q::Array{T, 3}, K::Array{T, 3}, V::Array{T, 3}, W::Int) where T
    N, d, batchsize = size(Q)
    τ = one(T) / T(sqrt(d))
    keys = circulant_keys(N, W)
    Threads.@threads for idx in CartesianIndices((N, W, batchsize))
        ii, ww, bb = idx.I
        jj = keys[ww, ii] + 1
        P[ww, ii, bb] = τ * sum(Q[ii, :, bb] .* K[jj, :, bb])
    end
    softmax!(P, dims=1)
    Ps = batch_circulant(P) |> transpose
    bV = vcat([V[:, :, b] for b = 1:batchsize]...)
    bO = Ps * bV
    return cat([bO[(b-1)*N+1:b*N, :] for b = 1:batchsize]...; dims=3), Ps
end
