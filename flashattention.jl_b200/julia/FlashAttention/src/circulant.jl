# circulant_fa / circulant_fa! -- reference src/circulant.jl:1,9.  The reference's allocating
# wrapper forgets to pass W (src/circulant.jl:6, a MethodError); this one passes it.
@inline function circulant_fa(Q, K, V, W)
    N, d, batchsize = size(Q)
    O = similar(Q, N, size(V, 2), batchsize)
    l = statarray(Q, N, 1, batchsize)
    m = statarray(Q, N, 1, batchsize)
    return circulant_fa!(O, l, m, Q, K, V, W)
end

function circulant_fa!(O::CuArray{T, 3}, l::CuArray{Float32, 3}, m::CuArray{Float32, 3},
                       Q::CuArray{T, 3}, K::CuArray{T, 3}, V::CuArray{T, 3}, W::Int; flags::Integer=0) where {T}
    N, d, batchsize = size(Q)
    rc = GC.@preserve K O Q V l m begin
        ccall(sym(:fa_circulant_fwd), Cint,
                   (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid},
                    Int64, Int64, Int64, Int64, Int64, Cint, Cint, Ptr{Cvoid}),
                   devptr(Q), devptr(K), devptr(V), devptr(O), devptr(l), devptr(m),
                   N, d, size(V, 2), batchsize, W, fa_dtype(T), Cint(flags), current_stream())
    end
    check(rc, "fa_circulant_fwd")
    return O, l, m
end

function circulant_fa!(O::Array{T, 3}, l::Array{Float32, 3}, m::Array{Float32, 3},
                       Q::Array{T, 3}, K::Array{T, 3}, V::Array{T, 3}, W::Int; flags::Integer=0, device::Integer=0) where {T}
    N, d, batchsize = size(Q)
    rc = ccall(sym(:fa_circulant_fwd_host), Cint,
               (Ptr{T}, Ptr{T}, Ptr{T}, Ptr{T}, Ptr{Float32}, Ptr{Float32},
                Int64, Int64, Int64, Int64, Int64, Cint, Cint, Cint),
               Q, K, V, O, l, m, N, d, size(V, 2), batchsize, W, fa_dtype(T), Cint(flags), Cint(device))
    check(rc, "fa_circulant_fwd_host")
    return O, l, m
end

function circulant_fa_backward(Q::CuArray{T, 3}, K::CuArray{T, 3}, V::CuArray{T, 3}, O::CuArray{T, 3},
                               dO::CuArray{T, 3}, l::CuArray{Float32, 3}, m::CuArray{Float32, 3}, W::Int;
                               flags::Integer=0) where T
    N, d, batchsize = size(Q)
    dv = size(V, 2)
    dQ, dK, dV = similar(Q), similar(K), similar(V)
    nws = ccall(sym(:fa_workspace_bytes_circulant_bwd), Csize_t, (Int64, Int64, Int64, Int64, Int64, Cint, Cint),
                N, d, dv, batchsize, W, fa_dtype(T), Cint(flags))
    ws = workspace(nws)
    rc = GC.@preserve K O Q V dK dO dQ dV l m ws begin
        ccall(sym(:fa_circulant_bwd), Cint,
                   (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid},
                    Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Int64, Int64, Cint, Cint,
                    Ptr{Cvoid}, Csize_t, Ptr{Cvoid}),
                   devptr(Q), devptr(K), devptr(V), devptr(O), devptr(dO), devptr(l), devptr(m),
                   devptr(dQ), devptr(dK), devptr(dV), N, d, dv, batchsize, W, fa_dtype(T), Cint(flags),
                   devptr(ws), length(ws), current_stream())
    end
    check(rc, "fa_circulant_bwd")
    return dQ, dK, dV
end

# ---- 2-D circulant (periodic neighbourhood) attention: the reference's todo (README.md:38-41,53).
# Q, K, V :: (X, Y, d, B); keys of query (x, y): (mod(x-p+s, X), mod(y-p+t, Y)), s, t = 0..W-1.
function circulant_fa(Q::CuArray{T, 4}, K::CuArray{T, 4}, V::CuArray{T, 4}, W::Int; flags::Integer=0) where {T}
    X, Y, d, batchsize = size(Q)
    dv = size(V, 3)
    O = similar(Q, X, Y, dv, batchsize)
    l = statarray(Q, X * Y, 1, batchsize)
    m = statarray(Q, X * Y, 1, batchsize)
    rc = GC.@preserve K O Q V l m begin
        ccall(sym(:fa_circulant2d_fwd), Cint,
                   (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid},
                    Int64, Int64, Int64, Int64, Int64, Int64, Cint, Cint, Ptr{Cvoid}),
                   devptr(Q), devptr(K), devptr(V), devptr(O), devptr(l), devptr(m),
                   X, Y, d, dv, batchsize, W, fa_dtype(T), Cint(flags), current_stream())
    end
    check(rc, "fa_circulant2d_fwd")
    return O, l, m
end

function circulant_fa_backward(Q::CuArray{T, 4}, K::CuArray{T, 4}, V::CuArray{T, 4}, O::CuArray{T, 4}, dO::CuArray{T, 4},
                               l::CuArray{Float32, 3}, m::CuArray{Float32, 3}, W::Int; flags::Integer=0) where {T}
    X, Y, d, batchsize = size(Q)
    dv = size(V, 3)
    dQ, dK, dV = similar(Q), similar(K), similar(V)
    nws = ccall(sym(:fa_workspace_bytes_circulant2d_bwd_ex), Csize_t, (Int64, Int64, Int64, Int64, Int64, Int64, Cint, Cint),
                X, Y, d, dv, batchsize, W, fa_dtype(T), Cint(flags))
    ws = workspace(nws)
    rc = GC.@preserve K O Q V dK dO dQ dV l m ws begin
        ccall(sym(:fa_circulant2d_bwd), Cint,
                   (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid},
                    Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Int64, Int64, Int64, Cint, Cint,
                    Ptr{Cvoid}, Csize_t, Ptr{Cvoid}),
                   devptr(Q), devptr(K), devptr(V), devptr(O), devptr(dO), devptr(l), devptr(m),
                   devptr(dQ), devptr(dK), devptr(dV), X, Y, d, dv, batchsize, W, fa_dtype(T), Cint(flags),
                   devptr(ws), length(ws), current_stream())
    end
    check(rc, "fa_circulant2d_bwd")
    return dQ, dK, dV
end

# host Arrays, backward
function circulant_fa_backward(Q::Array{T, 3}, K::Array{T, 3}, V::Array{T, 3}, O::Array{T, 3}, dO::Array{T, 3},
                               l::Array{Float32, 3}, m::Array{Float32, 3}, W::Int; flags::Integer=0, device::Integer=0) where T
    N, d, batchsize = size(Q)
    dv = size(V, 2)
    dQ, dK, dV = similar(Q), similar(K), similar(V)
    rc = ccall(sym(:fa_circulant_bwd_host), Cint,
               (Ptr{T}, Ptr{T}, Ptr{T}, Ptr{T}, Ptr{T}, Ptr{Float32}, Ptr{Float32}, Ptr{T}, Ptr{T}, Ptr{T},
                Int64, Int64, Int64, Int64, Int64, Cint, Cint, Cint),
               Q, K, V, O, dO, l, m, dQ, dK, dV, N, d, dv, batchsize, W, fa_dtype(T), Cint(flags), Cint(device))
    check(rc, "fa_circulant_bwd_host")
    return dQ, dK, dV
end

# Float64 arrays: computed in Float32 (see libfa.jl)
function circulant_fa!(O::AbstractArray{Float64, 3}, l, m, Q::AbstractArray{Float64, 3}, K::AbstractArray{Float64, 3},
                       V::AbstractArray{Float64, 3}, W::Int; kws...)
    O32, l32, m32 = circulant_fa(f32(Q), f32(K), f32(V), W)
    O .= O32; l .= l32; m .= m32
    return O, l, m
end
