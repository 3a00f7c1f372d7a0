# dense_fa / dense_fa! / dense_fa_backward -- same signatures as reference src/dense.jl:1,21,104.
function dense_fa(q::AbstractArray{T, D}, k::AbstractArray{T, D}, v::AbstractArray{T, D}) where {T, D}
    d  = size(q, D-1)
    dv = size(v, D-1)
    batchsize = size(q, D)
    Q = reshape(q, :, d, batchsize)                 # reference src/dense.jl:6-8
    K = reshape(k, :, d, batchsize)
    V = reshape(v, :, dv, batchsize)
    N = size(Q, 1)
    O = similar(Q, N, dv, batchsize)                # reference allocates d columns (:11); dv is what :17 needs
    l = statarray(Q, N, 1, batchsize)
    m = statarray(Q, N, 1, batchsize)
    dense_fa!(O, l, m, Q, K, V)
    y = reshape(O, size(q)[1:D-2]..., dv, :)
    return y, l, m
end

function dense_fa!(O::CuArray{T, 3}, l::CuArray{Float32, 3}, m::CuArray{Float32, 3},
                   Q::CuArray{T, 3}, K::CuArray{T, 3}, V::CuArray{T, 3}; flags::Integer=0) where {T}
    N, d, batchsize = size(Q)
    dv = size(V, 2)
    rc = ccall(sym(:fa_dense_fwd), Cint,
               (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid},
                Int64, Int64, Int64, Int64, Cint, Cint, Ptr{Cvoid}),
               devptr(Q), devptr(K), devptr(V), devptr(O), devptr(l), devptr(m),
               N, d, dv, batchsize, fa_dtype(T), Cint(flags), current_stream())
    check(rc, "fa_dense_fwd")
    return O, l, m
end

# host Arrays (what the reference's callers pass): H2D + kernels + D2H inside the library
function dense_fa!(O::Array{T, 3}, l::Array{Float32, 3}, m::Array{Float32, 3},
                   Q::Array{T, 3}, K::Array{T, 3}, V::Array{T, 3}; flags::Integer=0, device::Integer=0) where {T}
    N, d, batchsize = size(Q)
    dv = size(V, 2)
    rc = ccall(sym(:fa_dense_fwd_host), Cint,
               (Ptr{T}, Ptr{T}, Ptr{T}, Ptr{T}, Ptr{Float32}, Ptr{Float32},
                Int64, Int64, Int64, Int64, Cint, Cint, Cint),
               Q, K, V, O, l, m, N, d, dv, batchsize, fa_dtype(T), Cint(flags), Cint(device))
    check(rc, "fa_dense_fwd_host")
    return O, l, m
end

function dense_fa_backward(Q::CuArray{T, 3}, K::CuArray{T, 3}, V::CuArray{T, 3}, O::CuArray{T, 3},
                           dO::CuArray{T, 3}, l::CuArray{Float32, 3}, m::CuArray{Float32, 3}; flags::Integer=0) where T
    N, d, batchsize = size(Q)
    dv = size(V, 2)
    dQ, dK, dV = similar(Q), similar(K), similar(V)
    nws = ccall(sym(:fa_workspace_bytes_dense_bwd), Csize_t, (Int64, Int64, Int64, Int64, Cint, Cint),
                N, d, dv, batchsize, fa_dtype(T), Cint(flags))
    ws = CuArray{UInt8}(undef, max(nws, 256))
    rc = ccall(sym(:fa_dense_bwd), Cint,
               (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid},
                Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Int64, Cint, Cint,
                Ptr{Cvoid}, Csize_t, Ptr{Cvoid}),
               devptr(Q), devptr(K), devptr(V), devptr(O), devptr(dO), devptr(l), devptr(m),
               devptr(dQ), devptr(dK), devptr(dV), N, d, dv, batchsize, fa_dtype(T), Cint(flags),
               devptr(ws), length(ws), current_stream())
    check(rc, "fa_dense_bwd")
    return dQ, dK, dV
end
