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
    rc = GC.@preserve K O Q V l m begin
        ccall(sym(:fa_dense_fwd), Cint,
                   (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid},
                    Int64, Int64, Int64, Int64, Cint, Cint, Ptr{Cvoid}),
                   devptr(Q), devptr(K), devptr(V), devptr(O), devptr(l), devptr(m),
                   N, d, dv, batchsize, fa_dtype(T), Cint(flags), current_stream())
    end
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
    ws = workspace(nws)
    rc = GC.@preserve K O Q V dK dO dQ dV l m ws begin
        ccall(sym(:fa_dense_bwd), Cint,
                   (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid},
                    Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Int64, Cint, Cint,
                    Ptr{Cvoid}, Csize_t, Ptr{Cvoid}),
                   devptr(Q), devptr(K), devptr(V), devptr(O), devptr(dO), devptr(l), devptr(m),
                   devptr(dQ), devptr(dK), devptr(dV), N, d, dv, batchsize, fa_dtype(T), Cint(flags),
                   devptr(ws), length(ws), current_stream())
    end
    check(rc, "fa_dense_bwd")
    return dQ, dK, dV
end

# host Arrays: the reference's dense_fa_backward(Q, K, V, O, dO, l, m) takes Arrays (src/dense.jl:104-111)
function dense_fa_backward(Q::Array{T, 3}, K::Array{T, 3}, V::Array{T, 3}, O::Array{T, 3}, dO::Array{T, 3},
                           l::Array{Float32, 3}, m::Array{Float32, 3}; flags::Integer=0, device::Integer=0) where T
    N, d, batchsize = size(Q)
    dv = size(V, 2)
    dQ, dK, dV = similar(Q), similar(K), similar(V)
    rc = ccall(sym(:fa_dense_bwd_host), Cint,
               (Ptr{T}, Ptr{T}, Ptr{T}, Ptr{T}, Ptr{T}, Ptr{Float32}, Ptr{Float32}, Ptr{T}, Ptr{T}, Ptr{T},
                Int64, Int64, Int64, Int64, Cint, Cint, Cint),
               Q, K, V, O, dO, l, m, dQ, dK, dV, N, d, dv, batchsize, fa_dtype(T), Cint(flags), Cint(device))
    check(rc, "fa_dense_bwd_host")
    return dQ, dK, dV
end

# Float64 arrays (the eltype of every reference test and benchmark): computed in Float32, results returned as Float64
function dense_fa(q::AbstractArray{Float64, D}, k::AbstractArray{Float64, D}, v::AbstractArray{Float64, D}) where {D}
    y, l, m = dense_fa(f32(q), f32(k), f32(v))
    return f64(y), l, m
end
function dense_fa!(O::AbstractArray{Float64, 3}, l, m, Q::AbstractArray{Float64, 3}, K::AbstractArray{Float64, 3}, V::AbstractArray{Float64, 3}; kws...)
    O32, l32, m32 = dense_fa(f32(Q), f32(K), f32(V))
    O .= O32; l .= l32; m .= m32
    return O, l, m
end

"""
    dense_fa(q, k, v; via=Core.BFloat16)   (Float32 CuArrays)

Opt-in: Float32 arrays on the tensor cores.  q, k, v are cast on the device (`fa_cast`), the tcgen05 kernels run in
`via` with Float32 outputs (FA_FLAG_OUT_F32): results of the 16-bit compute class (2e-3) for callers that hold Float32
arrays, as bench/compare.jl:8-10 does.  Without `via` Float32 arrays take the exact FFMA kernels (1e-5).
"""
function dense_fa(q::CuArray{Float32, D}, k::CuArray{Float32, D}, v::CuArray{Float32, D}, via::Type) where {D}
    d, dv, B = size(q, D-1), size(v, D-1), size(q, D)
    Q, K, V = cast(reshape(q, :, d, B), via), cast(reshape(k, :, d, B), via), cast(reshape(v, :, dv, B), via)
    N = size(Q, 1)
    O = CUDA.zeros(Float32, N, dv, B); l = CUDA.zeros(Float32, N, 1, B); m = CUDA.zeros(Float32, N, 1, B)
    rc = GC.@preserve Q K V O l m begin
        ccall(sym(:fa_dense_fwd), Cint,
              (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Int64, Cint, Cint, Ptr{Cvoid}),
              devptr(Q), devptr(K), devptr(V), devptr(O), devptr(l), devptr(m), N, d, dv, B, fa_dtype(via), Cint(FA_FLAG_OUT_F32), current_stream())
    end
    check(rc, "fa_dense_fwd")
    return reshape(O, size(q)[1:D-2]..., dv, :), l, m
end

"`T.(x)` on the device through `fa_cast` (round to nearest even)."
function cast(x::CuArray{TI}, ::Type{TO}) where {TI, TO}
    TI === TO && return x
    out = similar(x, TO)
    rc = GC.@preserve x out begin
        ccall(sym(:fa_cast), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Cint, Cint, Ptr{Cvoid}),
              devptr(x), devptr(out), length(x), fa_dtype(TI), fa_dtype(TO), current_stream())
    end
    check(rc, "fa_cast")
    return out
end
