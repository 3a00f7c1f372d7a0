# windowed_fa / block_fa -- same signatures and defaults as reference src/windowed.jl:1,3 and
# src/utils.jl:36 (stride = windowsize, pad = (windowsize-1) ÷ 2); window/unwindow are fused
# into the kernel instead of three unfolds + fold + ones-unfold-fold.
@inline block_fa(q, k, v, windowsize; pad=0) = windowed_fa(q, k, v, windowsize; stride=windowsize, pad=pad)

function windowed_fa(q::CuArray{T, N}, k::CuArray{T, N}, v::CuArray{T, N}, windowsize;
                     stride=windowsize, pad=(windowsize-1)÷2, flags::Integer=0) where {T, N}
    D = N - 2
    dims = Int64[size(q, i) for i in 1:D]
    d, dv, B = size(q, N-1), size(v, N-1), size(q, N)
    nwin = [(dims[i] + 2pad - windowsize) ÷ stride + 1 for i in 1:D]
    L, WD = prod(nwin), windowsize^D
    y = similar(q, size(q)[1:D]..., dv, B)
    l = CUDA.zeros(Float32, WD, 1, L, B)
    m = CUDA.zeros(Float32, WD, 1, L, B)
    nws = ccall(sym(:fa_workspace_bytes_windowed_fwd), Csize_t,
                (Cint, Ptr{Int64}, Int64, Int64, Int64, Int64, Int64, Int64, Cint, Cint),
                D, dims, d, dv, B, windowsize, stride, pad, fa_dtype(T), Cint(flags))
    ws = workspace(nws)
    rc = GC.@preserve k l m q v ws y begin
        ccall(sym(:fa_windowed_fwd), Cint,
                   (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid},
                    Cint, Ptr{Int64}, Int64, Int64, Int64, Int64, Int64, Int64, Cint, Cint,
                    Ptr{Cvoid}, Csize_t, Ptr{Cvoid}),
                   devptr(q), devptr(k), devptr(v), devptr(y), devptr(l), devptr(m),
                   D, dims, d, dv, B, windowsize, stride, pad, fa_dtype(T), Cint(flags),
                   devptr(ws), length(ws), current_stream())
    end
    check(rc, "fa_windowed_fwd")
    return y, l, m
end

function windowed_fa(q::Array{T, N}, k::Array{T, N}, v::Array{T, N}, windowsize;
                     stride=windowsize, pad=(windowsize-1)÷2, flags::Integer=0, device::Integer=0) where {T, N}
    D = N - 2
    dims = Int64[size(q, i) for i in 1:D]
    d, dv, B = size(q, N-1), size(v, N-1), size(q, N)
    nwin = [(dims[i] + 2pad - windowsize) ÷ stride + 1 for i in 1:D]
    L, WD = prod(nwin), windowsize^D
    y = similar(q, size(q)[1:D]..., dv, B)
    l = zeros(Float32, WD, 1, L, B)
    m = zeros(Float32, WD, 1, L, B)
    rc = ccall(sym(:fa_windowed_fwd_host), Cint,
               (Ptr{T}, Ptr{T}, Ptr{T}, Ptr{T}, Ptr{Float32}, Ptr{Float32},
                Cint, Ptr{Int64}, Int64, Int64, Int64, Int64, Int64, Int64, Cint, Cint, Cint),
               q, k, v, y, l, m, D, dims, d, dv, B, windowsize, stride, pad, fa_dtype(T), Cint(flags), Cint(device))
    check(rc, "fa_windowed_fwd_host")
    return y, l, m
end

# backward (the reference has none; SURVEY A.5.2)
function windowed_fa_backward(q::CuArray{T, N}, k::CuArray{T, N}, v::CuArray{T, N}, dy::CuArray{T, N},
                              l::CuArray{Float32, 4}, m::CuArray{Float32, 4}, windowsize;
                              stride=windowsize, pad=(windowsize-1)÷2, flags::Integer=0) where {T, N}
    D = N - 2
    dims = Int64[size(q, i) for i in 1:D]
    d, dv, B = size(q, N-1), size(v, N-1), size(q, N)
    dq, dk, dvv = similar(q), similar(k), similar(v)
    nws = ccall(sym(:fa_workspace_bytes_windowed_bwd), Csize_t,
                (Cint, Ptr{Int64}, Int64, Int64, Int64, Int64, Int64, Int64, Cint, Cint),
                D, dims, d, dv, B, windowsize, stride, pad, fa_dtype(T), Cint(flags))
    ws = workspace(nws)
    rc = GC.@preserve dk dq dvv dy k l m q v ws begin
        ccall(sym(:fa_windowed_bwd), Cint,
                   (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid},
                    Cint, Ptr{Int64}, Int64, Int64, Int64, Int64, Int64, Int64, Cint, Cint, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}),
                   devptr(q), devptr(k), devptr(v), devptr(dy), devptr(l), devptr(m), devptr(dq), devptr(dk), devptr(dvv),
                   D, dims, d, dv, B, windowsize, stride, pad, fa_dtype(T), Cint(flags), devptr(ws), length(ws), current_stream())
    end
    check(rc, "fa_windowed_bwd")
    return dq, dk, dvv
end

# host Arrays, backward
function windowed_fa_backward(q::Array{T, N}, k::Array{T, N}, v::Array{T, N}, dy::Array{T, N},
                              l::Array{Float32, 4}, m::Array{Float32, 4}, windowsize;
                              stride=windowsize, pad=(windowsize-1)÷2, flags::Integer=0, device::Integer=0) where {T, N}
    D = N - 2
    dims = Int64[size(q, i) for i in 1:D]
    d, dv, B = size(q, N-1), size(v, N-1), size(q, N)
    dq, dk, dvv = similar(q), similar(k), similar(v)
    rc = ccall(sym(:fa_windowed_bwd_host), Cint,
               (Ptr{T}, Ptr{T}, Ptr{T}, Ptr{T}, Ptr{Float32}, Ptr{Float32}, Ptr{T}, Ptr{T}, Ptr{T},
                Cint, Ptr{Int64}, Int64, Int64, Int64, Int64, Int64, Int64, Cint, Cint, Cint),
               q, k, v, dy, l, m, dq, dk, dvv, D, dims, d, dv, B, windowsize, stride, pad, fa_dtype(T), Cint(flags), Cint(device))
    check(rc, "fa_windowed_bwd_host")
    return dq, dk, dvv
end

# Float64 arrays: computed in Float32 (see libfa.jl)
function windowed_fa(q::AbstractArray{Float64, N}, k::AbstractArray{Float64, N}, v::AbstractArray{Float64, N}, windowsize; kws...) where {N}
    y, l, m = windowed_fa(f32(q), f32(k), f32(v), windowsize; kws...)
    return f64(y), l, m
end
