# fused_softmax / fused_softmax! -- reference src/fused_softmax.jl:1-16 (CuArray methods).
fused_softmax(S; dims=1) = fused_softmax!(similar(S), S; dims=dims)
fused_softmax!(S; dims=1) = fused_softmax!(S, S; dims=dims)

function fused_softmax!(P::CuMatrix{T}, S::CuMatrix{T}; dims=1) where T
    fused_softmax!(reshape(P, size(P)..., 1), reshape(S, size(S)..., 1); dims=dims)
    return P
end

function fused_softmax!(P::CuArray{T, 3}, S::CuArray{T, 3}; dims=1) where T
    @assert dims in (1, 2) "only softmax in dims 1 or 2 supported"       # reference :12
    M, N, B = size(S)
    rc = GC.@preserve P S begin
        ccall(sym(:fa_softmax), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Cint, Cint, Ptr{Cvoid}),
                   devptr(P), devptr(S), M, N, B, Cint(dims), fa_dtype(T), current_stream())
    end
    check(rc, "fa_softmax")
    return P
end

# CPU arrays (the reference's src/fused_softmax.jl:1-39 is a CPU implementation): round trip through the device kernel
function fused_softmax!(P::Array{T}, S::Array{T}; dims=1) where T <: Union{Float32, Float16}
    P .= Array(fused_softmax!(similar(CuArray(S)), CuArray(S); dims=dims))
    return P
end
function fused_softmax!(P::Array{Float64}, S::Array{Float64}; dims=1)
    P .= Array(fused_softmax!(CuArray(Float32.(S)); dims=dims))
    return P
end
