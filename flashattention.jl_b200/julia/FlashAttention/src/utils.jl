# Index/layout utilities -- same names and results as reference src/utils.jl:4-54.  The integer
# index sets come from the library's host functions (bit-exact, tests/test_abi_cpu.py); window /
# unwindow on CuArrays are the on-device unfold / fold.
circshift_index(m, s, M) = mod(m - 1 - s, M) + 1

function circulant_keys(N::Integer, M::Integer)
    keys = Matrix{Int64}(undef, M, N)
    check(ccall(sym(:fa_circulant_index), Cint, (Int64, Int64, Ptr{Int64}), N, M, keys), "fa_circulant_index")
    return keys                      # 0-based key of nz-entry w of query j
end

function cartesian_circulant(n, N, M)
    j = cld(n, M)
    return circulant_keys(N, M)[mod(n - 1, M) + 1, j] + 1, j
end

function circulant(N::Int, M::Int, Tv=Float64, Ti=Int64)
    rowval = Ti.(vec(circulant_keys(N, M)) .+ 1)
    colptr = 1 .+ M .* collect(0:N) .|> Ti
    return SparseMatrixCSC{Tv, Ti}(N, N, colptr, rowval, ones(Tv, N*M))
end

function circulant(V::AbstractMatrix{Tv}, Ti=Int64) where Tv
    M, N = size(V)
    rowval = Ti.(vec(circulant_keys(N, M)) .+ 1)
    colptr = 1 .+ M .* collect(0:N) .|> Ti
    return SparseMatrixCSC{Tv, Ti}(N, N, colptr, rowval, reshape(V, :))
end

batch_circulant(bV::AbstractArray{Tv, 3}, Ti=Int64) where Tv = blockdiag([circulant(bV[:, :, b]) for b = 1:size(bV, 3)]...)

function window(x::CuArray{T, N}, windowsize; stride=windowsize, pad=(windowsize-1)÷2) where {T, N}
    D = N - 2
    dims = Int64[size(x, i) for i in 1:D]
    d, B = size(x, N-1), size(x, N)
    L = prod((dims[i] + 2pad - windowsize) ÷ stride + 1 for i in 1:D)
    X = similar(x, windowsize^D, d, L, B)
    rc = GC.@preserve X x begin
        ccall(sym(:fa_window), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Int64}, Int64, Int64, Int64, Int64, Int64, Cint, Ptr{Cvoid}),
                   devptr(x), devptr(X), D, dims, d, B, windowsize, stride, pad, fa_dtype(T), current_stream())
    end
    check(rc, "fa_window")
    return X
end

function unwindow(X::CuArray{T, N2}, outputsize::NTuple{N}, windowsize; stride=windowsize, pad=(windowsize-1)÷2) where {T, N, N2}
    D = N - 2
    dims = Int64[outputsize[i] for i in 1:D]
    d, B = outputsize[N-1], outputsize[N]
    x = similar(X, outputsize...)
    rc = GC.@preserve X x begin
        ccall(sym(:fa_unwindow), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Int64}, Int64, Int64, Int64, Int64, Int64, Cint, Ptr{Cvoid}),
                   devptr(X), devptr(x), D, dims, d, B, windowsize, stride, pad, fa_dtype(T), current_stream())
    end
    check(rc, "fa_unwindow")
    return x
end

# CPU Arrays keep the reference's NNlib formulation (src/utils.jl:36-54) for the naive oracles
function window(x::Array{T, N}, windowsize; stride=windowsize, pad=(windowsize-1)÷2) where {T, N}
    d = size(x, N-1)
    X = NNlib.unfold(x, (ntuple(i->windowsize, N-2)..., d, 1); stride=stride, pad=pad)
    X = permutedims(X, (2, 1, 3))
    return reshape(X, windowsize^(N-2), d, :, size(x, N))
end

function unwindow(X::Array{T, N2}, outputsize::NTuple{N}, windowsize; stride=windowsize, pad=(windowsize-1)÷2) where {T, N, N2}
    d = size(X, N2-2)
    X = reshape(X, windowsize^(N-2)*d, :, size(X, N2))
    X = permutedims(X, (2, 1, 3))
    return NNlib.fold(X, outputsize, (ntuple(i->windowsize, N-2)..., d, 1); stride=stride, pad=pad)
end
