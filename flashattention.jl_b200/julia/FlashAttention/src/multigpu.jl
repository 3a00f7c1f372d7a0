# Multi-GPU pieces (SURVEY 8e).  The path shards over the trailing batch*head dim with no collective
# (every slice is independent, reference src/dense.jl:45); one long sequence sharded by tokens uses the
# NCCL ring of libfa_sm100a (fa_ring_dense_fwd).

"(first, count) of the contiguous batch range (0-based first) rank `rank` of `nranks` owns."
function shard_batch(B::Integer, nranks::Integer, rank::Integer)
    b = Ref{Int64}(0); c = Ref{Int64}(0)
    rc = ccall(sym(:fa_shard_batch), Cint, (Int64, Cint, Cint, Ref{Int64}, Ref{Int64}), B, nranks, rank, b, c)
    check(rc, "fa_shard_batch")
    return b[], c[]
end

"""
    ring_dense_fa(q, k, v, comm::Ptr{Cvoid}, rank, nranks) -> (y, l, m)

`dense_fa` of ONE sequence whose tokens are sharded over `nranks` GPUs: every rank passes its
`(Nl, d, B)` shard and gets its `(Nl, dv, B)` slice of the output.  `comm` is the rank's `ncclComm_t`
(e.g. `NCCL.Communicator(...).handle`); K/V blocks travel round the ring with ncclSend/ncclRecv.
"""
function ring_dense_fa(q::CuArray{T, 3}, k::CuArray{T, 3}, v::CuArray{T, 3}, comm::Ptr{Cvoid},
                       rank::Integer, nranks::Integer; flags::Integer=0) where {T}
    Nl, d, batchsize = size(q)
    dv = size(v, 2)
    O = similar(q, Nl, dv, batchsize)
    l = statarray(q, Nl, 1, batchsize)
    m = statarray(q, Nl, 1, batchsize)
    nws = ccall(sym(:fa_workspace_bytes_ring_dense_fwd), Csize_t, (Int64, Int64, Int64, Int64, Cint),
                Nl, d, dv, batchsize, fa_dtype(T))
    ws = CuArray{UInt8}(undef, max(nws, 256))
    rc = ccall(sym(:fa_ring_dense_fwd), Cint,
               (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid},
                Int64, Int64, Int64, Int64, Cint, Cint, Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}),
               devptr(q), devptr(k), devptr(v), devptr(O), devptr(l), devptr(m),
               Nl, d, dv, batchsize, fa_dtype(T), Cint(flags), comm, Cint(rank), Cint(nranks),
               devptr(ws), length(ws), current_stream())
    check(rc, "fa_ring_dense_fwd")
    CUDA.synchronize()          # the workspace must outlive the enqueued work
    return O, l, m
end

"Backward of `ring_dense_fa`: pass the shards and the `(O, l, m)` the ring forward returned."
function ring_dense_fa_backward(q::CuArray{T, 3}, k::CuArray{T, 3}, v::CuArray{T, 3}, O::CuArray{T, 3}, dO::CuArray{T, 3},
                                l::CuArray{Float32, 3}, m::CuArray{Float32, 3}, comm::Ptr{Cvoid},
                                rank::Integer, nranks::Integer; flags::Integer=0) where {T}
    Nl, d, batchsize = size(q)
    dv = size(v, 2)
    dQ, dK, dV = similar(q), similar(k), similar(v)
    nws = ccall(sym(:fa_workspace_bytes_ring_dense_bwd), Csize_t, (Int64, Int64, Int64, Int64, Cint, Cint),
                Nl, d, dv, batchsize, fa_dtype(T), Cint(flags))
    ws = CuArray{UInt8}(undef, max(nws, 256))
    rc = ccall(sym(:fa_ring_dense_bwd), Cint,
               (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid},
                Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Int64, Cint, Cint,
                Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}),
               devptr(q), devptr(k), devptr(v), devptr(O), devptr(dO), devptr(l), devptr(m),
               devptr(dQ), devptr(dK), devptr(dV), Nl, d, dv, batchsize, fa_dtype(T), Cint(flags),
               comm, Cint(rank), Cint(nranks), devptr(ws), length(ws), current_stream())
    check(rc, "fa_ring_dense_bwd")
    CUDA.synchronize()
    return dQ, dK, dV
end
