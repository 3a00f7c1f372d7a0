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
    ws = workspace(nws)
    rc = GC.@preserve O k l m q v ws begin
        ccall(sym(:fa_ring_dense_fwd), Cint,
                   (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid},
                    Int64, Int64, Int64, Int64, Cint, Cint, Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}),
                   devptr(q), devptr(k), devptr(v), devptr(O), devptr(l), devptr(m),
                   Nl, d, dv, batchsize, fa_dtype(T), Cint(flags), comm, Cint(rank), Cint(nranks),
                   devptr(ws), length(ws), current_stream())
    end
    check(rc, "fa_ring_dense_fwd")
    CUDA.synchronize()          # the workspace must outlive the enqueued work
    return O, l, m
end

"`ring_dense_fa` on `(spatial..., d, B)` shards (e.g. a 3-D volume cut along its slowest spatial dim): flattened like `dense_fa` (src/dense.jl:1-19)."
function ring_dense_fa(q::CuArray{T, N}, k::CuArray{T, N}, v::CuArray{T, N}, comm::Ptr{Cvoid},
                       rank::Integer, nranks::Integer; flags::Integer=0) where {T, N}
    sz = size(q)[1:N-2]
    r3(x) = reshape(x, :, size(x, N-1), size(x, N))
    O, l, m = ring_dense_fa(r3(q), r3(k), r3(v), comm, rank, nranks; flags=flags)
    return reshape(O, sz..., size(O, 2), size(O, 3)), l, m
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
    ws = workspace(nws)
    rc = GC.@preserve O dK dO dQ dV k l m q v ws begin
        ccall(sym(:fa_ring_dense_bwd), Cint,
                   (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid},
                    Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Int64, Cint, Cint,
                    Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}),
                   devptr(q), devptr(k), devptr(v), devptr(O), devptr(dO), devptr(l), devptr(m),
                   devptr(dQ), devptr(dK), devptr(dV), Nl, d, dv, batchsize, fa_dtype(T), Cint(flags),
                   comm, Cint(rank), Cint(nranks), devptr(ws), length(ws), current_stream())
    end
    check(rc, "fa_ring_dense_bwd")
    CUDA.synchronize()
    return dQ, dK, dV
end

# ---- one volume over several GPUs: windowed attention with non-overlapping windows (stride >= W) splits into
# slabs along the slowest spatial dim on window boundaries; no exchange (SURVEY 8e).

"Slab of rank `rank` (0-based): token planes `plane_lo:plane_hi-1` (0-based, slowest spatial dim), windows `win_lo:win_hi-1`, `pad_lo` zero planes in front."
struct SlabPlan
    plane_lo::Int64
    plane_hi::Int64
    win_lo::Int64
    win_hi::Int64
    pad_lo::Int64
end

function windowed_slab_plan(spatial, windowsize; stride=windowsize, pad=(windowsize-1)÷2, rank::Integer=0, nranks::Integer=1)
    dims = Int64[spatial...]
    plan = zeros(Int64, 5)
    rc = ccall(sym(:fa_windowed_slab_plan), Cint, (Cint, Ptr{Int64}, Int64, Int64, Int64, Cint, Cint, Ptr{Int64}),
               length(dims), dims, windowsize, stride, pad, rank, nranks, plan)
    check(rc, "fa_windowed_slab_plan")
    return SlabPlan(plan...)
end

"The planes of `x :: (s.., d, B)` that `plan` assigns to its rank (a contiguous copy)."
slab_planes(x::AbstractArray{T, N}, plan::SlabPlan) where {T, N} =
    copy(selectdim(x, N - 2, plan.plane_lo+1:plan.plane_hi))

"""
    windowed_fa_slab(q, k, v, windowsize, plan; stride, pad) -> (y, l, m)

`windowed_fa` on one slab of a volume cut by `windowed_slab_plan`: `y` for the slab's planes, `l, m` for its
windows (a contiguous range of the volume's window index) -- the same bits the one-GPU call produces there.
"""
function windowed_fa_slab(q::CuArray{T, N}, k::CuArray{T, N}, v::CuArray{T, N}, windowsize, plan::SlabPlan;
                          stride=windowsize, pad=(windowsize-1)÷2, flags::Integer=0) where {T, N}
    D = N - 2
    dims = Int64[size(q, i) for i in 1:D]
    d, dv, B = size(q, N-1), size(v, N-1), size(q, N)
    nw = plan.win_hi - plan.win_lo
    nwin = [(dims[i] + 2pad - windowsize) ÷ stride + 1 for i in 1:D-1]
    L, WD = prod(nwin; init=1) * nw, windowsize^D
    y = similar(q, size(q)[1:D]..., dv, B)
    l = CUDA.zeros(Float32, WD, 1, L, B)
    m = CUDA.zeros(Float32, WD, 1, L, B)
    rc = GC.@preserve k l m q v y begin
        ccall(sym(:fa_windowed_slab_fwd), Cint,
                   (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid},
                    Cint, Ptr{Int64}, Int64, Int64, Int64, Int64, Int64, Int64, Int64, Int64, Cint, Cint, Ptr{Cvoid}),
                   devptr(q), devptr(k), devptr(v), devptr(y), devptr(l), devptr(m),
                   D, dims, d, dv, B, windowsize, stride, pad, plan.pad_lo, nw, fa_dtype(T), Cint(flags), current_stream())
    end
    check(rc, "fa_windowed_slab_fwd")
    return y, l, m
end

function windowed_fa_slab_backward(q::CuArray{T, N}, k::CuArray{T, N}, v::CuArray{T, N}, dy::CuArray{T, N},
                                   l::CuArray{Float32, 4}, m::CuArray{Float32, 4}, windowsize, plan::SlabPlan;
                                   stride=windowsize, pad=(windowsize-1)÷2, flags::Integer=0) where {T, N}
    D = N - 2
    dims = Int64[size(q, i) for i in 1:D]
    d, dv, B = size(q, N-1), size(v, N-1), size(q, N)
    nw = plan.win_hi - plan.win_lo
    dq, dk, dvv = similar(q), similar(k), similar(v)
    nws = ccall(sym(:fa_workspace_bytes_windowed_slab_bwd), Csize_t,
                (Cint, Ptr{Int64}, Int64, Int64, Int64, Int64, Int64, Int64, Int64, Int64),
                D, dims, d, dv, B, windowsize, stride, pad, plan.pad_lo, nw)
    ws = workspace(nws)
    rc = GC.@preserve dk dq dvv dy k l m q v ws begin
        ccall(sym(:fa_windowed_slab_bwd), Cint,
                   (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid},
                    Cint, Ptr{Int64}, Int64, Int64, Int64, Int64, Int64, Int64, Int64, Int64, Cint, Cint, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}),
                   devptr(q), devptr(k), devptr(v), devptr(dy), devptr(l), devptr(m), devptr(dq), devptr(dk), devptr(dvv),
                   D, dims, d, dv, B, windowsize, stride, pad, plan.pad_lo, nw, fa_dtype(T), Cint(flags),
                   devptr(ws), length(ws), current_stream())
    end
    check(rc, "fa_windowed_slab_bwd")
    return dq, dk, dvv
end

# ---- halo splits (SURVEY 8e): the neighbour exchanges are the caller's (NCCL.jl Send/Recv of plane ranges); the library
# supplies the plan, the un-normalised slab forward and the final division -- see fa_sm100a/halo.py for the complete
# sequence of calls (exchange q/k/v halo planes in, fa_windowed_slab_fwd_sums, exchange partial sums out, fa_window_divide)
"`(own_lo, own_hi, ext_hi, win_lo, win_hi, pad_lo)` (0-based, half-open) of rank `rank` for ONE volume with overlapping windows."
function windowed_halo_plan(dims::Vector{Int64}, windowsize::Integer, stride::Integer, pad::Integer, rank::Integer, nranks::Integer)
    plan = zeros(Int64, 6)
    rc = ccall(sym(:fa_windowed_halo_plan), Cint, (Cint, Ptr{Int64}, Int64, Int64, Int64, Cint, Cint, Ptr{Int64}),
               length(dims), dims, windowsize, stride, pad, rank, nranks, plan)
    check(rc, "fa_windowed_halo_plan")
    return Tuple(plan)
end
