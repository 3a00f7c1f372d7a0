# Library handle and calling conventions (include/fa_sm100a.h).
const LIBFA = Ref{Ptr{Cvoid}}(C_NULL)

function libfa()
    if LIBFA[] == C_NULL
        path = get(ENV, "FA_SM100A_LIB", joinpath(@__DIR__, "..", "..", "..", "lib", "libfa_sm100a.so"))
        LIBFA[] = Libdl.dlopen(path)          # throws if missing: there is no fallback implementation
    end
    return LIBFA[]
end
sym(name::Symbol) = Libdl.dlsym(libfa(), name)

const FA_F32, FA_F16, FA_BF16 = Cint(0), Cint(1), Cint(2)
# flags (include/fa_sm100a.h)
const FA_FLAG_FORCE_SIMT, FA_FLAG_BF16_INTERNALS, FA_FLAG_OUT_F32, FA_FLAG_HOST_NO_REGISTER = 1, 2, 4, 8
fa_dtype(T) = error("FlashAttention: unsupported eltype $T (Float32, Float16, BFloat16; Float64 arrays are converted by the wrappers)")
fa_dtype(::Type{Float32}) = FA_F32
fa_dtype(::Type{Float16}) = FA_F16
# Core.BFloat16 exists from Julia 1.11 on; on older versions the method is simply not defined (an unconditional
# `Core.BFloat16` would throw UndefVarError while the module loads and break every eltype).  BFloat16s.jl's type, when
# that package is loaded by the caller, is registered with `FlashAttention.register_bfloat16(BFloat16s.BFloat16)`.
@static if isdefined(Core, :BFloat16)
    fa_dtype(::Type{Core.BFloat16}) = FA_BF16
end
register_bfloat16(T::Type) = (@eval fa_dtype(::Type{$T}) = FA_BF16; nothing)

# Float64 callers (every test and benchmark of the reference uses Float64: test/test.jl:6-12, bench/compare.jl:8-10).
# The library computes in Float32 (exact FFMA path, 1e-5) or 16-bit tensor-core arithmetic, not in Float64: Float64
# arrays are converted to Float32 on the way in and the results back to Float64 -- a documented deviation.
f32(x::AbstractArray{Float64}) = Float32.(x)
f64(x::AbstractArray{Float32}) = Float64.(x)

function check(rc::Cint, what)
    rc == 0 && return nothing
    msg = unsafe_string(ccall(sym(:fa_last_error_string), Cstring, ()))
    error("$what failed (status $rc): $msg")
end

# CUDA.jl and the CUDA runtime API share the primary context, so the task-local stream handle
# can be handed to the library as a cudaStream_t.
current_stream() = Ptr{Cvoid}(UInt(CUDA.stream().handle))
devptr(x::CuArray) = Ptr{Cvoid}(UInt(pointer(x)))
statarray(Q::CuArray, dims...) = CUDA.zeros(Float32, dims...)     # l, m are Float32 (H8)
statarray(Q::Array, dims...) = zeros(Float32, dims...)

# Device workspace for one call.  The ccall only ENQUEUES work, so every array whose raw pointer crossed the boundary
# must stay rooted until the ccall has returned (GC.@preserve in the wrappers); after that the workspace may die:
# CUDA.jl returns pool memory in stream order on the task-local stream, which is the stream the library launched on.
workspace(nbytes::Integer) = CuArray{UInt8}(undef, max(Int(nbytes), 256))
