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
fa_dtype(::Type{Float32}) = FA_F32
fa_dtype(::Type{Float16}) = FA_F16
fa_dtype(::Type{Core.BFloat16}) = FA_BF16        # Julia >= 1.11 / BFloat16s.jl
fa_dtype(T) = error("FlashAttention: unsupported eltype $T (Float32, Float16, BFloat16)")

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
