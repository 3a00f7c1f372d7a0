# The reference's benchmark driver (bench/compare.jl:5-129: time_dense / time_windowed / time_circulant, runcompare,
# runwindow, runcirculant) on the libfa_sm100a-backed package.  Prints the reference's tables with a GPU column beside
# the CPU one and reports Threads.nthreads() / Sys.CPU_THREADS as BASELINE.md section 4 asks.  UNEXECUTED in the build
# image (no Julia runtime); the same protocol runs through Python in tools/bench_compare.py, whose numbers are the
# ones committed under profiles/.
using FlashAttention, CUDA, Printf, Test

function gpu_time(f; reps=100)
    f(); CUDA.synchronize()                       # warm-up (bench/compare.jl:18-20)
    t = CUDA.@elapsed begin
        for _ in 1:reps; f(); end
    end
    return t / reps
end

function time_dense(N, d, bs; reps=100, T=Float32)
    Q, K, V = (CUDA.randn(T, N, d, bs) for _ in 1:3)
    O = similar(Q); l = CUDA.zeros(Float32, N, 1, bs); m = similar(l)
    O1, _ = dense_dpa(Q, K, V); dense_fa!(O, l, m, Q, K, V)
    @test isapprox(Array(O1), Array(O); rtol=1e-3)                                           # bench/compare.jl:20
    return gpu_time(() -> dense_dpa(Q, K, V); reps=reps), gpu_time(() -> dense_fa!(O, l, m, Q, K, V); reps=reps)
end

function time_windowed(N, d, bs, W; stride=W, pad=0, reps=100, T=Float32)
    Q, K, V = (CUDA.randn(T, N, d, bs) for _ in 1:3)
    O1 = first(windowed_dpa(Q, K, V, W; stride=stride, pad=pad)); O2 = first(windowed_fa(Q, K, V, W; stride=stride, pad=pad))
    @test isapprox(Array(O1), Array(O2); rtol=1e-3)                                          # bench/compare.jl:47
    return gpu_time(() -> windowed_dpa(Q, K, V, W; stride=stride, pad=pad); reps=reps),
           gpu_time(() -> windowed_fa(Q, K, V, W; stride=stride, pad=pad); reps=reps)
end

function time_circulant(N, d, bs, W; reps=100, T=Float32)
    Q, K, V = (CUDA.randn(T, N, d, bs) for _ in 1:3)
    O = similar(Q); l = CUDA.zeros(Float32, N, 1, bs); m = similar(l)
    P = CUDA.zeros(T, W, N, bs)
    O1 = first(circulant_dpa!(similar(Q), P, Q, K, V, W)); circulant_fa!(O, l, m, Q, K, V, W)
    @test isapprox(Array(O1), Array(O); rtol=1e-3)                                           # bench/compare.jl:74
    return gpu_time(() -> circulant_dpa!(similar(Q), P, Q, K, V, W); reps=reps), gpu_time(() -> circulant_fa!(O, l, m, Q, K, V, W); reps=reps)
end

function runcompare(; N_range=2 .^ (8:14), d_range=(64,), bs_range=(1,), windowsize=64, reps=100, T=Float32)
    @printf("# %s, Threads.nthreads() = %d, Sys.CPU_THREADS = %d, eltype %s\n", CUDA.name(CUDA.device()), Threads.nthreads(), Sys.CPU_THREADS, T)
    @printf("%5s %5s %5s %10s %10s %10s %10s %10s %10s %10s %10s\n", "N", "d", "bs", "dense_dpa", "dense_fa", "block_dpa", "block_fa", "wind_dpa", "wind_fa", "circ_dpa", "circ_fa")
    for N in N_range, d in d_range, bs in bs_range
        tdense = time_dense(N, d, bs; reps=reps, T=T)
        tblock = time_windowed(N, d, bs, windowsize; reps=reps, T=T)
        twind = time_windowed(N, d, bs, windowsize; stride=16, reps=reps, T=T)
        tcirc = time_circulant(N, d, bs, windowsize + 1; reps=reps, T=T)
        @printf("%5d %5d %5d %10f %10f %10f %10f %10f %10f %10f %10f\n", N, d, bs, tdense..., tblock..., twind..., tcirc...)
    end
end

function runwindow(window_range=2 .^ (4:9); stride=8, N=4096, d=32, bs=1, reps=100, T=Float32)
    @printf("%5s %5s %5s %5s %10s %10s\n", "N", "d", "bs", "W", "wind_dpa", "wind_fa")
    for W in window_range
        @printf("%5d %5d %5d %5d %10f %10f\n", N, d, bs, W, time_windowed(N, d, bs, W; stride=stride, reps=reps, T=T)...)
    end
end

function runcirculant(window_range=2 .^ (4:10); N=4096, d=32, bs=1, reps=100, T=Float32)
    @printf("%5s %5s %5s %5s %10s %10s\n", "N", "d", "bs", "W", "circ_dpa", "circ_fa")
    for W in window_range
        @printf("%5d %5d %5d %5d %10f %10f\n", N, d, bs, W, time_circulant(N, d, bs, W; reps=reps, T=T)...)
    end
end

if abspath(PROGRAM_FILE) == @__FILE__
    runcompare(); runwindow(); runcirculant()
end
