# Port of the reference's only @testset (test/test.jl:5-21) plus the `@test fa ≈ dpa` checks of its benchmark driver
# (bench/compare.jl:20,47,74), for the libfa_sm100a-backed package.  UNEXECUTED in the build image (no Julia runtime
# there); run with `julia --project=. test/runtests.jl` on a machine with Julia, CUDA.jl and a B200.
#
# Deviations from the reference test, each stated: (1) dv == dqk in the dense_fa check -- the reference allocates
# O = similar(Q) (d columns) and reshapes with dv (src/dense.jl:11,17), so its own line 20 cannot pass with dv = 6,
# dqk = 12 (SURVEY B-3); this package supports dv != d, tested separately.  (2) Float64 arrays are computed in
# Float32 (libfa.jl), so Float64 comparisons use rtol = 1e-5, the tolerance of the exact-fp32 path.
using Test, CUDA, NNlib
using FlashAttention

@testset "FlashAttention.jl" begin
    Nq = 30; Nkv = 30; dqk = 12; dv = 6; bs = 2               # test/test.jl:6-10
    q, k, v = rand(Float64, Nq, dqk, bs), rand(Float64, Nkv, dqk, bs), rand(Float64, Nkv, dv, bs)

    # NNlib.dot_product_attention wants (features, tokens, batch): test/test.jl:13-17
    y1, α = NNlib.dot_product_attention(permutedims(q, (2, 1, 3)), permutedims(k, (2, 1, 3)), permutedims(v, (2, 1, 3)))
    y1 = permutedims(y1, (2, 1, 3))
    y2, P = dense_dpa(q, k, v)
    y3, l, m = dense_fa(q, k, v)                                # Array methods: fa_dense_fwd_host
    @test y1 ≈ y2                                               # test/test.jl:19
    @test isapprox(y3, y2; rtol=1e-5)                           # test/test.jl:20 (dv != dqk supported here)
    @test size(l) == (Nq, 1, bs) && size(m) == (Nq, 1, bs)

    @testset "CuArray, Float32: exact path 1e-5" begin
        Q, K, V = cu(Float32.(q)), cu(Float32.(k)), cu(Float32.(v))
        Y, L, M = dense_fa(Q, K, V)
        @test isapprox(Array(Y), Float32.(y2); rtol=1e-5)
        G = CUDA.randn(Float32, size(Y)...)
        dQ, dK, dV = FlashAttention.dense_fa_backward(Q, K, V, Y, G, L, M)
        @test all(isfinite, Array(dQ)) && size(dK) == size(K) && size(dV) == size(V)
    end

    @testset "bench/compare.jl checks" begin
        N, d = 1024, 64
        Q, K, V = (CUDA.randn(Float32, N, d, 1) for _ in 1:3)
        O1, _ = dense_dpa(Q, K, V); O2, _, _ = dense_fa(Q, K, V)
        @test isapprox(Array(O1), Array(O2); rtol=1e-4)                                      # bench/compare.jl:20
        for kws in ((stride=64, pad=0), (stride=16, pad=0))
            A, _ = windowed_dpa(Q, K, V, 64; kws...); B, _, _ = windowed_fa(Q, K, V, 64; kws...)
            @test isapprox(Array(A), Array(B); rtol=1e-4)                                    # bench/compare.jl:47
        end
        O = similar(Q); l = CUDA.zeros(Float32, N, 1, 1); m = similar(l)
        circulant_fa!(O, l, m, Q, K, V, 65)
        Od, _ = circulant_dpa!(similar(Q), CUDA.zeros(Float32, 65, N, 1), Q, K, V, 65)
        @test isapprox(Array(O), Array(Od); rtol=1e-4)                                       # bench/compare.jl:74
    end

    @testset "Float32 arrays on the tensor cores (via)" begin
        Q, K, V = (CUDA.randn(Float32, 1024, 64, 2) for _ in 1:3)
        Y, _, _ = dense_fa(Q, K, V); Yb, _, _ = dense_fa(Q, K, V, Float16)
        @test maximum(abs.(Array(Y) .- Array(Yb))) / maximum(abs.(Array(Y))) < 4e-3
    end

    @testset "golden vectors from the reference's C++ (tests/golden/ref_cpp_*.npz are the Python-side copy)" begin
        # the literal example of src_cpp/FlashAttention.cpp:319-356 (lambda = 1 there; the package scales by 1/sqrt(d))
        Qm = [1.2 2.3; 4.2 1.1; 2.2 2.3]; Km = [1.4 2.1; 4.6 1.0; 4.2 6.3]; Vm = [8.2 5.3; 1.2 0.1; 9.2 4.3]
        s = sqrt(2.0)
        S = Qm * Km'; P = exp.(S .- maximum(S, dims=2)); P ./= sum(P, dims=2)
        Y, _, _ = dense_fa(reshape(Qm .* s, 3, 2, 1), reshape(Km, 3, 2, 1), reshape(Vm, 3, 2, 1))
        @test isapprox(Y[:, :, 1], P * Vm; rtol=1e-5)
    end
end
