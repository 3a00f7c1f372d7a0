# Golden-vector exporter for maintainers WITH a Julia runtime: runs the ORIGINAL nikopj/FlashAttention.jl (CPU, Float64)
# on seeded inputs and writes flat little-endian Float64 files + a manifest that tests/test_julia_golden.py consumes
# (tests/golden/julia/).  This is the one artefact that would pin NNlib's unfold/fold semantics (window / unwindow,
# src/utils.jl:36-54) and the circulant order, which no reference test pins and which cannot be produced in the build
# image (no Julia there).  Usage, inside a checkout of the reference:
#     julia --project=. /path/to/export_golden.jl /path/to/repo/tests/golden/julia
using FlashAttention, Random

outdir = length(ARGS) >= 1 ? ARGS[1] : "golden_julia"
mkpath(outdir)
manifest = IOBuffer()

function dump(name, x)
    open(joinpath(outdir, name * ".f64"), "w") do io; write(io, Float64.(vec(collect(x)))); end
    println(manifest, name, " ", join(size(x), "x"))
end

Random.seed!(0)
# dense (src/dense.jl), windowed 1-D/2-D/3-D incl. overlap and default padding (src/windowed.jl), circulant (src/circulant.jl)
for (tag, shape) in (("dense_n96_d16", (96, 16, 2)), ("dense_2d_8x6_d12", (8, 6, 12, 2)))
    q, k, v = randn(shape...), randn(shape...), randn(shape...)
    y, l, m = dense_fa(q, k, v)
    for (n, x) in (("q", q), ("k", k), ("v", v), ("y", y), ("l", l), ("m", m)); dump("$(tag)_$n", x); end
end
for (tag, shape, W, kws) in (("win1d_n64_w16_s4", (64, 8, 2), 16, (stride=4, pad=0)), ("win2d_20x12_w7", (20, 12, 8, 2), 7, NamedTuple()),
                             ("win3d_6x7x8_w3", (6, 7, 8, 8, 2), 3, NamedTuple()), ("win1d_n22_w5_nan", (22, 8, 2), 5, (stride=5, pad=0)))
    q, k, v = randn(shape...), randn(shape...), randn(shape...)
    y, l, m = windowed_fa(q, k, v, W; kws...)
    xw = FlashAttention.window(q, W; kws...)
    for (n, x) in (("q", q), ("k", k), ("v", v), ("y", y), ("l", l), ("m", m), ("qw", xw)); dump("$(tag)_$n", x); end
end
for (tag, N, d, W) in (("circ_n128_d16_w16", 128, 16, 16), ("circ_n256_d8_w33", 256, 8, 33))
    Q, K, V = randn(N, d, 2), randn(N, d, 2), randn(N, d, 2)
    O = similar(Q); l = similar(Q, N, 1, 2); m = similar(l)
    circulant_fa!(O, l, m, Q, K, V, W)
    for (n, x) in (("q", Q), ("k", K), ("v", V), ("y", O), ("l", l), ("m", m)); dump("$(tag)_$n", x); end
end
write(joinpath(outdir, "MANIFEST.txt"), String(take!(manifest)))
println("wrote ", outdir)
