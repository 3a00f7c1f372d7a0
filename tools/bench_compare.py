"""The reference's own benchmark protocol (bench/compare.jl:86-129: runcompare / runwindow / runcirculant) on the B200
path, printed beside the reference's logged CPU numbers (logs/compare1.txt:3-9, logs/wind_t16.txt:3-8,
logs/circ_t16.txt:3-9; Float64, seconds per call, CPU model unrecorded).  GPU: bf16 on the tcgen05 kernels (and the
exact Float32 kernels with --f32), CUDA-event time per call over `reps` calls after a warm-up that also checks fa against
the naive dpa form (bench/compare.jl:19-20,45-47,72-74); `graph_s` = the same call replayed 20x from one CUDA graph (GPU
time without the host wrapper: the bs = 1 problems of the reference's tables are launch-latency bound).  One JSON line per row.
  python tools/bench_compare.py [--reps 50] [--f32]"""
import argparse, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flashattention.jl_b200"))
import fa_sm100a as fa

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=50)
ap.add_argument("--f32", action="store_true")
a = ap.parse_args()
dt = torch.float32 if a.f32 else torch.bfloat16

# reference logs (seconds): {N: (dense_fa, block_fa, wind_fa, circ_fa)} and {W: t}
COMPARE1 = {256: (0.000671, 0.001544, 0.001963, 0.010810), 512: (0.002392, 0.001754, 0.005196, 0.009673),
            1024: (0.003877, 0.003821, 0.014640, 0.009222), 2048: (0.011694, 0.006854, 0.027008, 0.016356),
            4096: (0.028642, 0.018076, 0.057028, 0.029273), 8192: (0.092271, 0.033633, 0.136042, 0.055418),
            16384: (0.348872, 0.084757, 0.244971, 0.101374)}
WIND_T16 = {16: 0.010367, 32: 0.019180, 64: 0.053079, 128: 0.115123, 256: 0.322063, 512: 0.805048}
CIRC_T16 = {16: 0.004450, 32: 0.007519, 64: 0.015007, 128: 0.027255, 256: 0.058335, 512: 0.133951, 1024: 0.282095}


def timeit(fn):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.reps * 1e-3


def graph_time(fn, reps=20):
    """GPU time per call without the Python wrapper / launch latency: `reps` calls captured in one CUDA graph."""
    fn(); torch.cuda.synchronize()
    st_ = torch.cuda.Stream()
    with torch.cuda.stream(st_):
        fn()
        g_ = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_, stream=st_):
            for _ in range(reps):
                fn()
    torch.cuda.synchronize()
    g_.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g_.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def close(x, y, tol=2e-2):
    x, y = x.float(), y.float()
    return float((x - y).abs().max() / y.abs().max()) < tol


def qkv(N, d):
    return [fa.jl_randn((N, d, 1), s, dt) for s in range(3)]


for N, ref in COMPARE1.items():                      # runcompare(N_range = 2 .^ (8:14), d = 64, windowsize = 64)
    q, k, v = qkv(N, 64)
    if N <= 4096:
        assert close(fa.dense_fa(q, k, v)[0], fa.dense_dpa(q, k, v)[0])
        assert close(fa.windowed_fa(q, k, v, 64, stride=16, pad=0)[0], fa.windowed_dpa(q, k, v, 64, stride=16, pad=0)[0])
        assert close(fa.circulant_fa(q, k, v, 65)[0], fa.circulant_dpa(q, k, v, 65)[0])
    td = timeit(lambda: fa.dense_fa(q, k, v)); pd = fa.last_path()
    tb = timeit(lambda: fa.block_fa(q, k, v, 64)); pb = fa.last_path()
    tw = timeit(lambda: fa.windowed_fa(q, k, v, 64, stride=16, pad=0)); pw = fa.last_path()
    tc = timeit(lambda: fa.circulant_fa(q, k, v, 65)); pc = fa.last_path()
    print(json.dumps({"table": "compare1", "N": N, "d": 64, "bs": 1, "dtype": str(dt)[6:],
                      "dense_fa_s": td, "block_fa_s": tb, "wind_fa_s": tw, "circ_fa_s": tc, "paths": [pd, pb, pw, pc],
                      "graph_s": [graph_time(lambda: fa.dense_fa(q, k, v)), graph_time(lambda: fa.block_fa(q, k, v, 64)),
                                  graph_time(lambda: fa.windowed_fa(q, k, v, 64, stride=16, pad=0)), graph_time(lambda: fa.circulant_fa(q, k, v, 65))],
                      "ref_cpu_f64_s": ref, "speedup": [r / t for r, t in zip(ref, (td, tb, tw, tc))]}), flush=True)
for W, ref in WIND_T16.items():                      # runwindow(2 .^ (4:9)): N = 4096, d = 32, stride 8, pad 0
    q, k, v = qkv(4096, 32)
    assert close(fa.windowed_fa(q, k, v, W, stride=8, pad=0)[0], fa.windowed_dpa(q, k, v, W, stride=8, pad=0)[0])
    t = timeit(lambda: fa.windowed_fa(q, k, v, W, stride=8, pad=0))
    print(json.dumps({"table": "wind_t16", "N": 4096, "d": 32, "W": W, "stride": 8, "dtype": str(dt)[6:], "wind_fa_s": t,
                      "graph_s": graph_time(lambda: fa.windowed_fa(q, k, v, W, stride=8, pad=0)),
                      "path": fa.last_path(), "ref_cpu_f64_16thr_s": ref, "speedup": ref / t}), flush=True)
for W, ref in CIRC_T16.items():                      # runcirculant(2 .^ (4:10)): N = 4096, d = 32
    q, k, v = qkv(4096, 32)
    assert close(fa.circulant_fa(q, k, v, W)[0], fa.circulant_dpa(q, k, v, W)[0])
    t = timeit(lambda: fa.circulant_fa(q, k, v, W))
    print(json.dumps({"table": "circ_t16", "N": 4096, "d": 32, "W": W, "dtype": str(dt)[6:], "circ_fa_s": t,
                      "graph_s": graph_time(lambda: fa.circulant_fa(q, k, v, W)),
                      "path": fa.last_path(), "ref_cpu_f64_16thr_s": ref, "speedup": ref / t}), flush=True)
