"""Times every BASELINE.json config on one B200 with CUDA events (warm-up 3, then `reps` launches)
and prints one JSON line per (config, pass): ms, TFLOP/s, algorithmic GB/s, roofline fractions.
Usage (GPU box): python tools/bench_configs.py [--only C4] [--reps 10]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flashattention.jl_b200"))
import fa_sm100a as fa  # noqa: E402

PEAK_TF, PEAK_GB = 1660.0, 6459.9
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    _p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    PEAK_TF, PEAK_GB = _p["bf16_tflops"], _p["hbm_gbs"]


def timeit(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def graph_time(fn, reps):
    """GPU-only time per call: `reps` calls captured into one CUDA graph and replayed (removes the Python
    wrapper and launch latency from the small configs)."""
    fn(); torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                fn()
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def rnd(shape, dtype):
    t = fa.jl_empty(shape, dtype)
    t.normal_()
    return t


def report(name, pss, ms, flops, nbytes, path, extra=None):
    line = {"config": name, "pass": pss, "ms": round(ms, 4), "tflops": round(flops / ms / 1e9, 2),
            "alg_gbs": round(nbytes / ms / 1e6, 1), "frac_tensor_peak": round(flops / ms / 1e9 / PEAK_TF, 4),
            "frac_hbm_peak": round(nbytes / ms / 1e6 / PEAK_GB, 4), "path": path}
    if extra:
        line.update(extra)
    print(json.dumps(line), flush=True)


def dense(name, N, d, B, dtype, reps, bwd=True):
    es = 4 if dtype == torch.float32 else 2
    q, k, v, g = (rnd((N, d, B), dtype) for _ in range(4))
    O = fa.jl_empty((N, d, B), dtype); l = fa.jl_empty((N, 1, B), torch.float32); m = fa.jl_empty((N, 1, B), torch.float32)
    ms = timeit(lambda: fa.dense_fa_(O, l, m, q, k, v), reps)
    report(name, "fwd", ms, 4.0 * N * N * d * B, 4 * N * d * B * es + 8 * N * B, fa.last_path(), {"tokens_per_s": B * N / ms * 1e3})
    if bwd:
        ms = timeit(lambda: fa.dense_fa_backward(q, k, v, O, g, l, m), max(1, reps // 3))
        report(name, "bwd", ms, 10.0 * N * N * d * B, 8 * N * d * B * es + 8 * N * B, fa.last_path())


def circulant(name, N, d, B, W, dtype, reps, bwd=False):
    es = 4 if dtype == torch.float32 else 2
    q, k, v, g = (rnd((N, d, B), dtype) for _ in range(4))
    O = fa.jl_empty((N, d, B), dtype); l = fa.jl_empty((N, 1, B), torch.float32); m = fa.jl_empty((N, 1, B), torch.float32)
    ms = timeit(lambda: fa.circulant_fa_(O, l, m, q, k, v, W), reps)
    report(name, "fwd", ms, 4.0 * N * W * d * B, 4 * N * d * B * es + 8 * N * B, fa.last_path(), {"tokens_per_s": B * N / ms * 1e3})
    if bwd:
        ms = timeit(lambda: fa.circulant_fa_backward(q, k, v, O, g, l, m, W), max(1, reps // 3))
        report(name, "bwd", ms, 10.0 * N * W * d * B, 8 * N * d * B * es + 8 * N * B, fa.last_path())


def windowed(name, spatial, d, B, W, stride, pad, dtype, reps):
    es = 4 if dtype == torch.float32 else 2
    q, k, v, g = (rnd(tuple(spatial) + (d, B), dtype) for _ in range(4))
    nw = fa.window_counts(spatial, W, stride, pad)
    L = 1
    for n in nw:
        L *= n
    WD = W ** len(spatial)
    Ntok = 1
    for s in spatial:
        Ntok *= s
    y, l, m = fa.windowed_fa(q, k, v, W, stride, pad)
    ms = timeit(lambda: fa.windowed_fa(q, k, v, W, stride, pad), reps)
    nbytes = 4 * Ntok * d * B * es + 8 * WD * L * B
    report(name, "fwd", ms, 4.0 * WD * WD * d * L * B, nbytes, fa.last_path(), {"windows": L, "tokens_per_s": B * Ntok / ms * 1e3})
    ms = timeit(lambda: fa.windowed_fa_backward(q, k, v, g, l, m, W, stride, pad), max(1, reps // 3))
    report(name, "bwd", ms, 10.0 * WD * WD * d * L * B, 8 * Ntok * d * B * es + 8 * WD * L * B, fa.last_path())
    try:   # GPU-only time (CUDA graph replay): what the kernels cost without the host wrapper
        gf = graph_time(lambda: fa.windowed_fa(q, k, v, W, stride, pad), 20)
        gb = graph_time(lambda: fa.windowed_fa_backward(q, k, v, g, l, m, W, stride, pad), 20)
        report(name, "fwd_graph", gf, 4.0 * WD * WD * d * L * B, nbytes, fa.last_path())
        report(name, "bwd_graph", gb, 10.0 * WD * WD * d * L * B, 8 * Ntok * d * B * es + 8 * WD * L * B, fa.last_path())
    except Exception as e:
        print(json.dumps({"config": name, "graph_error": repr(e)[:200]}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--reps", type=int, default=10)
    a = ap.parse_args()
    bf, f32 = torch.bfloat16, torch.float32
    runs = {
        "C1_dense_f32_N1024_d64_B4": lambda: dense("C1_dense_f32_N1024_d64_B4", 1024, 64, 4, f32, a.reps),
        "C1b_dense_bf16_N1024_d64_B4": lambda: dense("C1b_dense_bf16_N1024_d64_B4", 1024, 64, 4, bf, a.reps),
        "C2_win2d_64x64_W7_d64_B8_f32": lambda: windowed("C2_win2d_64x64_W7_d64_B8_f32", (64, 64), 64, 8, 7, 7, 3, f32, a.reps),
        "C2_win2d_64x64_W7_d64_B8_bf16": lambda: windowed("C2_win2d_64x64_W7_d64_B8_bf16", (64, 64), 64, 8, 7, 7, 3, bf, a.reps),
        "C2s_win2d_sliding_s1_bf16": lambda: windowed("C2s_win2d_sliding_s1_bf16", (64, 64), 64, 8, 7, 1, 3, bf, a.reps),
        "C3_dense_bf16_N8192_d128_B64": lambda: dense("C3_dense_bf16_N8192_d128_B64", 8192, 128, 64, bf, a.reps),
        "C3h_dense_f16_N8192_d128_B64": lambda: dense("C3h_dense_f16_N8192_d128_B64", 8192, 128, 64, torch.float16, a.reps, bwd=False),
        "C3d_dense_bf16_N8192_d64_B128": lambda: dense("C3d_dense_bf16_N8192_d64_B128", 8192, 64, 128, bf, a.reps, bwd=False),
        "C4_circ_bf16_N16384_W255_d64_B512": lambda: circulant("C4_circ_bf16_N16384_W255_d64_B512", 16384, 64, 512, 255, bf, a.reps),
        "C4b_circ_bf16_bwd_B32": lambda: circulant("C4b_circ_bf16_bwd_B32", 16384, 64, 32, 255, bf, a.reps, bwd=True),
        "C5_win3d_64c_W5_s5_p3_d64_B8_bf16": lambda: windowed("C5_win3d_64c_W5_s5_p3_d64_B8_bf16", (64, 64, 64), 64, 8, 5, 5, 3, bf, a.reps),
    }
    for name, fn in runs.items():
        if a.only and a.only not in name:
            continue
        try:
            fn()
        except Exception as e:  # keep going: one config must not hide the others
            print(json.dumps({"config": name, "error": repr(e)[:300]}), flush=True)
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
