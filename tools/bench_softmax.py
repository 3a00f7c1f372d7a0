"""GPU softmax benchmark with the protocol of the reference's bench/softmax.jl:8-78: vector softmax for
N = 2^10..2^16 and column softmax of (M=1024, N) matrices, warm-up then the mean of `reps` timed calls,
`fused_softmax!` next to the library softmax (NNlib.softmax! there, torch.softmax here), results checked
against each other first (`@test U2 ≈ U3`, bench/softmax.jl:21-22).  Adds the HBM roofline fraction on the
ALGORITHMIC bytes (one read + one write of the array per call, whatever the kernel executes), and the reference's logged
table (logs/sm_cuda.txt:3-8: column softmax, M = 256 .. 8192, N = 65536, Float32; GPU model unrecorded there).
Usage (GPU box): python tools/bench_softmax.py [--reps 100]"""
import argparse, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flashattention.jl_b200"))
import fa_sm100a as fa

PEAK_GB = 6459.9
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    PEAK_GB = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]


def elapsed(fn, reps):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):                      # CUDA.@elapsed CUDA.@sync per call, summed (bench/softmax.jl:27-31)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / reps


def run(shape, dims, reps, dtype=torch.float32, ref_ms=None):
    V = fa.jl_empty(shape, dtype).uniform_()
    U2 = fa.jl_empty(shape, dtype)
    ref = torch.softmax(V.float(), dim=dims - 1).to(dtype)
    fa.fused_softmax_(U2, V, dims)
    assert torch.allclose(U2.float(), ref.float(), rtol=1e-5 if dtype == torch.float32 else 2e-2, atol=1e-7)
    t_fused = elapsed(lambda: fa.fused_softmax_(U2, V, dims), reps)
    t_lib = elapsed(lambda: torch.softmax(V, dim=dims - 1), reps)
    nbytes = V.numel() * V.element_size()
    print(json.dumps({"shape": list(shape), "dims": dims, "dtype": str(dtype).split(".")[-1], "fused_ms": round(t_fused, 5),
                      "torch_ms": round(t_lib, 5), "fused_alg_gbs": round(2 * nbytes / t_fused / 1e6, 1),
                      "fused_frac_hbm_peak": round(2 * nbytes / t_fused / 1e6 / PEAK_GB, 4), "ref_logged_fused_ms": ref_ms}), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=100)
    a = ap.parse_args()
    for e in range(10, 17):                       # run_vec_softmax: N_range = 2 .^ (10:16)
        run((2 ** e, 1), 1, a.reps)
    for e in range(10, 17):                       # run_col_softmax: M = 1024, N_range = 2 .^ (10:16)
        run((1024, 2 ** e), 1, a.reps)
    for M, ref in ((256, 1.329), (512, 1.456), (1024, 1.843), (2048, 2.893), (4096, 4.958), (8192, 8.687)):
        run((M, 65536), 1, max(5, a.reps // 5), ref_ms=ref)     # logs/sm_cuda.txt:3-8 (their fused kernel, ms)
    for e in (12, 16):                            # the dims = 2 variant and a 16-bit type
        run((1024, 2 ** e), 2, a.reps)
        run((1024, 2 ** e), 1, a.reps, torch.bfloat16)
