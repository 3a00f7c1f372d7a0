"""Runs one hot-path case a few times (for ncu / quick timing).
Usage (GPU box): python tools/prof_case.py {dense_fwd|dense_bwd|circ_fwd|circ_bwd|win3d_fwd|win3d_bwd|win2d_fwd|win2d_bwd} [reps] [B]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flashattention.jl_b200"))
import fa_sm100a as fa

case = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
bf = torch.bfloat16


def rnd(shape):
    t = fa.jl_empty(shape, bf)
    t.normal_()
    return t


def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(case, "ms/iter", e0.elapsed_time(e1) / reps, "path", fa.last_path(), flush=True)


if case.startswith("dense"):
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 16
    q, k, v, g = (rnd((8192, 128, B)) for _ in range(4))
    O, l, m = fa.dense_fa(q, k, v)
    timed((lambda: fa.dense_fa(q, k, v)) if case == "dense_fwd" else (lambda: fa.dense_fa_backward(q, k, v, O, g, l, m)))
elif case.startswith("circ"):
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    q, k, v, g = (rnd((16384, 64, B)) for _ in range(4))
    O, l, m = fa.circulant_fa(q, k, v, 255)
    timed((lambda: fa.circulant_fa(q, k, v, 255)) if case == "circ_fwd" else (lambda: fa.circulant_fa_backward(q, k, v, O, g, l, m, 255)))
else:
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    if case.startswith("win3d"):
        shape, W, stride, pad = (64, 64, 64, 64, B), 5, 5, 3
    else:
        shape, W, stride, pad = (64, 64, 64, B), 7, 7, 3
    q, k, v, g = (rnd(shape) for _ in range(4))
    y, l, m = fa.windowed_fa(q, k, v, W, stride, pad)
    timed((lambda: fa.windowed_fa(q, k, v, W, stride, pad)) if case.endswith("fwd") else (lambda: fa.windowed_fa_backward(q, k, v, g, l, m, W, stride, pad)))
