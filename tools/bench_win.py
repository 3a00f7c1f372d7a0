"""Times windowed_fa forward (+ backward) at config 5 / config 2 / one 256^3 volume, CUDA events, for A/B runs of the
windowed kernels (FA_WINX=0/1 etc. are read once per process: run it once per setting).
  python tools/bench_win.py [B5=64] [reps=10]"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flashattention.jl_b200"))
import fa_sm100a as fa

B5 = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
bf = torch.bfloat16
peak = 6459.9
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timeit(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


out = {"FA_WINX": os.environ.get("FA_WINX"), "FA_WINXB": os.environ.get("FA_WINXB")}
for name, shape, W, kw, bwd in (("C5", (64, 64, 64, 64, B5), 5, dict(stride=5, pad=3), True),
                                ("C5_default_pad", (64, 64, 64, 64, max(1, B5 // 4)), 5, {}, False),
                                ("C2x64", (64, 64, 64, 512), 7, {}, True),
                                ("V256", (256, 256, 256, 64, 1), 5, dict(stride=5, pad=3), True)):
    q, k, v, g = (fa.jl_empty(shape, bf, "cuda").normal_() for _ in range(4))
    y, l, m = fa.windowed_fa(q, k, v, W, **kw)
    tf = timeit(lambda: fa.windowed_fa(q, k, v, W, **kw), reps)
    ntok = 1
    for s in shape[:-2]:
        ntok *= s
    byf = 4 * ntok * 64 * 2 * shape[-1] + 8 * l.numel()
    rec = {"fwd_ms": round(tf, 4), "fwd_frac_hbm": round(byf / tf / 1e6 / peak, 4)}
    if bwd:
        tb = timeit(lambda: fa.windowed_fa_backward(q, k, v, g, l, m, W, **kw), max(3, reps // 2))
        byb = 7 * ntok * 64 * 2 * shape[-1] + 8 * l.numel()
        rec.update({"bwd_ms": round(tb, 4), "bwd_frac_hbm": round(byb / tb / 1e6 / peak, 4)})
    out[name] = rec
    del q, k, v, g, y, l, m
    torch.cuda.empty_cache()
print(json.dumps(out))
