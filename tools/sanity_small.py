"""One small call of every tcgen05 / SIMT kernel family (for compute-sanitizer memcheck runs).
Usage (GPU box): compute-sanitizer --tool memcheck python tools/sanity_small.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flashattention.jl_b200"))
import fa_sm100a as fa
bf, f32 = torch.bfloat16, torch.float32
r = lambda shape, dt: fa.jl_empty(shape, dt).normal_()
for dt, shape in ((bf, (200, 128, 2)), (bf, (520, 64, 1)), (f32, (100, 24, 2))):
    q, k, v, g = (r(shape, dt) for _ in range(4))
    y, l, m = fa.dense_fa(q, k, v)
    fa.dense_fa_backward(q, k, v, y, g, l, m)
    print("dense", dt, shape, fa.last_path(), flush=True)
for dt, shape, W in ((bf, (384, 64, 2), 65), (bf, (256, 128, 1), 33), (f32, (64, 8, 2), 9)):
    q, k, v, g = (r(shape, dt) for _ in range(4))
    O, l, m = fa.circulant_fa(q, k, v, W)
    fa.circulant_fa_backward(q, k, v, O, g, l, m, W)
    print("circulant", dt, shape, W, fa.last_path(), flush=True)
for dt, shape, W, kw in ((bf, (20, 12, 64, 3), 7, {}), (bf, (12, 11, 10, 64, 2), 5, dict(stride=5, pad=3)), (bf, (64, 64, 2), 16, dict(stride=4, pad=0)),
                         (f32, (6, 7, 8, 8, 2), 3, {})):
    q, k, v, g = (r(shape, dt) for _ in range(4))
    y, l, m = fa.windowed_fa(q, k, v, W, **kw)
    fa.windowed_fa_backward(q, k, v, g, l, m, W, **kw)
    print("windowed", dt, shape, W, kw, fa.last_path(), flush=True)
q, k, v, g = (r((9, 10, 16, 2), bf) for _ in range(4))
O, l, m = fa.circulant_fa(q, k, v, 5)
fa.circulant_fa_backward(q, k, v, O, g, l, m, 5)
fa.fused_softmax(r((70000, 1), f32), 1); fa.fused_softmax(r((37, 1000, 2), bf), 2)
O, l, m = fa.ring_dense_fa(*(r((256, 128, 2), bf) for _ in range(3)))
torch.cuda.synchronize()
print("all families ran", flush=True)
