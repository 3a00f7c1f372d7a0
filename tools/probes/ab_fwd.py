"""Time + check the dense forward at N=8192, d=128 (B from argv, default 64) in THIS process' environment.
Usage: FA_FWD_MAXAHEAD=0 python tools/probes/ab_fwd.py [B] [label]   (A/B = several processes inside one gpurun call)"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "flashattention.jl_b200"))
import fa_sm100a as fa
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
label = sys.argv[2] if len(sys.argv) > 2 else ""
bf = torch.bfloat16
torch.manual_seed(0)
q, k, v = (fa.jl_empty((8192, 128, B), bf).normal_() for _ in range(3))
O, l, m = fa.jl_empty((8192, 128, B), bf), fa.jl_empty((8192, 1, B), torch.float32), fa.jl_empty((8192, 1, B), torch.float32)
for _ in range(3):
    fa.dense_fa_(O, l, m, q, k, v)
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fa.dense_fa_(O, l, m, q, k, v)
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 10)
# check batch 0 against torch SDPA in fp32
qq, kk, vv = (x[:, :, 0].float() for x in (q, k, v))            # (N, d)
ref = torch.softmax((qq @ kk.T) / (128 ** 0.5), dim=-1) @ vv
err = ((O[:, :, 0].float() - ref).abs().max() / ref.abs().max()).item()
print(f"{label:24s} B={B} {best:.3f} ms  {4*8192*8192*128*B/best/1e9:.0f} TFLOP/s  max-rel-err {err:.2e}", flush=True)
