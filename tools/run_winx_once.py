"""Runs windowed_fa (config 5 geometry, bf16, B given) a few times through the streamed kernel -- the command line the
ncu captures of profiles/r2* use.  Usage: FA_WINX=1 python tools/run_winx_once.py [B=8] [reps=3] [fwd|bwd]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flashattention.jl_b200"))
import fa_sm100a as fa
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mode = sys.argv[3] if len(sys.argv) > 3 else "fwd"
bf = torch.bfloat16
q, k, v, g = (fa.jl_empty((64, 64, 64, 64, B), bf).normal_() for _ in range(4))
y, l, m = fa.windowed_fa(q, k, v, 5, 5, 3)
for _ in range(reps):
    if mode == "fwd":
        fa.windowed_fa(q, k, v, 5, 5, 3)
    else:
        fa.windowed_fa_backward(q, k, v, g, l, m, 5, 5, 3)
torch.cuda.synchronize()
print("ok", fa.last_path())
