"""Clock trace of one CTA of tc_band_kernel (config 4: N=16384, W=255, d=64) from the middle of the grid.
Needs lib/libfa_sm100a_trace.so (make -C flashattention.jl_b200 trace).
Usage: python tools/trace_band.py [B]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("FA_SM100A_LIB", os.path.join(ROOT, "flashattention.jl_b200", "lib", "libfa_sm100a_trace.so"))
buf = torch.zeros(2 * 128, dtype=torch.int64, device="cuda")
os.environ["FA_TRACE_PTR"] = str(buf.data_ptr())
sys.path.insert(0, os.path.join(ROOT, "flashattention.jl_b200"))
import fa_sm100a as fa
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
bf = torch.bfloat16
q, k, v = (fa.jl_empty((16384, 64, B), bf).normal_() for _ in range(3))
for _ in range(3):
    fa.circulant_fa(q, k, v, 255)
torch.cuda.synchronize()
t = buf.cpu().reshape(2, 16, 8)
base = int(t[0, 15, 0])
rel = lambda x: int(x) - base if x > 0 else -1
print("CTA: entry 0, set-up done", rel(t[0, 15, 1]), "Q landed", rel(t[0, 15, 2]), "all MMAs done", rel(t[1, 15, 0]),
      "epilogue stores issued", rel(t[1, 15, 1]))
print("issuer  [K ready, QK issued (after commits), P seen, PV issued (after commit), QK MMAs handed over, V ready, PV MMAs handed over]")
for j in range(8):
    if t[0, j, 0] > 0:
        print("  step", j, [rel(t[0, j, i]) for i in range(7)])
print("softmax warp 4 [S seen, max done, rescale done, exps done, P published]")
for j in range(8):
    if t[1, j, 0] > 0:
        print("  step", j, [rel(t[1, j, i]) for i in range(5)])
