"""Phase trace of CTA 1 of tc_winx_fwd_kernel (csrc/fa_tc_winx.cu) at config 5 (64^3, W 5, stride 5, pad 3, d 64, bf16)
or config 2 x B (64x64, W 7).  Needs lib/libfa_sm100a_trace.so (make -C flashattention.jl_b200 trace).
Usage: python tools/trace_winx.py [B=8] [c5|c2]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("FA_SM100A_LIB", os.path.join(ROOT, "flashattention.jl_b200", "lib", "libfa_sm100a_trace.so"))
os.environ.setdefault("FA_WINX", "1")
buf = torch.zeros(16 * 20, dtype=torch.int64, device="cuda")
os.environ["FA_TRACE_PTR"] = str(buf.data_ptr())
sys.path.insert(0, os.path.join(ROOT, "flashattention.jl_b200"))
import fa_sm100a as fa
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
which = sys.argv[2] if len(sys.argv) > 2 else "c5"
bf = torch.bfloat16
shape, W, kw = ((64, 64, 64, 64, B), 5, dict(stride=5, pad=3)) if which == "c5" else ((64, 64, 64, B), 7, {})
q, k, v = (fa.jl_empty(shape, bf).normal_() for _ in range(3))
for _ in range(3):
    fa.windowed_fa(q, k, v, W, **kw)
torch.cuda.synchronize()
t = buf.cpu().reshape(16, 20)
names = ["QK0", "QK1", "QK2", "QK3", "S wait", "V0", "V1", "V2", "V3", "softmax", "PV", "stage O", "stores"]
print("per group: clk in", names, "| total | gap to next group")
for it in range(3, 12):
    r = [int(x) for x in t[it][:14]]
    print("  it", it, [r[i + 1] - r[i] for i in range(13)], "|", r[13] - r[0], "|", int(t[it + 1][0]) - r[13])
