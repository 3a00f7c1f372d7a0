"""Phase trace of CTA 1 of tc_win_fwd_kernel at the config-5 geometry (64^3, W = 5, stride 5, pad 3, d = 64, bf16).
Needs lib/libfa_sm100a_trace.so (make -C flashattention.jl_b200 trace).  Usage: python tools/trace_win.py [B]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("FA_SM100A_LIB", os.path.join(ROOT, "flashattention.jl_b200", "lib", "libfa_sm100a_trace.so"))
buf = torch.zeros(16 * 8, dtype=torch.int64, device="cuda")
os.environ["FA_TRACE_PTR"] = str(buf.data_ptr())
sys.path.insert(0, os.path.join(ROOT, "flashattention.jl_b200"))
import fa_sm100a as fa
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
bf = torch.bfloat16
q, k, v = (fa.jl_empty((64, 64, 64, 64, B), bf).normal_() for _ in range(3))
for _ in range(3):
    fa.windowed_fa(q, k, v, 5, 5, 3)
torch.cuda.synchronize()
t = buf.cpu().reshape(16, 8)
names = ["tables", "gather", "QK", "softmax", "PV", "stage O", "scatter"]
mode = sys.argv[2] if len(sys.argv) > 2 else "fwd"
if mode == "bwd":
    g = torch.randn_like(q)
    y, l, m = fa.windowed_fa(q, k, v, 5, 5, 3)
    buf.zero_()
    for _ in range(3):
        fa.windowed_fa_backward(q, k, v, g, l, m, 5, 5, 3)
    torch.cuda.synchronize()
    t = buf.cpu().reshape(16, 8)
    print("backward per iteration: clk in [tables, gather(4 tensors), re-encode, pass B element-wise, dQ MMA, pass C, stage+scatter x3] | total")
    for it in range(4, 12):
        r = [int(x) for x in t[it]]
        seq = [r[6], r[0], r[1], r[2], r[3], r[4], r[5], r[7]]
        print("  it", it, [seq[i + 1] - seq[i] for i in range(7)], "|", seq[7] - seq[0])
    sys.exit(0)
print("per iteration: clk in", names, "| total")
for it in range(4, 12):
    r = [int(x) for x in t[it]]
    print("  it", it, [r[i + 1] - r[i] for i in range(7)], "|", r[7] - r[0], " gap to next", int(t[it + 1][0]) - r[7])
