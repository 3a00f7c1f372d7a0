"""Per-step clock trace of tc_fwd_kernel (CTA (1,0), steps 32..95): issuer and softmax events.
Needs lib/libfa_sm100a_trace.so (built with -DFA_TRACE).  Usage: FA_SM100A_LIB=.../libfa_sm100a_trace.so python tools/trace_fwd.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("FA_SM100A_LIB", os.path.join(ROOT, "flashattention.jl_b200", "lib", "libfa_sm100a_trace.so"))
buf = torch.zeros(4 * 64 * 8, dtype=torch.int64, device="cuda")
os.environ["FA_TRACE_PTR"] = str(buf.data_ptr())
sys.path.insert(0, os.path.join(ROOT, "flashattention.jl_b200"))
import fa_sm100a as fa
bf = torch.bfloat16
q, k, v = (fa.jl_empty((8192, 128, 64), bf).normal_() for _ in range(3))
for _ in range(3):
    fa.dense_fa(q, k, v)
torch.cuda.synchronize()
t = buf.cpu().reshape(4, 64, 8)
base = int(t[t > 0].min())
names = {0: "issuer t0", 1: "issuer t1", 2: "softmax t0", 3: "softmax t1"}
ev = {0: ["pv_wait_start", "pv_p_ready", "qk_wait_start", "qk_k_ready", "qk_issued"],
      2: ["step_start", "max_done", "exp_st_issued", "p_published", "token_held", "s_next_in_regs", "half0_done", "midfetch_issued"]}
for role in range(4):
    e = ev[0] if role < 2 else ev[2]
    print(names[role], e)
    for s in range(8, 20):
        row = [int(t[role, s, i]) - base if t[role, s, i] > 0 else -1 for i in range(len(e))]
        extra = ""
        if role >= 2:   # softmax: max | wait token | half 0 | mid fetch | half 1 | publish | tail
            r = row
            extra = f"  max {r[1]-r[0]} token_wait {r[4]-r[1]} half0 {r[6]-r[4]} midfetch {r[7]-r[6]} half1 {r[2]-r[7]} publish {r[3]-r[2]} regs {r[5]-r[3]}"
        print("  step", 32 + s, row, " d_step", int(t[role, s, 0] - t[role, s - 1, 0]), extra)
