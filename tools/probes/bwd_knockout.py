"""Which resource paces tc_bwd_kernel<128>?  Timing-only knock-outs in the -DFA_TRACE build (results are WRONG by
construction): FA_BWD_DBG bit 1 = half the ex2 replaced by moves, 2 = half the T (S / dP) MMAs, 4 = half the accumulating
MMAs.  Usage: python tools/probes/bwd_knockout.py   (needs lib/libfa_sm100a_trace.so: make trace)"""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    os.environ["FA_SM100A_LIB"] = os.path.join(ROOT, "flashattention.jl_b200", "lib", "libfa_sm100a_trace.so")
    sys.path.insert(0, os.path.join(ROOT, "flashattention.jl_b200"))
    import torch
    import fa_sm100a as fa
    bf = torch.bfloat16
    q, k, v, g = (fa.jl_empty((8192, 128, 32), bf).normal_() for _ in range(4))
    O, l, m = fa.dense_fa(q, k, v)
    for _ in range(2):
        fa.dense_fa_backward(q, k, v, O, g, l, m)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fa.dense_fa_backward(q, k, v, O, g, l, m)
    e1.record(); torch.cuda.synchronize()
    print(f"dbg={os.environ.get('FA_BWD_DBG', '0')} {sys.argv[1]:28s} {e0.elapsed_time(e1) / 5:.3f} ms", flush=True)
else:
    names = {0: "baseline", 1: "half ex2", 2: "half T MMAs", 4: "half acc MMAs", 6: "half of all MMAs", 3: "half ex2 + half T", 7: "all three"}
    for dbg, name in names.items():
        subprocess.run([sys.executable, os.path.abspath(__file__), name], env=dict(os.environ, FA_BWD_DBG=str(dbg)))
