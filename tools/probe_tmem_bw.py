"""Measures tcgen05.ld / tcgen05.st throughput on one SM (fa_debug_tmem_bw): bytes per SM clock for 1, 4
and 8 warps.  Usage (GPU box): python tools/probe_tmem_bw.py"""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flashattention.jl_b200"))
import fa_sm100a as fa
sys.path.insert(0, os.path.join(ROOT, "tools"))
from _probe_lib import probe_lib
f = probe_lib().fa_debug_tmem_bw
f.restype = ctypes.c_int
f.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
out = torch.zeros(8, dtype=torch.int64, device="cuda")
for mode, name in ((0, "tcgen05.ld.32x32b.x32 (+wait per load)"), (1, "tcgen05.st.32x32b.x32 (wait per 4)")):
    for nw in (1, 4, 8):
        iters = 2000
        for _ in range(2):
            assert f(mode, nw, iters, out.data_ptr(), None) == 0
            torch.cuda.synchronize()
        clk = int(out[:nw].max())
        print(f"{name}: {nw} warps: {nw * iters * 4 * 4096 / clk:8.1f} B/clk/SM  ({clk / (iters * 4):.1f} clk per x32 op per warp)", flush=True)
