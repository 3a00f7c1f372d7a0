"""tcgen05.mma (kind::f16, M = 128, K = 16) throughput per SM by operand source and N: clocks per MMA and the
fraction of the 8192 flop/clk/SM math rate, for A in shared memory (SS) vs A in TMEM (TS).
Usage (GPU box): python tools/probe_umma_rate.py"""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flashattention.jl_b200"))
import fa_sm100a as fa
sys.path.insert(0, os.path.join(ROOT, "tools"))
from _probe_lib import probe_lib
f = probe_lib().fa_debug_umma_rate
f.restype = ctypes.c_int
f.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
out = torch.zeros(148, dtype=torch.int64, device="cuda")
iters = 2000
for blocks in (1,):
    for mode, name in ((0, "SS A MN-major, B MN-major"), (2, "SS A K-major,  B MN-major"), (4, "SS A MN-major, B K-major"),
                       (6, "SS A K-major,  B K-major"), (1, "TS A TMEM,     B MN-major"), (5, "TS A TMEM,     B K-major")):
        for n in (64, 128, 256):
            for _ in range(2):
                rc = f(mode, n, iters, blocks, out.data_ptr(), None)
                assert rc == 0, probe_lib().fa_last_error_string()
                torch.cuda.synchronize()
            clk = int(out[:blocks].max()) / (iters * 8)
            ideal = 128 * n * 16 * 2 / 8192
            smem_bytes = (128 * 16 * 2 if not (mode & 1) else 0) + n * 16 * 2
            print(f"{blocks:3d} CTAs  {name:28s} N={n:3d}: {clk:6.1f} clk/MMA (math {ideal:5.1f})  {100 * ideal / clk:5.1f} % of peak, smem operand bytes/clk {smem_bytes / clk:6.1f}", flush=True)
