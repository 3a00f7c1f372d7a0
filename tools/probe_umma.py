"""Diagnostics: pin the UMMA shared-memory descriptor conventions on real hardware.

Runs one 128x128 tcgen05.mma tile through `fa_debug_umma_probe` with the canonical descriptor
fields used by csrc/fa_tc_fwd.cu and with a few alternatives, and prints which variants
reproduce a float64 matmul.  Usage (GPU box): python tools/probe_umma.py
"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "flashattention.jl_b200"))
import fa_sm100a as fa  # noqa: E402
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _probe_lib import probe_lib  # noqa: E402


def run(mode, D, dtype, lbo, sbo, kstep, kbox, afmt=-1):
    torch.manual_seed(0)
    code = fa.FA_BF16 if dtype == torch.bfloat16 else fa.FA_F16
    a = torch.randn(D, 128, device="cuda").to(dtype)          # [D][tokens]  (token contiguous)
    b = torch.randn(D, 128, device="cuda").to(dtype)
    if mode == 0:
        out = torch.zeros(128, 128, device="cuda")
        want = a.double().T @ b.double()
        p = None
    else:
        pdt = dtype if afmt < 0 else (torch.bfloat16 if afmt == 1 else torch.float16)
        p = torch.rand(128, 128, device="cuda").to(pdt).float().contiguous()
        out = torch.zeros(128, D, device="cuda")
        want = p.double() @ b.double().T                           # O[i][c] = sum_j P[i][j] V[c][j]
    rc = probe_lib().fa_debug_umma_probe(mode, ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(b.data_ptr()),
                                    None if p is None else ctypes.c_void_p(p.data_ptr()),
                                    ctypes.c_void_p(out.data_ptr()), D, code, lbo, sbo, kstep, kbox, afmt, None)
    if rc != 0:
        return f"rc={rc} {probe_lib().fa_last_error_string().decode()}"
    torch.cuda.synchronize()
    err = (out.double() - want).abs().max().item() / want.abs().max().item()
    return err


def main():
    for dtype in (torch.bfloat16, torch.float16):
        for D in (64, 128):
            box = 64 * D * 2
            print(f"== {dtype} D={D}")
            variants0 = {"canonical lbo=box sbo=1024 kstep=2048": (box, 1024, 2048),
                         "swapped   lbo=1024 sbo=box kstep=2048": (1024, box, 2048),
                         "lbo=box sbo=1024 kstep=1024": (box, 1024, 1024)}
            for name, (lbo, sbo, ks) in variants0.items():
                print(f"  QK  {name:45s} rel_err={run(0, D, dtype, lbo, sbo, ks, 0)}")
            variants1 = {"canonical lbo=16 sbo=1024 kstep=32 kbox=4": (16, 1024, 32, 4),
                         "lbo=box sbo=1024 kstep=32 kbox=4": (box, 1024, 32, 4),
                         "lbo=16 sbo=1024 kstep=32 kbox=0": (16, 1024, 32, 0)}
            for name, (lbo, sbo, ks, kb) in variants1.items():
                print(f"  PV  {name:45s} rel_err={run(1, D, dtype, lbo, sbo, ks, kb)}")


if __name__ == "__main__":
    main()
