"""Loads lib/libfa_sm100a_probe.so (`make -C flashattention.jl_b200 probe`): the product objects plus the hardware
probes of csrc/fa_tc_probe.cu.  The probes are not exported by the product library libfa_sm100a.so."""
import ctypes
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "flashattention.jl_b200", "lib", "libfa_sm100a_probe.so")
_lib = None


def probe_lib():
    global _lib
    if _lib is None:
        if not os.path.exists(PATH):
            raise ImportError(f"{PATH} missing: make -C flashattention.jl_b200 probe")
        _lib = ctypes.CDLL(PATH)
        _lib.fa_last_error_string.restype = ctypes.c_char_p
        vp, ci = ctypes.c_void_p, ctypes.c_int
        _lib.fa_debug_umma_probe.argtypes = [ci, vp, vp, vp, vp, ci, ci, ci, ci, ci, ci, ci, vp]
        _lib.fa_debug_umma_probe.restype = ci
    return _lib
