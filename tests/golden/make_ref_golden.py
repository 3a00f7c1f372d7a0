"""Golden vectors produced by the REFERENCE'S OWN CODE: src_cpp/FlashAttention.cpp, unmodified, built by
oracle/ref_build/Makefile into oracle/_ref/libfa_ref_cpp.so.  /root/reference does not exist on the GPU box, so the
outputs of its OneDNaive / OneDFast / OneDNaiveBack / OneDFastBack on seeded inputs are frozen here:

    python tests/golden/make_ref_golden.py        # needs /root/reference (builds oracle/_ref first)

writes tests/golden/ref_cpp_*.npz.  Unlike the files written by make_golden.py these do not come from
oracle/fa_oracle.py: tests/test_ref_pin.py checks the oracle AND (on the GPU) the CUDA kernels against them.
Inputs: the literal 3x2 example of the reference's commented-out main() (src_cpp/FlashAttention.cpp:319-356,
lambda = 1) and seeded randn cases with lambda = 1/sqrt(d) (the Julia package's scale, src/dense.jl:43).
"""
import math
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

# src_cpp/FlashAttention.cpp:322-333
LITERAL = {
    "Q": [[1.2, 2.3], [4.2, 1.1], [2.2, 2.3]],
    "K": [[1.4, 2.1], [4.6, 1.0], [4.2, 6.3]],
    "V": [[8.2, 5.3], [1.2, 0.1], [9.2, 4.3]],
    "dO": [[0.2, 0.3], [0.2, 0.1], [0.2, 0.3]],
}


def softmax_stats(Q, K, lam):
    """P, l, m exactly as the reference's main() prepares them (src_cpp/FlashAttention.cpp:421-433)."""
    S = lam * (Q @ K.T)
    m = S.max(axis=1)
    P = np.exp(S - m[:, None])
    l = P.sum(axis=1)
    return P / l[:, None], l, m


def run_case(Q, K, V, dO, lam, cache, wsize=0):
    from oracle import ref_cpp as rc
    out = {"Q": Q, "K": K, "V": V, "dO": dO, "lam": np.float64(lam), "cache": np.int64(cache), "wsize": np.int64(wsize)}
    out["O_naive"] = rc.one_d_naive(Q, K, V, wsize, lam)
    out["O_fast"] = rc.one_d_fast(Q, K, V, cache, wsize, lam)
    if wsize == 0:
        P, l, m = softmax_stats(Q, K, lam)
        out["l"], out["m"] = l, m
        out["dQ_naive"], out["dK_naive"], out["dV_naive"] = rc.one_d_naive_back(Q, K, V, P, dO, lam)
        out["dQ_fast"], out["dK_fast"], out["dV_fast"] = rc.one_d_fast_back(Q, K, V, out["O_naive"], dO, l, m, cache, lam)
    return out


def cases():
    lit = {k: np.asfortranarray(np.array(v, np.float64)) for k, v in LITERAL.items()}
    yield "ref_cpp_literal_3x2", run_case(lit["Q"], lit["K"], lit["V"], lit["dO"], 1.0, 5)
    for name, N, d, cache, wsize, seed in (("ref_cpp_n96_d16", 96, 16, 1000, 0, 11),
                                           ("ref_cpp_n200_d64", 200, 64, 4000, 0, 12),
                                           ("ref_cpp_n128_d32_block16", 128, 32, 4000, 16, 13)):
        rng = np.random.default_rng(seed)
        Q, K, V, dO = (np.asfortranarray(rng.standard_normal((N, d)).astype(np.float32).astype(np.float64)) for _ in range(4))
        yield name, run_case(Q, K, V, dO, 1.0 / math.sqrt(d), cache, wsize)


if __name__ == "__main__":
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle", "ref_build")], check=True)
    for name, c in cases():
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **c)
        print(name, {k: getattr(v, "shape", v) for k, v in c.items()})
