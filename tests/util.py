"""Shared helpers for the parity tests."""
import math

import numpy as np
import torch

# BASELINE.md section 5 / BASELINE.json north_star: max-abs error relative to max-abs of the oracle
TOL_FP32 = 1e-5      # exact-fp32 (TF32-off FFMA) path
TOL_16 = 2e-3        # bf16 / fp16 compute with fp32 accumulation


def randn_np(shape, seed, round_to=None):
    """Seeded Float32 randn in Julia shape / Fortran order; optionally rounded so the values are
    exactly representable in ``round_to`` (then GPU and oracle really see the same inputs)."""
    x = np.random.default_rng(seed).standard_normal(shape).astype(np.float32)
    if round_to is not None and round_to != torch.float32:
        x = torch.from_numpy(x).to(round_to).float().numpy()
    return np.asfortranarray(x)


def to_dev(x, dtype=torch.float32, device="cuda"):
    import fa_sm100a as fa
    return fa.jl_array(x, dtype=dtype, device=device)


def to_np(t):
    return np.asfortranarray(t.detach().float().cpu().numpy().astype(np.float64))


# round-to-nearest half-ulp (relative) of bf16 STORAGE: bf16 keeps 8 significand bits, so a result stored in bf16
# can be off by 2^-8 = 3.9e-3 of its own magnitude no matter how it was computed -- more than the 2e-3 compute
# tolerance.  The two are kept apart here, and the compute error alone is PROVEN under 2e-3 with no allowance by the
# FA_FLAG_OUT_F32 legs (tests/test_gpu_parity_r2.py: same kernels, fp32 accumulators stored unrounded).
# fp16 storage (2^-11 = 4.9e-4) gets NO allowance: fp16 results must meet 2e-3 as stored.
STORAGE_HALF_ULP = {torch.bfloat16: 2.0 ** -8}
# `exact_math=True` is for a different claim -- "the exact-fp32 kernels computed this, only the store rounded it"
# (threshold 1e-5) -- where the fp16 half-ulp must be discounted too or the test would measure the dtype, not the math.
EXACT_MATH_HALF_ULP = {torch.bfloat16: 2.0 ** -8, torch.float16: 2.0 ** -11}


def rel_err(got, want, storage=None, want_rounded=False, exact_math=False):
    """max|got-want| / max|want| over the finite entries; NaN patterns must coincide.

    With ``storage`` (a 16-bit torch dtype the result was STORED in), the unavoidable rounding of
    that storage type, half_ulp * |want_i|, is subtracted element-wise first, so the number
    returned is the error of the COMPUTE (north_star: 2e-3 for bf16/fp16 compute with fp32
    accumulation) and not of the output quantisation."""
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    assert got.shape == want.shape, (got.shape, want.shape)
    ng, nw = np.isnan(got), np.isnan(want)
    assert np.array_equal(ng, nw), f"NaN pattern differs: {ng.sum()} vs {nw.sum()}"
    if nw.all():
        return 0.0
    scale = np.abs(want[~nw]).max()
    diff = np.abs(got[~nw] - want[~nw])
    table = EXACT_MATH_HALF_ULP if exact_math else STORAGE_HALF_ULP
    if storage in table:
        # want_rounded: `want` is itself a result stored in the same 16-bit type (two roundings)
        diff = np.maximum(diff - (2.0 if want_rounded else 1.0) * table[storage] * np.abs(want[~nw]), 0.0)
    return float(diff.max() / (scale if scale > 0 else 1.0))


def tol_for(dtype):
    return TOL_FP32 if dtype == torch.float32 else TOL_16


def sampled_dense_rows(q, k, v, rows, batches):
    """Exact float64 attention for a few (row, batch) pairs of a big problem: O(N d) each."""
    d = q.shape[1]
    out, ls, ms = [], [], []
    for i, b in zip(rows, batches):
        s = (k[:, :, b].astype(np.float64) @ q[i, :, b].astype(np.float64)) / math.sqrt(d)
        m = s.max()
        p = np.exp(s - m)
        out.append((p / p.sum()) @ v[:, :, b].astype(np.float64))
        ls.append(p.sum())
        ms.append(m)
    return np.array(out), np.array(ls), np.array(ms)
