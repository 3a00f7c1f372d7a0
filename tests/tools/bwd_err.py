"""Prints the max-abs relative error (tests/util.rel_err) of the tcgen05 backward in both bf16 modes
(default: inputs re-encoded as scaled fp16; FA_FLAG_BF16_INTERNALS: P/dS kept in bf16) against the
float64 oracle evaluated on the same (Q,K,V,O,dO,l,m).  Usage (GPU box): python tests/tools/bwd_err.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "flashattention.jl_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import fa_sm100a as fa
from oracle import fa_oracle as fo
from util import randn_np, rel_err, to_dev, to_np

for dtype in (torch.bfloat16, torch.float16):
    for (N, d, B) in [(128, 64, 1), (256, 128, 2), (520, 128, 1), (1024, 128, 2), (2048, 64, 1)]:
        q, k, v, g = (randn_np((N, d, B), s, dtype) for s in range(4))
        Q, K, V, G = (to_dev(t, dtype) for t in (q, k, v, g))
        y, l, m = fa.dense_fa(Q, K, V)
        want = fo.dense_fa_backward_blocked(*(t.astype(np.float64) for t in (q, k, v)), to_np(y), g.astype(np.float64), to_np(l), to_np(m))
        out = []
        for flags in (0, fa.FA_FLAG_BF16_INTERNALS, fa.FA_FLAG_FORCE_SIMT):
            got = fa.dense_fa_backward(Q, K, V, y, G, l, m, flags=flags)
            out.append([round(rel_err(to_np(x), x0, dtype) * 1e3, 3) for x, x0 in zip(got, want)])
        print(str(dtype), (N, d, B), "default", out[0], "bf16-internals", out[1], "simt", out[2], "(x1e-3; dq dk dv)", flush=True)
