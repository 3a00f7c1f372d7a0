"""Checker for the TMA-gather variant of the windowed forward (FA_WIN_TMA=1, read once per process):
1-D / 2-D / 3-D exact-cover windows against the oracle.  Exit code 1 on a mismatch.  Run by
tests/test_gpu_parity.py::test_windowed_tma_gather_variant in a subprocess."""
import os, sys
os.environ.setdefault("FA_WIN_TMA", "1")
import numpy as np, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path[:0] = [ROOT, os.path.join(ROOT, "flashattention.jl_b200"), os.path.join(ROOT, "tests")]
import fa_sm100a as fa
from oracle import fa_oracle as fo
from util import randn_np, rel_err, to_dev, to_np
bf = torch.bfloat16
bad = 0
for spatial, W, kw in (((64,), 8, {}), ((16, 12), 7, {}), ((24, 9), 4, dict(stride=4, pad=0)), ((16, 11, 10), 5, dict(stride=5, pad=3)),
                       ((64, 64, 64), 5, dict(stride=5, pad=3))):
    B = 2 if len(spatial) < 3 or spatial[0] < 64 else 1
    q, k, v = (randn_np(spatial + (64, B), s, bf) for s in range(3))
    y, l, m = fa.windowed_fa(*(to_dev(t, bf) for t in (q, k, v)), W, **kw)
    torch.cuda.synchronize()
    y0, l0, m0 = fo.windowed_fa(*(t.astype(np.float64) for t in (q, k, v)), W, **kw)
    ey, el = rel_err(to_np(y), y0, bf), rel_err(to_np(l), l0)
    print(spatial, W, kw, fa.last_path(), "err", ey, el, flush=True)
    bad += not (ey <= 2e-3 and el <= 2e-3 and fa.last_path() == "tc")
sys.exit(1 if bad else 0)
