"""Checker for the streamed windowed kernels of csrc/fa_tc_winx.cu, forced on with FA_WINX=1 (the flag is read once per
process; by default the kernel only takes problems with at least four groups per SM): 1-D / 2-D / 3-D exact-cover
windows incl. short last groups, every box shift, default padding (uncovered planes -> NaN), fp16 and bf16, against
the float64 oracle AND bit for bit against the round-1 kernel is not required (different summation order inside the
tensor core is not involved: the operands are identical, so S, P and O are the same bits -- checked as equality of y
with FA_WINX=0 run in a second process by the caller).  Prints one line per case; exit code 1 on a mismatch.
  python tests/tools/check_winx.py [dump.pt]     # with a path: also saves the outputs for the bitwise comparison"""
import os, sys
import numpy as np, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path[:0] = [ROOT, os.path.join(ROOT, "flashattention.jl_b200"), os.path.join(ROOT, "tests")]
import fa_sm100a as fa
from oracle import fa_oracle as fo
from util import randn_np, rel_err, to_dev, to_np

CASES = [
    ((256,), 32, {}, 3),                                 # 1-D defaults (stride 32, pad 15): shifted boxes, 4 windows per tile
    ((512,), 48, dict(stride=48, pad=0), 2),             # block_fa 1-D, aligned starts, a short last group (10 windows: 8 + 2)
    ((16, 12), 7, {}, 2),                                # config-2 geometry in small: G = 2, one short group per row
    ((64, 64), 7, {}, 3),                                # config 2 (64x64, W 7, pad 3): 10 windows per row = groups of 8 + 2
    ((24, 9), 4, dict(stride=4, pad=0), 2),              # 16 slots per window, G = 8
    ((16, 11, 10), 5, dict(stride=5, pad=3), 2),         # config-5 geometry in small: 125 slots, G = 1
    ((40, 12, 11), 5, {}, 1),                            # 3-D defaults (pad 2): uncovered last positions -> NaN
    ((64, 64, 64), 5, dict(stride=5, pad=3), 1),         # config 5 volume: 14 windows per row = groups 4 + 4 + 4 + 2
    ((32, 10, 9), 3, dict(stride=3, pad=1), 2),          # 27 slots, G = 4, 16 windows per group
]


def main():
    dump = {}
    bad = 0
    for dt in (torch.bfloat16, torch.float16):
        for spatial, W, kw, B in CASES:
            if dt == torch.float16 and np.prod(spatial) > 70000:
                continue
            q, k, v = (randn_np(spatial + (64, B), s, dt) for s in range(3))
            y, l, m = fa.windowed_fa(*(to_dev(t, dt) for t in (q, k, v)), W, **kw)
            torch.cuda.synchronize()
            y0, l0, m0 = fo.windowed_fa(*(t.astype(np.float64) for t in (q, k, v)), W, **kw)
            try:
                ey, el = rel_err(to_np(y), y0, dt), rel_err(to_np(l), l0)
                em = float(np.abs(to_np(m) - m0).max())
            except AssertionError as e:
                ey = el = em = float("inf")
                print("  ", str(e)[:200])
            ok = ey <= 2e-3 and el <= 2e-3 and em <= 2e-3 * max(1.0, np.abs(m0).max()) and fa.last_path() == "tc"
            print(str(dt)[6:], spatial, W, kw, "B", B, fa.last_path(), "err y %.2e l %.2e m %.2e" % (ey, el, em), "ok" if ok else "MISMATCH", flush=True)
            bad += not ok
            dump[f"{dt}{spatial}{W}{kw}"] = (y.cpu(), l.cpu(), m.cpu())
    if len(sys.argv) > 1:
        torch.save(dump, sys.argv[1])
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
