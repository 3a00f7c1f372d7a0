"""Generates tests/golden/*.npz: seeded inputs and the oracle's outputs for every hot-path op.

The reference cannot run here (no Julia) and ships no golden vectors, so these fixtures freeze the
outputs of the PINNED restatement (oracle/fa_oracle.py, pinned in tests/test_oracle.py against torch
SDPA/unfold/fold, brute force and finite differences).  Regenerate with
    python tests/golden/make_golden.py
Inputs are Float32 randn (seed = case index * 10 + tensor index), rounded to bf16 so the same fixture
serves the exact-fp32 path and the 16-bit tensor-core path; outputs are float64 oracle results stored
as float32.  Arrays are stored in their Julia shapes, Fortran order.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import fa_oracle as fo  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def rnd(shape, seed):
    x = np.random.default_rng(seed).standard_normal(shape).astype(np.float32)
    return np.asfortranarray(torch.from_numpy(x).to(torch.bfloat16).float().numpy())


def f64(*xs):
    return [x.astype(np.float64) for x in xs]


CASES = {
    # name: (kind, shape(q), dv, kwargs)
    "dense_n192_d64": ("dense", (192, 64, 2), 64, {}),
    "dense_n40_d12_dv6": ("dense", (40, 12, 2), 6, {}),
    "dense_2d_8x6_d16": ("dense", (8, 6, 16, 2), 16, {}),
    "circ_n256_d64_w33": ("circulant", (256, 64, 2), 64, dict(W=33)),
    "circ_n128_d16_w16": ("circulant", (128, 16, 1), 16, dict(W=16)),
    "win1d_n64_w16_s4": ("windowed", (64, 8, 2), 8, dict(W=16, stride=4, pad=0)),
    "win1d_n22_w5_nan": ("windowed", (22, 8, 1), 8, dict(W=5, stride=5, pad=0)),
    "win2d_20x12_w7": ("windowed", (20, 12, 8, 2), 8, dict(W=7)),
    "win3d_6x7x8_w3": ("windowed", (6, 7, 8, 8, 1), 8, dict(W=3)),
    "circ2d_16x6_d8_w5": ("circulant2d", (16, 6, 8, 2), 8, dict(W=5)),       # SURVEY 8f-2 (no reference code: README todo)
}


def main():
    for ci, (name, (kind, shape, dv, kw)) in enumerate(CASES.items()):
        vshape = shape[:-2] + (dv, shape[-1])
        q, k, v, g = rnd(shape, 10 * ci), rnd(shape, 10 * ci + 1), rnd(vshape, 10 * ci + 2), rnd(vshape, 10 * ci + 3)
        Q, K, V, G = f64(q, k, v, g)
        if kind == "dense":
            y, l, m = fo.dense_fa(Q, K, V)
            dq, dk, dvv = fo.dense_backward(Q, K, V, G)
        elif kind == "circulant":
            y, l, m = fo.circulant_fa(Q, K, V, kw["W"])
            dq, dk, dvv = fo.circulant_backward(Q, K, V, G, kw["W"])
        elif kind == "circulant2d":
            y, l, m = fo.circulant2d_fa(Q, K, V, kw["W"])
            dq, dk, dvv = fo.circulant2d_backward(Q, K, V, G, kw["W"])
        else:
            y, l, m = fo.windowed_fa(Q, K, V, kw["W"], kw.get("stride"), kw.get("pad"))
            dq, dk, dvv = fo.windowed_backward(Q, K, V, G, kw["W"], kw.get("stride"), kw.get("pad"))
        dq, dk, dvv = (np.reshape(t, s, order="F") for t, s in ((dq, shape), (dk, shape), (dvv, vshape)))
        meta = dict(kind=kind, **{k_: (-1 if v_ is None else v_) for k_, v_ in kw.items()})
        np.savez_compressed(os.path.join(OUT, name + ".npz"), q=q, k=k, v=v, g=g,
                            y=y.astype(np.float32), l=l.astype(np.float32), m=m.astype(np.float32),
                            dq=dq.astype(np.float32), dk=dk.astype(np.float32), dv=dvv.astype(np.float32),
                            meta=np.array(repr(meta)))
        print(name, "ok", y.shape)
    # integer index sets (bit-exact)
    np.savez_compressed(os.path.join(OUT, "index_sets.npz"),
                        circ_16_5=fo.circulant_keys(16, 5), circ_32_8=fo.circulant_keys(32, 8),
                        win_9x8_w3_s2_p1=fo.window_index((9, 8), 3, 2, 1), win_6x6x6_w5_s5_p2=fo.window_index((6, 6, 6), 5, 5, 2),
                        circ2d_6x5_w3=fo.circulant2d_keys(6, 5, 3),
                        # slab plans (plane_lo, plane_hi, win_lo, win_hi, pad_lo) per rank: config-5 volume over 8, config-2 image over 3
                        slab_64c_w5_s5_p3_g8=np.array([fo.windowed_slab_plan((64, 64, 64), 5, 5, 3, r, 8) for r in range(8)], dtype=np.int64),
                        slab_64x64_w7_s7_p3_g3=np.array([fo.windowed_slab_plan((64, 64), 7, 7, 3, r, 3) for r in range(3)], dtype=np.int64))


if __name__ == "__main__":
    main()
