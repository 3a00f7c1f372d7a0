"""Consumes golden vectors exported by a maintainer who HAS Julia from the ORIGINAL reference package
(flashattention.jl_b200/julia/FlashAttention/bench/export_golden.jl -> tests/golden/julia/*.f64 + MANIFEST.txt).  None can be
produced in the build image (no Julia runtime), so without the directory these tests skip; with it they pin the oracle --
including NNlib's unfold / fold semantics, which no reference test pins -- to real outputs of the reference."""
import os

import numpy as np
import pytest

from oracle import fa_oracle as fo

DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "julia")
HAVE = os.path.exists(os.path.join(DIR, "MANIFEST.txt"))
pytestmark = pytest.mark.skipif(not HAVE, reason="no Julia-exported golden vectors (tests/golden/julia/MANIFEST.txt); see export_golden.jl")


def load():
    out = {}
    for line in open(os.path.join(DIR, "MANIFEST.txt")):
        name, shape = line.split()
        shape = tuple(int(s) for s in shape.split("x"))
        out[name] = np.fromfile(os.path.join(DIR, name + ".f64"), dtype="<f8").reshape(shape, order="F")
    return out


KW = {"win1d_n64_w16_s4": (16, 4, 0), "win2d_20x12_w7": (7, None, None), "win3d_6x7x8_w3": (3, None, None), "win1d_n22_w5_nan": (5, 5, 0),
      "circ_n128_d16_w16": (16,), "circ_n256_d8_w33": (33,)}


def test_oracle_matches_julia_reference_outputs():
    g = load()
    tags = sorted({n.rsplit("_", 1)[0] for n in g})
    assert tags
    for tag in tags:
        q, k, v = g[tag + "_q"], g[tag + "_k"], g[tag + "_v"]
        if tag.startswith("dense"):
            y, l, m = fo.dense_fa(q, k, v)
        elif tag.startswith("win"):
            W, stride, pad = KW[tag]
            y, l, m = fo.windowed_fa(q, k, v, W, stride, pad)
            assert np.array_equal(fo.window(q, W, stride, pad), g[tag + "_qw"])        # unfold semantics, bit-exact
        else:
            y, l, m = fo.circulant_fa(q, k, v, KW[tag][0])
        for got, name in ((y, "y"), (l, "l"), (m, "m")):
            want = g[f"{tag}_{name}"]
            assert np.array_equal(np.isnan(got), np.isnan(want))
            assert np.nanmax(np.abs(got - want)) <= 1e-10 * max(1.0, np.nanmax(np.abs(want))), (tag, name)
