"""Guard-band test of every kernel family (compute-sanitizer is closed on this pool): every OUTPUT and WORKSPACE buffer the
host mirror allocates is carved out of a larger sentinel-filled arena (64 KiB of 0xA5 in front and behind); after each
operator the bands must be untouched -- an out-of-bounds store of a kernel (ragged last tile, padded window slots,
staging / TMA box overhang, workspace sub-buffers) lands in them.  The shapes are the ragged ones of tools/sanity_small.py
and tests/test_gpu_fuzz.py (parity of the same shapes is checked there; here only the bands and finiteness)."""
import os
import subprocess
import sys

import pytest
import torch

import fa_sm100a as fa  # loads lib/libfa_sm100a.so, fails loudly without it

pytestmark = pytest.mark.gpu

BAND = 65536
SENT = 0xA5


class Arena:
    """Replaces fa.jl_empty / fa._workspace: tensors are views of the middle of a sentinel-filled byte buffer."""

    def __init__(self):
        self.blocks = []

    def _carve(self, nbytes, device):
        nbytes = int(nbytes)
        pad = (-nbytes) % 256
        buf = torch.full((BAND + nbytes + pad + BAND,), SENT, dtype=torch.uint8, device=device)
        self.blocks.append((buf, nbytes))
        return buf[BAND:BAND + nbytes]

    def jl_empty(self, shape, dtype=torch.float32, device="cuda"):
        shape = tuple(int(s) for s in shape)
        n = 1
        for s in shape:
            n *= s
        esz = torch.empty((), dtype=dtype).element_size()
        mid = self._carve(n * esz, device)
        return mid.view(dtype).reshape(shape[::-1]).permute(*range(len(shape) - 1, -1, -1))

    def workspace(self, nbytes, device):
        return self._carve(max(int(nbytes), 256), device)

    def check(self, what):
        torch.cuda.synchronize()
        for buf, nbytes in self.blocks:
            front = buf[:BAND]
            back = buf[BAND + nbytes:]
            assert bool((front == SENT).all()), f"{what}: bytes written in FRONT of a {nbytes}-byte buffer"
            assert bool((back == SENT).all()), f"{what}: bytes written BEHIND a {nbytes}-byte buffer"
        n = len(self.blocks)
        self.blocks = []
        return n


@pytest.fixture
def arena(monkeypatch):
    a = Arena()
    monkeypatch.setattr(fa, "jl_empty", a.jl_empty)
    monkeypatch.setattr(fa, "_workspace", a.workspace)
    return a


def _rand(shape, dt, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    t = torch.empty(tuple(shape)[::-1], dtype=torch.float32, device="cuda").normal_(generator=g).to(dt)
    return t.permute(*range(len(shape) - 1, -1, -1))


BF, F16, F32 = torch.bfloat16, torch.float16, torch.float32


@pytest.mark.parametrize("dt,shape", [(BF, (200, 128, 2)), (BF, (520, 64, 1)), (F16, (136, 128, 3)), (BF, (72, 32, 2)),
                                      (F32, (100, 24, 2)), (BF, (1031 * 8, 64, 1))])
def test_guard_dense(arena, dt, shape):
    q, k, v, g = (_rand(shape, dt, s) for s in range(4))
    O, l, m = fa.dense_fa(q, k, v)
    assert arena.check(f"dense_fa {dt} {shape}") >= 3
    dq, dk, dv = fa.dense_fa_backward(q, k, v, O, g, l, m)
    assert arena.check(f"dense_fa_backward {dt} {shape}") >= 3
    assert all(bool(torch.isfinite(t.float()).all()) for t in (O, l, m, dq, dk, dv))


@pytest.mark.parametrize("dt,shape,W", [(BF, (384, 64, 2), 65), (BF, (256, 128, 1), 33), (F16, (448, 64, 3), 255), (BF, (192, 32, 2), 17),
                                        (F32, (64, 8, 2), 9)])
def test_guard_circulant(arena, dt, shape, W):
    q, k, v, g = (_rand(shape, dt, s) for s in range(4))
    O, l, m = fa.circulant_fa(q, k, v, W)
    assert arena.check(f"circulant_fa {dt} {shape} W={W}") >= 3
    fa.circulant_fa_backward(q, k, v, O, g, l, m, W)
    assert arena.check(f"circulant_fa_backward {dt} {shape} W={W}") >= 3


def test_guard_circulant_2d(arena):
    q, k, v, g = (_rand((64, 10, 64, 2), BF, s) for s in range(4))
    O, l, m = fa.circulant_fa(q, k, v, 5)
    assert arena.check("circulant_fa 2-D") >= 3
    fa.circulant_fa_backward(q, k, v, O, g, l, m, 5)
    assert arena.check("circulant_fa_backward 2-D") >= 3


WIN_CASES = [(BF, (20, 12, 64, 3), 7, {}), (BF, (12, 11, 10, 64, 2), 5, dict(stride=5, pad=3)), (BF, (64, 64, 2), 16, dict(stride=4, pad=0)),
             (F16, (24, 9, 64, 2), 4, dict(stride=4, pad=0)), (BF, (16, 11, 10, 64, 2), 5, {}), (BF, (40, 32, 2), 8, dict(stride=8, pad=0)),
             (BF, (256, 32, 3), 32, {}), (F32, (6, 7, 8, 8, 2), 3, {}), (BF, (18, 15, 128, 1), 5, dict(stride=3, pad=1))]


@pytest.mark.parametrize("dt,shape,W,kw", WIN_CASES)
def test_guard_windowed(arena, dt, shape, W, kw):
    q, k, v, g = (_rand(shape, dt, s) for s in range(4))
    y, l, m = fa.windowed_fa(q, k, v, W, **kw)
    assert arena.check(f"windowed_fa {dt} {shape} W={W} {kw}") >= 3
    fa.windowed_fa_backward(q, k, v, g, l, m, W, **kw)
    assert arena.check(f"windowed_fa_backward {dt} {shape} W={W} {kw}") >= 3


def test_guard_small_ops(arena):
    fa.fused_softmax(_rand((70000, 1), F32, 1), 1)
    fa.fused_softmax(_rand((37, 1000, 2), BF, 2), 2)
    fa.fused_softmax(_rand((513, 129), F16, 3), 1)
    assert arena.check("fused_softmax") >= 3
    x = _rand((13, 11, 24, 2), BF, 4)
    xw = fa.window(x, 5, 3, 2)
    assert arena.check("window") >= 1
    fa.unwindow(xw, (13, 11, 24, 2), 5, 3, 2)
    assert arena.check("unwindow") >= 1
    fa.cast(_rand((1001, 3), F32, 5), BF)
    assert arena.check("cast") >= 1


def test_guard_streamed_windowed_kernels_subprocess():
    """fa_tc_winx.cu is selected by environment variables read once per process: the same guard-band check in a child with
    FA_WINX=1 (4-byte store output) and FA_WINX=1 FA_WINX_OUT=1 (TMA reduce-add output; the box overhang beyond the group's
    own columns must carry zeros and out-of-volume coordinates must be dropped)."""
    code = r"""
import sys, os
sys.path.insert(0, os.path.join(%r, "tests"))
sys.path.insert(0, os.path.join(%r, "flashattention.jl_b200"))
import torch
import test_gpu_guard as tg
import fa_sm100a as fa
a = tg.Arena()
fa.jl_empty = a.jl_empty
fa._workspace = a.workspace
for dt, shape, W, kw in [(tg.BF, (16, 11, 10, 64, 2), 5, dict(stride=5, pad=3)), (tg.BF, (40, 12, 11, 64, 1), 5, {}), (tg.F16, (32, 10, 9, 64, 2), 3, dict(stride=3, pad=1)),
                         (tg.BF, (64, 64, 64, 2), 7, {}), (tg.BF, (24, 9, 64, 2), 4, dict(stride=4, pad=0)), (tg.BF, (64, 20, 15, 64, 1), 5, dict(stride=5, pad=3))]:
    q, k, v = (tg._rand(shape, dt, s) for s in range(3))
    y, l, m = fa.windowed_fa(q, k, v, W, **kw)
    assert fa.last_path() == "tc"
    assert a.check(f"winx {shape} {W} {kw}") >= 3
    cover = ~torch.isnan(y.float())
    assert bool(torch.isfinite(y.float()[cover]).all()) and bool((l > 0).all())
print("guard ok")
""" % ((os.path.dirname(os.path.dirname(os.path.abspath(__file__))),) * 2)
    for extra in ({"FA_WINX": "1"}, {"FA_WINX": "1", "FA_WINX_OUT": "1"}):
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=dict(os.environ, **extra))
        assert r.returncode == 0 and "guard ok" in r.stdout, f"{extra}: " + r.stdout[-2000:] + r.stderr[-3000:]
