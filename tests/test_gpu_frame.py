"""gpu: whole-frame kernels (border, hpel, integral, lowres) vs the oracle, bit-exact, through the C ABI"""
import numpy as np
import pytest
import xo_api as X
from helpers import pad_view

pytestmark = pytest.mark.gpu

SIZES = [(64, 48), (100, 70), (352, 288), (720, 480), (1920, 1080)]


def _pictures(w, h, seed):
    from x264_vs2008_b200 import synth
    yield "synthetic", synth.Clip(w, h, seed=seed).luma(1)
    rng = np.random.default_rng(seed)
    yield "noise", rng.integers(0, 256, (h, w), dtype=np.uint8)
    yield "zeros", np.zeros((h, w), np.uint8)
    yield "white", np.full((h, w), 255, np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    yield "checker", (((xx + yy) & 1) * 255).astype(np.uint8)  # max-difference pattern (checkasm.c:246-256)


@pytest.mark.parametrize("w,h", SIZES)
def test_border_hpel_integral(pkg, ctx, port, w, h):
    g = port.geometry(w, h)
    for sub8 in (0, 1):
        flags = pkg.FRAME_HPEL | pkg.FRAME_INTEGRAL | (pkg.FRAME_INTEGRAL4 if sub8 else 0)
        f = ctx.frame(w, h, flags)
        for name, pic in _pictures(w, h, 5 + sub8):
            if (w * h > 500000) and name not in ("synthetic", "checker"):
                continue
            f.upload(pic)
            f.expand_border()
            f.filter()
            want_plane = port.plane_from_picture(g, pic)
            assert np.array_equal(f.download(pkg.PLANE_FULL), pad_view(g, want_plane)), (name, "border")
            oh, ov, oc, oi = port.frame_filter(g, want_plane, sub8)
            for pl, want in ((pkg.PLANE_H, oh), (pkg.PLANE_V, ov), (pkg.PLANE_C, oc)):
                assert np.array_equal(f.download(pl), pad_view(g, want)), (name, "hpel plane", pl)
            # integral: the reference defines rows [-31, lines+23] x cols [-32, w16+24) (see frame_filter.cu)
            r0, r1, c1 = 1, g.lines + 24 + 32, g.mb_width * 16 + 24 + 32
            got = f.download(pkg.PLANE_INTEGRAL)
            want8 = oi[:g.plane_size].reshape(-1, g.stride)
            assert np.array_equal(got[r0:r1, :c1], want8[r0:r1, :c1]), (name, "integral8")
            if sub8:
                got4 = f.download(pkg.PLANE_INTEGRAL4)
                want4 = oi[g.plane_size:].reshape(-1, g.stride)
                assert np.array_equal(got4[r0:r1, :c1], want4[r0:r1, :c1]), (name, "integral4")
        f.close()


@pytest.mark.parametrize("w,h", SIZES)
def test_lowres(pkg, ctx, port, w, h):
    g = port.geometry(w, h)
    f = ctx.frame(w, h, pkg.FRAME_LOWRES)
    for name, pic in _pictures(w, h, 9):
        if (w * h > 500000) and name not in ("synthetic", "checker"):
            continue
        f.upload(pic)
        f.expand_border()
        f.init_lowres()
        want = port.init_lowres(g, port.plane_from_picture(g, pic))
        for i in range(4):
            got = f.download(pkg.PLANE_LOWRES + i)
            assert np.array_equal(got, pad_view(g, want[i], lowres=True)), (name, "lowres", i)
    f.close()


def test_full_size_properties_4k(pkg, ctx):
    """size-independent properties at 4K: a constant picture filters to itself; hpel of a frame shifted by whole
    pixels is the shifted hpel (interior); integral of ones is 64 / 16."""
    w, h = 3840, 2160
    f = ctx.frame(w, h, pkg.FRAME_HPEL | pkg.FRAME_INTEGRAL | pkg.FRAME_INTEGRAL4 | pkg.FRAME_LOWRES)
    f.upload(np.full((h, w), 77, np.uint8)); f.expand_border(); f.filter(); f.init_lowres()
    for pl in (pkg.PLANE_H, pkg.PLANE_V, pkg.PLANE_C, pkg.PLANE_LOWRES, pkg.PLANE_LOWRES + 3):
        assert np.all(f.download(pl) == 77)
    f.upload(np.ones((h, w), np.uint8)); f.expand_border(); f.filter()
    i8, i4 = f.download(pkg.PLANE_INTEGRAL), f.download(pkg.PLANE_INTEGRAL4)
    assert np.all(i8[1:h + 24 + 32, :w + 24 + 32] == 64) and np.all(i4[1:h + 24 + 32, :w + 24 + 32] == 16)
    rng = np.random.default_rng(1)
    big = rng.integers(0, 256, (h + 8, w + 8), dtype=np.uint8)
    f.upload(np.ascontiguousarray(big[:h, :w])); f.expand_border(); f.filter()
    a = [f.download(p) for p in (pkg.PLANE_H, pkg.PLANE_V, pkg.PLANE_C)]
    f.upload(np.ascontiguousarray(big[3:h + 3, 5:w + 5])); f.expand_border(); f.filter()
    b = [f.download(p) for p in (pkg.PLANE_H, pkg.PLANE_V, pkg.PLANE_C)]
    m = 48  # stay away from the replicated borders
    for x, y in zip(a, b):
        assert np.array_equal(x[32 + 3 + m:32 + h - m, 32 + 5 + m:32 + w - m], y[32 + m:32 + h - 3 - m, 32 + m:32 + w - 5 - m])
    f.close()
