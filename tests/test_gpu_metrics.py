"""gpu: whole-frame analysis metrics — SSD (PSNR), SSIM, AQ macroblock energies, hadamard_ac — vs the oracle.
Integer outputs bit-exact; the float tails are the product's host helpers and must equal the oracle's floats bit for bit."""
import numpy as np
import pytest
import xo_api as X
from test_oracle_vs_ref import _metric_inputs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("size", [(176, 144), (100, 70), (352, 288), (20, 12), (1920, 1080)])
def test_frame_metrics(pkg, ctx, port, size):
    w, h = size
    g, (y, u, v), (ry, ru, rv) = _metric_inputs(port, w, h, seed=9)
    fa, fb = ctx.frame(w, h, pkg.FRAME_CHROMA), ctx.frame(w, h, pkg.FRAME_CHROMA)
    fa.upload(y); fa.upload_chroma(u, v)
    fb.upload(ry); fb.upload_chroma(ru, rv)
    for plane, a, b, ww, hh in ((pkg.PLANE_FULL, y, ry, w, h), (pkg.PLANE_CB, u, ru, w // 2, h // 2), (pkg.PLANE_CR, v, rv, w // 2, h // 2),
                                (pkg.PLANE_FULL, y, ry, w - 3, h - 5)):
        assert ctx.frame_ssd(fa, fb, plane, ww, hh) == port.frame_ssd(a, b, ww, hh)
        if ww >= 8 and hh >= 8:
            val, sums = ctx.frame_ssim(fa, fb, plane, ww, hh)
            assert np.array_equal(sums, port.frame_ssim_sums(a, b, ww, hh))
            assert val == port.frame_ssim(a, b, ww, hh)
    if h >= 48:  # the encoder's SSIM slabs: x0 = 2, rows [min_y, max_y) (encoder.c:1047-1056)
        for (y0, hh) in ((2, 22), (10, 32), (h - 30, 30)):
            val, sums = ctx.frame_ssim(fa, fb, pkg.PLANE_FULL, w - 2, hh, x0=2, y0=y0)
            a, b = np.ascontiguousarray(y[y0:, 2:]), np.ascontiguousarray(ry[y0:, 2:])
            assert np.array_equal(sums, port.frame_ssim_sums(a, b, w - 2, hh)) and val == port.frame_ssim(a, b, w - 2, hh)
            assert ctx.frame_ssd(fa, fb, pkg.PLANE_FULL, w - 2, hh, x0=2, y0=y0) == port.frame_ssd(a, b, w - 2, hh)
    py = port.new_plane(g)
    py.reshape(-1, g.stride)[X.PADV:X.PADV + y.shape[0], X.PADH:X.PADH + y.shape[1]] = y
    e = ctx.frame_mb_energy(fa)
    assert np.array_equal(e, port.frame_mb_energy(g, py, u, v))
    assert np.array_equal(ctx.frame_mb_hadamard_ac(fa), port.frame_mb_hadamard_ac(g, py))
    q, inv = pkg.host_aq(e, 1.0)
    q2, inv2 = port.frame_aq(g, py, u, v, 1.0)
    assert np.array_equal(q, q2) and np.array_equal(inv, inv2)
    fa.close(); fb.close()


def test_metric_extremes(pkg, ctx, port):
    """all-0 vs all-255 (largest SSD / sums), checkerboards (largest Hadamard energy), constant pictures (zero variance -> energy 1)"""
    w, h = 64, 48
    g = port.geometry(w, h)
    yy, xx = np.mgrid[0:h, 0:w]
    pats = [(np.zeros((h, w), np.uint8), np.full((h, w), 255, np.uint8)), ((((xx + yy) & 1) * 255).astype(np.uint8), (((xx + yy + 1) & 1) * 255).astype(np.uint8)),
            (np.full((h, w), 77, np.uint8), np.full((h, w), 77, np.uint8))]
    for a, b in pats:
        fa, fb = ctx.frame(w, h, pkg.FRAME_CHROMA), ctx.frame(w, h, pkg.FRAME_CHROMA)
        ca, cb = np.ascontiguousarray(a[::2, ::2]), np.ascontiguousarray(b[::2, ::2])
        fa.upload(a); fa.upload_chroma(ca, ca); fb.upload(b); fb.upload_chroma(cb, cb)
        assert ctx.frame_ssd(fa, fb, pkg.PLANE_FULL, w, h) == port.frame_ssd(a, b, w, h)
        val, sums = ctx.frame_ssim(fa, fb, pkg.PLANE_FULL, w, h)
        assert np.array_equal(sums, port.frame_ssim_sums(a, b, w, h)) and val == port.frame_ssim(a, b, w, h)
        py = port.new_plane(g)
        py.reshape(-1, g.stride)[X.PADV:X.PADV + h, X.PADH:X.PADH + w] = a
        assert np.array_equal(ctx.frame_mb_energy(fa), port.frame_mb_energy(g, py, ca, ca))
        assert np.array_equal(ctx.frame_mb_hadamard_ac(fa), port.frame_mb_hadamard_ac(g, py))
        fa.close(); fb.close()
