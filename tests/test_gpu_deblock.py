"""gpu: in-loop deblocking (x264_frame_deblock_row, S/common/frame.c:621-792) — wavefront kernel vs the oracle, whole frames,
encoder-like and fully random macroblock state, P and B slices, CAVLC+8x8dct nnz munging, offset extremes."""
import numpy as np
import pytest
import xo_api as X
from helpers import make_deblock_info, blocky_recon

pytestmark = pytest.mark.gpu


def run_both(pkg, ctx, port, w, h, info, seed):
    from x264_vs2008_b200 import synth
    g = port.geometry(w, h)
    y, u, v = blocky_recon(synth.Clip(w, h, seed=5), g, seed=seed)
    py = port.new_plane(g)
    pad = py.reshape(-1, g.stride)
    pad[X.PADV:X.PADV + y.shape[0], X.PADH:X.PADH + y.shape[1]] = y
    uu, vv = u.copy(), v.copy()
    port.frame_deblock(g, info, py, uu, vv)
    want_y = pad[X.PADV:X.PADV + y.shape[0], X.PADH:X.PADH + y.shape[1]]
    f = ctx.frame(w, h, pkg.FRAME_CHROMA)
    f.upload(y); f.upload_chroma(u, v)
    ctx.frame_deblock(f, info)
    gy = f.download(pkg.PLANE_FULL)[pkg.PADV:pkg.PADV + y.shape[0], pkg.PADH:pkg.PADH + y.shape[1]]
    gu = f.download(pkg.PLANE_CB)[16:16 + u.shape[0], 16:16 + u.shape[1]]
    gv = f.download(pkg.PLANE_CR)[16:16 + v.shape[0], 16:16 + v.shape[1]]
    f.close()
    assert w < 64 or not np.array_equal(want_y, y)  # something was filtered
    for name, a, b in (("y", gy, want_y), ("u", gu, uu), ("v", gv, vv)):
        bad = np.argwhere(a != b)
        assert len(bad) == 0, (name, len(bad), bad[:6], a[tuple(bad[0])], b[tuple(bad[0])])


@pytest.mark.parametrize("size,kw", [((176, 144), dict()), ((176, 144), dict(slice_b=1)), ((208, 112), dict(cavlc_8x8dct=1, alpha=-2, beta=2, chroma_off=3)),
                                     ((64, 48), dict(chaos=True)), ((96, 80), dict(chaos=True, slice_b=1, cavlc_8x8dct=1, alpha=6, beta=-4, chroma_off=-5)),
                                     ((176, 144), dict(psub8x8=0, qp_centre=18, alpha=-6, beta=-6)), ((32, 16), dict(qp_centre=45, alpha=12, beta=12)),
                                     ((16, 16), dict()), ((1920, 1080), dict()), ((1920, 1080), dict(chaos=True, slice_b=1))])
def test_frame_deblock(pkg, ctx, port, size, kw):
    w, h = size
    g = port.geometry(w, h)
    for seed in range(2 if w < 1000 else 1):
        run_both(pkg, ctx, port, w, h, make_deblock_info(g, seed=100 + seed, **kw), seed)
