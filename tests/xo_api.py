"""ctypes binding of the ORACLE api (oracle/src/xo.h).  Test infrastructure only.

Two shared objects export this same symbol set:
  oracle/liboracle.so            -- our C restatement ("port")
  oracle/_ref/libref_harness.so  -- the unmodified reference compiled from /root/reference ("reference");
                                    present only when built in the container (it travels to the GPU box
                                    inside the gpurun snapshot).
"""
import ctypes as C
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PORT_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_harness.so")

PADH = PADV = 32
SAD, SSD, SATD, SA8D = 0, 1, 2, 3
P16x16, P16x8, P8x16, P8x8, P8x4, P4x8, P4x4 = range(7)
BLK_W = [16, 16, 8, 8, 8, 4, 4]
BLK_H = [16, 8, 16, 8, 4, 8, 4]
ME_DIA, ME_HEX, ME_UMH, ME_ESA, ME_TESA = range(5)
COST_MAX = 1 << 28


class Geom(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "width", "height", "mb_width", "mb_height", "stride", "lines", "plane_size", "origin",
        "stride_lowres", "width_lowres", "lines_lowres", "plane_size_lowres", "origin_lowres")]


class MeIn(C.Structure):
    _fields_ = [("me_method", C.c_int), ("me_range", C.c_int), ("qp", C.c_int), ("fpel_satd", C.c_int),
                ("i_pixel", C.c_int), ("bx", C.c_int), ("by", C.c_int),
                ("mv_min_fpel", C.c_int * 2), ("mv_max_fpel", C.c_int * 2),
                ("mv_min_spel", C.c_int * 2), ("mv_max_spel", C.c_int * 2),
                ("mvp", C.c_int16 * 2), ("i_mvc", C.c_int), ("mvc", (C.c_int16 * 2) * 16), ("b_sub8x8", C.c_int)]


class MeOut(C.Structure):
    _fields_ = [("mv", C.c_int16 * 2), ("cost", C.c_int), ("cost_mv", C.c_int),
                ("bmx", C.c_int), ("bmy", C.c_int), ("bcost", C.c_int),
                ("seed_mx", C.c_int), ("seed_my", C.c_int), ("seed_cost", C.c_int)]


class LowresIn(C.Structure):
    _fields_ = [("p0", C.c_int), ("p1", C.c_int), ("b", C.c_int), ("me_method", C.c_int), ("me_range", C.c_int),
                ("mbcmp_satd", C.c_int), ("fpel_satd", C.c_int), ("b_weighted_bipred", C.c_int),
                ("do_search", C.c_int * 2), ("b_intra_calculated", C.c_int)]


class LowresOut(C.Structure):
    _fields_ = [("score", C.c_int), ("score_aq", C.c_int), ("intra_mbs", C.c_int), ("intra_cost_sum", C.c_int)]


class Chroma(C.Structure):
    _fields_ = [("fenc_u", C.c_void_p), ("fenc_v", C.c_void_p), ("fref_u", C.c_void_p), ("fref_v", C.c_void_p), ("stride_c", C.c_int)]


class DeblockIn(C.Structure):
    _fields_ = [("alpha_c0_offset", C.c_int), ("beta_offset", C.c_int), ("chroma_qp_offset", C.c_int), ("b_slice_b", C.c_int),
                ("b_psub8x8", C.c_int), ("b_cavlc_8x8dct", C.c_int),
                ("type", C.c_void_p), ("qp", C.c_void_p), ("transform8x8", C.c_void_p), ("nnz", C.c_void_p),
                ("ref", C.c_void_p * 2), ("mv", C.c_void_p * 2)]


class ResidIn(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("qp", "chroma_qp", "b_transform_8x8", "b_decimate", "cqm")]


class ResidOut(C.Structure):
    _fields_ = [("luma4x4", (C.c_int16 * 16) * 24), ("luma8x8", (C.c_int16 * 64) * 4), ("chroma_dc", (C.c_int16 * 4) * 2),
                ("nnz", C.c_uint8 * 27), ("pad", C.c_uint8), ("cbp_luma", C.c_int), ("cbp_chroma", C.c_int)]


class IntraIn(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("neighbour", "lambda_", "mbcmp_satd", "b_slice_b")]


class IntraOut(C.Structure):
    _fields_ = [("cost16", C.c_int * 7), ("cost_chroma", C.c_int * 7), ("best16", C.c_int), ("best_chroma", C.c_int),
                ("mode16", C.c_int), ("mode_chroma", C.c_int)]

    def astuple(self):
        return (tuple(self.cost16), tuple(self.cost_chroma), self.best16, self.best_chroma, self.mode16, self.mode_chroma)


u8p = C.POINTER(C.c_uint8)
u16p = C.POINTER(C.c_uint16)
i16p = C.POINTER(C.c_int16)
i32p = C.POINTER(C.c_int)


def _ptr(a, ty=u8p, off=0):
    """pointer to element `off` of a contiguous numpy array"""
    assert a.flags["C_CONTIGUOUS"]
    return C.cast(a.ctypes.data + off * a.itemsize, ty)


class Oracle:
    def __init__(self, path):
        self.lib = L = C.CDLL(path)
        L.xo_backend.restype = C.c_char_p
        L.xo_pixel_cmp.argtypes = [C.c_int, C.c_int, u8p, C.c_int, u8p, C.c_int]
        L.xo_pixel_var.argtypes = [C.c_int, u8p, C.c_int]
        L.xo_pixel_hadamard_ac.argtypes = [C.c_int, u8p, C.c_int]
        L.xo_pixel_hadamard_ac.restype = C.c_uint64
        L.xo_pixel_ads.argtypes = [C.c_int, i32p, u16p, C.c_int, u16p, i16p, C.c_int, C.c_int]
        L.xo_cost_mv_table.argtypes = [C.c_int, i16p]
        L.xo_geometry.argtypes = [C.c_int, C.c_int, C.POINTER(Geom)]
        L.xo_frame_expand_border.argtypes = [C.POINTER(Geom), u8p]
        L.xo_frame_filter.argtypes = [C.POINTER(Geom), u8p, u8p, u8p, u8p, u16p, C.c_int]
        L.xo_frame_init_lowres.argtypes = [C.POINTER(Geom), u8p, u8p, u8p, u8p, u8p]
        L.xo_mc_luma.argtypes = [u8p, C.c_int, C.POINTER(u8p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.xo_mc_chroma.argtypes = [u8p, C.c_int, u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.xo_pixel_avg.argtypes = [C.c_int, u8p, C.c_int, u8p, C.c_int, u8p, C.c_int, C.c_int]
        L.xo_me_search_fpel.argtypes = [C.POINTER(Geom), u8p, u8p, u16p, C.POINTER(MeIn), C.POINTER(MeOut)]
        L.xo_me_search_fpel_batch.argtypes = [C.POINTER(Geom), u8p, u8p, u16p, C.POINTER(MeIn), C.c_int, C.POINTER(MeOut)]
        L.xo_me_search_subpel.argtypes = [C.POINTER(Geom), u8p, C.POINTER(u8p), u16p, C.POINTER(MeIn), C.c_int,
                                          C.c_int, C.POINTER(MeOut)]
        for n in ("xo_sub4x4_dct", "xo_sub8x8_dct8"):
            getattr(L, n).argtypes = [i16p, u8p, u8p]
        for n in ("xo_add4x4_idct", "xo_add8x8_idct8"):
            getattr(L, n).argtypes = [u8p, i16p]
        L.xo_dct4x4dc.argtypes = [i16p]
        L.xo_idct4x4dc.argtypes = [i16p]
        L.xo_add_idct_dc.argtypes = [u8p, i16p, C.c_int]
        L.xo_quant4_tables.argtypes = [C.c_int, C.c_int, C.c_int, u16p, u16p]
        L.xo_quant8_tables.argtypes = [C.c_int, C.c_int, C.c_int, u16p, u16p]
        L.xo_dequant4_table.argtypes = [C.c_int, C.c_int, i32p]
        L.xo_dequant8_table.argtypes = [C.c_int, C.c_int, i32p]
        L.xo_quant_4x4.argtypes = [i16p, u16p, u16p]
        L.xo_quant_8x8.argtypes = [i16p, u16p, u16p]
        L.xo_quant_4x4_dc.argtypes = [i16p, C.c_int, C.c_int]
        L.xo_quant_2x2_dc.argtypes = [i16p, C.c_int, C.c_int]
        for n in ("xo_dequant_4x4", "xo_dequant_8x8", "xo_dequant_4x4_dc"):
            getattr(L, n).argtypes = [i16p, i32p, C.c_int]
        L.xo_lowres_frame_cost.argtypes = [C.POINTER(Geom), C.POINTER(LowresIn), C.POINTER(u8p), C.POINTER(u8p), C.POINTER(u8p),
                                           i16p, i32p, i16p, i32p, i16p, u16p, C.POINTER(LowresOut)]
        L.xo_lowres_frame_cost_vbv.argtypes = list(L.xo_lowres_frame_cost.argtypes) + [u16p, i32p]
        L.xo_lowres_intra_pred.argtypes = [C.c_int, u8p, C.c_int, C.c_int, C.c_int, u8p]
        L.xo_lowres_intra_cost.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.xo_me_search_subpel_chroma.argtypes = [C.POINTER(Geom), u8p, C.POINTER(u8p), u16p, C.POINTER(Chroma), C.POINTER(MeIn), C.c_int,
                                                 C.c_int, C.POINTER(MeOut)]
        L.xo_frame_ssd.argtypes = [u8p, C.c_int, u8p, C.c_int, C.c_int, C.c_int]
        L.xo_frame_ssd.restype = C.c_int64
        L.xo_frame_ssim.argtypes = [u8p, C.c_int, u8p, C.c_int, C.c_int, C.c_int]
        L.xo_frame_ssim.restype = C.c_float
        L.xo_frame_ssim_sums.argtypes = [u8p, C.c_int, u8p, C.c_int, C.c_int, C.c_int, i32p]
        L.xo_frame_mb_energy.argtypes = [C.POINTER(Geom), u8p, u8p, u8p, C.c_int, C.POINTER(C.c_uint32)]
        L.xo_frame_mb_hadamard_ac.argtypes = [C.POINTER(Geom), u8p, C.POINTER(C.c_uint64)]
        L.xo_frame_aq.argtypes = [C.POINTER(Geom), u8p, u8p, u8p, C.c_int, C.c_float, C.POINTER(C.c_float), u16p]
        L.xo_me_refine_qpel.argtypes = [C.POINTER(Geom), u8p, C.POINTER(u8p), C.POINTER(Chroma), C.POINTER(MeIn), C.c_int, C.c_int, i16p, C.c_int,
                                        C.POINTER(MeOut)]
        L.xo_me_refine_bidir_satd.argtypes = [C.POINTER(Geom), u8p, C.POINTER(u8p), C.POINTER(u8p), C.POINTER(MeIn), i16p, i16p, C.c_int, C.c_int, i16p, i16p]
        L.xo_frame_deblock.argtypes = [C.POINTER(Geom), C.POINTER(DeblockIn), u8p, u8p, u8p, C.c_int]
        self.backend = L.xo_backend().decode()

    # ---- convenience wrappers (numpy in / numpy out) ----
    def geometry(self, w, h):
        g = Geom()
        self.lib.xo_geometry(w, h, C.byref(g))
        return g

    def pixel_cmp(self, metric, i_pixel, a, sa, b, sb, offa=0, offb=0):
        return self.lib.xo_pixel_cmp(metric, i_pixel, _ptr(a, u8p, offa), sa, _ptr(b, u8p, offb), sb)

    def cost_mv_table(self, qp):
        t = np.zeros(4 * 4 * 2048 + 1, np.int16)
        self.lib.xo_cost_mv_table(qp, _ptr(t, i16p))
        return t

    def new_plane(self, g, fill=0):
        """padded luma plane buffer; pixel (0,0) is at flat index g.origin"""
        return np.full(g.plane_size, fill, np.uint8)

    def plane_from_picture(self, g, pic):
        """pic: (height,width) uint8 -> padded plane with mod16 + 32px border replication done by the oracle"""
        p = self.new_plane(g)
        v = p[g.origin - 0:].view()
        for y in range(g.height):
            o = g.origin + y * g.stride
            p[o:o + g.width] = pic[y]
        self.lib.xo_frame_expand_border(C.byref(g), _ptr(p, u8p, g.origin))
        return p

    def frame_filter(self, g, plane, sub8x8=0, want_integral=True):
        h = np.zeros(g.plane_size, np.uint8)
        v = np.zeros(g.plane_size, np.uint8)
        c = np.zeros(g.plane_size, np.uint8)
        integ = np.zeros(g.plane_size << sub8x8, np.uint16) if want_integral else None
        self.lib.xo_frame_filter(C.byref(g), _ptr(plane, u8p, g.origin), _ptr(h, u8p, g.origin), _ptr(v, u8p, g.origin),
                                 _ptr(c, u8p, g.origin), _ptr(integ, u16p, g.origin) if want_integral else None, sub8x8)
        return h, v, c, integ

    def init_lowres(self, g, plane):
        outs = [np.zeros(g.plane_size_lowres, np.uint8) for _ in range(4)]
        self.lib.xo_frame_init_lowres(C.byref(g), _ptr(plane, u8p, g.origin), *[_ptr(o, u8p, g.origin_lowres) for o in outs])
        return outs

    def me_search_fpel(self, g, fenc, fref, integral, mi):
        out = MeOut()
        self.lib.xo_me_search_fpel(C.byref(g), _ptr(fenc, u8p, g.origin), _ptr(fref, u8p, g.origin),
                                   _ptr(integral, u16p, g.origin) if integral is not None else None, C.byref(mi), C.byref(out))
        return out

    def me_search_fpel_batch(self, g, fenc, fref, integral, mis):
        """mis: ctypes array (MeIn * n) or list of MeIn"""
        n = len(mis)
        arr = mis if isinstance(mis, C.Array) else (MeIn * n)(*mis)
        outs = (MeOut * n)()
        self.lib.xo_me_search_fpel_batch(C.byref(g), _ptr(fenc, u8p, g.origin), _ptr(fref, u8p, g.origin),
                                         _ptr(integral, u16p, g.origin) if integral is not None else None, arr, n, outs)
        return outs

    def me_search_subpel(self, g, fenc, planes4, integral, mi, subme, mbcmp_satd):
        out = MeOut()
        arr = (u8p * 4)(*[_ptr(p, u8p, g.origin) for p in planes4])
        self.lib.xo_me_search_subpel(C.byref(g), _ptr(fenc, u8p, g.origin), arr,
                                     _ptr(integral, u16p, g.origin) if integral is not None else None, C.byref(mi),
                                     subme, mbcmp_satd, C.byref(out))
        return out


    def me_search_subpel_chroma(self, g, fenc, planes4, integral, chroma, mi, subme, mbcmp_satd):
        """chroma: (fenc_u, fenc_v, fref_u, fref_v) 2-D padded chroma planes (16-px borders, as Frame.download(PLANE_CB) returns them)"""
        out = MeOut()
        arr = (u8p * 4)(*[_ptr(p, u8p, g.origin) for p in planes4])
        sc = chroma[0].shape[1]
        ch = Chroma(*[a.ctypes.data + 16 * sc + 16 for a in chroma], sc)
        self.lib.xo_me_search_subpel_chroma(C.byref(g), _ptr(fenc, u8p, g.origin), arr,
                                            _ptr(integral, u16p, g.origin) if integral is not None else None, C.byref(ch), C.byref(mi),
                                            subme, mbcmp_satd, C.byref(out))
        return out

    # whole-frame metrics; a, b: 2-D uint8 arrays (row stride = shape[1])
    def frame_ssd(self, a, b, width, height):
        return int(self.lib.xo_frame_ssd(_ptr(a), a.shape[1], _ptr(b), b.shape[1], width, height))

    def frame_ssim(self, a, b, width, height):
        return float(self.lib.xo_frame_ssim(_ptr(a), a.shape[1], _ptr(b), b.shape[1], width, height))

    def frame_ssim_sums(self, a, b, width, height):
        sums = np.zeros((height // 4, width // 4, 4), np.int32)
        self.lib.xo_frame_ssim_sums(_ptr(a), a.shape[1], _ptr(b), b.shape[1], width, height, _ptr(sums, i32p))
        return sums

    def frame_mb_energy(self, g, y, u, v):
        """y: padded luma plane (flat); u, v: 2-D chroma"""
        out = np.zeros(g.mb_width * g.mb_height, np.uint32)
        self.lib.xo_frame_mb_energy(C.byref(g), _ptr(y, u8p, g.origin), _ptr(u), _ptr(v), u.shape[1], out.ctypes.data_as(C.POINTER(C.c_uint32)))
        return out

    def frame_mb_hadamard_ac(self, g, y):
        out = np.zeros(g.mb_width * g.mb_height, np.uint64)
        self.lib.xo_frame_mb_hadamard_ac(C.byref(g), _ptr(y, u8p, g.origin), out.ctypes.data_as(C.POINTER(C.c_uint64)))
        return out

    def frame_aq(self, g, y, u, v, strength):
        n = g.mb_width * g.mb_height
        qp, inv = np.zeros(n, np.float32), np.zeros(n, np.uint16)
        self.lib.xo_frame_aq(C.byref(g), _ptr(y, u8p, g.origin), _ptr(u), _ptr(v), u.shape[1], strength, qp.ctypes.data_as(C.POINTER(C.c_float)), _ptr(inv, u16p))
        return qp, inv

    def me_refine_qpel(self, g, fenc, planes4, chroma, mi, subme, mbcmp_satd, mv, cost):
        out = MeOut()
        arr = (u8p * 4)(*[_ptr(p, u8p, g.origin) for p in planes4])
        chp = None
        if chroma is not None:
            sc = chroma[0].shape[1]
            chp = C.byref(Chroma(*[a.ctypes.data + 16 * sc + 16 for a in chroma], sc))
        mvv = np.array(mv, np.int16)
        self.lib.xo_me_refine_qpel(C.byref(g), _ptr(fenc, u8p, g.origin), arr, chp, C.byref(mi), subme, mbcmp_satd, _ptr(mvv, i16p), int(cost), C.byref(out))
        return out

    def me_refine_bidir_satd(self, g, fenc, planes0, planes1, mi, mvp0, mvp1, weight, mbcmp_satd, mv0, mv1):
        """-> (mv0, mv1, best cost or -1)"""
        a0 = (u8p * 4)(*[_ptr(p, u8p, g.origin) for p in planes0])
        a1 = (u8p * 4)(*[_ptr(p, u8p, g.origin) for p in planes1])
        p0, p1, v0, v1 = (np.array(x, np.int16) for x in (mvp0, mvp1, mv0, mv1))
        c = self.lib.xo_me_refine_bidir_satd(C.byref(g), _ptr(fenc, u8p, g.origin), a0, a1, C.byref(mi), _ptr(p0, i16p), _ptr(p1, i16p), weight, mbcmp_satd,
                                             _ptr(v0, i16p), _ptr(v1, i16p))
        return (int(v0[0]), int(v0[1])), (int(v1[0]), int(v1[1])), c

    def residual_inter_mb(self, rin, fy, fu, fv, py, pu, pv):
        """x264_macroblock_encode's inter branch on one macroblock -> (ResidOut, rec_y, rec_u, rec_v); inputs are contiguous uint8 tiles"""
        o = ResidOut()
        ry, ru, rv = py.copy(), pu.copy(), pv.copy()
        self.lib.xo_residual_inter_mb(C.byref(rin), _ptr(fy), _ptr(fu), _ptr(fv), _ptr(ry), _ptr(ru), _ptr(rv), C.byref(o))
        return o, ry, ru, rv

    def intra_mb_costs(self, iin, fy, fu, fv, nby, nbu, nbv):
        """Intra16x16 + chroma mode costs from neighbour pixels (nb*: corner, row above, left column) -> IntraOut"""
        o = IntraOut()
        self.lib.xo_intra_mb_costs(C.byref(iin), _ptr(fy), _ptr(fu), _ptr(fv), _ptr(nby), _ptr(nbu), _ptr(nbv), C.byref(o))
        return o

    def predict(self, chroma, mode, nb):
        n = 8 if chroma else 16
        out = np.zeros((n, n), np.uint8)
        (self.lib.xo_predict_8x8c if chroma else self.lib.xo_predict_16x16)(mode, _ptr(nb), _ptr(out))
        return out

    def residual_intra16_mb(self, rin, mode16, mode_chroma, fy, fu, fv, nby, nbu, nbv):
        """one I_16x16 macroblock through x264_macroblock_encode -> (ResidOut, luma_dc int16[16], rec_y, rec_u, rec_v)"""
        ry, ru, rv = np.zeros(256, np.uint8), np.zeros(64, np.uint8), np.zeros(64, np.uint8)
        dc = np.zeros(16, np.int16)
        o = ResidOut()
        self.lib.xo_residual_intra16_mb(C.byref(rin), mode16, mode_chroma, _ptr(fy), _ptr(fu), _ptr(fv), _ptr(nby), _ptr(nbu), _ptr(nbv),
                                        _ptr(ry), _ptr(ru), _ptr(rv), C.byref(o), _ptr(dc, i16p))
        return o, dc, ry, ru, rv

    def probe_skip_mb(self, rin, fy, fu, fv, py, pu, pv):
        """x264_macroblock_probe_skip with the prediction supplied -> 0/1"""
        return int(self.lib.xo_probe_skip_mb(C.byref(rin), _ptr(fy), _ptr(fu), _ptr(fv), _ptr(py), _ptr(pu), _ptr(pv)))

    def frame_deblock(self, g, info, y, u, v):
        """y: padded luma plane (flat, pixel 0,0 at g.origin); u, v: 2-D chroma arrays (contiguous).  Filtered in place."""
        d = deblock_in(info)
        self.lib.xo_frame_deblock(C.byref(g), C.byref(d), _ptr(y, u8p, g.origin), _ptr(u), _ptr(v), u.shape[1])

    def lowres_frame_cost(self, g, fenc4, fref0_4, fref1_4, p0, p1, b, state, me_method=ME_HEX, me_range=16, mbcmp_satd=1,
                          fpel_satd=0, weighted=0, do_search=(1, 1), b_intra_calculated=0, vbv=False, inv_qscale=None, row_satd=None):
        """x264_slicetype_frame_cost on lowres planes.  state: dict of per-frame lookahead arrays that persist between calls —
        'mvs0','mvs1' int16[n_mb,2]; 'costs0','costs1' int32[n_mb]; 'intra' uint16[n_mb]; 'ref1_mvs' int16[n_mb,2] (B only).
        Updated in place.  Returns LowresOut (score is the raw sum for the port; the reference scales B scores, see xo.h).
        vbv: the rc.i_vbv_buffer_size form (every block, row_satd int32[mb_height] filled, inv_qscale uint16[n_mb] or None = AQ off)."""
        li = LowresIn(p0, p1, b, me_method, me_range, mbcmp_satd, fpel_satd, weighted, (C.c_int * 2)(*do_search), b_intra_calculated)
        mk = lambda planes: (u8p * 4)(*[_ptr(p, u8p, g.origin_lowres) for p in planes])
        out = LowresOut()
        args = (C.byref(g), C.byref(li), mk(fenc4), mk(fref0_4), mk(fref1_4),
                _ptr(state["mvs0"], i16p), _ptr(state["costs0"], i32p), _ptr(state["mvs1"], i16p),
                _ptr(state["costs1"], i32p), _ptr(state["ref1_mvs"], i16p), _ptr(state["intra"], u16p), C.byref(out))
        if vbv or inv_qscale is not None:
            self.lib.xo_lowres_frame_cost_vbv(*args, _ptr(inv_qscale, u16p) if inv_qscale is not None else None, _ptr(row_satd, i32p) if vbv else None)
        else:
            self.lib.xo_lowres_frame_cost(*args)
        return out


def deblock_in(info):
    """info: dict from helpers.make_deblock_info -> (DeblockIn, keepalive)"""
    d = DeblockIn(info["alpha_c0_offset"], info["beta_offset"], info["chroma_qp_offset"], info["b_slice_b"], info["b_psub8x8"],
                  info["b_cavlc_8x8dct"])
    d.type, d.qp, d.transform8x8, d.nnz = (info[k].ctypes.data for k in ("type", "qp", "transform8x8", "nnz"))
    d.ref[0], d.ref[1] = info["ref0"].ctypes.data, info["ref1"].ctypes.data
    d.mv[0], d.mv[1] = info["mv0"].ctypes.data, info["mv1"].ctypes.data
    return d


def lowres_state(g):
    n = g.mb_width * g.mb_height
    return {"mvs0": np.zeros((n, 2), np.int16), "mvs1": np.zeros((n, 2), np.int16), "costs0": np.zeros(n, np.int32),
            "costs1": np.zeros(n, np.int32), "intra": np.zeros(n, np.uint16), "ref1_mvs": np.zeros((n, 2), np.int16)}


_cache = {}


def port():
    if "port" not in _cache:
        _cache["port"] = Oracle(PORT_SO)
    return _cache["port"]


def have_ref():
    return os.path.exists(REF_SO)


def ref():
    if "ref" not in _cache:
        _cache["ref"] = Oracle(REF_SO)
    return _cache["ref"]


def mv_limits_fpel(g, mb_x, mb_y, mv_range=512):
    """h->mb.mv_{min,max}_{spel,fpel} for a progressive, single-thread encode (S/encoder/analyse.c:258-304)."""
    fr = 4 * mv_range
    clip = lambda v: max(-fr, min(fr - 1, v))
    mn = [4 * (-16 * mb_x - 24), 4 * (-16 * mb_y - 24)]
    mx = [4 * (16 * (g.mb_width - mb_x - 1) + 24), 4 * (16 * (g.mb_height - mb_y - 1) + 24)]
    min_spel = [clip(mn[0]), max(mn[1], max(4 * (-512 + 8), -fr))]
    min_spel[1] = min(min_spel[1], fr)
    max_spel = [clip(mx[0]), clip(mx[1])]
    min_fpel = [(min_spel[0] >> 2) + 5, (min_spel[1] >> 2) + 5]
    max_fpel = [(max_spel[0] >> 2) - 5, (max_spel[1] >> 2) - 5]
    return min_fpel, max_fpel, min_spel, max_spel
