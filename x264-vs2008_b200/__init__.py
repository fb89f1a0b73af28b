"""x264-vs2008_b200 — Python face of the B200 back-end (ctypes over lib/libx264_cuda.so, the C ABI of
include/x264_cuda.h).  Used by tests/ and bench.py; the product itself is the shared library.

No fallback of any kind: if the library or a CUDA device is missing, calls raise."""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libx264_cuda.so")

PADH = PADV = 32
FRAME_HPEL, FRAME_INTEGRAL, FRAME_INTEGRAL4, FRAME_LOWRES, FRAME_CHROMA = 1, 2, 4, 8, 16
PLANE_FULL, PLANE_H, PLANE_V, PLANE_C, PLANE_LOWRES, PLANE_INTEGRAL, PLANE_INTEGRAL4, PLANE_CB, PLANE_CR = 0, 1, 2, 3, 4, 8, 9, 10, 11
RESID_8x8DCT, RESID_DECIMATE = 1, 2
MC_JOB = np.dtype([("bx", "<i2"), ("by", "<i2"), ("mvx", "<i2"), ("mvy", "<i2"), ("w", "u1"), ("h", "u1"), ("reserved", "u1", (2,))], align=True)
assert MC_JOB.itemsize == 12
RESID_JOB = np.dtype([("mb_x", "<i2"), ("mb_y", "<i2"), ("qp", "u1"), ("chroma_qp", "u1"), ("flags", "u1"), ("reserved", "u1")], align=True)
MB_COEFFS = np.dtype([("luma", "<i2", (256,)), ("chroma_ac", "<i2", (8, 16)), ("chroma_dc", "<i2", (2, 4)), ("nnz", "u1", (27,)),
                      ("cbp_luma", "u1"), ("cbp_chroma", "u1"), ("reserved", "u1", (3,))], align=True)
assert RESID_JOB.itemsize == 8 and MB_COEFFS.itemsize == 816
INTRA16_JOB = np.dtype([("mb_x", "<i2"), ("mb_y", "<i2"), ("qp", "u1"), ("chroma_qp", "u1"), ("mode16", "u1"), ("mode_chroma", "u1"), ("flags", "u1"),
                        ("reserved", "u1", (3,))], align=True)
MB_COEFFS_I16 = np.dtype([("c", MB_COEFFS), ("luma_dc", "<i2", (16,))], align=True)
assert INTRA16_JOB.itemsize == 12 and MB_COEFFS_I16.itemsize == 848
SKIP_JOB = np.dtype([("mb_x", "<i2"), ("mb_y", "<i2"), ("mvx", "<i2"), ("mvy", "<i2"), ("qp", "u1"), ("chroma_qp", "u1"), ("flags", "u1"),
                     ("reserved", "u1")], align=True)
assert SKIP_JOB.itemsize == 12
SKIP_PRED_IN_FDEC, SKIP_STORE_PRED = 1, 2
INTRA_JOB = np.dtype([("mb_x", "<i2"), ("mb_y", "<i2"), ("neighbour", "u1"), ("flags", "u1"), ("lambda", "<u2")], align=True)
INTRA_RESULT = np.dtype([("cost16", "<i4", (7,)), ("cost_chroma", "<i4", (7,)), ("best16", "<i4"), ("best_chroma", "<i4"), ("mode16", "u1"),
                         ("mode_chroma", "u1"), ("reserved", "u1", (2,))], align=True)
assert INTRA_JOB.itemsize == 8 and INTRA_RESULT.itemsize == 68
INTRA_SATD, INTRA_SLICE_B = 1, 2
MB_LEFT, MB_TOP, MB_TOPRIGHT, MB_TOPLEFT = 1, 2, 4, 8
ME_SEEDED, ME_TESA, ME_FPEL_SATD = 1, 2, 4
ME_MAX_MVC = 12


class Geom(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("width", "height", "mb_width", "mb_height", "stride", "lines",
                                       "stride_lowres", "width_lowres", "lines_lowres", "flags")]


# numpy mirrors of x264_cuda_me_job_t / x264_cuda_me_result_t (include/x264_cuda.h)
ME_JOB = np.dtype([("bx", "<i2"), ("by", "<i2"), ("i_pixel", "u1"), ("qp", "u1"), ("i_mvc", "u1"), ("flags", "u1"),
                   ("mvp", "<i2", (2,)), ("mv_min_fpel", "<i2", (2,)), ("mv_max_fpel", "<i2", (2,)),
                   ("seed_mv", "<i2", (2,)), ("seed_cost", "<i4"), ("mvc", "<i2", (ME_MAX_MVC, 2)),
                   ("mv_min_spel", "<i2", (2,)), ("mv_max_spel", "<i2", (2,))], align=True)
MC_BI_JOB = np.dtype([("bx", "<i2"), ("by", "<i2"), ("mv0", "<i2", (2,)), ("mv1", "<i2", (2,)), ("w", "u1"), ("h", "u1"), ("weight", "u1"),
                      ("reserved", "u1")])
assert MC_BI_JOB.itemsize == 16
BIDIR_JOB = np.dtype([("bx", "<i2"), ("by", "<i2"), ("i_pixel", "u1"), ("qp", "u1"), ("weight", "u1"), ("flags", "u1"), ("mv0", "<i2", (2,)),
                      ("mv1", "<i2", (2,)), ("mvp0", "<i2", (2,)), ("mvp1", "<i2", (2,)), ("mv_min_spel", "<i2", (2,)), ("mv_max_spel", "<i2", (2,))])
BIDIR_RESULT = np.dtype([("mv0", "<i2", (2,)), ("mv1", "<i2", (2,)), ("cost", "<i4")])
assert BIDIR_JOB.itemsize == 32 and BIDIR_RESULT.itemsize == 12
GRID_JOB = np.dtype([("mb_x", "<i2"), ("mb_y", "<i2"), ("cx", "<i2"), ("cy", "<i2"), ("mv_min_fpel", "<i2", (2,)), ("mv_max_fpel", "<i2", (2,)),
                     ("part_mask", "<u2"), ("reserved", "<u2")])
assert GRID_JOB.itemsize == 20


def grid_w(radius):
    return (2 * radius + 1 + 3) & ~3


def grid_h(radius):
    return 2 * radius + 1


ME_FINAL = np.dtype([("mv", "<i2", (2,)), ("cost", "<i4"), ("cost_mv", "<i4"), ("bmx", "<i2"), ("bmy", "<i2")], align=True)
ME_METHOD_DIA, ME_METHOD_HEX, ME_METHOD_UMH, ME_METHOD_TESA, ME_METHOD_SEEDED, ME_METHOD_REFINE_QPEL = 0, 1, 2, 4, 8, 16
ME_MBCMP_SATD = 8
ME_CHROMA = 32
LOWRES_WEIGHTED_BIPRED = 16
LOWRES_VBV = 32
ME_RESULT = np.dtype([("bmx", "<i2"), ("bmy", "<i2"), ("bcost", "<i4"), ("seed_mx", "<i2"), ("seed_my", "<i2"),
                      ("seed_cost", "<i4")], align=True)
ME_MB_PARTS, ME_MB_MVC = 9, 4
# partition p of a macroblock job: (i_pixel, x offset, y offset)
ME_MB_PART_GEOM = [(0, 0, 0), (1, 0, 0), (1, 0, 8), (2, 0, 0), (2, 8, 0), (3, 0, 0), (3, 8, 0), (3, 0, 8), (3, 8, 8)]
ME_MB_JOB = np.dtype([("mb_x", "<i2"), ("mb_y", "<i2"), ("part_mask", "<u2"), ("qp", "u1"), ("flags", "u1"),
                      ("mv_min_fpel", "<i2", (2,)), ("mv_max_fpel", "<i2", (2,)), ("i_mvc", "u1", (ME_MB_PARTS,)),
                      ("reserved", "u1", (3,)), ("mvp", "<i2", (ME_MB_PARTS, 2)), ("mvc", "<i2", (ME_MB_PARTS, ME_MB_MVC, 2)),
                      ("seed_mv", "<i2", (ME_MB_PARTS, 2)), ("seed_cost", "<i4", (ME_MB_PARTS,))], align=True)
ME_MB_RESULT = np.dtype([("part", ME_RESULT, (ME_MB_PARTS,))], align=True)
assert ME_JOB.itemsize == 84 and ME_RESULT.itemsize == 16 and ME_FINAL.itemsize == 16
assert ME_MB_JOB.itemsize == 280 and ME_MB_RESULT.itemsize == 144

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libx264_cuda.so is not built (run __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        vp, ip = C.c_void_p, C.c_int
        L.x264_cuda_open.argtypes = [C.POINTER(vp), ip]
        L.x264_cuda_close.argtypes = [vp]
        L.x264_cuda_error.argtypes = [vp]
        L.x264_cuda_error.restype = C.c_char_p
        L.x264_cuda_set_stream.argtypes = [vp, vp]
        L.x264_cuda_get_stream.argtypes = [vp]
        L.x264_cuda_get_stream.restype = vp
        L.x264_cuda_synchronize.argtypes = [vp]
        L.x264_cuda_set_blocking_wait.argtypes = [vp, ip]
        L.x264_cuda_launch_count.argtypes = [vp]
        L.x264_cuda_launch_count.restype = C.c_longlong
        L.x264_cuda_sm_count.argtypes = [vp]
        L.x264_cuda_measure_int_pipe.argtypes = [vp, C.POINTER(C.c_double)]
        L.x264_cuda_frame_new.argtypes = [vp, ip, ip, ip]
        L.x264_cuda_frame_new.restype = vp
        L.x264_cuda_frame_delete.argtypes = [vp]
        L.x264_cuda_frame_geometry.argtypes = [vp, C.POINTER(Geom)]
        L.x264_cuda_frame_plane.argtypes = [vp, ip]
        L.x264_cuda_frame_plane.restype = vp
        L.x264_cuda_frame_upload.argtypes = [vp, vp, vp, ip, ip, ip]
        L.x264_cuda_frame_upload_dev.argtypes = [vp, vp, vp, ip, ip, ip]
        L.x264_cuda_frame_upload_chroma.argtypes = [vp, vp, ip, vp, ip, ip, ip]
        L.x264_cuda_set_quant_preset.argtypes = [vp, ip]
        L.x264_cuda_sad_grid_quad.argtypes = [vp, vp, vp, ip, vp, ip, vp, ip]
        L.x264_cuda_fence_record.argtypes = [vp]
        L.x264_cuda_fence_record.restype = vp
        L.x264_cuda_fence_wait.argtypes = [vp, vp]
        L.x264_cuda_mc_blocks.argtypes = [vp, vp, vp, vp, ip]
        L.x264_cuda_mc_blocks_dev.argtypes = [vp, vp, vp, vp, ip]
        L.x264_cuda_mc_blocks_bi.argtypes = [vp, vp, vp, vp, vp, ip]
        L.x264_cuda_mc_blocks_bi_dev.argtypes = [vp, vp, vp, vp, vp, ip]
        L.x264_cuda_block_residual.argtypes = [vp, ip, ip, vp, vp, vp, vp, vp, vp, vp, vp]
        L.x264_cuda_block_dc.argtypes = [vp, ip, vp, vp, vp, vp, vp, vp, vp]
        L.x264_cuda_residual_inter.argtypes = [vp, vp, vp, vp, ip, vp]
        L.x264_cuda_residual_inter_dev.argtypes = [vp, vp, vp, vp, ip, vp]
        L.x264_cuda_residual_intra16.argtypes = [vp, vp, vp, vp, ip, vp]
        L.x264_cuda_residual_intra16_dev.argtypes = [vp, vp, vp, vp, ip, vp]
        L.x264_cuda_frame_download.argtypes = [vp, vp, ip, vp, ip]
        for n in ("x264_cuda_frame_expand_border", "x264_cuda_frame_expand_border_mod16", "x264_cuda_frame_filter", "x264_cuda_frame_init_lowres"):
            if hasattr(L, n):
                getattr(L, n).argtypes = [vp, vp]
        L.x264_cuda_set_cost_mv.argtypes = [vp, ip, vp]
        L.x264_cuda_host_cost_mv.argtypes = [ip, vp]
        L.x264_cuda_host_lambda.argtypes = [ip]
        L.x264_cuda_host_lambda2.argtypes = [ip]
        L.x264_cuda_probe_skip.argtypes = [vp, vp, vp, vp, vp, ip, vp]
        L.x264_cuda_intra_mb_costs.argtypes = [vp, vp, vp, vp, ip, vp]
        L.x264_cuda_intra_mb_costs_dev.argtypes = [vp, vp, vp, vp, ip, vp]
        L.x264_cuda_probe_skip_dev.argtypes = [vp, vp, vp, vp, vp, ip, ip, ip, vp]
        L.x264_cuda_me_search.argtypes = [vp, vp, vp, ip, vp, ip, vp]
        L.x264_cuda_me_search_dev.argtypes = [vp, vp, vp, ip, vp, ip, vp]
        L.x264_cuda_me_search_mb.argtypes = [vp, vp, vp, ip, vp, ip, vp]
        L.x264_cuda_me_search_small.argtypes = [vp, vp, vp, ip, ip, ip, vp, ip, vp]
        L.x264_cuda_me_search_small_dev.argtypes = [vp, vp, vp, ip, ip, ip, vp, ip, vp]
        L.x264_cuda_block_cmp.argtypes = [vp, ip, ip, ip, vp, vp, vp]
        L.x264_cuda_me_search_mb_dev.argtypes = [vp, vp, vp, ip, vp, ip, vp]
        L.x264_cuda_me_finish.argtypes = [vp, vp, vp, ip, vp, vp, vp]
        L.x264_cuda_me_finish.restype = None
        L.x264_cuda_me_refine_bidir.argtypes = [vp, vp, vp, vp, vp, ip, vp]
        L.x264_cuda_me_refine_bidir_dev.argtypes = [vp, vp, vp, vp, vp, ip, vp]
        L.x264_cuda_sad_grid.argtypes = [vp, vp, vp, ip, vp, ip, vp]
        L.x264_cuda_sad_grid_dev.argtypes = [vp, vp, vp, ip, vp, ip, vp]
        L.x264_cuda_host_esa_replay.argtypes = [vp, ip, ip, ip, vp, ip, vp, vp]
        L.x264_cuda_frame_lookahead_alloc.argtypes = [vp, vp, ip]
        L.x264_cuda_frame_lookahead_get.argtypes = [vp, vp, ip, ip, vp, vp, vp]
        L.x264_cuda_frame_lookahead_set.argtypes = [vp, vp, ip, ip, vp, vp, vp]
        L.x264_cuda_lowres_frame_cost.argtypes = [vp, vp, vp, vp, vp, vp]
        L.x264_cuda_lowres_frame_cost_batch.argtypes = [vp, ip, vp, vp, vp, vp, vp]
        L.x264_cuda_lowres_frame_cost_rc.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp]
        L.x264_cuda_lowres_frame_cost_batch_rc.argtypes = [vp, ip, vp, vp, vp, vp, vp, vp, vp]
        L.x264_cuda_frame_deblock.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
        L.x264_cuda_frame_deblock_dev.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
        L.x264_cuda_frame_ssd.argtypes = [vp, vp, vp, ip, ip, ip, ip, ip, vp]
        L.x264_cuda_frame_ssim_sums.argtypes = [vp, vp, vp, ip, ip, ip, ip, ip, vp]
        L.x264_cuda_host_ssim_end.argtypes = [vp, ip, ip]
        L.x264_cuda_host_ssim_end.restype = C.c_float
        L.x264_cuda_frame_mb_energy.argtypes = [vp, vp, vp]
        L.x264_cuda_host_aq.argtypes = [vp, ip, C.c_float, vp, vp]
        L.x264_cuda_frame_mb_hadamard_ac.argtypes = [vp, vp, vp]
        L.x264_cuda_host_alloc.argtypes = [C.c_size_t]
        L.x264_cuda_host_alloc.restype = vp
        L.x264_cuda_host_free.argtypes = [vp]
        L.x264_cuda_host_register.argtypes = [vp, C.c_size_t]
        L.x264_cuda_host_unregister.argtypes = [vp]
        _lib = L
    return _lib


class CudaError(RuntimeError):
    pass


def host_esa_replay(grid_part, radius, cx, cy, job, me_range, cost_table):
    """x264_me_search_ref's predictor stage + ESA loop on one partition's grid plane -> ME_RESULT record or None (outside grid)"""
    res = np.zeros(1, ME_RESULT)
    j = np.ascontiguousarray(job)
    gp = np.ascontiguousarray(grid_part)
    rc = lib().x264_cuda_host_esa_replay(gp.ctypes.data, radius, cx, cy, j.ctypes.data, me_range, cost_table.ctypes.data, res.ctypes.data)
    return res[0] if rc == 0 else None


def host_aq(energy, aq_strength=1.0):
    """x264_adaptive_quant_frame's float part -> (f_qp_offset, i_inv_qscale_factor)"""
    e = np.ascontiguousarray(energy, np.uint32)
    qp, inv = np.zeros(len(e), np.float32), np.zeros(len(e), np.uint16)
    lib().x264_cuda_host_aq(e.ctypes.data, len(e), aq_strength, qp.ctypes.data, inv.ctypes.data)
    return qp, inv


def host_cost_mv(qp):
    t = np.zeros(4 * 4 * 2048 + 1, np.int16)
    lib().x264_cuda_host_cost_mv(qp, t.ctypes.data)
    return t


class Frame:
    def __init__(self, ctx, width, height, flags):
        self.ctx = ctx
        self.h = lib().x264_cuda_frame_new(ctx.h, width, height, flags)
        if not self.h:
            raise CudaError(ctx.error())
        self.g = Geom()
        lib().x264_cuda_frame_geometry(self.h, C.byref(self.g))

    def close(self):
        if self.h:
            lib().x264_cuda_frame_delete(self.h)
            self.h = None

    def plane_ptr(self, plane):
        return lib().x264_cuda_frame_plane(self.h, plane)

    def upload(self, pic, cols=None, rows=None):
        """pic: 2-D uint8 numpy array whose [0,0] is pixel (0,0) (host memory)"""
        assert pic.dtype == np.uint8 and pic.strides[1] == 1
        rows = pic.shape[0] if rows is None else rows
        cols = pic.shape[1] if cols is None else cols
        self.ctx.check(lib().x264_cuda_frame_upload(self.ctx.h, self.h, pic.ctypes.data, pic.strides[0], cols, rows))

    def upload_chroma(self, cb, cr):
        for plane, pic in ((PLANE_CB, cb), (PLANE_CR, cr)):
            assert pic.dtype == np.uint8 and pic.strides[1] == 1
            self.ctx.check(lib().x264_cuda_frame_upload_chroma(self.ctx.h, self.h, plane, pic.ctypes.data, pic.strides[0],
                                                               pic.shape[1], pic.shape[0]))

    def upload_dev(self, dptr, stride, cols, rows):
        self.ctx.check(lib().x264_cuda_frame_upload_dev(self.ctx.h, self.h, dptr, stride, cols, rows))

    def expand_border(self):
        self.ctx.check(lib().x264_cuda_frame_expand_border(self.ctx.h, self.h))

    def expand_border_mod16(self):
        self.ctx.check(lib().x264_cuda_frame_expand_border_mod16(self.ctx.h, self.h))

    def filter(self):
        self.ctx.check(lib().x264_cuda_frame_filter(self.ctx.h, self.h))

    def init_lowres(self):
        self.ctx.check(lib().x264_cuda_frame_init_lowres(self.ctx.h, self.h))

    def lookahead_alloc(self, n_dist):
        """lowres_mvs / lowres_mv_costs / i_intra_cost of this frame (S/common/frame.c:85-98), zero-filled"""
        self.ctx.check(lib().x264_cuda_frame_lookahead_alloc(self.ctx.h, self.h, n_dist))

    def lookahead_get(self, lst, dist):
        n = self.g.mb_width * self.g.mb_height
        mvs, costs, intra = np.zeros((n, 2), np.int16), np.zeros(n, np.int32), np.zeros(n, np.uint16)
        self.ctx.check(lib().x264_cuda_frame_lookahead_get(self.ctx.h, self.h, lst, dist, mvs.ctypes.data, costs.ctypes.data, intra.ctypes.data))
        return mvs, costs, intra

    def lookahead_set(self, lst, dist, mvs=None, costs=None, intra=None):
        a = [None if x is None else np.ascontiguousarray(x, t) for x, t in ((mvs, np.int16), (costs, np.int32), (intra, np.uint16))]
        self.ctx.check(lib().x264_cuda_frame_lookahead_set(self.ctx.h, self.h, lst, dist, *[None if x is None else x.ctypes.data for x in a]))

    def download(self, plane):
        """whole padded plane as a 2-D array (rows -32.., cols -32..)"""
        g = self.g
        lowres = PLANE_LOWRES <= plane < PLANE_LOWRES + 4
        chroma = plane in (PLANE_CB, PLANE_CR)
        lines = g.lines_lowres if lowres else g.lines // 2 if chroma else g.lines
        pad = 16 if chroma else PADH
        w = (g.width_lowres if lowres else g.mb_width * 8 if chroma else g.mb_width * 16) + 2 * pad
        out = np.zeros((lines + 2 * pad, w), np.uint16 if plane in (PLANE_INTEGRAL, PLANE_INTEGRAL4) else np.uint8)
        self.ctx.check(lib().x264_cuda_frame_download(self.ctx.h, self.h, plane, out.ctypes.data, w))
        return out


class Context:
    def __init__(self, device=0):
        self.h = C.c_void_p()
        if lib().x264_cuda_open(C.byref(self.h), device) != 0:
            raise CudaError(lib().x264_cuda_error(None).decode())
        self._qps = set()

    def close(self):
        if self.h:
            lib().x264_cuda_close(self.h)
            self.h = C.c_void_p()

    def error(self):
        return lib().x264_cuda_error(self.h).decode()

    def check(self, rc):
        if rc != 0:
            raise CudaError(self.error())

    def set_stream(self, s):
        self.check(lib().x264_cuda_set_stream(self.h, s))

    def synchronize(self):
        self.check(lib().x264_cuda_synchronize(self.h))

    def set_blocking_wait(self, on):
        self.check(lib().x264_cuda_set_blocking_wait(self.h, int(on)))

    def launches(self):
        return lib().x264_cuda_launch_count(self.h)

    def measure_int_pipe(self):
        v = C.c_double()
        self.check(lib().x264_cuda_measure_int_pipe(self.h, C.byref(v)))
        return v.value

    def sm_count(self):
        return lib().x264_cuda_sm_count(self.h)

    def frame(self, width, height, flags=0):
        return Frame(self, width, height, flags)

    def set_cost_mv(self, qp, table=None):
        t = host_cost_mv(qp) if table is None else np.ascontiguousarray(table, np.int16)
        self.check(lib().x264_cuda_set_cost_mv(self.h, qp, t.ctypes.data))
        self._qps.add(qp)

    def me_search(self, fenc, fref, me_range, jobs):
        """jobs: numpy array of ME_JOB (host) -> numpy array of ME_RESULT"""
        jobs = np.ascontiguousarray(jobs, ME_JOB)
        for qp in np.unique(jobs["qp"]):
            if int(qp) not in self._qps:
                self.set_cost_mv(int(qp))
        res = np.zeros(len(jobs), ME_RESULT)
        self.check(lib().x264_cuda_me_search(self.h, fenc.h, fref.h, me_range, jobs.ctypes.data, len(jobs), res.ctypes.data))
        return res

    def mc_blocks(self, fref, fdec, jobs):
        jobs = np.ascontiguousarray(jobs, MC_JOB)
        self.check(lib().x264_cuda_mc_blocks(self.h, fref.h, fdec.h, jobs.ctypes.data, len(jobs)))

    def mc_blocks_bi(self, fref0, fref1, fdec, jobs):
        assert jobs.dtype == MC_BI_JOB
        self.check(lib().x264_cuda_mc_blocks_bi(self.h, fref0.h, fref1.h, fdec.h, jobs.ctypes.data, len(jobs)))

    def me_refine_bidir(self, fenc, fref0, fref1, jobs):
        assert jobs.dtype == BIDIR_JOB
        out = np.zeros(len(jobs), BIDIR_RESULT)
        self.check(lib().x264_cuda_me_refine_bidir(self.h, fenc.h, fref0.h, fref1.h, jobs.ctypes.data, len(jobs), out.ctypes.data))
        return out

    def set_quant_preset(self, cqm):
        self.check(lib().x264_cuda_set_quant_preset(self.h, cqm))

    def block_residual(self, kind, fenc, pred, qp, cat):
        """packed blocks (n, 16|64) uint8 -> dict(dct, level, nz, recon)"""
        n, bs = fenc.shape[0], 64 if kind else 16
        fenc, pred = np.ascontiguousarray(fenc, np.uint8), np.ascontiguousarray(pred, np.uint8)
        qp, cat = np.ascontiguousarray(qp, np.uint8), np.ascontiguousarray(cat, np.uint8)
        dct, level = np.zeros((n, bs), np.int16), np.zeros((n, bs), np.int16)
        nz, recon = np.zeros(n, np.uint8), np.zeros((n, bs), np.uint8)
        self.check(lib().x264_cuda_block_residual(self.h, kind, n, fenc.ctypes.data, pred.ctypes.data, qp.ctypes.data, cat.ctypes.data,
                                                  dct.ctypes.data, level.ctypes.data, nz.ctypes.data, recon.ctypes.data))
        return dict(dct=dct, level=level, nz=nz, recon=recon)

    def block_dc(self, dc, qp, cat):
        n = dc.shape[0]
        dc, qp, cat = np.ascontiguousarray(dc, np.int16), np.ascontiguousarray(qp, np.uint8), np.ascontiguousarray(cat, np.uint8)
        fwd, level, deq, nz = np.zeros((n, 16), np.int16), np.zeros((n, 16), np.int16), np.zeros((n, 16), np.int16), np.zeros(n, np.uint8)
        self.check(lib().x264_cuda_block_dc(self.h, n, dc.ctypes.data, qp.ctypes.data, cat.ctypes.data, fwd.ctypes.data, level.ctypes.data,
                                            nz.ctypes.data, deq.ctypes.data))
        return dict(fwd=fwd, level=level, nz=nz, deq=deq)

    def intra_mb_costs(self, fenc, fdec, jobs):
        """Intra16x16 + chroma 8x8 candidate costs of the listed macroblocks from the neighbours in fdec -> INTRA_RESULT[n]"""
        jobs = np.ascontiguousarray(jobs, INTRA_JOB)
        out = np.zeros(len(jobs), INTRA_RESULT)
        self.check(lib().x264_cuda_intra_mb_costs(self.h, fenc.h, fdec.h, jobs.ctypes.data, len(jobs), out.ctypes.data))
        return out

    def probe_skip(self, fenc, fref, fdec, jobs):
        """x264_macroblock_probe_skip for a list of macroblocks -> uint8[n] (1 = skippable); fref / fdec may be None (see the header)"""
        jobs = np.ascontiguousarray(jobs, SKIP_JOB)
        out = np.zeros(len(jobs), np.uint8)
        self.check(lib().x264_cuda_probe_skip(self.h, fenc.h, fref.h if fref else None, fdec.h if fdec else None, jobs.ctypes.data,
                                              len(jobs), out.ctypes.data))
        return out

    def residual_inter(self, fenc, fdec, jobs):
        jobs = np.ascontiguousarray(jobs, RESID_JOB)
        out = np.zeros(len(jobs), MB_COEFFS)
        self.check(lib().x264_cuda_residual_inter(self.h, fenc.h, fdec.h, jobs.ctypes.data, len(jobs), out.ctypes.data))
        return out

    def residual_intra16(self, fenc, fdec, jobs):
        """I_16x16 macroblocks through x264_macroblock_encode, wavefront in list order (S/encoder/macroblock.c:184-270, :272-363)"""
        jobs = np.ascontiguousarray(jobs, INTRA16_JOB)
        out = np.zeros(len(jobs), MB_COEFFS_I16)
        self.check(lib().x264_cuda_residual_intra16(self.h, fenc.h, fdec.h, jobs.ctypes.data, len(jobs), out.ctypes.data))
        return out

    def me_search_small(self, fenc, fref, method, me_range, subme, jobs):
        jobs = np.ascontiguousarray(jobs, ME_JOB)
        for qp in np.unique(jobs["qp"]):
            if int(qp) not in self._qps:
                self.set_cost_mv(int(qp))
        res = np.zeros(len(jobs), ME_FINAL)
        self.check(lib().x264_cuda_me_search_small(self.h, fenc.h, fref.h, method, me_range, subme, jobs.ctypes.data, len(jobs),
                                                   res.ctypes.data))
        return res

    def block_cmp(self, metric, i_pixel, pix1, pix2):
        """pix1, pix2: (n, 16, 16) uint8 tiles -> int32[n]"""
        pix1, pix2 = np.ascontiguousarray(pix1, np.uint8), np.ascontiguousarray(pix2, np.uint8)
        out = np.zeros(pix1.shape[0], np.int32)
        self.check(lib().x264_cuda_block_cmp(self.h, metric, i_pixel, pix1.shape[0], pix1.ctypes.data, pix2.ctypes.data, out.ctypes.data))
        return out

    def me_search_mb(self, fenc, fref, me_range, jobs):
        """jobs: numpy array of ME_MB_JOB (host) -> numpy array of ME_MB_RESULT"""
        jobs = np.ascontiguousarray(jobs, ME_MB_JOB)
        for qp in np.unique(jobs["qp"]):
            if int(qp) not in self._qps:
                self.set_cost_mv(int(qp))
        res = np.zeros(len(jobs), ME_MB_RESULT)
        self.check(lib().x264_cuda_me_search_mb(self.h, fenc.h, fref.h, me_range, jobs.ctypes.data, len(jobs), res.ctypes.data))
        return res

    def lowres_frame_cost(self, fenc, fref0, fref1, p0, p1, b, me_method=1, me_range=16, flags=ME_MBCMP_SATD, do_search=(1, 1),
                          b_intra_calculated=0):
        """x264_slicetype_frame_cost (S/encoder/slicetype.c:248-355) -> (score, intra_mbs, intra_cost_sum); score is the raw sum"""
        pm = np.array([p0, p1, b, me_method, me_range, flags, do_search[0], do_search[1], b_intra_calculated], np.int32)
        res = np.zeros(4, np.int32)
        self.check(lib().x264_cuda_lowres_frame_cost(self.h, fenc.h, fref0.h, fref1.h, pm.ctypes.data, res.ctypes.data))
        return int(res[0]), int(res[1]), int(res[2])

    def lowres_frame_cost_rc(self, fenc, fref0, fref1, p0, p1, b, inv_qscale=None, vbv=True, me_method=1, me_range=16, flags=ME_MBCMP_SATD,
                             do_search=(1, 1), b_intra_calculated=0):
        """the rate-control forms of x264_slicetype_frame_cost (S/encoder/slicetype.c:300-316): -> (score, intra_mbs, intra_cost_sum,
        score_aq, row_satd int32[mb_height] or None).  inv_qscale: uint16[n_mb] (rc.i_aq_mode) or None; vbv: evaluate every block."""
        pm = np.array([p0, p1, b, me_method, me_range, flags | (LOWRES_VBV if vbv else 0), do_search[0], do_search[1], b_intra_calculated], np.int32)
        res = np.zeros(4, np.int32)
        rows = np.zeros(fenc.g.mb_height, np.int32) if vbv else None
        iq = np.ascontiguousarray(inv_qscale, np.uint16) if inv_qscale is not None else None
        self.check(lib().x264_cuda_lowres_frame_cost_rc(self.h, fenc.h, fref0.h, fref1.h, pm.ctypes.data, iq.ctypes.data if iq is not None else None,
                                                        res.ctypes.data, rows.ctypes.data if vbv else None))
        return int(res[0]), int(res[1]), int(res[2]), int(res[3]), rows

    def frame_deblock(self, fdec, info):
        """x264_frame_deblock (S/common/frame.c:621-799) in place on fdec's luma + chroma planes.  info: dict with the reference's
        per-macroblock arrays (type, qp, transform8x8, nnz[n,24], ref0/ref1 [2H,2W], mv0/mv1 [4H,4W,2]) and the slice parameters."""
        pm = np.array([info["alpha_c0_offset"], info["beta_offset"], info["chroma_qp_offset"], info["b_slice_b"], info["b_psub8x8"],
                       info["b_cavlc_8x8dct"]], np.int32)
        arr = {k: np.ascontiguousarray(info[k], t) for k, t in (("type", np.int8), ("qp", np.int8), ("transform8x8", np.int8), ("nnz", np.uint8),
                                                               ("ref0", np.int8), ("mv0", np.int16), ("ref1", np.int8), ("mv1", np.int16))}
        self.check(lib().x264_cuda_frame_deblock(self.h, fdec.h, pm.ctypes.data, *[arr[k].ctypes.data for k in
                                                 ("type", "qp", "transform8x8", "nnz", "ref0", "mv0", "ref1", "mv1")]))

    def frame_ssd(self, a, b, plane, width, height, x0=0, y0=0):
        out = np.zeros(1, np.int64)
        self.check(lib().x264_cuda_frame_ssd(self.h, a.h, b.h, plane, x0, y0, width, height, out.ctypes.data))
        return int(out[0])

    def frame_ssim(self, a, b, plane, width, height, x0=0, y0=0):
        """-> (x264_pixel_ssim_wxh value, sums[h4, w4, 4])"""
        sums = np.zeros((height // 4, width // 4, 4), np.int32)
        self.check(lib().x264_cuda_frame_ssim_sums(self.h, a.h, b.h, plane, x0, y0, width, height, sums.ctypes.data))
        return float(lib().x264_cuda_host_ssim_end(sums.ctypes.data, width // 4, height // 4)), sums

    def frame_mb_energy(self, f):
        out = np.zeros(f.g.mb_width * f.g.mb_height, np.uint32)
        self.check(lib().x264_cuda_frame_mb_energy(self.h, f.h, out.ctypes.data))
        return out

    def frame_mb_hadamard_ac(self, f):
        out = np.zeros(f.g.mb_width * f.g.mb_height, np.uint64)
        self.check(lib().x264_cuda_frame_mb_hadamard_ac(self.h, f.h, out.ctypes.data))
        return out

    def lowres_frame_cost_batch(self, evals, me_method=1, me_range=16, flags=ME_MBCMP_SATD):
        """evals: list of (fenc, fref0, fref1, p0, p1, b, do_search, b_intra_calculated) — independent evaluations, one launch.
        -> list of (score, intra_mbs, intra_cost_sum)"""
        n = len(evals)
        ptrs = [(C.c_void_p * n)(*[e[k].h for e in evals]) for k in range(3)]
        pm = np.array([[e[3], e[4], e[5], me_method, me_range, flags, e[6][0], e[6][1], e[7]] for e in evals], np.int32)
        res = np.zeros((n, 4), np.int32)
        self.check(lib().x264_cuda_lowres_frame_cost_batch(self.h, n, ptrs[0], ptrs[1], ptrs[2], pm.ctypes.data, res.ctypes.data))
        return [(int(r[0]), int(r[1]), int(r[2])) for r in res]

    def lowres_frame_cost_batch_rc(self, evals, inv_qscales=None, vbv=True, me_method=1, me_range=16, flags=ME_MBCMP_SATD):
        """the rate-control forms for a batch (all VBV or all default): evals as for lowres_frame_cost_batch, inv_qscales: list of uint16[n_mb]
        or None.  -> list of (score, intra_mbs, intra_cost_sum, score_aq, row_satd or None)"""
        n = len(evals)
        ptrs = [(C.c_void_p * n)(*[e[k].h for e in evals]) for k in range(3)]
        pm = np.array([[e[3], e[4], e[5], me_method, me_range, flags | (LOWRES_VBV if vbv else 0), e[6][0], e[6][1], e[7]] for e in evals], np.int32)
        res = np.zeros((n, 4), np.int32)
        rows = [np.zeros(evals[i][0].g.mb_height, np.int32) for i in range(n)] if vbv else None
        iqs = [np.ascontiguousarray(q, np.uint16) for q in inv_qscales] if inv_qscales is not None else None
        iq_p = (C.c_void_p * n)(*[q.ctypes.data for q in iqs]) if iqs is not None else None
        row_p = (C.c_void_p * n)(*[r.ctypes.data for r in rows]) if vbv else None
        self.check(lib().x264_cuda_lowres_frame_cost_batch_rc(self.h, n, ptrs[0], ptrs[1], ptrs[2], pm.ctypes.data, iq_p, res.ctypes.data, row_p))
        return [(int(r[0]), int(r[1]), int(r[2]), int(r[3]), rows[i] if vbv else None) for i, r in enumerate(res)]

    def sad_grid(self, fenc, fref, radius, jobs):
        """-> uint16 [n_jobs, 9, GH, GW]: SAD of every partition at every integer vector of the window (0xffff = not available)"""
        assert jobs.dtype == GRID_JOB
        out = np.zeros((len(jobs), 9, grid_h(radius), grid_w(radius)), np.uint16)
        self.check(lib().x264_cuda_sad_grid(self.h, fenc.h, fref.h, radius, jobs.ctypes.data, len(jobs), out.ctypes.data))
        return out

    def sad_grid_quad(self, fenc, fref, radius, jobs):
        """-> uint16 [n_jobs, GH, GW, 4]: SADs of the four 8x8 quadrants (TL, TR, BL, BR) at every integer vector of the window"""
        assert jobs.dtype == GRID_JOB
        out = np.zeros((len(jobs), grid_h(radius), grid_w(radius), 4), np.uint16)
        self.check(lib().x264_cuda_sad_grid_quad(self.h, fenc.h, fref.h, radius, jobs.ctypes.data, len(jobs), out.ctypes.data, 0))
        return out

    def me_search_mb_dev(self, fenc, fref, me_range, d_jobs, n, d_results):
        self.check(lib().x264_cuda_me_search_mb_dev(self.h, fenc.h, fref.h, me_range, d_jobs, n, d_results))

    def me_search_dev(self, fenc, fref, me_range, d_jobs, n, d_results):
        self.check(lib().x264_cuda_me_search_dev(self.h, fenc.h, fref.h, me_range, d_jobs, n, d_results))
