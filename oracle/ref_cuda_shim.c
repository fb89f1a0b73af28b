/* ref_cuda_shim.c — TEST INFRASTRUCTURE ONLY.  Lets the UNMODIFIED reference encoder run with its four function tables overridden by
 * the CUDA back-end, the way INTEGRATION.md describes, without touching a reference source file: oracle/Makefile compiles
 * S/common/{pixel,mc,dct,quant}.c a second time with -Dx264_<t>_init=x264_<t>_init_c (a rename of the one symbol each file exports for
 * its table), and this file supplies x264_<t>_init: C table first, then the x264_<t>_init_cuda overrides on top
 * (include/x264_cuda_tables.h).  The resulting CLI (oracle/_ref/x264_cuda) must write a byte-identical stream
 * (tests/test_gpu_stream.py): SURVEY 8c "stream level" parity.  Further down the same trick wraps frame-level functions, the motion
 * searches and the macroblock residual coder so that the frame-batched C ABI is driven by the live encoder's data as well. */
#include <stdio.h>
#include <stdlib.h>
#include "common/common.h"
#include "x264_cuda_tables.h"

void x264_pixel_init_c(int cpu, x264_pixel_function_t *pixf);
void x264_mc_init_c(int cpu, x264_mc_functions_t *pf);
void x264_dct_init_c(int cpu, x264_dct_function_t *dctf);
void x264_quant_init_c(x264_t *h, int cpu, x264_quant_function_t *pf);

/* the mirrors must be layout-identical to the reference's structs */
typedef char chk_pixel[sizeof(x264_cuda_pixel_function_t) == sizeof(x264_pixel_function_t) ? 1 : -1];
typedef char chk_mc[sizeof(x264_cuda_mc_functions_t) == sizeof(x264_mc_functions_t) ? 1 : -1];
typedef char chk_dct[sizeof(x264_cuda_dct_function_t) == sizeof(x264_dct_function_t) ? 1 : -1];
typedef char chk_quant[sizeof(x264_cuda_quant_function_t) == sizeof(x264_quant_function_t) ? 1 : -1];

static void need(int rc, const char *what)
{
    if (rc) { fprintf(stderr, "ref_cuda_shim: %s failed (no CUDA device?)\n", what); exit(3); }
}
static void report(void) { fprintf(stderr, "ref_cuda_shim: %lld device launches\n", x264_cuda_tables_launches()); }
static int enabled(const char *table) /* X264_CUDA_TABLES=pixel,mc,dct,quant (default: all) selects which tables are overridden */
{
    const char *e = getenv("X264_CUDA_TABLES");
    return !e || strstr(e, table);
}

void x264_pixel_init(int cpu, x264_pixel_function_t *pixf)
{
    static int once;
    if (!once++) atexit(report);
    x264_pixel_init_c(cpu, pixf);
    if (enabled("pixel")) need(x264_pixel_init_cuda((x264_cuda_pixel_function_t *)pixf), "x264_pixel_init_cuda");
}
void x264_mc_init(int cpu, x264_mc_functions_t *pf)
{
    x264_mc_init_c(cpu, pf);
    if (enabled("mc")) need(x264_mc_init_cuda((x264_cuda_mc_functions_t *)pf), "x264_mc_init_cuda");
}
void x264_dct_init(int cpu, x264_dct_function_t *dctf)
{
    x264_dct_init_c(cpu, dctf);
    if (enabled("dct")) need(x264_dct_init_cuda((x264_cuda_dct_function_t *)dctf), "x264_dct_init_cuda");
}
void x264_quant_init(x264_t *h, int cpu, x264_quant_function_t *pf)
{
    x264_quant_init_c(h, cpu, pf);
    if (enabled("quant")) need(x264_quant_init_cuda((x264_cuda_quant_function_t *)pf), "x264_quant_init_cuda");
}

/* ------------------------------------------------------------------------------------------------------------------------------
 * Frame-level hooks exercised inside the live encoder: x264_frame_init_lowres, the half-pel / integral
 * planes of a finished reconstruction, and in-loop deblocking.  Each wrapper lets the reference's C do its work, has the device
 * recompute the same planes from the same inputs through the frame-batched C ABI, demands byte equality, and then REPLACES the host
 * planes with the device's, so the rest of the encode (and the bitstream) runs on device-produced data.  Progressive, one thread.
 * Off with X264_CUDA_FRAME_HOOKS=0. */
void x264_frame_init_lowres_c(x264_t *h, x264_frame_t *frame);
void x264_frame_expand_border_filtered_c(x264_t *h, x264_frame_t *frame, int mb_y, int b_end);
void x264_frame_deblock_row_c(x264_t *h, int mb_y);

static x264_cuda_t *fctx;
static long long slot_clock;
static x264_cuda_frame_t *lowres_slot(x264_t *h, x264_frame_t *f, int frame_no, int create);
static x264_t *g_h; /* for the hooks whose reference signature does not carry the encoder handle */
static x264_cuda_frame_t *ffr;
static uint8_t *tmp_plane, *pre[3];
static long long n_lowres, n_filter, n_deblock, n_bytes_checked;

static int hooks_on(void)
{
    const char *e = getenv("X264_CUDA_FRAME_HOOKS");
    return !e || atoi(e);
}
static void report_frames(void)
{
    fprintf(stderr, "ref_cuda_shim: frame hooks: %lld lowres, %lld filter, %lld deblock frames recomputed on the device, %lld bytes compared equal\n",
            n_lowres, n_filter, n_deblock, n_bytes_checked);
}
static void ck(int rc, const char *what)
{
    if (rc) { fprintf(stderr, "ref_cuda_shim: %s: %s\n", what, x264_cuda_error(fctx)); exit(3); }
}
static void frame_ctx(x264_t *h, x264_frame_t *fr)
{
    g_h = h;
    if (fctx) return;
    need(x264_cuda_open(&fctx, 0), "x264_cuda_open");
    int flags = X264_CUDA_FRAME_CHROMA;
    if (h->param.analyse.i_subpel_refine) flags |= X264_CUDA_FRAME_HPEL;
    if (h->frames.b_have_lowres) flags |= X264_CUDA_FRAME_LOWRES;
    if (h->param.analyse.i_me_method >= X264_ME_ESA) flags |= X264_CUDA_FRAME_INTEGRAL | (h->frames.b_have_sub8x8_esa ? X264_CUDA_FRAME_INTEGRAL4 : 0);
    /* the padded (mod 16) size is "the picture" here: reconstructed frames carry real data up to the macroblock grid */
    ffr = x264_cuda_frame_new(fctx, fr->i_width[0], fr->i_lines[0], flags);
    if (!ffr) ck(-1, "x264_cuda_frame_new");
    const size_t sz = (size_t)fr->i_stride[0] * (fr->i_lines[0] + 2 * PADV) * 2;
    tmp_plane = malloc(sz);
    for (int i = 0; i < 3; i++) pre[i] = malloc((size_t)fr->i_stride[i] * (fr->i_lines[i] + 2 * PADV));
    atexit(report_frames);
}
/* rows x cols bytes (elem size es) of two pitched buffers must agree; then dst takes the device's bytes */
static void same_then_take(const char *what, uint8_t *host, const uint8_t *dev, int stride, int es, int row0, int rows, int col0, int cols)
{
    for (int y = row0; y < row0 + rows; y++) {
        uint8_t *a = host + ((size_t)y * stride + col0) * es;
        const uint8_t *b = dev + ((size_t)y * stride + col0) * es;
        if (memcmp(a, b, (size_t)cols * es)) {
            int x = 0;
            while (a[x] == b[x]) x++;
            fprintf(stderr, "ref_cuda_shim: %s differs at padded row %d, byte %d: reference %d, device %d\n", what, y, col0 * es + x, a[x], b[x]);
            exit(4);
        }
        memcpy(a, b, (size_t)cols * es);
        n_bytes_checked += (long long)cols * es;
    }
}

void x264_frame_init_lowres(x264_t *h, x264_frame_t *frame)
{
    x264_frame_init_lowres_c(h, frame); /* pixel work + the per-frame lookahead bookkeeping (mc.c:306-331) */
    if (!hooks_on()) return;
    frame_ctx(h, frame);
    x264_cuda_frame_t *dl = lowres_slot(h, frame, frame->i_frame, 1); /* kept per input frame: the lookahead hook evaluates costs on these planes */
    ck(x264_cuda_frame_upload(fctx, dl, frame->plane[0], frame->i_stride[0], frame->i_width[0], frame->i_lines[0]), "upload");
    ck(x264_cuda_frame_expand_border(fctx, dl), "expand_border");
    ck(x264_cuda_frame_init_lowres(fctx, dl), "init_lowres");
    const int s = frame->i_stride_lowres, rows = frame->i_lines_lowres + 2 * PADV, cols = frame->i_width_lowres + 2 * PADH;
    for (int k = 0; k < 4; k++) {
        ck(x264_cuda_frame_download(fctx, dl, X264_CUDA_PLANE_LOWRES + k, tmp_plane, s), "download lowres");
        uint8_t *host = frame->lowres[k] - (s * PADV + PADH);
        if (k < 3)
            same_then_take("lowres plane", host, tmp_plane, s, 1, 0, rows, 0, cols);
        else { /* the last pixel of the centre plane reads plane[lines][width], which the reference never writes for an input frame (it
                * duplicates the last column for rows < lines and the last row for columns < width, mc.c:315-317): that pixel and the
                * bottom-right border replicated from it are stale heap memory there — left as the reference has them */
            const int r_last = PADV + frame->i_lines_lowres - 1, c_last = PADH + frame->i_width_lowres - 1;
            same_then_take("lowres plane", host, tmp_plane, s, 1, 0, r_last, 0, cols);
            same_then_take("lowres plane", host, tmp_plane, s, 1, r_last, rows - r_last, 0, c_last);
        }
    }
    n_lowres++;
}

void x264_frame_expand_border_filtered(x264_t *h, x264_frame_t *frame, int mb_y, int b_end)
{
    x264_frame_expand_border_filtered_c(h, frame, mb_y, b_end);
    if (!hooks_on() || !b_end) return;
    /* end of the frame: plane[0] is deblocked and border-expanded, filtered[1..3] and the integral are complete (encoder.c:1009-1023) */
    frame_ctx(h, frame);
    const int s = frame->i_stride[0], rows = frame->i_lines[0] + 2 * PADV, cols = frame->i_width[0] + 2 * PADH;
    ck(x264_cuda_frame_upload(fctx, ffr, frame->plane[0], s, frame->i_width[0], frame->i_lines[0]), "upload");
    ck(x264_cuda_frame_expand_border(fctx, ffr), "expand_border");
    ck(x264_cuda_frame_filter(fctx, ffr), "frame_filter");
    for (int k = 0; k < 4; k++) {
        ck(x264_cuda_frame_download(fctx, ffr, k, tmp_plane, s), "download hpel");
        same_then_take(k ? "half-pel plane" : "border-expanded luma", frame->filtered[k] - (s * PADV + PADH), tmp_plane, s, 1, 0, rows, 0, cols);
    }
    if (frame->integral) { /* defined area of the reference's integral: padded rows [1, lines+56), columns [0, width+56) (mc.c:436-461) */
        uint16_t *base = frame->integral - (s * PADV + PADH);
        ck(x264_cuda_frame_download(fctx, ffr, X264_CUDA_PLANE_INTEGRAL, tmp_plane, s), "download integral");
        same_then_take("integral (8x8 sums)", (uint8_t *)base, tmp_plane, s, 2, 1, frame->i_lines[0] + 55, 0, frame->i_width[0] + 56);
        if (h->frames.b_have_sub8x8_esa) {
            ck(x264_cuda_frame_download(fctx, ffr, X264_CUDA_PLANE_INTEGRAL4, tmp_plane, s), "download integral4");
            same_then_take("integral (4x4 sums)", (uint8_t *)(base + (size_t)s * (frame->i_lines[0] + 2 * PADV)), tmp_plane, s, 2, 1,
                           frame->i_lines[0] + 55, 0, frame->i_width[0] + 56);
        }
    }
    n_filter++;
}

void x264_frame_deblock_row(x264_t *h, int mb_y)
{
    x264_frame_t *f = h->fdec;
    const int on = hooks_on() && !h->sh.b_mbaff;
    if (on) { /* macroblock row mb_y is still unfiltered when its turn comes: keep a copy (frame.c:621-792 only reaches 3 px upwards) */
        frame_ctx(h, f);
        for (int i = 0; i < 3; i++) {
            const int rows = 16 >> !!i;
            memcpy(pre[i] + (size_t)mb_y * rows * f->i_stride[i], f->plane[i] + (size_t)mb_y * rows * f->i_stride[i], (size_t)rows * f->i_stride[i]);
        }
    }
    x264_frame_deblock_row_c(h, mb_y);
    if (!on || mb_y != h->sps->i_mb_height - 1) return;
    /* last row done on the host: deblock the whole unfiltered picture on the device with the encoder's own per-macroblock arrays */
    ck(x264_cuda_frame_upload(fctx, ffr, pre[0], f->i_stride[0], f->i_width[0], f->i_lines[0]), "upload");
    ck(x264_cuda_frame_upload_chroma(fctx, ffr, X264_CUDA_PLANE_CB, pre[1], f->i_stride[1], f->i_width[1], f->i_lines[1]), "upload cb");
    ck(x264_cuda_frame_upload_chroma(fctx, ffr, X264_CUDA_PLANE_CR, pre[2], f->i_stride[2], f->i_width[2], f->i_lines[2]), "upload cr");
    x264_cuda_deblock_params_t p = { h->sh.i_alpha_c0_offset, h->sh.i_beta_offset, h->pps->i_chroma_qp_index_offset, h->sh.i_type == SLICE_TYPE_B,
                                     !!(h->param.analyse.inter & X264_ANALYSE_PSUB8x8), !h->pps->b_cabac && h->pps->b_transform_8x8_mode };
    ck(x264_cuda_frame_deblock(fctx, ffr, &p, h->mb.type, h->mb.qp, h->mb.mb_transform_size, (const uint8_t(*)[24])h->mb.non_zero_count, h->mb.ref[0],
                               (const int16_t(*)[2])h->mb.mv[0], h->mb.ref[1], (const int16_t(*)[2])h->mb.mv[1]), "frame_deblock");
    static const int ids[3] = { X264_CUDA_PLANE_FULL, X264_CUDA_PLANE_CB, X264_CUDA_PLANE_CR };
    for (int i = 0; i < 3; i++) {
        const int s = f->i_stride[i], padv = PADV >> !!i, padh = PADH >> !!i;
        ck(x264_cuda_frame_download(fctx, ffr, ids[i], tmp_plane, s), "download deblocked");
        same_then_take(i ? "deblocked chroma" : "deblocked luma", f->plane[i] - (s * padv + padh), tmp_plane, s, 1, padv, f->i_lines[i], padh, f->i_width[i]);
    }
    n_deblock++;
}

/* ------------------------------------------------------------------------------------------------------------------------------
 * Motion-search hooks: every full-resolution x264_me_search_ref, x264_me_refine_qpel and x264_me_refine_bidir_satd call of the live
 * encoder is repeated on the device as a one-job batch through the C ABI (x264_cuda_me_search / _me_search_small /
 * _me_refine_bidir) with the encoder's own predictors, limits and cost table, and must return the same vector and costs.
 * Device frames are kept per (x264_frame_t, frame number): source pictures as they are, references border-expanded and filtered
 * on the device.  Lowres (lookahead) searches and searches with a half-pel early-exit threshold (multi-reference P16x16, analyse.c:
 * 1095-1100: the threshold is carried across calls) are left to the C path.  Off with X264_CUDA_ME_HOOKS=0. */
#include "encoder/me.h"
void x264_me_search_ref_c(x264_t *h, x264_me_t *m, int16_t (*mvc)[2], int i_mvc, int *p_halfpel_thresh);
void x264_me_refine_qpel_c(x264_t *h, x264_me_t *m);
void x264_me_refine_bidir_satd_c(x264_t *h, x264_me_t *m0, x264_me_t *m1, int i_weight);
extern int16_t *g_cost_mv[52];

typedef struct { x264_frame_t *f; int frame, poc, is_ref; x264_cuda_frame_t *d; long long used; } dev_slot;
static dev_slot slots[8];
static long long n_search, n_qpel, n_bidir, n_skipped;
static int cost_uploaded[52];

static int me_hooks_on(void)
{
    const char *e = getenv("X264_CUDA_ME_HOOKS");
    return (!e || atoi(e)) && hooks_on();
}
static void report_me(void)
{
    fprintf(stderr, "ref_cuda_shim: me hooks: %lld searches, %lld qpel refinements, %lld bidir refinements repeated on the device and equal; %lld left to C\n",
            n_search, n_qpel, n_bidir, n_skipped);
}
static x264_cuda_frame_t *dev_frame(x264_t *h, x264_frame_t *f, int is_ref)
{
    static int once;
    if (!once++) atexit(report_me);
    frame_ctx(h, f);
    dev_slot *victim = &slots[0];
    for (int i = 0; i < 8; i++) {
        dev_slot *s = &slots[i];
        if (s->d && s->f == f && s->frame == f->i_frame && s->poc == f->i_poc && s->is_ref == is_ref) { s->used = ++slot_clock; return s->d; }
        if (s->used < victim->used) victim = s;
    }
    if (!victim->d) {
        x264_cuda_geom_t g;
        x264_cuda_frame_geometry(ffr, &g);
        victim->d = x264_cuda_frame_new(fctx, f->i_width[0], f->i_lines[0], g.flags);
        if (!victim->d) ck(-1, "x264_cuda_frame_new");
    }
    victim->f = f; victim->frame = f->i_frame; victim->poc = f->i_poc; victim->is_ref = is_ref; victim->used = ++slot_clock;
    ck(x264_cuda_frame_upload(fctx, victim->d, f->plane[0], f->i_stride[0], f->i_width[0], f->i_lines[0]), "upload");
    ck(x264_cuda_frame_upload_chroma(fctx, victim->d, X264_CUDA_PLANE_CB, f->plane[1], f->i_stride[1], f->i_width[1], f->i_lines[1]), "upload cb");
    ck(x264_cuda_frame_upload_chroma(fctx, victim->d, X264_CUDA_PLANE_CR, f->plane[2], f->i_stride[2], f->i_width[2], f->i_lines[2]), "upload cr");
    if (is_ref) {
        ck(x264_cuda_frame_expand_border(fctx, victim->d), "expand_border");
        if (h->param.analyse.i_subpel_refine) ck(x264_cuda_frame_filter(fctx, victim->d), "frame_filter");
    }
    return victim->d;
}
/* which reference frame (and block position) does m->p_fref[0] point into?  NULL for lowres / unknown */
static x264_frame_t *find_ref(x264_t *h, const x264_me_t *m, int *bx, int *by)
{
    for (int l = 0; l < 2; l++)
        for (int i = 0; i < (l ? h->i_ref1 : h->i_ref0); i++) {
            x264_frame_t *f = l ? h->fref1[i] : h->fref0[i];
            const ptrdiff_t off = m->p_fref[0] - f->plane[0];
            const ptrdiff_t lim = (ptrdiff_t)f->i_stride[0] * f->i_lines[0];
            if (off >= 0 && off < lim && m->i_stride[0] == f->i_stride[0]) { *bx = (int)(off % f->i_stride[0]); *by = (int)(off / f->i_stride[0]); return f; }
        }
    return NULL;
}
static int qp_of(const x264_me_t *m)
{
    for (int q = 0; q < 52; q++)
        if (g_cost_mv[q] && g_cost_mv[q] + 2 * 4 * 2048 == m->p_cost_mv) {
            if (!cost_uploaded[q]) { ck(x264_cuda_set_cost_mv(fctx, q, g_cost_mv[q]), "set_cost_mv"); cost_uploaded[q] = 1; } /* the reference's own table */
            return q;
        }
    return -1;
}
static int fenc_pos_ok(x264_t *h, const x264_me_t *m, int bx, int by)
{
    const ptrdiff_t off = m->p_fenc[0] - h->mb.pic.p_fenc[0];
    return off >= 0 && off < 16 * FENC_STRIDE && bx == 16 * h->mb.i_mb_x + (int)(off % FENC_STRIDE) && by == 16 * h->mb.i_mb_y + (int)(off / FENC_STRIDE);
}
static int me_flags(x264_t *h)
{
    return (h->pixf.mbcmp[0] == h->pixf.satd[0] ? X264_CUDA_ME_MBCMP_SATD : 0) | (h->pixf.fpelcmp[0] == h->pixf.satd[0] ? X264_CUDA_ME_FPEL_SATD : 0) |
           (h->mb.b_chroma_me ? X264_CUDA_ME_CHROMA : 0);
}
static void fill_limits(x264_t *h, x264_cuda_me_job_t *j)
{
    for (int k = 0; k < 2; k++) {
        j->mv_min_fpel[k] = h->mb.mv_min_fpel[k]; j->mv_max_fpel[k] = h->mb.mv_max_fpel[k];
        j->mv_min_spel[k] = h->mb.mv_min_spel[k]; j->mv_max_spel[k] = h->mb.mv_max_spel[k];
    }
}
static void me_differs(const char *what, x264_t *h, const x264_me_t *m, const x264_cuda_me_final_t *d, int bx, int by)
{
    fprintf(stderr, "ref_cuda_shim: %s differs at frame %d block (%d,%d) pixel %d method %d subme %d: reference mv (%d,%d) cost %d cost_mv %d, device mv (%d,%d) "
            "cost %d cost_mv %d\n", what, h->fenc->i_frame, bx, by, m->i_pixel, h->mb.i_me_method, h->mb.i_subpel_refine, m->mv[0], m->mv[1], m->cost, m->cost_mv,
            d->mv[0], d->mv[1], d->cost, d->cost_mv);
    exit(5);
}

static void record_search(x264_t *h, x264_frame_t *ref, const x264_cuda_me_job_t *j, const x264_me_t *m);
static long long n_grid, n_grid_outside;
static void report_grid(void)
{
    fprintf(stderr, "ref_cuda_shim: grid replay: %lld ESA searches replayed on device SAD grids and equal; %lld needed a vector outside the grid\n", n_grid, n_grid_outside);
}
/* the sequential-predictor path of INTEGRATION.md: SAD grids of the macroblock around a guessed centre (here: the rounded mvp of the 16x16
 * search would do; we use this search's own mvp), then x264_cuda_host_esa_replay with the exact predictors must give the per-block
 * kernel's full-pel result (which the caller has just shown to lead to the C result) */
static void grid_replay_check(x264_t *h, const x264_me_t *m, const x264_cuda_me_job_t *j, const x264_cuda_me_result_t *want, x264_cuda_frame_t *denc,
                              x264_cuda_frame_t *dref, int range)
{
    static uint16_t *grid;
    static int once;
    const int R = range + 8, gw = X264_CUDA_GRID_W(R), gh = X264_CUDA_GRID_H(R);
    if (!once++) { atexit(report_grid); }
    if (!grid) grid = malloc((size_t)9 * X264_CUDA_GRID_W(64 + 8) * X264_CUDA_GRID_H(64 + 8) * sizeof(uint16_t));
    if (range > 64) return;
    const int ox = j->bx & 15, oy = j->by & 15;
    const int part = m->i_pixel == PIXEL_16x16 ? 0 : m->i_pixel == PIXEL_16x8 ? 1 + (oy >> 3) : m->i_pixel == PIXEL_8x16 ? 3 + (ox >> 3) : 5 + (oy >> 3) * 2 + (ox >> 3);
    x264_cuda_grid_job_t g;
    memset(&g, 0, sizeof(g));
    g.mb_x = j->bx >> 4; g.mb_y = j->by >> 4; g.part_mask = 1 << part;
    g.cx = x264_clip3((j->mvp[0] + 2) >> 2, j->mv_min_fpel[0], j->mv_max_fpel[0]);
    g.cy = x264_clip3((j->mvp[1] + 2) >> 2, j->mv_min_fpel[1], j->mv_max_fpel[1]);
    for (int k = 0; k < 2; k++) { g.mv_min_fpel[k] = j->mv_min_fpel[k]; g.mv_max_fpel[k] = j->mv_max_fpel[k]; }
    ck(x264_cuda_sad_grid(fctx, denc, dref, R, &g, 1, grid), "sad_grid");
    x264_cuda_me_result_t r;
    const int rc = x264_cuda_host_esa_replay(grid + (size_t)part * gw * gh, R, g.cx, g.cy, j, range, m->p_cost_mv - 2 * 4 * 2048, &r);
    if (rc) { n_grid_outside++; return; }
    if (r.bmx != want->bmx || r.bmy != want->bmy || r.bcost != want->bcost) {
        fprintf(stderr, "ref_cuda_shim: grid replay differs at block (%d,%d) part %d: search (%d,%d) %d, replay (%d,%d) %d\n", j->bx, j->by, part, want->bmx, want->bmy,
                want->bcost, r.bmx, r.bmy, r.bcost);
        exit(5);
    }
    n_grid++;
}
static void flush_batched(x264_t *h);

void x264_me_search_ref(x264_t *h, x264_me_t *m, int16_t (*mvc)[2], int i_mvc, int *p_halfpel_thresh)
{
    x264_cuda_me_job_t j;
    x264_cuda_frame_t *dref = NULL, *denc = NULL;
    int bx = 0, by = 0, qp = -1, ok = me_hooks_on() && !h->sh.b_mbaff && !p_halfpel_thresh && i_mvc <= X264_CUDA_ME_MAX_MVC;
    const int method = h->mb.i_me_method, subme = h->mb.i_subpel_refine;
    x264_frame_t *fr = NULL;
    if (ok) {
        fr = find_ref(h, m, &bx, &by);
        ok = fr && fenc_pos_ok(h, m, bx, by) && !(method == X264_ME_ESA && subme >= 3);
        if (ok) { frame_ctx(h, fr); qp = qp_of(m); ok = qp >= 0; }
        if (ok) {
            dref = dev_frame(h, fr, 1); denc = dev_frame(h, h->fenc, 0);
            memset(&j, 0, sizeof(j));
            j.bx = bx; j.by = by; j.i_pixel = m->i_pixel; j.qp = qp; j.i_mvc = i_mvc; j.flags = me_flags(h);
            j.mvp[0] = m->mvp[0]; j.mvp[1] = m->mvp[1];
            for (int k = 0; k < i_mvc; k++) { j.mvc[k][0] = mvc[k][0]; j.mvc[k][1] = mvc[k][1]; }
            fill_limits(h, &j);
        }
    }
    x264_me_search_ref_c(h, m, mvc, i_mvc, p_halfpel_thresh);
    if (!ok) { n_skipped++; return; }
    x264_cuda_me_final_t fin;
    const int range = h->param.analyse.i_me_range;
    if (method == X264_ME_ESA) { /* full-pel stage as its own job, then the "-> qpel" + refine_subpel tail seeded with its winner */
        x264_cuda_me_result_t r;
        ck(x264_cuda_me_search(fctx, denc, dref, range, &j, 1, &r), "me_search");
        j.seed_mv[0] = r.bmx; j.seed_mv[1] = r.bmy; j.seed_cost = r.bcost;
        ck(x264_cuda_me_search_small(fctx, denc, dref, X264_CUDA_ME_METHOD_SEEDED, range, subme, &j, 1, &fin), "me_search_small (seeded)");
        if (m->i_pixel <= PIXEL_8x8) grid_replay_check(h, m, &j, &r, denc, dref, range);
    } else
        ck(x264_cuda_me_search_small(fctx, denc, dref, method == X264_ME_TESA ? X264_CUDA_ME_METHOD_TESA : method, range, subme, &j, 1, &fin), "me_search_small");
    if (fin.mv[0] != m->mv[0] || fin.mv[1] != m->mv[1] || fin.cost != m->cost || fin.cost_mv != m->cost_mv) me_differs("x264_me_search_ref", h, m, &fin, bx, by);
    n_search++;
    record_search(h, fr, &j, m);
}

void x264_me_refine_qpel(x264_t *h, x264_me_t *m)
{
    x264_cuda_me_job_t j;
    x264_cuda_frame_t *dref = NULL, *denc = NULL;
    int bx = 0, by = 0, qp = -1, ok = me_hooks_on() && !h->sh.b_mbaff;
    if (ok) {
        x264_frame_t *fr = find_ref(h, m, &bx, &by);
        ok = fr && fenc_pos_ok(h, m, bx, by);
        if (ok) { frame_ctx(h, fr); qp = qp_of(m); ok = qp >= 0; }
        if (ok) {
            dref = dev_frame(h, fr, 1); denc = dev_frame(h, h->fenc, 0);
            memset(&j, 0, sizeof(j));
            j.bx = bx; j.by = by; j.i_pixel = m->i_pixel; j.qp = qp; j.flags = me_flags(h);
            j.mvp[0] = m->mvp[0]; j.mvp[1] = m->mvp[1];
            j.seed_mv[0] = m->mv[0]; j.seed_mv[1] = m->mv[1];
            j.seed_cost = m->cost - ((m->i_pixel <= PIXEL_8x8 && h->sh.i_type == SLICE_TYPE_P) ? m->i_ref_cost : 0); /* me.c:639-640 */
            fill_limits(h, &j);
        }
    }
    x264_me_refine_qpel_c(h, m);
    if (!ok) { n_skipped++; return; }
    x264_cuda_me_final_t fin;
    ck(x264_cuda_me_search_small(fctx, denc, dref, X264_CUDA_ME_METHOD_REFINE_QPEL, h->param.analyse.i_me_range, h->mb.i_subpel_refine, &j, 1, &fin), "refine_qpel");
    if (fin.mv[0] != m->mv[0] || fin.mv[1] != m->mv[1] || fin.cost != m->cost || fin.cost_mv != m->cost_mv) me_differs("x264_me_refine_qpel", h, m, &fin, bx, by);
    n_qpel++;
}

void x264_me_refine_bidir_satd(x264_t *h, x264_me_t *m0, x264_me_t *m1, int i_weight)
{
    x264_cuda_bidir_job_t j;
    x264_cuda_frame_t *d0 = NULL, *d1 = NULL, *denc = NULL;
    int bx = 0, by = 0, bx1 = 0, by1 = 0, qp = -1, ok = me_hooks_on() && !h->sh.b_mbaff && i_weight >= 0 && i_weight < 256;
    if (ok) {
        x264_frame_t *f0 = find_ref(h, m0, &bx, &by), *f1 = find_ref(h, m1, &bx1, &by1);
        ok = f0 && f1 && bx == bx1 && by == by1 && fenc_pos_ok(h, m0, bx, by);
        if (ok) { frame_ctx(h, f0); qp = qp_of(m0); ok = qp >= 0; }
        if (ok) {
            d0 = dev_frame(h, f0, 1); d1 = dev_frame(h, f1, 1); denc = dev_frame(h, h->fenc, 0);
            memset(&j, 0, sizeof(j));
            j.bx = bx; j.by = by; j.i_pixel = m0->i_pixel; j.qp = qp; j.weight = i_weight; j.flags = me_flags(h) & X264_CUDA_ME_MBCMP_SATD;
            for (int k = 0; k < 2; k++) {
                j.mv0[k] = m0->mv[k]; j.mv1[k] = m1->mv[k]; j.mvp0[k] = m0->mvp[k]; j.mvp1[k] = m1->mvp[k];
                j.mv_min_spel[k] = h->mb.mv_min_spel[k]; j.mv_max_spel[k] = h->mb.mv_max_spel[k];
            }
        }
    }
    x264_me_refine_bidir_satd_c(h, m0, m1, i_weight);
    if (!ok) { n_skipped++; return; }
    x264_cuda_bidir_result_t r;
    ck(x264_cuda_me_refine_bidir(fctx, denc, d0, d1, &j, 1, &r), "me_refine_bidir");
    if (r.mv0[0] != m0->mv[0] || r.mv0[1] != m0->mv[1] || r.mv1[0] != m1->mv[0] || r.mv1[1] != m1->mv[1]) {
        fprintf(stderr, "ref_cuda_shim: x264_me_refine_bidir_satd differs at frame %d block (%d,%d) pixel %d weight %d: reference (%d,%d)/(%d,%d), device (%d,%d)/(%d,%d)\n",
                h->fenc->i_frame, bx, by, m0->i_pixel, i_weight, m0->mv[0], m0->mv[1], m1->mv[0], m1->mv[1], r.mv0[0], r.mv0[1], r.mv1[0], r.mv1[1]);
        exit(5);
    }
    n_bidir++;
}

/* ------------------------------------------------------------------------------------------------------------------------------
 * Residual hooks: x264_macroblock_probe_skip and the inter branch of x264_macroblock_encode (no trellis / noise reduction / lossless)
 * repeated on the device for the live macroblock, with the encoder's own quantiser tables (x264_cuda_set_quant_tables from
 * h->quant4_mf ...).  The prediction the C code works from (formed by x264_mb_mc, or by the probe's own mc) is placed at the
 * macroblock's position of a device frame; coefficients, non-zero counts, cbp and the reconstructed pixels must agree (exit 6). */
#include "encoder/macroblock.h"
void x264_macroblock_encode_c(x264_t *h);
int x264_macroblock_probe_skip_c(x264_t *h, int b_bidir);
static void mc_check(x264_t *h);
static x264_cuda_frame_t *fpred;
static uint8_t *shadow[3];
static long long n_resid, n_probe, n_resid_c;

static void report_resid(void)
{
    fprintf(stderr, "ref_cuda_shim: residual hooks: %lld inter macroblock encodes, %lld skip probes repeated on the device and equal; %lld encodes left to C\n",
            n_resid, n_probe, n_resid_c);
}
static int resid_hooks_on(void)
{
    const char *e = getenv("X264_CUDA_RESID_HOOKS");
    return (!e || atoi(e)) && hooks_on();
}
static void resid_ctx(x264_t *h)
{
    if (fpred) return;
    x264_frame_t *f = h->fdec;
    frame_ctx(h, f);
    fpred = x264_cuda_frame_new(fctx, f->i_width[0], f->i_lines[0], X264_CUDA_FRAME_CHROMA);
    if (!fpred) ck(-1, "x264_cuda_frame_new");
    for (int i = 0; i < 3; i++) shadow[i] = calloc((size_t)f->i_stride[i] * (f->i_lines[i] + 2 * PADV), 1);
    ck(x264_cuda_set_quant_tables(fctx, (const uint16_t *const *)h->quant4_mf, (const uint16_t *const *)h->quant4_bias, (const int *const *)h->dequant4_mf,
                                  (const uint16_t *const *)h->quant8_mf, (const uint16_t *const *)h->quant8_bias, (const int *const *)h->dequant8_mf),
       "set_quant_tables");
    atexit(report_resid);
}
/* the macroblock's prediction (p_fdec tiles) -> device frame at the macroblock's position */
static void upload_pred(x264_t *h)
{
    x264_frame_t *f = h->fdec;
    for (int i = 0; i < 3; i++) {
        const int n = 16 >> !!i;
        for (int y = 0; y < n; y++)
            memcpy(shadow[i] + ((size_t)h->mb.i_mb_y * n + y) * f->i_stride[i] + h->mb.i_mb_x * n, h->mb.pic.p_fdec[i] + y * FDEC_STRIDE, n);
    }
    ck(x264_cuda_frame_upload(fctx, fpred, shadow[0], f->i_stride[0], f->i_width[0], f->i_lines[0]), "upload pred");
    ck(x264_cuda_frame_upload_chroma(fctx, fpred, X264_CUDA_PLANE_CB, shadow[1], f->i_stride[1], f->i_width[1], f->i_lines[1]), "upload pred cb");
    ck(x264_cuda_frame_upload_chroma(fctx, fpred, X264_CUDA_PLANE_CR, shadow[2], f->i_stride[2], f->i_width[2], f->i_lines[2]), "upload pred cr");
}

int x264_macroblock_probe_skip(x264_t *h, int b_bidir)
{
    uint8_t dev_skip = 0xff;
    if (resid_hooks_on() && !h->sh.b_mbaff) {
        resid_ctx(h);
        x264_cuda_skip_job_t j;
        memset(&j, 0, sizeof(j));
        j.mb_x = h->mb.i_mb_x; j.mb_y = h->mb.i_mb_y; j.qp = h->mb.i_qp; j.chroma_qp = h->mb.i_chroma_qp;
        x264_cuda_frame_t *denc = dev_frame(h, h->fenc, 0);
        if (b_bidir) { /* the caller has put the direct prediction into p_fdec (analyse.c:2486-2489) */
            upload_pred(h);
            j.flags = X264_CUDA_SKIP_PRED_IN_FDEC;
            ck(x264_cuda_probe_skip(fctx, denc, NULL, fpred, &j, 1, &dev_skip), "probe_skip (bidir)");
        } else {
            j.mvx = x264_clip3(h->mb.cache.pskip_mv[0], h->mb.mv_min[0], h->mb.mv_max[0]); /* macroblock.c:812-813 */
            j.mvy = x264_clip3(h->mb.cache.pskip_mv[1], h->mb.mv_min[1], h->mb.mv_max[1]);
            ck(x264_cuda_probe_skip(fctx, denc, dev_frame(h, h->fref0[0], 1), NULL, &j, 1, &dev_skip), "probe_skip");
        }
    }
    const int r = x264_macroblock_probe_skip_c(h, b_bidir);
    if (dev_skip != 0xff) {
        if (dev_skip != r) {
            fprintf(stderr, "ref_cuda_shim: x264_macroblock_probe_skip differs at frame %d mb (%d,%d) bidir %d: reference %d, device %d\n", h->fenc->i_frame,
                    h->mb.i_mb_x, h->mb.i_mb_y, b_bidir, r, dev_skip);
            exit(6);
        }
        n_probe++;
    }
    return r;
}

void x264_macroblock_encode(x264_t *h)
{
    const int t = h->mb.i_type;
    const int inter = !IS_INTRA(t) && t != P_SKIP && t != B_SKIP;
    if (!(resid_hooks_on() && inter && !h->sh.b_mbaff && !h->mb.b_lossless && !h->mb.b_trellis && !h->mb.b_noise_reduction)) {
        n_resid_c++;
        x264_macroblock_encode_c(h);
        if (h->mb.i_mb_xy == h->mb.i_mb_count - 1) flush_batched(h);
        return;
    }
    resid_ctx(h);
    if (!h->mb.b_skip_mc) x264_mb_mc(h); /* the prediction x264_macroblock_encode is about to form itself (macroblock.c:596-598) */
    mc_check(h);
    upload_pred(h);
    x264_cuda_resid_job_t j;
    memset(&j, 0, sizeof(j));
    j.mb_x = h->mb.i_mb_x; j.mb_y = h->mb.i_mb_y; j.qp = h->mb.i_qp; j.chroma_qp = h->mb.i_chroma_qp;
    j.flags = (h->mb.b_transform_8x8 ? X264_CUDA_RESID_8x8DCT : 0) |
              ((h->sh.i_type == SLICE_TYPE_B || h->param.analyse.b_dct_decimate) ? X264_CUDA_RESID_DECIMATE : 0);
    static x264_cuda_mb_coeffs_t co;
    ck(x264_cuda_residual_inter(fctx, dev_frame(h, h->fenc, 0), fpred, &j, 1, &co), "residual_inter");
    const int was8 = h->mb.b_transform_8x8;
    memset(&h->dct, 0, sizeof(h->dct)); /* blocks the C code does not code keep stale coefficients otherwise; nothing reads them */
    x264_macroblock_encode_c(h);
    int bad = 0;
    if (was8) bad |= memcmp(co.luma, h->dct.luma8x8, sizeof(h->dct.luma8x8)) ? 1 : 0;
    else bad |= memcmp(co.luma, h->dct.luma4x4, 16 * 16 * sizeof(int16_t)) ? 1 : 0;
    bad |= memcmp(co.chroma_ac, h->dct.luma4x4[16], 8 * 16 * sizeof(int16_t)) ? 2 : 0;
    bad |= memcmp(co.chroma_dc, h->dct.chroma_dc, sizeof(h->dct.chroma_dc)) ? 4 : 0;
    for (int i = 0; i < 27; i++) bad |= co.nnz[i] != h->mb.cache.non_zero_count[x264_scan8[i]] ? 8 : 0;
    bad |= (co.cbp_luma != h->mb.i_cbp_luma || co.cbp_chroma != h->mb.i_cbp_chroma) ? 16 : 0;
    static const int ids[3] = { X264_CUDA_PLANE_FULL, X264_CUDA_PLANE_CB, X264_CUDA_PLANE_CR };
    x264_frame_t *f = h->fdec;
    for (int i = 0; i < 3 && !bad; i++) {
        const int s = f->i_stride[i], padv = PADV >> !!i, padh = PADH >> !!i, n = 16 >> !!i;
        ck(x264_cuda_frame_download(fctx, fpred, ids[i], tmp_plane, s), "download recon");
        for (int y = 0; y < n; y++)
            if (memcmp(tmp_plane + (size_t)(padv + h->mb.i_mb_y * n + y) * s + padh + h->mb.i_mb_x * n, h->mb.pic.p_fdec[i] + y * FDEC_STRIDE, n)) bad |= 32 << i;
    }
    if (bad) {
        fprintf(stderr, "ref_cuda_shim: x264_macroblock_encode (inter) differs at frame %d mb (%d,%d) type %d qp %d/%d 8x8dct %d: mask %d\n", h->fenc->i_frame,
                h->mb.i_mb_x, h->mb.i_mb_y, t, h->mb.i_qp, h->mb.i_chroma_qp, was8, bad);
        exit(6);
    }
    n_resid++;
    if (h->mb.i_mb_xy == h->mb.i_mb_count - 1) flush_batched(h);
}

/* ------------------------------------------------------------------------------------------------------------------------------
 * Frame-metric hooks on live data: x264_adaptive_quant_frame (device macroblock energies + the host float tail
 * must give the same f_qp_offset / i_inv_qscale_factor bit for bit — these steer every macroblock's qp), x264_pixel_ssd_wxh and
 * x264_pixel_ssim_wxh (the PSNR / SSIM row slabs of x264_fdec_filter_row, encoder.c:1034-1056).  Exit 7 on a difference. */
void x264_adaptive_quant_frame_c(x264_t *h, x264_frame_t *frame);
int64_t x264_pixel_ssd_wxh_c(x264_pixel_function_t *pf, uint8_t *pix1, int i_pix1, uint8_t *pix2, int i_pix2, int i_width, int i_height);
float x264_pixel_ssim_wxh_c(x264_pixel_function_t *pf, uint8_t *pix1, int i_pix1, uint8_t *pix2, int i_pix2, int i_width, int i_height, void *buf);
static x264_cuda_frame_t *fm[2];
static long long n_aq, n_ssd, n_ssim;

static void report_metrics(void)
{
    fprintf(stderr, "ref_cuda_shim: metric hooks: %lld AQ frames, %lld SSD slabs, %lld SSIM slabs repeated on the device and equal\n", n_aq, n_ssd, n_ssim);
}
static void metric_ctx(x264_t *h, x264_frame_t *f)
{
    g_h = h;
    if (fm[0]) return;
    frame_ctx(h, f);
    for (int k = 0; k < 2; k++) {
        fm[k] = x264_cuda_frame_new(fctx, f->i_width[0], f->i_lines[0], X264_CUDA_FRAME_CHROMA);
        if (!fm[k]) ck(-1, "x264_cuda_frame_new");
    }
    atexit(report_metrics);
}
static void upload_all(x264_cuda_frame_t *d, x264_frame_t *f)
{
    ck(x264_cuda_frame_upload(fctx, d, f->plane[0], f->i_stride[0], f->i_width[0], f->i_lines[0]), "upload");
    ck(x264_cuda_frame_upload_chroma(fctx, d, X264_CUDA_PLANE_CB, f->plane[1], f->i_stride[1], f->i_width[1], f->i_lines[1]), "upload cb");
    ck(x264_cuda_frame_upload_chroma(fctx, d, X264_CUDA_PLANE_CR, f->plane[2], f->i_stride[2], f->i_width[2], f->i_lines[2]), "upload cr");
}

void x264_adaptive_quant_frame(x264_t *h, x264_frame_t *frame)
{
    x264_adaptive_quant_frame_c(h, frame);
    if (!hooks_on() || h->mb.b_interlaced) return;
    metric_ctx(h, frame);
    upload_all(fm[0], frame);
    const int n = h->mb.i_mb_count;
    uint32_t *energy = malloc(n * sizeof(uint32_t));
    float *qp = malloc(n * sizeof(float));
    uint16_t *inv = malloc(n * sizeof(uint16_t));
    ck(x264_cuda_frame_mb_energy(fctx, fm[0], energy), "frame_mb_energy");
    x264_cuda_host_aq(energy, n, h->param.rc.f_aq_strength, qp, inv);
    if (memcmp(qp, frame->f_qp_offset, n * sizeof(float)) || (h->frames.b_have_lowres && memcmp(inv, frame->i_inv_qscale_factor, n * sizeof(uint16_t)))) {
        fprintf(stderr, "ref_cuda_shim: x264_adaptive_quant_frame differs\n");
        exit(7);
    }
    memcpy(frame->f_qp_offset, qp, n * sizeof(float));
    free(energy); free(qp); free(inv);
    n_aq++;
}

/* pix1 inside a plane of the frame being reconstructed, pix2 at the same offset of the source frame (encoder.c:1038-1055)? */
static int locate(uint8_t *pix1, uint8_t *pix2, int s1, int s2, int *plane, int *x0, int *y0)
{
    x264_t *h = g_h;
    if (!h || !h->fdec || !h->fenc) return 0;
    for (int i = 0; i < 3; i++) {
        const ptrdiff_t off = pix1 - h->fdec->plane[i], lim = (ptrdiff_t)h->fdec->i_stride[i] * h->fdec->i_lines[i];
        if (off >= 0 && off < lim && s1 == h->fdec->i_stride[i] && s2 == h->fenc->i_stride[i] && pix2 - h->fenc->plane[i] == off) {
            *plane = i ? (i == 1 ? X264_CUDA_PLANE_CB : X264_CUDA_PLANE_CR) : X264_CUDA_PLANE_FULL;
            *x0 = (int)(off % s1); *y0 = (int)(off / s1);
            return 1;
        }
    }
    return 0;
}

int64_t x264_pixel_ssd_wxh(x264_pixel_function_t *pf, uint8_t *pix1, int i_pix1, uint8_t *pix2, int i_pix2, int i_width, int i_height)
{
    const int64_t r = x264_pixel_ssd_wxh_c(pf, pix1, i_pix1, pix2, i_pix2, i_width, i_height);
    int plane, x0, y0;
    if (hooks_on() && i_width > 0 && i_height > 0 && locate(pix1, pix2, i_pix1, i_pix2, &plane, &x0, &y0)) {
        metric_ctx(g_h, g_h->fdec);
        upload_all(fm[0], g_h->fdec); upload_all(fm[1], g_h->fenc);
        int64_t d = -1;
        ck(x264_cuda_frame_ssd(fctx, fm[0], fm[1], plane, x0, y0, i_width, i_height, &d), "frame_ssd");
        if (d != r) { fprintf(stderr, "ref_cuda_shim: x264_pixel_ssd_wxh differs: reference %lld, device %lld\n", (long long)r, (long long)d); exit(7); }
        n_ssd++;
    }
    return r;
}

float x264_pixel_ssim_wxh(x264_pixel_function_t *pf, uint8_t *pix1, int i_pix1, uint8_t *pix2, int i_pix2, int i_width, int i_height, void *buf)
{
    const float r = x264_pixel_ssim_wxh_c(pf, pix1, i_pix1, pix2, i_pix2, i_width, i_height, buf);
    int plane, x0, y0;
    if (hooks_on() && i_width >= 8 && i_height >= 8 && locate(pix1, pix2, i_pix1, i_pix2, &plane, &x0, &y0)) {
        metric_ctx(g_h, g_h->fdec);
        upload_all(fm[0], g_h->fdec); upload_all(fm[1], g_h->fenc);
        const int w4 = i_width >> 2, h4 = i_height >> 2;
        int (*sums)[4] = malloc((size_t)w4 * h4 * sizeof(*sums));
        ck(x264_cuda_frame_ssim_sums(fctx, fm[0], fm[1], plane, x0, y0, i_width, i_height, sums), "frame_ssim_sums");
        const float d = x264_cuda_host_ssim_end((const int(*)[4])sums, w4, h4);
        free(sums);
        if (memcmp(&d, &r, sizeof(float))) { fprintf(stderr, "ref_cuda_shim: x264_pixel_ssim_wxh differs: reference %.9g, device %.9g\n", r, d); exit(7); }
        n_ssim++;
    }
    return r;
}

/* ------------------------------------------------------------------------------------------------------------------------------
 * The macroblock-batched ESA kernel (x264_cuda_me_search_mb, the headline kernel) on live data: for --me esa with subme < 3 and no RD,
 * every 16x16 / 16x8 / 8x16 / 8x8 search of a frame is RECORDED by the x264_me_search_ref hook (predictors, limits, the C result);
 * when the frame's last macroblock has been encoded, one x264_cuda_me_search_mb launch per reference frame searches all recorded
 * partitions of all macroblocks at once, one x264_cuda_me_search_small (SEEDED) launch adds the sub-pel tail, and every result must
 * equal what the C code returned at the time — the data flow of INTEGRATION.md section 3 / bench.py, inside the real encoder.
 * (A 16x16 search may carry up to twelve extra predictors, a sub-partition three: what the batched job can hold.) */
typedef struct { x264_frame_t *ref; int mb, part; x264_cuda_me_job_t job; int16_t mv[2]; int cost, cost_mv; } rec_t;
static rec_t *recs;
static int n_recs, cap_recs, rec_frame = -1;
static long long n_mb_batched, n_mb_launches, n_mb_unrep;

static void report_batched(void)
{
    fprintf(stderr, "ref_cuda_shim: batched ESA: %lld partition searches in %lld x264_cuda_me_search_mb launches equal to the C results; %lld not representable\n",
            n_mb_batched, n_mb_launches, n_mb_unrep);
}
static void record_search(x264_t *h, x264_frame_t *ref, const x264_cuda_me_job_t *j, const x264_me_t *m)
{
    static int once;
    if (!once++) atexit(report_batched);
    if (h->mb.i_me_method != X264_ME_ESA || h->mb.i_subpel_refine >= 3 || m->i_pixel > PIXEL_8x8) return;
    if (rec_frame != h->fenc->i_frame) { n_recs = 0; rec_frame = h->fenc->i_frame; }
    const int ox = j->bx & 15, oy = j->by & 15;
    const int part = m->i_pixel == PIXEL_16x16 ? 0 : m->i_pixel == PIXEL_16x8 ? 1 + (oy >> 3) : m->i_pixel == PIXEL_8x16 ? 3 + (ox >> 3) : 5 + (oy >> 3) * 2 + (ox >> 3);
    /* the 16x16 search may carry up to 4 + 7 predictors (X264_CUDA_ME_MB_MVC16); a sub-partition whose fourth slot it borrows only three */
    if (j->i_mvc > (part ? X264_CUDA_ME_MB_MVC : X264_CUDA_ME_MB_MVC + X264_CUDA_ME_MB_MVC16_EXTRA)) { n_mb_unrep++; return; }
    if (part && j->i_mvc == X264_CUDA_ME_MB_MVC)
        for (int i = 0; i < n_recs; i++)
            if (recs[i].ref == ref && recs[i].mb == h->mb.i_mb_xy && recs[i].part == 0 && recs[i].job.i_mvc - X264_CUDA_ME_MB_MVC > part - 1) { n_mb_unrep++; return; }
    const int mb = h->mb.i_mb_xy;
    for (int i = 0; i < n_recs; i++)
        if (recs[i].ref == ref && recs[i].mb == mb && recs[i].part == part) { n_mb_unrep++; return; } /* searched twice: keep the first */
    if (n_recs == cap_recs) { cap_recs = cap_recs ? 2 * cap_recs : 1024; recs = realloc(recs, cap_recs * sizeof(rec_t)); }
    rec_t *r = &recs[n_recs++];
    r->ref = ref; r->mb = mb; r->part = part; r->job = *j; r->mv[0] = m->mv[0]; r->mv[1] = m->mv[1]; r->cost = m->cost; r->cost_mv = m->cost_mv;
}
static void flush_batched(x264_t *h)
{
    if (!n_recs || rec_frame != h->fenc->i_frame) return;
    const int n_mb = h->mb.i_mb_count, range = h->param.analyse.i_me_range, subme = h->mb.i_subpel_refine;
    x264_cuda_me_mb_job_t *jobs = calloc(n_mb, sizeof(*jobs));
    x264_cuda_me_mb_result_t *res = malloc(n_mb * sizeof(*res));
    x264_cuda_me_job_t *tail = malloc(n_recs * sizeof(*tail));
    x264_cuda_me_final_t *fin = malloc(n_recs * sizeof(*fin));
    int *idx = malloc(n_recs * sizeof(int));
    x264_cuda_frame_t *denc = dev_frame(h, h->fenc, 0);
    for (int done = 0; done < n_recs;) {
        x264_frame_t *ref = NULL;
        int n = 0, n_jobs = 0;
        for (int i = 0; i < n_recs; i++) /* next reference frame with unprocessed records */
            if (recs[i].mb >= 0 && (!ref || recs[i].ref == ref)) { ref = recs[i].ref; idx[n++] = i; }
        int *slot = malloc(n_mb * sizeof(int));
        for (int i = 0; i < n_mb; i++) slot[i] = -1;
        for (int k = 0; k < n; k++) {
            const rec_t *r = &recs[idx[k]];
            if (slot[r->mb] < 0) {
                slot[r->mb] = n_jobs;
                x264_cuda_me_mb_job_t *J = &jobs[n_jobs++];
                memset(J, 0, sizeof(*J));
                J->mb_x = r->job.bx >> 4; J->mb_y = r->job.by >> 4; J->qp = r->job.qp;
                for (int c = 0; c < 2; c++) { J->mv_min_fpel[c] = r->job.mv_min_fpel[c]; J->mv_max_fpel[c] = r->job.mv_max_fpel[c]; }
            }
            x264_cuda_me_mb_job_t *J = &jobs[slot[r->mb]];
            J->part_mask |= 1 << r->part;
            J->i_mvc[r->part] = r->job.i_mvc;
            J->mvp[r->part][0] = r->job.mvp[0]; J->mvp[r->part][1] = r->job.mvp[1];
            for (int c = 0; c < r->job.i_mvc; c++) {
                int16_t *dst = r->part ? J->mvc[r->part][c] : X264_CUDA_ME_MB_MVC16(J, c);
                dst[0] = r->job.mvc[c][0]; dst[1] = r->job.mvc[c][1];
            }
        }
        x264_cuda_frame_t *dref = dev_frame(h, ref, 1);
        ck(x264_cuda_me_search_mb(fctx, denc, dref, range, jobs, n_jobs, res), "me_search_mb");
        n_mb_launches++;
        for (int k = 0; k < n; k++) { /* sub-pel tail of every search, seeded with the batched kernel's full-pel winner (me.c:603-631) */
            const rec_t *r = &recs[idx[k]];
            const x264_cuda_me_result_t *w = &res[slot[r->mb]].part[r->part];
            tail[k] = r->job;
            tail[k].seed_mv[0] = w->bmx; tail[k].seed_mv[1] = w->bmy; tail[k].seed_cost = w->bcost;
        }
        ck(x264_cuda_me_search_small(fctx, denc, dref, X264_CUDA_ME_METHOD_SEEDED, range, subme, tail, n, fin), "me_search_small (seeded batch)");
        for (int k = 0; k < n; k++) {
            rec_t *r = &recs[idx[k]];
            if (fin[k].mv[0] != r->mv[0] || fin[k].mv[1] != r->mv[1] || fin[k].cost != r->cost || fin[k].cost_mv != r->cost_mv) {
                fprintf(stderr, "ref_cuda_shim: batched ESA differs at frame %d mb %d part %d: reference mv (%d,%d) cost %d, device mv (%d,%d) cost %d\n", rec_frame,
                        r->mb, r->part, r->mv[0], r->mv[1], r->cost, fin[k].mv[0], fin[k].mv[1], fin[k].cost);
                exit(5);
            }
            r->mb = -1;
            n_mb_batched++;
        }
        done += n;
        free(slot);
    }
    n_recs = 0;
    free(jobs); free(res); free(tail); free(fin); free(idx);
}

/* ------------------------------------------------------------------------------------------------------------------------------
 * Intra hook on live data: after x264_macroblock_analyse has decided an intra macroblock without RD, the device
 * evaluates the Intra16x16 and chroma candidates of that macroblock (x264_cuda_intra_mb_costs) from the same source pixels and the
 * same neighbouring reconstruction (still around p_fdec), with the macroblock's neighbour mask and lambda; its best chroma mode must be
 * the encoder's h->mb.i_chroma_pred_mode and, when Intra16x16 won, its best 16x16 mode h->mb.i_intra16x16_pred_mode.  Exit 8. */
void x264_macroblock_analyse_c(x264_t *h);
extern const int x264_lambda_tab[52];
static long long n_intra16, n_intra_chroma;

static void report_intra(void)
{
    fprintf(stderr, "ref_cuda_shim: intra hooks: %lld Intra16x16 decisions, %lld chroma mode decisions repeated on the device and equal\n", n_intra16, n_intra_chroma);
}

void x264_macroblock_analyse(x264_t *h)
{
    x264_macroblock_analyse_c(h);
    const int subme = h->param.analyse.i_subpel_refine - (h->sh.i_type == SLICE_TYPE_B);
    if (!resid_hooks_on() || h->sh.b_mbaff || h->mb.b_lossless || subme >= 6 || !IS_INTRA(h->mb.i_type) || h->mb.i_type == I_PCM) return;
    static int once;
    if (!once++) atexit(report_intra);
    resid_ctx(h);
    x264_frame_t *f = h->fdec;
    /* neighbours of this macroblock as the encoder sees them: row above, column left, corner of each p_fdec tile */
    for (int i = 0; i < 3; i++) {
        const int n = 16 >> !!i, s = f->i_stride[i];
        uint8_t *dst = shadow[i] + (size_t)(PADV >> !!i) * s + (PADH >> !!i) + (size_t)h->mb.i_mb_y * n * s + h->mb.i_mb_x * n; /* picture origin inside the bordered plane */
        const uint8_t *src = h->mb.pic.p_fdec[i];
        memcpy(dst - s - 1, src - FDEC_STRIDE - 1, n + 1);
        for (int y = 0; y < n; y++) dst[(ptrdiff_t)y * s - 1] = src[y * FDEC_STRIDE - 1];
    }
    static x264_cuda_frame_t *fnb;
    if (!fnb) {
        fnb = x264_cuda_frame_new(fctx, f->i_width[0] + 2 * PADH, f->i_lines[0] + 2 * PADV, X264_CUDA_FRAME_CHROMA);
        if (!fnb) ck(-1, "x264_cuda_frame_new");
    }
    /* the bordered shadow planes are uploaded as one bigger picture, so that row -1 / column -1 of the first macroblocks exist too;
     * the source frame goes into a frame of the same geometry at the same offset */
    static x264_cuda_frame_t *fsrc;
    static uint8_t *src_shadow[3];
    if (!fsrc) {
        fsrc = x264_cuda_frame_new(fctx, f->i_width[0] + 2 * PADH, f->i_lines[0] + 2 * PADV, X264_CUDA_FRAME_CHROMA);
        if (!fsrc) ck(-1, "x264_cuda_frame_new");
        for (int i = 0; i < 3; i++) src_shadow[i] = calloc((size_t)f->i_stride[i] * (f->i_lines[i] + 2 * PADV), 1);
    }
    for (int i = 0; i < 3; i++) {
        const int n = 16 >> !!i, s = f->i_stride[i];
        uint8_t *dst = src_shadow[i] + (size_t)(PADV >> !!i) * s + (PADH >> !!i) + (size_t)h->mb.i_mb_y * n * s + h->mb.i_mb_x * n;
        for (int y = 0; y < n; y++) memcpy(dst + (size_t)y * s, h->mb.pic.p_fenc[i] + y * FENC_STRIDE, n);
    }
    const int W = f->i_width[0] + 2 * PADH, H = f->i_lines[0] + 2 * PADV;
    ck(x264_cuda_frame_upload(fctx, fnb, shadow[0], f->i_stride[0], W, H), "upload nb");
    ck(x264_cuda_frame_upload_chroma(fctx, fnb, X264_CUDA_PLANE_CB, shadow[1], f->i_stride[1], W / 2, H / 2), "upload nb cb");
    ck(x264_cuda_frame_upload_chroma(fctx, fnb, X264_CUDA_PLANE_CR, shadow[2], f->i_stride[2], W / 2, H / 2), "upload nb cr");
    ck(x264_cuda_frame_upload(fctx, fsrc, src_shadow[0], f->i_stride[0], W, H), "upload src");
    ck(x264_cuda_frame_upload_chroma(fctx, fsrc, X264_CUDA_PLANE_CB, src_shadow[1], f->i_stride[1], W / 2, H / 2), "upload src cb");
    ck(x264_cuda_frame_upload_chroma(fctx, fsrc, X264_CUDA_PLANE_CR, src_shadow[2], f->i_stride[2], W / 2, H / 2), "upload src cr");
    x264_cuda_intra_job_t j;
    memset(&j, 0, sizeof(j));
    j.mb_x = h->mb.i_mb_x + PADH / 16; j.mb_y = h->mb.i_mb_y + PADV / 16; /* position inside the bordered picture */
    j.neighbour = h->mb.i_neighbour;
    j.flags = (h->pixf.mbcmp[0] == h->pixf.satd[0] ? X264_CUDA_INTRA_SATD : 0) | (h->sh.i_type == SLICE_TYPE_B ? X264_CUDA_INTRA_SLICE_B : 0);
    j.lambda = x264_lambda_tab[h->mb.i_qp];
    x264_cuda_intra_result_t r;
    ck(x264_cuda_intra_mb_costs(fctx, fsrc, fnb, &j, 1, &r), "intra_mb_costs");
    if (r.mode_chroma != h->mb.i_chroma_pred_mode || (h->mb.i_type == I_16x16 && r.mode16 != h->mb.i_intra16x16_pred_mode)) {
        fprintf(stderr, "ref_cuda_shim: intra decision differs at frame %d mb (%d,%d) type %d: reference 16x16 mode %d chroma mode %d, device %d / %d\n",
                h->fenc->i_frame, h->mb.i_mb_x, h->mb.i_mb_y, h->mb.i_type, h->mb.i_intra16x16_pred_mode, h->mb.i_chroma_pred_mode, r.mode16, r.mode_chroma);
        exit(8);
    }
    n_intra_chroma++;
    n_intra16 += h->mb.i_type == I_16x16;
}

/* ------------------------------------------------------------------------------------------------------------------------------
 * Motion-compensation hook: the prediction x264_mb_mc has just formed in p_fdec for an inter macroblock is formed again on the device
 * with the frame-batched x264_cuda_mc_blocks / x264_cuda_mc_blocks_bi — same partition walk as S/common/macroblock.c:560-655, vectors
 * and reference indices from the macroblock cache, clipped like x264_mb_mc_0xywh (:462-467) — into a device frame, and the macroblock's
 * luma and chroma pixels must be equal (exit 9).  Called from the x264_macroblock_encode wrapper. */
static x264_cuda_frame_t *fmc;
static long long n_mc_mb, n_mc_rect;
static void report_mc(void)
{
    fprintf(stderr, "ref_cuda_shim: mc hooks: %lld inter macroblocks (%lld partition rectangles) predicted on the device and equal\n", n_mc_mb, n_mc_rect);
}
static void mc_rect(x264_t *h, int x, int y, int w, int hh, int lists)
{
    const int i8 = x264_scan8[0] + x + 8 * y;
    int mv[2][2], ref[2];
    for (int l = 0; l < 2; l++) {
        ref[l] = h->mb.cache.ref[l][i8];
        mv[l][0] = x264_clip3(h->mb.cache.mv[l][i8][0], h->mb.mv_min[0], h->mb.mv_max[0]);
        mv[l][1] = x264_clip3(h->mb.cache.mv[l][i8][1], h->mb.mv_min[1], h->mb.mv_max[1]);
    }
    if (lists == 4) lists = ref[0] >= 0 ? (ref[1] >= 0 ? 3 : 1) : 2; /* x264_mb_mc_direct8x8 */
    const int bx = 16 * h->mb.i_mb_x + 4 * x, by = 16 * h->mb.i_mb_y + 4 * y;
    if (lists == 3) {
        x264_cuda_mc_bi_job_t j;
        memset(&j, 0, sizeof(j));
        j.bx = bx; j.by = by; j.w = 4 * w; j.h = 4 * hh; j.weight = h->mb.bipred_weight[ref[0]][ref[1]];
        j.mv0[0] = mv[0][0]; j.mv0[1] = mv[0][1]; j.mv1[0] = mv[1][0]; j.mv1[1] = mv[1][1];
        ck(x264_cuda_mc_blocks_bi(fctx, dev_frame(h, h->fref0[ref[0]], 1), dev_frame(h, h->fref1[ref[1]], 1), fmc, &j, 1), "mc_blocks_bi");
    } else {
        const int l = lists == 2;
        x264_cuda_mc_job_t j;
        memset(&j, 0, sizeof(j));
        j.bx = bx; j.by = by; j.w = 4 * w; j.h = 4 * hh; j.mvx = mv[l][0]; j.mvy = mv[l][1];
        ck(x264_cuda_mc_blocks(fctx, dev_frame(h, l ? h->fref1[ref[l]] : h->fref0[ref[l]], 1), fmc, &j, 1), "mc_blocks");
    }
    n_mc_rect++;
}
static void mc_check(x264_t *h)
{
    x264_frame_t *f = h->fdec;
    if (!fmc) {
        fmc = x264_cuda_frame_new(fctx, f->i_width[0], f->i_lines[0], X264_CUDA_FRAME_CHROMA);
        if (!fmc) ck(-1, "x264_cuda_frame_new");
        atexit(report_mc);
    }
    const int t = h->mb.i_type;
    if (t == P_L0 || (t != P_8x8 && t != B_8x8 && t != B_SKIP && t != B_DIRECT)) {
        const uint8_t *l0 = x264_mb_type_list_table[t][0], *l1 = x264_mb_type_list_table[t][1];
        const int n = h->mb.i_partition == D_16x16 ? 1 : 2;
        for (int k = 0; k < n; k++) {
            const int lists = t == P_L0 ? 1 : (l0[k] ? 1 : 0) | (l1[k] ? 2 : 0);
            if (h->mb.i_partition == D_16x16) mc_rect(h, 0, 0, 4, 4, lists);
            else if (h->mb.i_partition == D_16x8) mc_rect(h, 0, 2 * k, 4, 2, lists);
            else mc_rect(h, 2 * k, 0, 2, 4, lists);
        }
    } else
        for (int i8 = 0; i8 < 4; i8++) {
            const int x = 2 * (i8 & 1), y = 2 * (i8 >> 1);
            if (t == B_SKIP || t == B_DIRECT) { mc_rect(h, x, y, 2, 2, 4); continue; }
            switch (h->mb.i_sub_partition[i8]) {
            case D_L0_8x8: mc_rect(h, x, y, 2, 2, 1); break;
            case D_L0_8x4: mc_rect(h, x, y, 2, 1, 1); mc_rect(h, x, y + 1, 2, 1, 1); break;
            case D_L0_4x8: mc_rect(h, x, y, 1, 2, 1); mc_rect(h, x + 1, y, 1, 2, 1); break;
            case D_L0_4x4: mc_rect(h, x, y, 1, 1, 1); mc_rect(h, x + 1, y, 1, 1, 1); mc_rect(h, x, y + 1, 1, 1, 1); mc_rect(h, x + 1, y + 1, 1, 1, 1); break;
            case D_L1_8x8: mc_rect(h, x, y, 2, 2, 2); break;
            case D_BI_8x8: mc_rect(h, x, y, 2, 2, 3); break;
            case D_DIRECT_8x8: mc_rect(h, x, y, 2, 2, 4); break;
            default: return; /* (B sub-8x8 partitions do not exist in this encoder) */
            }
        }
    static const int ids[3] = { X264_CUDA_PLANE_FULL, X264_CUDA_PLANE_CB, X264_CUDA_PLANE_CR };
    for (int i = 0; i < 3; i++) {
        const int s = f->i_stride[i], padv = PADV >> !!i, padh = PADH >> !!i, n = 16 >> !!i;
        ck(x264_cuda_frame_download(fctx, fmc, ids[i], tmp_plane, s), "download prediction");
        for (int y = 0; y < n; y++)
            if (memcmp(tmp_plane + (size_t)(padv + h->mb.i_mb_y * n + y) * s + padh + h->mb.i_mb_x * n, h->mb.pic.p_fdec[i] + y * FDEC_STRIDE, n)) {
                fprintf(stderr, "ref_cuda_shim: x264_mb_mc differs at frame %d mb (%d,%d) type %d partition %d plane %d row %d\n", h->fenc->i_frame, h->mb.i_mb_x,
                        h->mb.i_mb_y, t, h->mb.i_partition, i, y);
                exit(9);
            }
    }
    n_mc_mb++;
}

/* ------------------------------------------------------------------------------------------------------------------------------
 * Lookahead hook: x264_slicetype_frame_cost is static, but rate control reaches it through the extern x264_rc_analyse_slice
 * (S/encoder/slicetype.c:638-679; CRF / ABR only).  After the C call, the cost of the same (p0, p1, b) is evaluated from scratch on the
 * device (x264_cuda_lowres_frame_cost on the lowres planes the device itself produced when each frame came in) and must reproduce what
 * the reference holds for that frame: i_cost_est, i_intra_mbs, and the motion vectors / costs of every interior block.  Exit 10.
 * B frames would be checked when the index arithmetic of :661-662 agrees with the frame's true position, but rate control does not ask for
 * them in this version; the x264_slicetype_decide wrapper below covers B-type estimates. */
int x264_rc_analyse_slice_c(x264_t *h);
typedef struct { int frame; x264_cuda_frame_t *d; long long used; } lslot_t;
static lslot_t lslots[12];
static long long n_la_p, n_la_b, n_la_skipped;

static void report_la(void)
{
    fprintf(stderr, "ref_cuda_shim: lookahead hooks: %lld P and %lld B frame costs re-evaluated on the device and equal; %lld left to C\n", n_la_p, n_la_b, n_la_skipped);
}
static x264_cuda_frame_t *lowres_slot(x264_t *h, x264_frame_t *f, int frame_no, int create)
{
    lslot_t *victim = &lslots[0];
    for (int i = 0; i < 12; i++) {
        if (lslots[i].d && lslots[i].frame == frame_no) { lslots[i].used = ++slot_clock; return lslots[i].d; }
        if (lslots[i].used < victim->used) victim = &lslots[i];
    }
    if (!create) return NULL;
    if (!victim->d) {
        victim->d = x264_cuda_frame_new(fctx, f->i_width[0], f->i_lines[0], X264_CUDA_FRAME_LOWRES);
        if (!victim->d) ck(-1, "x264_cuda_frame_new");
    }
    victim->frame = frame_no; victim->used = ++slot_clock;
    return victim->d;
}

/* re-evaluates cost(p0 = 0, p1, b) of source frame nb against frames n0 (and n1) on the device, from scratch, and compares it with what the
 * frame object fb holds; f1 = the later reference's object (B only).  Returns 0 when the frames are not on the device (nothing compared). */
static int check_lowres_cost(x264_t *h, x264_frame_t *fb, x264_frame_t *f1, int nb, int n0, int n1, int p1, int b)
{
    x264_cuda_frame_t *db = lowres_slot(h, fb, nb, 0), *d0 = lowres_slot(h, fb, n0, 0), *d1 = lowres_slot(h, fb, n1, 0);
    if (!db || !d0 || !d1 || nb - n0 != b || n1 - n0 != p1) return 0;
    const int n_dist = h->param.i_bframe + 1, n_mb = h->mb.i_mb_count;
    ck(x264_cuda_frame_lookahead_alloc(fctx, db, n_dist), "lookahead_alloc"); /* fresh state: everything is searched again */
    if (b != p1) { /* the direct-like candidate of a B block reads the later reference's own list-0 vectors (slicetype.c:96-112) */
        ck(x264_cuda_frame_lookahead_alloc(fctx, d1, n_dist), "lookahead_alloc");
        ck(x264_cuda_frame_lookahead_set(fctx, d1, 0, p1 - 1, &f1->lowres_mvs[0][p1 - 1][0][0], f1->lowres_mv_costs[0][p1 - 1], NULL), "lookahead_set");
    }
    x264_cuda_lowres_params_t pm;
    memset(&pm, 0, sizeof(pm));
    pm.p0 = 0; pm.p1 = p1; pm.b = b; pm.me_method = h->param.analyse.i_me_method; pm.me_range = h->param.analyse.i_me_range;
    pm.flags = (h->pixf.mbcmp[0] == h->pixf.satd[0] ? X264_CUDA_ME_MBCMP_SATD : 0) | (h->pixf.fpelcmp[0] == h->pixf.satd[0] ? X264_CUDA_ME_FPEL_SATD : 0) |
               (h->param.analyse.b_weighted_bipred ? X264_CUDA_LOWRES_WEIGHTED_BIPRED : 0);
    pm.do_search[0] = 1; pm.do_search[1] = b != p1;
    x264_cuda_lowres_result_t r;
    ck(x264_cuda_lowres_frame_cost(fctx, db, d0, d1, &pm, &r), "lowres_frame_cost");
    int score = r.score;
    if (b != p1) score = score * 100 / (120 + h->param.i_bframe_bias);
    int bad = score != fb->i_cost_est[b][p1 - b];
    if (b == p1) bad |= (r.intra_mbs != fb->i_intra_mbs[b]) << 1;
    int16_t *mv = malloc(n_mb * 4);
    int *cs = malloc(n_mb * sizeof(int));
    const int W = h->sps->i_mb_width, H = h->sps->i_mb_height;
    for (int l = 0; l < 1 + (b != p1) && !bad; l++) {
        const int d = l ? p1 - b - 1 : b - 1;
        ck(x264_cuda_frame_lookahead_get(fctx, db, l, d, mv, cs, NULL), "lookahead_get");
        for (int y = 1; y < H - 1; y++)
            for (int x = 1; x < W - 1; x++) {
                const int i = y * W + x;
                if (mv[2 * i] != fb->lowres_mvs[l][d][i][0] || mv[2 * i + 1] != fb->lowres_mvs[l][d][i][1] || cs[i] != fb->lowres_mv_costs[l][d][i]) bad |= 4 << l;
            }
    }
    free(mv); free(cs);
    if (bad) {
        fprintf(stderr, "ref_cuda_shim: lookahead cost differs for frame %d (p0 %d, p1 %d, b %d): reference score %d intra mbs %d, device %d / %d, mask %d\n", nb, 0,
                p1, b, fb->i_cost_est[b][p1 - b], fb->i_intra_mbs[b], score, r.intra_mbs, bad);
        exit(10);
    }
    if (b == p1) n_la_p++; else n_la_b++;
    return 1;
}

int x264_rc_analyse_slice(x264_t *h)
{
    const int cost = x264_rc_analyse_slice_c(h);
    static int once;
    if (!hooks_on() || h->sh.b_mbaff || h->param.rc.i_vbv_buffer_size || !h->frames.b_have_lowres) return cost;
    if (!once++) atexit(report_la);
    if (IS_X264_TYPE_I(h->fenc->i_type)) return cost;
    x264_frame_t *f0 = h->fref0[0], *f1 = NULL, *fb = h->fenc;
    int p1, b;
    if (h->fenc->i_type == X264_TYPE_P) {
        p1 = 0;
        while (h->frames.current[p1] && IS_X264_TYPE_B(h->frames.current[p1]->i_type)) p1++;
        b = ++p1;
    } else {
        f1 = h->fref1[0];
        p1 = (f1->i_poc - f0->i_poc) / 2;
        b = (f1->i_poc - fb->i_poc) / 2;
        if (b != (fb->i_poc - f0->i_poc) / 2) { n_la_skipped++; return cost; }
    }
    /* the source frames' numbers: reconstructed frames carry poc = 2 * (frame - last idr) (encoder.c:1514-1515) */
    const int nb = fb->i_frame, n0 = f0->i_poc / 2 + h->frames.i_last_idr, n1 = f1 ? f1->i_poc / 2 + h->frames.i_last_idr : nb;
    if (!check_lowres_cost(h, fb, f1, nb, n0, n1, p1, b)) n_la_skipped++;
    return cost;
}

/* x264_slicetype_decide (slicetype.c:577-636) runs the B-adapt analysis, which leaves cost estimates cached in the undecided frames
 * (i_cost_est[b - p0][p1 - b]).  Every estimate this call adds — P-type (p1 == b: a frame predicted from an earlier one at distance
 * 1..bframes+1) and B-type (a frame between two others) — is re-evaluated from scratch on the device and compared. */
void x264_slicetype_decide_c(x264_t *h);
void x264_slicetype_decide(x264_t *h)
{
    x264_frame_t *objs[X264_BFRAME_MAX * 4 + 4] = { NULL };
    int n = 0, n0 = 0;
    const int on = hooks_on() && !h->sh.b_mbaff && h->frames.b_have_lowres && !h->param.rc.i_vbv_buffer_size && h->frames.last_nonb && h->frames.next[0];
    if (on) {
        static int once;
        if (!once++) atexit(report_la);
        n0 = h->frames.last_nonb->i_poc / 2 + h->frames.i_last_idr;
        for (int j = 0; h->frames.next[j] && h->frames.next[j]->i_type == X264_TYPE_AUTO && n < X264_BFRAME_MAX * 4 + 2; j++) objs[++n] = h->frames.next[j];
    }
    int before[X264_BFRAME_MAX * 4 + 4][X264_BFRAME_MAX + 2];
    static int beforeb[X264_BFRAME_MAX * 4 + 4][X264_BFRAME_MAX + 2][X264_BFRAME_MAX + 2];
    for (int k = 1; k <= n; k++)
        for (int d = 1; d <= h->param.i_bframe + 1 && d <= X264_BFRAME_MAX + 1; d++) {
            before[k][d] = objs[k]->i_cost_est[d][0];
            for (int e = 1; e <= h->param.i_bframe + 1 && e <= X264_BFRAME_MAX + 1; e++) beforeb[k][d][e] = objs[k]->i_cost_est[d][e];
        }
    x264_slicetype_decide_c(h);
    /* B-type estimates: frame k between k - d and k + e, provided the later frame's own vectors towards k - d exist (they feed the direct-like
     * candidate, slicetype.c:96-112, and the B-adapt analysis evaluates that P cost first) */
    if (on)
        for (int k = 1; k < n; k++) {
            x264_frame_t *fb = objs[k];
            if (fb->i_frame != n0 + k) continue;
            for (int d = 1; d <= k && d <= h->param.i_bframe + 1; d++)
                for (int e = 1; k + e <= n && d + e <= h->param.i_bframe + 1; e++) {
                    x264_frame_t *f1 = objs[k + e];
                    if (fb->i_cost_est[d][e] < 0 || beforeb[k][d][e] >= 0 || f1->i_frame != n0 + k + e) continue;
                    if (f1->lowres_mvs[0][d + e - 1][0][0] == 0x7FFF) { n_la_skipped++; continue; }
                    if (!check_lowres_cost(h, fb, f1, fb->i_frame, fb->i_frame - d, fb->i_frame + e, d + e, d)) n_la_skipped++;
                }
        }
    for (int k = 1; k <= n; k++) {
        x264_frame_t *fb = objs[k];
        if (fb->i_frame != n0 + k) continue; /* (consecutive numbering expected: last non-B, then the undecided frames) */
        for (int d = 1; d <= k && d <= h->param.i_bframe + 1; d++)
            if (fb->i_cost_est[d][0] >= 0 && before[k][d] < 0)   /* newly evaluated: frame n0 + k predicted from frame n0 + k - d */
                if (!check_lowres_cost(h, fb, NULL, fb->i_frame, fb->i_frame - d, fb->i_frame, d, d)) n_la_skipped++;
    }
}
