/* x264_b200_hooks.c — performance-mode integration of the B200 back-end into the reference encoder
 * (x264-snapshot-20090216-2245; S/ = the reference tree).
 *
 * This is the code a maintainer of the reference adds.  No reference source is edited: integration/Makefile
 * compiles the files of S/ where they lie and renames, at compile time, the handful of functions this file takes over
 * (-Dx264_me_search_ref=x264_me_search_ref_c, ...), exactly the way S/common/pixel.c:781-791 lets a back-end override table entries.
 * Unlike the verification shim under oracle/ (which lets the C code compute and only compares), the device results are USED here:
 *
 *   x264_me_search_ref, --me esa (S/encoder/me.c:156-631)
 *       The exhaustive loop (me.c:449-600: up to 1056 SAD + ADS evaluations per search, 5-9 searches per macroblock) never runs on
 *       the host.  Before a slice's macroblock loop reaches a macroblock row, the device has written the SADs of that row's
 *       macroblocks at every integer vector of a window (x264_cuda_sad_grid_quad, four 8x8 quadrant SADs per position, page-locked host
 *       memory, copied back while the host encodes earlier rows).  The search itself — predictor stage, strict-'<' raster argmin with
 *       the lambda-weighted mv cost relative to the sequentially known mvp, "-> qpel", sub-pel refinement — stays in the reference's
 *       order on the host and reads its SADs from the grid, so vectors, costs and the bitstream are unchanged.  A window that leaves
 *       the grid (the guessed centre was off) is recomputed for that macroblock by a one-job launch around the exact centre.
 *       Sub-8x8 partitions (no quadrant sums) run as one-job x264_cuda_me_search calls seeded with the host's predictor stage.
 *   end of a reconstructed frame (x264_fdec_filter_row, S/encoder/encoder.c:983-1057)
 *       The per-row x264_frame_deblock_row / x264_frame_expand_border / x264_frame_filter / x264_frame_expand_border_filtered calls
 *       become ONE device pass when the last row is reached: upload the unfiltered reconstruction and the encoder's own per-macroblock
 *       arrays, x264_cuda_frame_deblock, x264_cuda_frame_expand_border, x264_cuda_frame_filter; the deblocked planes and the three
 *       half-pel planes come back to the host frame (its sub-pel code reads them) and the device copy stays resident as the reference
 *       picture of the next frames' grids.  Whole-frame order is byte-exact for --threads 1.
 *
 * There is no CPU fallback: with the back-end enabled (default; X264_B200=0 gives the plain reference) a missing device is fatal.
 * Requirements: --threads 1 (the bit-exact configuration), progressive.  Anything else runs the reference's code untouched.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "common/common.h"
#include "encoder/me.h"
#include "x264_cuda.h"

void x264_me_search_ref_c(x264_t *h, x264_me_t *m, int16_t (*mvc)[2], int i_mvc, int *p_halfpel_thresh);
void x264_frame_deblock_row_c(x264_t *h, int mb_y);
void x264_frame_expand_border_c(x264_t *h, x264_frame_t *frame, int mb_y, int b_end);
void x264_frame_expand_border_filtered_c(x264_t *h, x264_frame_t *frame, int mb_y, int b_end);
void x264_frame_filter_c(x264_t *h, x264_frame_t *frame, int mb_y, int b_end);
int64_t x264_pixel_ssd_wxh_c(x264_pixel_function_t *pf, uint8_t *pix1, int i_pix1, uint8_t *pix2, int i_pix2, int i_width, int i_height);
float x264_pixel_ssim_wxh_c(x264_pixel_function_t *pf, uint8_t *pix1, int i_pix1, uint8_t *pix2, int i_pix2, int i_width, int i_height, void *buf);
extern int16_t *g_cost_mv[52]; /* S/encoder/analyse.c:179 */

#define N_SLOTS 20 /* device mirrors of reconstructed frames: i_frame_reference <= 16 plus the frames in flight */
#define N_GRIDSETS 4 /* (source frame, reference frame) pairs with grids in host memory */
#define MAX_CHUNKS 512
#define N_EXTRA 8
#define N_RING 3   /* chunks of macroblock rows resident in host memory per grid set: the one being read, the next one arriving, one spare */

typedef struct {
    x264_frame_t *f; int i_frame, i_poc; /* which host frame content this mirrors */
    long long used;
    x264_cuda_frame_t *d;
} dev_slot_t;

typedef struct {
    x264_frame_t *ref; int ref_frame, ref_poc; /* identity of the reference picture */
    int enc_frame, enc_type;                   /* h->fenc->i_frame / slice type the grids were made for; -1: empty */
    long long used;
    uint16_t *grid;                            /* page-locked ring of N_RING chunks of rows_per_chunk * mb_w macroblock grids (GW * GH * 4 each) */
    x264_cuda_grid_job_t *jobs;                /* page-locked: n_mb (centre + limits of every macroblock) */
    void *fence[MAX_CHUNKS];                   /* completion of each chunk of macroblock rows (NULL: not in flight) */
    uint8_t issued[MAX_CHUNKS];
    int n_chunks, rows_per_chunk, list;
    /* grids recomputed around an exact centre (the guess was off): a few of them are kept beside the frame's grids, because the partitions
     * of a macroblock on a motion boundary alternate between two centres */
    uint16_t *extra_grid;                      /* page-locked: N_EXTRA grids */
    x264_cuda_grid_job_t *extra_job;           /* page-locked: the job (macroblock, centre) each one was computed for */
    long long extra_used[N_EXTRA];
} gridset_t;

/* all state is per host thread: several encoder instances may run side by side in one process (x264_b200_gops.c), each with its own
 * device context and stream */
static __thread struct {
    int state; /* 0: undecided, 1: on, -1: off (plain reference) */
    int verbose, me_on, frame_on, check, avx2, reported; /* X264_B200_ME=0 / X264_B200_FRAME=0 switch one of the two hook groups off (diagnosis) */
    x264_cuda_t *ctx;
    int radius, flags, chunk_mbs;
    int mb_w, mb_h;
    dev_slot_t slot[N_SLOTS];
    gridset_t gs[N_GRIDSETS];
    long long clock;
    x264_cuda_frame_t *denc; x264_frame_t *enc_f; int enc_frame; /* the source picture on the device */
    double t_pin_pending; int deblock_pending;   /* x264_frame_deblock_row was asked for rows of the current fdec */
    x264_frame_t *end_done; int end_done_frame; /* fdec whose end-of-frame pass has run */
    int cost_uploaded[52];
    /* statistics */
    long long n_search, n_extra_hit, n_relaunch, n_percall, n_outside_pred, n_gridsets, n_frame_end, n_left_to_c;
    double t_grid_issue, t_grid_wait, t_frame_end, t_relaunch, t_open, t_alloc, t_alloc_in_issue, t_alloc_in_end;
    /* deferred PSNR / SSIM slabs (see x264_pixel_ssd_wxh below) */
    struct { int y0, h; } ssim_slab[256]; int n_ssim_slab;
} B;

static double now_ms(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
static void die(const char *what)
{
    fprintf(stderr, "x264_b200: %s: %s\n", what, x264_cuda_error(B.ctx));
    exit(3);
}
#define CK(call) do { if ((call) < 0) die(#call); } while (0)

static void report(void)
{
    if (!B.verbose || B.reported) return;
    B.reported = 1;
    fprintf(stderr, "x264_b200: %lld ESA searches read device grids (%lld macroblock grids recomputed around the exact centre and %lld reused, %lld predictor SADs outside a grid "
            "taken from the table entry, %lld sub-8x8 searches as one-job device calls, %lld searches left to the reference), %lld frame grid sets, "
            "%lld end-of-frame device passes; %lld kernel launches\n", B.n_search, B.n_relaunch, B.n_extra_hit, B.n_outside_pred, B.n_percall, B.n_left_to_c, B.n_gridsets,
            B.n_frame_end, B.ctx ? x264_cuda_launch_count(B.ctx) : 0);
    fprintf(stderr, "x264_b200: host time in device calls: open %.1f ms, page-locking and allocation %.1f ms, grid issue %.1f ms, grid wait %.1f ms, recompute %.1f ms, "
            "end of frame %.1f ms\n", B.t_open, B.t_alloc, B.t_grid_issue - B.t_alloc_in_issue, B.t_grid_wait, B.t_relaunch, B.t_frame_end - B.t_alloc_in_end);
}

/* Is the back-end in charge of this encoder?  Decided once. */
static int b200_on(x264_t *h)
{
    if (B.state) return B.state > 0;
    const char *e = getenv("X264_B200");
    B.verbose = getenv("X264_B200_VERBOSE") ? atoi(getenv("X264_B200_VERBOSE")) : 0;
    B.state = -1;
    if (e && !atoi(e)) return 0;
    if (h->param.i_threads > 1 || h->param.b_interlaced) {
        fprintf(stderr, "x264_b200: --threads 1 and progressive input are required (SURVEY.md F3); running the reference's own code\n");
        return 0;
    }
    const double t_open = now_ms();
    if (x264_cuda_open(&B.ctx, getenv("X264_B200_DEVICE") ? atoi(getenv("X264_B200_DEVICE")) : 0) < 0) {
        fprintf(stderr, "x264_b200: %s\n", x264_cuda_error(NULL));
        exit(3); /* no CPU fallback */
    }
    B.mb_w = h->sps->i_mb_width; B.mb_h = h->sps->i_mb_height;
    const int slack = getenv("X264_B200_GRID_SLACK") ? atoi(getenv("X264_B200_GRID_SLACK")) : 8;
    B.chunk_mbs = getenv("X264_B200_CHUNK_MBS") ? atoi(getenv("X264_B200_CHUNK_MBS")) : 960; /* macroblocks per grid launch (whole rows) */
    B.radius = h->param.analyse.i_me_range + (slack < 2 ? 2 : slack); /* >= merange + 2: the width rounding of me.c:457 */
    if (B.radius > 64) B.radius = 64;
    B.flags = X264_CUDA_FRAME_CHROMA | (h->param.analyse.i_subpel_refine ? X264_CUDA_FRAME_HPEL : 0);
    if (h->param.analyse.i_me_method == X264_ME_TESA) /* the reference's own TESA reads the integral image on the host */
        B.flags |= X264_CUDA_FRAME_INTEGRAL | (h->frames.b_have_sub8x8_esa ? X264_CUDA_FRAME_INTEGRAL4 : 0);
    B.me_on = !getenv("X264_B200_ME") || atoi(getenv("X264_B200_ME"));
    B.avx2 = __builtin_cpu_supports("avx2") && !(getenv("X264_B200_AVX2") && !atoi(getenv("X264_B200_AVX2")));
    B.check = getenv("X264_B200_CHECK") && atoi(getenv("X264_B200_CHECK"));
    B.frame_on = !getenv("X264_B200_FRAME") || atoi(getenv("X264_B200_FRAME"));
    for (int i = 0; i < N_GRIDSETS; i++) B.gs[i].enc_frame = -1;
    B.enc_frame = -1;
    B.state = 1;
    B.t_open = now_ms() - t_open;
    atexit(report);
    return 1;
}

/* GOP-sharded encoding (x264-vs2008_b200/gop_shard.py): a worker that encodes closed GOP k of a longer stream must number its IDR
 * pictures from k, as the single-process encoder would (S/encoder/encoder.c:1107-1110), and start its count of coded frames where that
 * encoder would be (h->i_frame selects the signature bit the CABAC flush embeds in every slice, S/common/cabac.c:917); everything else
 * restarts at an IDR anyway. */
static __thread int seed_set, seed_idr_pic_id, seed_coded_frames;
void x264_b200_set_gop_seed(int idr_pic_id, int coded_frames) { seed_set = 1; seed_idr_pic_id = idr_pic_id; seed_coded_frames = coded_frames; }
void x264_b200_disable_for_this_thread(void) { B.state = -1; }
void x264_b200_report(void) { report(); }
/* creates the process's CUDA context (1-3 s on a freshly started box) from a helper thread, so that a front end can overlap it with its own
 * start-up work; the encoder threads' x264_cuda_open calls then find the context in place */
void x264_b200_warm_device(void)
{
    const char *e = getenv("X264_B200");
    if (e && !atoi(e)) return;
    x264_cuda_t *c = NULL;
    if (x264_cuda_open(&c, getenv("X264_B200_DEVICE") ? atoi(getenv("X264_B200_DEVICE")) : 0) == 0 && c) { x264_cuda_synchronize(c); x264_cuda_close(c); }
}
x264_t *x264_encoder_open_c(x264_param_t *param);
x264_t *x264_encoder_open(x264_param_t *param)
{
    x264_t *h = x264_encoder_open_c(param);
    if (!h) return h;
    if (seed_set) { /* an in-process front end (x264_b200_gops.c) */
        h->i_idr_pic_id = seed_idr_pic_id & 0xffff; h->i_frame = seed_coded_frames;
        seed_set = 0;
        return h;
    }
    const char *e = getenv("X264_B200_IDR_PIC_ID"); /* a CLI worker started by gop_shard.py */
    if (e) h->i_idr_pic_id = atoi(e) & 0xffff;
    e = getenv("X264_B200_CODED_FRAMES");
    if (e) h->i_frame = atoi(e);
    return h;
}

/* ------------------------------------------------------------------------------------------------------------------------------
 * device mirrors of frames */
static dev_slot_t *slot_find(x264_frame_t *f)
{
    for (int i = 0; i < N_SLOTS; i++)
        if (B.slot[i].d && B.slot[i].f == f && B.slot[i].i_frame == f->i_frame && B.slot[i].i_poc == f->i_poc) { B.slot[i].used = ++B.clock; return &B.slot[i]; }
    return NULL;
}
static dev_slot_t *slot_take(x264_frame_t *f)
{
    dev_slot_t *v = NULL;
    for (int i = 0; i < N_SLOTS && !v; i++) /* a mirror of the same host frame object is stale by construction: reuse it first */
        if (B.slot[i].d && B.slot[i].f == f) v = &B.slot[i];
    for (int i = 0; i < N_SLOTS && !v; i++)
        if (!B.slot[i].d) v = &B.slot[i];
    if (!v) {
        v = &B.slot[0];
        for (int i = 1; i < N_SLOTS; i++) if (B.slot[i].used < v->used) v = &B.slot[i];
    }
    if (!v->d && !(v->d = x264_cuda_frame_new(B.ctx, f->i_width[0], f->i_lines[0], B.flags))) die("x264_cuda_frame_new");
    v->f = f; v->i_frame = f->i_frame; v->i_poc = f->i_poc; v->used = ++B.clock;
    return v;
}
/* x264 frames live in a small recycled pool (S/common/frame.c:898-975): page-lock each buffer the first time it is seen, so that the
 * plane copies are DMA'd straight from / to it (a 2-D copy through pageable memory costs ~10 ms per 1080p frame) */
static void pin_frame(x264_t *h, x264_frame_t *f)
{
    static __thread void *seen[256];
    static __thread int n_seen;
    for (int i = 0; i < n_seen; i++) if (seen[i] == f->buffer[0]) return;
    if (n_seen == 256) return;
    const double t0 = now_ms();
    seen[n_seen++] = f->buffer[0];
    const size_t luma = (size_t)f->i_stride[0] * (f->i_lines[0] + 2 * PADV), chroma = (size_t)f->i_stride[1] * (f->i_lines[1] + 2 * PADV);
    x264_cuda_host_register(f->buffer[0], (h->param.analyse.i_subpel_refine ? 4 : 1) * luma); /* failure just leaves the slower pageable path */
    x264_cuda_host_register(f->buffer[1], chroma);
    x264_cuda_host_register(f->buffer[2], chroma);
    B.t_alloc += now_ms() - t0; B.t_pin_pending += now_ms() - t0;
}
static void upload_picture(x264_cuda_frame_t *d, x264_frame_t *f, int chroma)
{
    CK(x264_cuda_frame_upload(B.ctx, d, f->plane[0], f->i_stride[0], f->i_width[0], f->i_lines[0]));
    if (chroma) {
        CK(x264_cuda_frame_upload_chroma(B.ctx, d, X264_CUDA_PLANE_CB, f->plane[1], f->i_stride[1], f->i_width[1], f->i_lines[1]));
        CK(x264_cuda_frame_upload_chroma(B.ctx, d, X264_CUDA_PLANE_CR, f->plane[2], f->i_stride[2], f->i_width[2], f->i_lines[2]));
    }
}
/* a reference picture that did not come through the end-of-frame pass (should not happen; kept so that a miss is not fatal): upload the
 * host's finished planes */
static dev_slot_t *slot_for_ref(x264_frame_t *f)
{
    dev_slot_t *s = slot_find(f);
    if (s) return s;
    s = slot_take(f);
    upload_picture(s->d, f, 1);
    CK(x264_cuda_frame_expand_border(B.ctx, s->d));
    return s;
}
static x264_cuda_frame_t *source_on_device(x264_t *h)
{
    x264_frame_t *f = h->fenc;
    if (B.denc && B.enc_f == f && B.enc_frame == f->i_frame) return B.denc;
    if (!B.denc && !(B.denc = x264_cuda_frame_new(B.ctx, f->i_width[0], f->i_lines[0], 0))) die("x264_cuda_frame_new");
    pin_frame(h, f);
    upload_picture(B.denc, f, 0); /* already padded to the macroblock grid by x264_frame_expand_border_mod16 (encoder.c:1411) */
    B.enc_f = f; B.enc_frame = f->i_frame;
    return B.denc;
}

/* ------------------------------------------------------------------------------------------------------------------------------
 * end of a reconstructed frame */
static void frame_end(x264_t *h, x264_frame_t *f)
{
    const double t0 = now_ms();
    dev_slot_t *s = slot_take(f);
    pin_frame(h, f);
    upload_picture(s->d, f, 1);
    if (B.deblock_pending) {
        x264_cuda_deblock_params_t p = { h->sh.i_alpha_c0_offset, h->sh.i_beta_offset, h->pps->i_chroma_qp_index_offset, h->sh.i_type == SLICE_TYPE_B,
                                         !!(h->param.analyse.inter & X264_ANALYSE_PSUB8x8), !h->pps->b_cabac && h->pps->b_transform_8x8_mode };
        CK(x264_cuda_frame_deblock(B.ctx, s->d, &p, h->mb.type, h->mb.qp, h->mb.mb_transform_size, (const uint8_t(*)[24])h->mb.non_zero_count, h->mb.ref[0],
                                   (const int16_t(*)[2])h->mb.mv[0], h->mb.ref[1], (const int16_t(*)[2])h->mb.mv[1]));
    }
    CK(x264_cuda_frame_expand_border(B.ctx, s->d));
    if (h->param.analyse.i_subpel_refine) CK(x264_cuda_frame_filter(B.ctx, s->d));
    /* back to the host frame: the padded planes its motion compensation and sub-pel search read */
    static const int ids[3] = { X264_CUDA_PLANE_FULL, X264_CUDA_PLANE_CB, X264_CUDA_PLANE_CR };
    for (int i = 0; i < 3; i++) {
        const int st = f->i_stride[i], padv = PADV >> !!i, padh = PADH >> !!i;
        CK(x264_cuda_frame_download(B.ctx, s->d, ids[i], f->plane[i] - (st * padv + padh), st));
    }
    if (h->param.analyse.i_subpel_refine)
        for (int k = 1; k < 4; k++)
            CK(x264_cuda_frame_download(B.ctx, s->d, k, f->filtered[k] - (f->i_stride[0] * PADV + PADH), f->i_stride[0]));
    if (f->integral && (B.flags & X264_CUDA_FRAME_INTEGRAL)) {
        CK(x264_cuda_frame_download(B.ctx, s->d, X264_CUDA_PLANE_INTEGRAL, f->integral - (f->i_stride[0] * PADV + PADH), f->i_stride[0]));
        if (B.flags & X264_CUDA_FRAME_INTEGRAL4)
            CK(x264_cuda_frame_download(B.ctx, s->d, X264_CUDA_PLANE_INTEGRAL4,
                                        f->integral + (size_t)f->i_stride[0] * (f->i_lines[0] + 2 * PADV) - (f->i_stride[0] * PADV + PADH), f->i_stride[0]));
    }
    B.deblock_pending = 0;
    B.end_done = f; B.end_done_frame = f->i_frame;
    B.n_frame_end++;
    B.t_frame_end += now_ms() - t0;
    B.t_alloc_in_end += B.t_pin_pending; B.t_pin_pending = 0;
}
static __thread x264_t *g_h; /* the encoder handle, for the hooks whose reference signature does not carry it (one thread) */
static int frame_hooks_on(x264_t *h) { return b200_on(h) && B.frame_on && h->fdec->b_kept_as_ref && !h->sh.b_mbaff; }

void x264_frame_deblock_row(x264_t *h, int mb_y)
{
    g_h = h;
    if (!frame_hooks_on(h)) { x264_frame_deblock_row_c(h, mb_y); return; }
    B.deblock_pending = 1; /* filtered with the whole frame when x264_frame_expand_border sees the last row (same x264_fdec_filter_row call) */
}
void x264_frame_expand_border(x264_t *h, x264_frame_t *frame, int mb_y, int b_end)
{
    g_h = h;
    if (!b200_on(h) || frame != h->fdec || !frame_hooks_on(h)) { x264_frame_expand_border_c(h, frame, mb_y, b_end); return; }
    if (b_end) frame_end(h, frame);
}
void x264_frame_filter(x264_t *h, x264_frame_t *frame, int mb_y, int b_end)
{
    if (!b200_on(h) || frame != h->fdec || !frame_hooks_on(h)) x264_frame_filter_c(h, frame, mb_y, b_end);
}
void x264_frame_expand_border_filtered(x264_t *h, x264_frame_t *frame, int mb_y, int b_end)
{
    if (!b200_on(h) || frame != h->fdec || !frame_hooks_on(h)) x264_frame_expand_border_filtered_c(h, frame, mb_y, b_end);
}

/* PSNR / SSIM statistics (encoder.c:1034-1056) are taken per slab of rows right after that slab was deblocked — which now happens at the
 * end of the frame.  The slabs before the last one therefore contribute nothing when they are asked for, and the last call (made after
 * frame_end) evaluates the whole picture: SSD is an integer sum over pixels, so one whole-plane evaluation equals the sum of the slabs;
 * SSIM is a float sum, so the recorded slabs are evaluated one by one and accumulated in their original order. */
static int stats_deferred(void)
{
    return B.state > 0 && B.frame_on && g_h && g_h->fdec->b_kept_as_ref && !g_h->sh.b_mbaff && !g_h->sh.i_disable_deblocking_filter_idc;
}
static int locate_plane(x264_t *h, uint8_t *pix_dec, int stride, int *plane, int *y0)
{
    for (int i = 0; i < 3; i++) {
        const ptrdiff_t off = pix_dec - h->fdec->plane[i];
        if (stride == h->fdec->i_stride[i] && off >= 0 && off < (ptrdiff_t)stride * h->fdec->i_lines[i]) { *plane = i; *y0 = (int)(off / stride); return 1; }
    }
    return 0;
}
int64_t x264_pixel_ssd_wxh(x264_pixel_function_t *pf, uint8_t *pix1, int i_pix1, uint8_t *pix2, int i_pix2, int i_width, int i_height)
{
    int plane, y0;
    if (!stats_deferred() || !locate_plane(g_h, pix1, i_pix1, &plane, &y0)) return x264_pixel_ssd_wxh_c(pf, pix1, i_pix1, pix2, i_pix2, i_width, i_height);
    x264_t *h = g_h;
    if (!(B.end_done == h->fdec && B.end_done_frame == h->fdec->i_frame)) return 0;
    return x264_pixel_ssd_wxh_c(pf, h->fdec->plane[plane], i_pix1, h->fenc->plane[plane], i_pix2, i_width, y0 + i_height);
}
float x264_pixel_ssim_wxh(x264_pixel_function_t *pf, uint8_t *pix1, int i_pix1, uint8_t *pix2, int i_pix2, int i_width, int i_height, void *buf)
{
    int plane, y0;
    if (!stats_deferred() || !locate_plane(g_h, pix1, i_pix1, &plane, &y0) || plane) return x264_pixel_ssim_wxh_c(pf, pix1, i_pix1, pix2, i_pix2, i_width, i_height, buf);
    x264_t *h = g_h;
    if (B.n_ssim_slab < 256) { B.ssim_slab[B.n_ssim_slab].y0 = y0; B.ssim_slab[B.n_ssim_slab].h = i_height; B.n_ssim_slab++; }
    if (!(B.end_done == h->fdec && B.end_done_frame == h->fdec->i_frame)) return 0.f;
    /* the caller accumulates the returned floats in h->stat.frame.f_ssim: add the earlier slabs there in their original order and
     * return the last one, so that the sum is formed exactly as the reference forms it */
    const int x0 = (int)((pix1 - h->fdec->plane[0]) % i_pix1);
    float v = 0.f;
    for (int i = 0; i < B.n_ssim_slab; i++) {
        v = x264_pixel_ssim_wxh_c(pf, h->fdec->plane[0] + x0 + B.ssim_slab[i].y0 * i_pix1, i_pix1, h->fenc->plane[0] + x0 + B.ssim_slab[i].y0 * i_pix2, i_pix2, i_width,
                                  B.ssim_slab[i].h, buf);
        if (i < B.n_ssim_slab - 1) h->stat.frame.f_ssim += v;
    }
    B.n_ssim_slab = 0;
    return v;
}

/* ------------------------------------------------------------------------------------------------------------------------------
 * candidate grids of a (source, reference) pair */

/* h->mb.mv_{min,max}_fpel of macroblock (x, y) as x264_mb_analyse_init sets them for --threads 1, progressive (S/encoder/analyse.c:259-305) */
static void fpel_limits(x264_t *h, int x, int y, int16_t lo[2], int16_t hi[2])
{
    const int fmv = 4 * h->param.analyse.i_mv_range;
    const int min_x = x264_clip3(4 * (-16 * x - 24), -fmv, fmv - 1), max_x = x264_clip3(4 * (16 * (B.mb_w - x - 1) + 24), -fmv, fmv - 1);
    const int min_y = x264_clip3(4 * (-16 * y - 24), X264_MAX(4 * (-512 + 8), -fmv), fmv);
    const int max_y = X264_MIN(x264_clip3(4 * (16 * (B.mb_h - y - 1) + 24), -fmv, fmv - 1), fmv * 4);
    lo[0] = (min_x >> 2) + 5; hi[0] = (max_x >> 2) - 5;
    lo[1] = (min_y >> 2) + 5; hi[1] = (max_y >> 2) - 5;
}

/* Where will the searches of macroblock (x, y) be centred?  Only a guess is needed (a wrong one costs a one-macroblock relaunch):
 * the lookahead's vector when the slice-type decision ran one (the same vector x264_mb_predict_mv_ref16x16 offers, S/common/macroblock.c:
 * 393-398), else the co-located vector of the reference picture (constant motion), else the reference picture's median vector. */
static void guess_centres(x264_t *h, x264_frame_t *ref, int list, gridset_t *g)
{
    const int n_mb = B.mb_w * B.mb_h;
    int16_t (*lowres)[2] = NULL;
    if (h->frames.b_have_lowres) {
        const int dist = list ? ref->i_frame - h->fenc->i_frame : h->fenc->i_frame - ref->i_frame; /* display-order distance */
        if (dist >= 1 && dist <= h->param.i_bframe + 1 && h->fenc->lowres_mvs[list][dist - 1][0][0] != 0x7FFF) lowres = h->fenc->lowres_mvs[list][dist - 1];
    }
    /* median vector of the reference's own inter macroblocks, for macroblocks without a better hint */
    int gx = 0, gy = 0;
    if (!lowres && ref->mv[0] && ref->ref[0] && ref->i_type != X264_TYPE_I && ref->i_type != X264_TYPE_IDR) {
        int hist_x[129] = { 0 }, hist_y[129] = { 0 }, n = 0;
        for (int y = 0; y < B.mb_h; y++)
            for (int x = 0; x < B.mb_w; x++)
                if (ref->ref[0][(2 * y) * 2 * B.mb_w + 2 * x] >= 0) {
                    const int16_t *mv = ref->mv[0][(4 * y) * 4 * B.mb_w + 4 * x];
                    hist_x[x264_clip3((mv[0] + 2) >> 2, -64, 64) + 64]++; hist_y[x264_clip3((mv[1] + 2) >> 2, -64, 64) + 64]++; n++;
                }
        if (n) {
            int a = 0, i;
            for (i = 0; i < 129 && (a += hist_x[i]) * 2 < n; i++);
            gx = i - 64;
            for (a = 0, i = 0; i < 129 && (a += hist_y[i]) * 2 < n; i++);
            gy = i - 64;
        }
        if (list) { gx = -gx; gy = -gy; } /* a later picture: the motion towards it runs the other way */
    }
    for (int y = 0, i = 0; y < B.mb_h; y++)
        for (int x = 0; x < B.mb_w; x++, i++) {
            x264_cuda_grid_job_t *j = &g->jobs[i];
            int cx = gx, cy = gy;
            if (lowres) { cx = (2 * lowres[i][0] + 2) >> 2; cy = (2 * lowres[i][1] + 2) >> 2; }
            else if (!list && ref->mv[0] && ref->ref[0] && ref->i_type != X264_TYPE_I && ref->i_type != X264_TYPE_IDR && ref->ref[0][(2 * y) * 2 * B.mb_w + 2 * x] >= 0) {
                const int16_t *mv = ref->mv[0][(4 * y) * 4 * B.mb_w + 4 * x];
                cx = (mv[0] + 2) >> 2; cy = (mv[1] + 2) >> 2;
            }
            j->mb_x = x; j->mb_y = y; j->part_mask = 0x1ff; j->reserved = 0;
            fpel_limits(h, x, y, j->mv_min_fpel, j->mv_max_fpel);
            j->cx = x264_clip3(cx, j->mv_min_fpel[0], j->mv_max_fpel[0]);
            j->cy = x264_clip3(cy, j->mv_min_fpel[1], j->mv_max_fpel[1]);
        }
    (void)n_mb;
}

/* Launch chunk c of a grid set.  src_row >= 0: macroblock row src_row of the CURRENT frame is finished — where its macroblocks point into
 * this reference picture, their vectors replace the guessed centres of the macroblocks below them (h->mb.mv / h->mb.ref are the frame-wide
 * arrays x264_macroblock_cache_save fills, S/common/macroblock.c:1219-1372). */
static void issue_chunk(x264_t *h, gridset_t *g, int c, int src_row)
{
    const int n_mb = B.mb_w * B.mb_h, mb0 = c * g->rows_per_chunk * B.mb_w, mb1 = X264_MIN(n_mb, mb0 + g->rows_per_chunk * B.mb_w);
    if (src_row >= 0 && h->mb.mv[g->list] && h->mb.ref[g->list]) {
        x264_frame_t **fref = g->list ? h->fref1 : h->fref0;
        const int n_ref = g->list ? h->i_ref1 : h->i_ref0;
        for (int i = mb0; i < mb1; i++) {
            const int x = i % B.mb_w;
            const int r = h->mb.ref[g->list][(2 * src_row + 1) * 2 * B.mb_w + 2 * x];   /* bottom-left 8x8 block of the macroblock above */
            if (r < 0 || r >= n_ref || fref[r] != g->ref) continue;
            const int16_t *mv = h->mb.mv[g->list][(4 * src_row + 3) * 4 * B.mb_w + 4 * x];
            x264_cuda_grid_job_t *j = &g->jobs[i];
            j->cx = x264_clip3((mv[0] + 2) >> 2, j->mv_min_fpel[0], j->mv_max_fpel[0]);
            j->cy = x264_clip3((mv[1] + 2) >> 2, j->mv_min_fpel[1], j->mv_max_fpel[1]);
        }
    }
    if (c >= N_RING && g->fence[c - N_RING]) { CK(x264_cuda_fence_wait(B.ctx, g->fence[c - N_RING])); g->fence[c - N_RING] = NULL; } /* the slot's previous occupant */
    CK(x264_cuda_sad_grid_quad(B.ctx, source_on_device(h), slot_for_ref(g->ref)->d, B.radius, g->jobs + mb0, mb1 - mb0,
                               (uint16_t *)((uint8_t *)g->grid + X264_CUDA_GRID_QUAD_BYTES(B.radius) * (size_t)(c % N_RING) * g->rows_per_chunk * B.mb_w), 1));
    if (!(g->fence[c] = x264_cuda_fence_record(B.ctx))) die("x264_cuda_fence_record");
    g->issued[c] = 1;
}

static gridset_t *gridset_for(x264_t *h, x264_frame_t *ref, int list)
{
    gridset_t *v = &B.gs[0];
    for (int i = 0; i < N_GRIDSETS; i++) {
        gridset_t *g = &B.gs[i];
        if (g->enc_frame == h->fenc->i_frame && g->enc_type == h->sh.i_type && g->ref == ref && g->ref_frame == ref->i_frame && g->ref_poc == ref->i_poc) {
            g->used = ++B.clock;
            return g;
        }
        if (g->used < v->used) v = g;
    }
    const double t0 = now_ms();
    gridset_t *g = v;
    const int n_mb = B.mb_w * B.mb_h, R = B.radius;
    const size_t per_mb = X264_CUDA_GRID_QUAD_BYTES(R);
    for (int c = 0; c < g->n_chunks; c++) if (g->fence[c]) { CK(x264_cuda_fence_wait(B.ctx, g->fence[c])); g->fence[c] = NULL; }
    const int rows_per_chunk = x264_clip3(B.chunk_mbs / B.mb_w, 1, 16);
    B.t_pin_pending = 0;
    if (!g->grid) {
        const double ta = now_ms();
        CK(x264_cuda_grid_ring_reserve(B.ctx, (size_t)(2.25 * per_mb * n_mb) + (1 << 20))); /* two frames' worth of grids + job copies */
        /* the macroblock loop is strictly raster order, so only a short ring of row chunks has to be resident: ~15 MB per grid set at
         * 1080p instead of the whole frame's 166 MB (+ room for the last 32-byte load of a window's last row) */
        g->grid = x264_cuda_host_alloc(per_mb * rows_per_chunk * B.mb_w * N_RING + 64);
        g->jobs = x264_cuda_host_alloc(sizeof(x264_cuda_grid_job_t) * n_mb);
        g->extra_grid = x264_cuda_host_alloc(per_mb * N_EXTRA + 64);
        g->extra_job = x264_cuda_host_alloc(sizeof(x264_cuda_grid_job_t) * N_EXTRA);
        if (!g->grid || !g->jobs || !g->extra_grid || !g->extra_job) { fprintf(stderr, "x264_b200: cannot page-lock %zu MB for the candidate grids\n", (per_mb * rows_per_chunk * B.mb_w * N_RING) >> 20); exit(3); }
        B.t_alloc += now_ms() - ta; B.t_alloc_in_issue += now_ms() - ta;
    }
    g->ref = ref; g->ref_frame = ref->i_frame; g->ref_poc = ref->i_poc; g->enc_frame = h->fenc->i_frame; g->enc_type = h->sh.i_type; g->used = ++B.clock;
    guess_centres(h, ref, list, g);
    /* A few macroblock rows per launch, each followed by its own copy back and fence.  Only the first two chunks are issued now; chunk
     * c + 1 goes out when the host starts on chunk c (gridset_wait_row), with its centres refreshed from the vectors the rows above have
     * just been given — the device computes and copies it while the host encodes chunk c. */
    g->list = list;
    g->rows_per_chunk = rows_per_chunk;
    g->n_chunks = (B.mb_h + g->rows_per_chunk - 1) / g->rows_per_chunk;
    if (g->n_chunks > MAX_CHUNKS) { fprintf(stderr, "x264_b200: picture too tall (%d macroblock rows)\n", B.mb_h); exit(3); }
    memset(g->issued, 0, sizeof(g->issued));
    for (int k = 0; k < N_EXTRA; k++) { g->extra_job[k].mb_x = -1; g->extra_used[k] = 0; }
    issue_chunk(h, g, 0, -1);
    if (g->n_chunks > 1) issue_chunk(h, g, 1, -1);
    B.n_gridsets++;
    B.t_grid_issue += now_ms() - t0;
    B.t_alloc_in_issue += B.t_pin_pending; B.t_pin_pending = 0;
    return g;
}
/* the host is about to search macroblock row mb_y: its chunk must have arrived; the next chunk is sent on its way */
static inline void gridset_wait_row(x264_t *h, gridset_t *g, int mb_y)
{
    const int c = mb_y / g->rows_per_chunk;
    if (!g->issued[c]) issue_chunk(h, g, c, mb_y - 1);
    if (g->fence[c]) {
        const double t0 = now_ms();
        for (int k = 0; k <= c; k++) if (g->fence[k]) { CK(x264_cuda_fence_wait(B.ctx, g->fence[k])); g->fence[k] = NULL; }
        B.t_grid_wait += now_ms() - t0;
    }
    if (c + 1 < g->n_chunks && !g->issued[c + 1]) {
        const double t0 = now_ms();
        issue_chunk(h, g, c + 1, mb_y - 1);
        B.t_grid_issue += now_ms() - t0;
    }
}

/* ------------------------------------------------------------------------------------------------------------------------------
 * x264_me_search_ref for --me esa */

/* which reference frame (and block position) does m->p_fref[0] point into?  NULL for the lookahead's half-resolution planes */
static x264_frame_t *find_ref(x264_t *h, const x264_me_t *m, int *bx, int *by, int *list)
{
    for (int l = 0; l < 2; l++)
        for (int i = 0; i < (l ? h->i_ref1 : h->i_ref0); i++) {
            x264_frame_t *f = l ? h->fref1[i] : h->fref0[i];
            const ptrdiff_t off = m->p_fref[0] - f->plane[0];
            if (off >= 0 && off < (ptrdiff_t)f->i_stride[0] * f->i_lines[0] && m->i_stride[0] == f->i_stride[0]) {
                *bx = (int)(off % f->i_stride[0]); *by = (int)(off / f->i_stride[0]); *list = l;
                return f;
            }
        }
    return NULL;
}

typedef struct {
    const uint64_t *quad; /* the macroblock's grid: [GH][GW] positions of four uint16 */
    uint64_t mask;        /* which quadrants make up this partition */
    int gx0, gy0, gw, gh; /* window covered */
} grid_view_t;
#define QUAD_SUM(v, mask) ((int)((((v) & (mask)) * 0x0001000100010001ULL) >> 48)) /* sum of the selected 16-bit lanes: at most 4 * 64 * 255 */

static inline int view_has(const grid_view_t *v, int mx, int my) { return (unsigned)(mx - v->gx0) < (unsigned)v->gw && (unsigned)(my - v->gy0) < (unsigned)v->gh; }
static void view_of(grid_view_t *v, const gridset_t *g, int mb_xy, uint64_t mask)
{
    const int R = B.radius;
    v->gw = X264_CUDA_GRID_W(R); v->gh = X264_CUDA_GRID_H(R);
    const int per_chunk = g->rows_per_chunk * B.mb_w, c = mb_xy / per_chunk;
    v->quad = (const uint64_t *)((const uint8_t *)g->grid + X264_CUDA_GRID_QUAD_BYTES(R) * ((size_t)(c % N_RING) * per_chunk + (mb_xy - c * per_chunk)));
    v->gx0 = g->jobs[mb_xy].cx - R; v->gy0 = g->jobs[mb_xy].cy - R;
    v->mask = mask;
}

/* Raster argmin of SAD + x cost + y cost over rows [min_y, max_y] x columns [min_x, min_x + width) of a quadrant grid, strict '<', seeded
 * with (*bcost, *bmx, *bmy) — the loop of me.c:580-598 without the SAD calls.  xc[i] = vector cost of column min_x + i, padded with a huge
 * value up to a multiple of 8.  Eight positions per step: the 16-bit quadrant sums are selected by the partition mask, added pairwise
 * (pmaddwd with ones) and across (phaddd); only when some lane beats the best cost so far are the eight positions revisited in raster
 * order by the scalar code, so ties resolve exactly as in the reference. */
#include <immintrin.h>
__attribute__((target("avx2"))) static void scan_window_avx2(const uint64_t *quad, int gw, int gx0, int gy0, uint64_t mask, const int *xc, const int16_t *cost_y,
                                                             int min_x, int width, int min_y, int max_y, int *pbcost, int *pbmx, int *pbmy)
{
    int bcost = *pbcost, bmx = *pbmx, bmy = *pbmy;
    const __m256i vm = _mm256_set1_epi64x((long long)mask), ones = _mm256_set1_epi16(1);
    int xcp[48]; /* x costs in the lane order phaddd leaves the sums in: p0 p1 p4 p5 | p2 p3 p6 p7 */
    static const int order[8] = { 0, 1, 4, 5, 2, 3, 6, 7 };
    for (int i = 0; i < width; i += 8)
        for (int k = 0; k < 8; k++) xcp[i + k] = xc[i + order[k]];
    for (int my = min_y; my <= max_y; my++) {
        const int ycost = cost_y[my << 2];
        if (bcost <= ycost) continue;
        const uint64_t *row = quad + (my - gy0) * gw + (min_x - gx0);
        __m256i vb = _mm256_set1_epi32(bcost - ycost);
        for (int i = 0; i < width; i += 8) {
            const __m256i a = _mm256_and_si256(_mm256_loadu_si256((const __m256i *)(row + i)), vm);
            const __m256i b = _mm256_and_si256(_mm256_loadu_si256((const __m256i *)(row + i + 4)), vm);
            const __m256i sums = _mm256_hadd_epi32(_mm256_madd_epi16(a, ones), _mm256_madd_epi16(b, ones));
            const __m256i c = _mm256_add_epi32(sums, _mm256_loadu_si256((const __m256i *)(xcp + i)));
            if (_mm256_movemask_epi8(_mm256_cmpgt_epi32(vb, c))) {
                for (int k = 0; k < 8 && i + k < width; k++) {
                    const int cc = QUAD_SUM(row[i + k], mask) + xc[i + k] + ycost;
                    if (cc < bcost) { bcost = cc; bmx = min_x + i + k; bmy = my; }
                }
                vb = _mm256_set1_epi32(bcost - ycost);
            }
        }
    }
    *pbcost = bcost; *pbmx = bmx; *pbmy = bmy;
}

/* the sub-pel stage of a search: refine_subpel( h, m, hpel, qpel, p_halfpel_thresh, 0 ) of S/encoder/me.c:680-778, through the
 * reference's own function tables */
static void subpel_stage(x264_t *h, x264_me_t *m, int hpel_iters, int qpel_iters, int *p_halfpel_thresh)
{
    const int i_pixel = m->i_pixel, bw = x264_pixel_size[i_pixel].w, bh = x264_pixel_size[i_pixel].h, stride_ref = m->i_stride[0];
    const int16_t *cost_x = m->p_cost_mv - m->mvp[0], *cost_y = m->p_cost_mv - m->mvp[1];
    const int chroma_me = h->mb.b_chroma_me && i_pixel <= PIXEL_8x8;
    DECLARE_ALIGNED_16(uint8_t pix[2][32 * 18]);
    int bmx = m->mv[0], bmy = m->mv[1], bcost = m->cost, odir = -1, bdir;

    if (hpel_iters && h->mb.i_subpel_refine < 3) { /* the sub-pel part of the predicted vector (me.c:699-705) */
        const int mx = x264_clip3(m->mvp[0], h->mb.mv_min_spel[0], h->mb.mv_max_spel[0]), my = x264_clip3(m->mvp[1], h->mb.mv_min_spel[1], h->mb.mv_max_spel[1]);
        if ((mx - bmx) | (my - bmy)) {
            int st = 16;
            uint8_t *src = h->mc.get_ref(pix[0], &st, m->p_fref, stride_ref, mx, my, bw, bh);
            const int c = h->pixf.fpelcmp[i_pixel](m->p_fenc[0], FENC_STRIDE, src, st) + cost_x[mx] + cost_y[my];
            if (c < bcost) { bcost = c; bmx = mx; bmy = my; }
        }
    }
    for (int it = hpel_iters; it > 0; it--) { /* half-pel diamond (me.c:708-726) */
        const int ox = bmx, oy = bmy;
        int costs[4], st = 32;
        uint8_t *s0 = h->mc.get_ref(pix[0], &st, m->p_fref, stride_ref, ox, oy - 2, bw, bh + 1);
        uint8_t *s2 = h->mc.get_ref(pix[1], &st, m->p_fref, stride_ref, ox - 2, oy, bw + 4, bh);
        h->pixf.fpelcmp_x4[i_pixel](m->p_fenc[0], s0, s0 + st, s2, s2 + 1, st, costs);
        int c;
        c = costs[0] + cost_x[ox] + cost_y[oy - 2]; if (c < bcost) { bcost = c; bmy = oy - 2; }
        c = costs[1] + cost_x[ox] + cost_y[oy + 2]; if (c < bcost) { bcost = c; bmy = oy + 2; }
        c = costs[2] + cost_x[ox - 2] + cost_y[oy]; if (c < bcost) { bcost = c; bmx = ox - 2; bmy = oy; }
        c = costs[3] + cost_x[ox + 2] + cost_y[oy]; if (c < bcost) { bcost = c; bmx = ox + 2; bmy = oy; }
        if (bmx == ox && bmy == oy) break;
    }
#define TRY_MBCMP(mx_, my_, dir_) do { /* COST_MV_SATD with b_refine_qpel = 0 (me.c:654-678) */ \
        const int mx = (mx_), my = (my_), dir = (dir_); \
        if ((dir ^ 1) != odir) { \
            int st = 16; \
            uint8_t *src = h->mc.get_ref(pix[0], &st, m->p_fref, stride_ref, mx, my, bw, bh); \
            int c = h->pixf.mbcmp_unaligned[i_pixel](m->p_fenc[0], FENC_STRIDE, src, st) + cost_x[mx] + cost_y[my]; \
            if (chroma_me && c < bcost) { \
                h->mc.mc_chroma(pix[0], 8, m->p_fref[4], m->i_stride[1], mx, my, bw / 2, bh / 2); \
                c += h->pixf.mbcmp[i_pixel + 3](m->p_fenc[1], FENC_STRIDE, pix[0], 8); \
                if (c < bcost) { \
                    h->mc.mc_chroma(pix[0], 8, m->p_fref[5], m->i_stride[1], mx, my, bw / 2, bh / 2); \
                    c += h->pixf.mbcmp[i_pixel + 3](m->p_fenc[2], FENC_STRIDE, pix[0], 8); \
                } \
            } \
            if (c < bcost) { bcost = c; bmx = mx; bmy = my; bdir = dir; } \
        } } while (0)
    if (bmy > h->mb.mv_max_spel[1]) bmy = h->mb.mv_max_spel[1]; /* me.c:730-735 */
    bcost = COST_MAX;
    TRY_MBCMP(bmx, bmy, -1);
    if (p_halfpel_thresh) { /* early termination over several reference frames (me.c:738-750) */
        if ((bcost * 7) >> 3 > *p_halfpel_thresh) { m->cost = bcost; m->mv[0] = bmx; m->mv[1] = bmy; return; }
        if (bcost < *p_halfpel_thresh) *p_halfpel_thresh = bcost;
    }
    bdir = -1;
    for (int it = qpel_iters; it > 0; it--) { /* quarter-pel diamond (me.c:753-765) */
        const int ox = bmx, oy = bmy;
        odir = bdir;
        TRY_MBCMP(ox, oy - 1, 0);
        TRY_MBCMP(ox, oy + 1, 1);
        TRY_MBCMP(ox - 1, oy, 2);
        TRY_MBCMP(ox + 1, oy, 3);
        if (bmx == ox && bmy == oy) break;
    }
    if (bmy > h->mb.mv_max_spel[1]) { /* me.c:768-773 */
        bmy = h->mb.mv_max_spel[1];
        bcost = COST_MAX;
        TRY_MBCMP(bmx, bmy, -1);
    }
#undef TRY_MBCMP
    m->cost = bcost; m->mv[0] = bmx; m->mv[1] = bmy;
    m->cost_mv = cost_x[bmx] + cost_y[bmy];
}

static const int subpel_search_iters[10][2] = { {0,0}, {0,0}, {1,0}, {1,0}, {1,1}, {1,2}, {2,2}, {2,2}, {4,10}, {4,10} }; /* me.c:34-44, columns 2 and 3 */

void x264_me_search_ref(x264_t *h, x264_me_t *m, int16_t (*mvc)[2], int i_mvc, int *p_halfpel_thresh)
{
    int bx, by, list;
    x264_frame_t *ref;
    g_h = h;
    /* --subme 0 is left alone: the reference never builds the integral image then (x264_frame_filter is only called when
     * i_subpel_refine != 0, encoder.c:1019), so its own ESA prunes with ADS values of uninitialised memory — not a defined result */
    if (h->mb.i_me_method != X264_ME_ESA || !b200_on(h) || !B.me_on || h->sh.b_mbaff || !h->param.analyse.i_subpel_refine || !(ref = find_ref(h, m, &bx, &by, &list))) {
        if (B.state > 0 && h->mb.i_me_method >= X264_ME_ESA) B.n_left_to_c++;
        x264_me_search_ref_c(h, m, mvc, i_mvc, p_halfpel_thresh);
        return;
    }
    const x264_me_t m_in = *m;
    const int thr_in = p_halfpel_thresh ? *p_halfpel_thresh : 0;
    const int i_pixel = m->i_pixel, bw = x264_pixel_size[i_pixel].w, bh = x264_pixel_size[i_pixel].h, stride = m->i_stride[0];
    const int range = h->param.analyse.i_me_range, subme = h->mb.i_subpel_refine;
    const int x_min = h->mb.mv_min_fpel[0], y_min = h->mb.mv_min_fpel[1], x_max = h->mb.mv_max_fpel[0], y_max = h->mb.mv_max_fpel[1];
    const int16_t *cost_x = m->p_cost_mv - m->mvp[0], *cost_y = m->p_cost_mv - m->mvp[1];
    uint8_t *p_fref = m->p_fref[0];
    const int mb_xy = (by >> 4) * B.mb_w + (bx >> 4), ox = bx & 15, oy = by & 15;
    const int use_grid = i_pixel <= PIXEL_8x8 && range + 2 <= B.radius;
    gridset_t *g = NULL;
    grid_view_t gv;
    memset(&gv, 0, sizeof(gv));
    if (use_grid) {
        static const uint64_t Q[4] = { 0xffffULL, 0xffffULL << 16, 0xffffULL << 32, 0xffffULL << 48 };
        const uint64_t mask = i_pixel == PIXEL_16x16 ? ~0ULL : i_pixel == PIXEL_16x8 ? (oy ? Q[2] | Q[3] : Q[0] | Q[1])
                            : i_pixel == PIXEL_8x16 ? (ox ? Q[1] | Q[3] : Q[0] | Q[2]) : Q[(oy >> 3) * 2 + (ox >> 3)];
        g = gridset_for(h, ref, list);
        gridset_wait_row(h, g, by >> 4);
        view_of(&gv, g, mb_xy, mask);
    }
    /* SAD of the block at integer vector (mx, my): from the grid, or — for a predictor candidate the grid does not cover — the table entry */
#define SAD_AT(mx, my, out) do { \
        if (view_has(&gv, (mx), (my))) (out) = QUAD_SUM(gv.quad[((my) - gv.gy0) * gv.gw + ((mx) - gv.gx0)], gv.mask); \
        else { (out) = h->pixf.fpelcmp[i_pixel](m->p_fenc[0], FENC_STRIDE, &p_fref[(my) * stride + (mx)], stride); B.n_outside_pred += use_grid; } } while (0)
#define TRY_FPEL(mx_, my_) do { const int tx = (mx_), ty = (my_); int sad_; SAD_AT(tx, ty, sad_); \
        const int c = sad_ + cost_x[tx << 2] + cost_y[ty << 2]; if (c < bcost) { bcost = c; bmx = tx; bmy = ty; } } while (0)

    /* predictor stage (me.c:182-229) */
    int bmx = x264_clip3(m->mvp[0], x_min * 4, x_max * 4), bmy = x264_clip3(m->mvp[1], y_min * 4, y_max * 4);
    const int pmx = (bmx + 2) >> 2, pmy = (bmy + 2) >> 2;
    int bcost = COST_MAX, bpred_mx = 0, bpred_my = 0, bpred_cost = COST_MAX;
    if (subme >= 3) {
        DECLARE_ALIGNED_16(uint8_t pix[16 * 16]);
        const uint32_t bmv = pack16to32_mask(bmx, bmy);
#define TRY_QPEL_PRED(mx_, my_) do { const int qx = (mx_), qy = (my_); int st = 16; \
            uint8_t *src = h->mc.get_ref(pix, &st, m->p_fref, stride, qx, qy, bw, bh); \
            const int c = h->pixf.fpelcmp[i_pixel](m->p_fenc[0], FENC_STRIDE, src, st) + cost_x[qx] + cost_y[qy]; \
            if (c < bpred_cost) { bpred_cost = c; bpred_mx = qx; bpred_my = qy; } } while (0)
        TRY_QPEL_PRED(bmx, bmy);
        for (int i = 0; i < i_mvc; i++) {
            const uint32_t v = *(uint32_t *)mvc[i];
            if (v && (bmv - v)) TRY_QPEL_PRED(x264_clip3(mvc[i][0], x_min * 4, x_max * 4), x264_clip3(mvc[i][1], y_min * 4, y_max * 4));
        }
#undef TRY_QPEL_PRED
        bmx = (bpred_mx + 2) >> 2; bmy = (bpred_my + 2) >> 2;
        const int sx = bmx, sy = bmy;
        TRY_FPEL(sx, sy);
    } else {
        bmx = pmx; bmy = pmy;
        TRY_FPEL(pmx, pmy);
        bcost -= cost_x[pmx << 2] + cost_y[pmy << 2]; /* the rounded prediction carries no vector cost (me.c:209-216) */
        for (int i = 0; i < i_mvc; i++) {
            const int mx = (mvc[i][0] + 2) >> 2, my = (mvc[i][1] + 2) >> 2;
            if ((mx | my) && ((mx - bmx) | (my - bmy))) TRY_FPEL(x264_clip3(mx, x_min, x_max), x264_clip3(my, y_min, y_max));
        }
    }
    TRY_FPEL(0, 0);

    /* exhaustive stage (me.c:449-492, :580-598): strict-'<' argmin in raster order over the window, seeded with the predictor stage */
    const int min_x = X264_MAX(bmx - range, x_min), min_y = X264_MAX(bmy - range, y_min);
    const int max_x = X264_MIN(bmx + range, x_max), max_y = X264_MIN(bmy + range, y_max);
    const int width = (max_x - min_x + 3) & ~3;
    if (use_grid) {
        if (min_x < gv.gx0 || min_x + width > gv.gx0 + gv.gw || min_y < gv.gy0 || max_y >= gv.gy0 + gv.gh) {
            /* the guessed centre was off.  A grid recomputed for this macroblock earlier may cover the window; else recompute one around
             * the exact centre (one-job launch) into the least recently used spare slot */
            const int R = B.radius;
            int slot = -1;
            for (int k = 0; k < N_EXTRA && slot < 0; k++) {
                const x264_cuda_grid_job_t *e = &g->extra_job[k];
                if (e->mb_x == (bx >> 4) && e->mb_y == (by >> 4) && min_x >= e->cx - R && min_x + width <= e->cx - R + gv.gw && min_y >= e->cy - R && max_y <= e->cy + R) slot = k;
            }
            if (slot < 0) {
                const double t0 = now_ms();
                slot = 0;
                for (int k = 1; k < N_EXTRA; k++) if (g->extra_used[k] < g->extra_used[slot]) slot = k;
                x264_cuda_grid_job_t *e = &g->extra_job[slot];
                *e = g->jobs[mb_xy];
                e->cx = bmx; e->cy = bmy;
                if (B.verbose > 1) fprintf(stderr, "x264_b200: miss frame %d mb (%d,%d) pixel %d off (%d,%d): centre (%d,%d), needed (%d,%d), mvp (%d,%d)\n", h->fenc->i_frame, bx >> 4,
                                           by >> 4, i_pixel, ox, oy, g->jobs[mb_xy].cx, g->jobs[mb_xy].cy, bmx, bmy, m->mvp[0], m->mvp[1]);
                CK(x264_cuda_sad_grid_quad_direct(B.ctx, source_on_device(h), slot_for_ref(ref)->d, R, e, 1, (uint16_t *)((uint8_t *)g->extra_grid + X264_CUDA_GRID_QUAD_BYTES(R) * slot)));
                B.n_relaunch++;
                B.t_relaunch += now_ms() - t0;
            } else
                B.n_extra_hit++;
            g->extra_used[slot] = ++B.clock;
            gv.quad = (const uint64_t *)((const uint8_t *)g->extra_grid + X264_CUDA_GRID_QUAD_BYTES(R) * slot);
            gv.gx0 = g->extra_job[slot].cx - R; gv.gy0 = g->extra_job[slot].cy - R;
        }

        int xc[136 + 8];
        for (int i = 0; i < width; i++) xc[i] = cost_x[(min_x + i) << 2];
        for (int i = width; i < ((width + 7) & ~7); i++) xc[i] = COST_MAX;
        if (B.avx2)
            scan_window_avx2(gv.quad, gv.gw, gv.gx0, gv.gy0, gv.mask, xc, cost_y, min_x, width, min_y, max_y, &bcost, &bmx, &bmy);
        else
            for (int my = min_y; my <= max_y; my++) {
                const int ycost = cost_y[my << 2];
                if (bcost <= ycost) continue; /* nothing in this row can be strictly better (the reference skips it too, me.c:585-587) */
                const uint64_t *row = gv.quad + (my - gv.gy0) * gv.gw + (min_x - gv.gx0);
                const uint64_t mask = gv.mask;
                for (int i = 0; i < width; i++) {
                    const int c = QUAD_SUM(row[i], mask) + xc[i] + ycost;
                    if (c < bcost) { bcost = c; bmx = min_x + i; bmy = my; }
                }
            }
        B.n_search++;
    } else { /* sub-8x8 partition or a range beyond the grid: the device searches this block alone, seeded with the predictor stage */
        x264_cuda_me_job_t j;
        x264_cuda_me_result_t r;
        int qp = -1;
        for (int q = 0; q < 52 && qp < 0; q++) if (g_cost_mv[q] && g_cost_mv[q] + 2 * 4 * 2048 == m->p_cost_mv) qp = q;
        if (qp < 0) { fprintf(stderr, "x264_b200: cost table of the search not found\n"); exit(3); }
        if (!B.cost_uploaded[qp]) { CK(x264_cuda_set_cost_mv(B.ctx, qp, g_cost_mv[qp])); B.cost_uploaded[qp] = 1; }
        memset(&j, 0, sizeof(j));
        j.bx = bx; j.by = by; j.i_pixel = i_pixel; j.qp = qp; j.flags = X264_CUDA_ME_SEEDED;
        j.mvp[0] = m->mvp[0]; j.mvp[1] = m->mvp[1];
        j.mv_min_fpel[0] = x_min; j.mv_min_fpel[1] = y_min; j.mv_max_fpel[0] = x_max; j.mv_max_fpel[1] = y_max;
        j.seed_mv[0] = bmx; j.seed_mv[1] = bmy; j.seed_cost = bcost;
        CK(x264_cuda_me_search(B.ctx, source_on_device(h), slot_for_ref(ref)->d, range, &j, 1, &r));
        bmx = r.bmx; bmy = r.bmy; bcost = r.bcost;
        B.n_percall++;
    }
#undef TRY_FPEL
#undef SAD_AT

    /* "-> qpel mv" and the sub-pel stage (me.c:602-630) */
    if (bpred_cost < bcost) { m->mv[0] = bpred_mx; m->mv[1] = bpred_my; m->cost = bpred_cost; }
    else { m->mv[0] = bmx << 2; m->mv[1] = bmy << 2; m->cost = bcost; }
    m->cost_mv = cost_x[m->mv[0]] + cost_y[m->mv[1]];
    if (bmx == pmx && bmy == pmy && subme < 3) m->cost += m->cost_mv;
    if (subme >= 2) subpel_stage(h, m, subpel_search_iters[subme][0], subpel_search_iters[subme][1], p_halfpel_thresh);
    else if (m->mv[1] > h->mb.mv_max_spel[1]) m->mv[1] = h->mb.mv_max_spel[1];
    if (B.check) { /* diagnosis (X264_B200_CHECK=1 with X264_B200_FRAME=0, so that the host integral image exists): compare with the reference's search */
        x264_me_t ref_m = m_in;
        int thr = thr_in;
        x264_me_search_ref_c(h, &ref_m, mvc, i_mvc, p_halfpel_thresh ? &thr : NULL);
        if (ref_m.mv[0] != m->mv[0] || ref_m.mv[1] != m->mv[1] || ref_m.cost != m->cost || (p_halfpel_thresh && thr != *p_halfpel_thresh)) {
            fprintf(stderr, "x264_b200: CHECK: frame %d block (%d,%d) pixel %d subme %d: reference mv (%d,%d) cost %d cost_mv %d, here mv (%d,%d) cost %d cost_mv %d; fpel (%d,%d) %d, "
                    "mvp (%d,%d), pm (%d,%d), limits x %d..%d y %d..%d\n", h->fenc->i_frame, bx, by, i_pixel, subme, ref_m.mv[0], ref_m.mv[1], ref_m.cost, ref_m.cost_mv,
                    m->mv[0], m->mv[1], m->cost, m->cost_mv, bmx, bmy, bcost, m->mvp[0], m->mvp[1], pmx, pmy, x_min, x_max, y_min, y_max);
            exit(5);
        }
    }
}
