/* ref_cuda_shim.c — TEST INFRASTRUCTURE ONLY.  Lets the UNMODIFIED reference encoder run with its four function tables overridden by
 * the CUDA back-end, the way INTEGRATION.md describes, without touching a reference source file: oracle/Makefile compiles
 * S/common/{pixel,mc,dct,quant}.c a second time with -Dx264_<t>_init=x264_<t>_init_c (a rename of the one symbol each file exports for
 * its table), and this file supplies x264_<t>_init: C table first, then the x264_<t>_init_cuda overrides on top
 * (include/x264_cuda_tables.h).  The resulting CLI (oracle/_ref/x264_cuda) must write a byte-identical stream
 * (tests/test_gpu_stream.py): SURVEY 8c "stream level" parity. */
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
    ck(x264_cuda_frame_upload(fctx, ffr, frame->plane[0], frame->i_stride[0], frame->i_width[0], frame->i_lines[0]), "upload");
    ck(x264_cuda_frame_expand_border(fctx, ffr), "expand_border");
    ck(x264_cuda_frame_init_lowres(fctx, ffr), "init_lowres");
    const int s = frame->i_stride_lowres, rows = frame->i_lines_lowres + 2 * PADV, cols = frame->i_width_lowres + 2 * PADH;
    for (int k = 0; k < 4; k++) {
        ck(x264_cuda_frame_download(fctx, ffr, X264_CUDA_PLANE_LOWRES + k, tmp_plane, s), "download lowres");
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
