/* ref_harness.c — TEST INFRASTRUCTURE ONLY.  Serves the xo_* oracle API (oracle/src/xo.h) with the
 * UNMODIFIED reference: it is compiled against the headers under S/ and linked to oracle/_ref/libx264ref.so
 * (the reference's own .c files built in place, see oracle/Makefile).  Nothing here computes pixels: every
 * xo_* function only marshals buffers into the reference's structures and calls the reference's function.
 * Exists only in the build container (S/ is absent on the GPU box; the built .so travels with the snapshot). */
#include <stdlib.h>
#include <string.h>
#include "common/common.h"
#include "encoder/me.h"
#include "src/xo.h"

extern int16_t *g_cost_mv[52];                 /* S/encoder/analyse.c:179 (this fork's exported copy) */
extern uint16_t *x264_cost_mv_fpel[52][4];     /* S/encoder/analyse.c:177 */

const char *xo_backend(void) { return "reference"; }

/* ------------------------------------------------------------------------------------------------ */
/* encoder handles, keyed by what shapes the function tables and frame layout                        */
typedef struct {
    int width, height, me_method, subme_tables, sub8x8, lowres;
    x264_t *h;
    x264_frame_t *fenc;
} hkey;
static hkey g_h[64];
static int g_nh;

static x264_t *get_h(int width, int height, int me_method, int subme_tables, int sub8x8, int lowres, x264_frame_t **fenc)
{
    for (int i = 0; i < g_nh; i++)
        if (g_h[i].width == width && g_h[i].height == height && g_h[i].me_method == me_method &&
            g_h[i].subme_tables == subme_tables && g_h[i].sub8x8 == sub8x8 && g_h[i].lowres == lowres) {
            if (fenc) *fenc = g_h[i].fenc;
            return g_h[i].h;
        }
    x264_param_t p;
    x264_param_default(&p);
    p.i_width = width;
    p.i_height = height;
    p.i_threads = 1;
    p.i_log_level = X264_LOG_NONE;
    p.rc.i_rc_method = X264_RC_CQP;
    p.rc.i_qp_constant = 26;
    p.analyse.i_me_method = me_method;
    p.analyse.i_me_range = 16;
    p.analyse.i_subpel_refine = subme_tables;
    p.analyse.b_psnr = 0;
    p.analyse.b_ssim = 0;
    if (sub8x8) p.analyse.inter |= X264_ANALYSE_PSUB8x8 | X264_ANALYSE_PSUB16x16;
    if (lowres) { p.i_bframe = 3; p.i_bframe_adaptive = X264_B_ADAPT_FAST; }
    x264_t *h = x264_encoder_open(&p);
    if (!h || g_nh == 64) abort();
    hkey *k = &g_h[g_nh++];
    k->width = width; k->height = height; k->me_method = me_method; k->subme_tables = subme_tables;
    k->sub8x8 = sub8x8; k->lowres = lowres; k->h = h;
    k->fenc = x264_frame_new(h);
    h->fenc = k->fenc;
    if (fenc) *fenc = k->fenc;
    return h;
}

/* make the reference build its own MV cost tables for qp: run its real encoder for I+P at --qp qp --me esa */
static void ensure_costs(int qp)
{
    if (g_cost_mv[qp] && x264_cost_mv_fpel[qp][0])
        return;
    x264_param_t p;
    x264_param_default(&p);
    p.i_width = 32; p.i_height = 32; p.i_threads = 1; p.i_log_level = X264_LOG_NONE;
    p.rc.i_rc_method = X264_RC_CQP;
    p.rc.i_qp_constant = qp;
    p.rc.f_ip_factor = 1.0; p.rc.f_pb_factor = 1.0;
    p.analyse.i_me_method = X264_ME_ESA;
    p.analyse.i_subpel_refine = 1;
    p.i_bframe = 0;
    p.i_scenecut_threshold = -1;
    x264_t *h = x264_encoder_open(&p);
    x264_picture_t pic, out;
    x264_nal_t *nal; int nnal;
    x264_picture_alloc(&pic, X264_CSP_I420, 32, 32);
    for (int f = 0; f < 3; f++) {
        for (int i = 0; i < 32 * 32; i++) pic.img.plane[0][i] = (uint8_t)((i * 7 + f * 13) ^ (i >> 5));
        memset(pic.img.plane[1], 128, 16 * 16); memset(pic.img.plane[2], 128, 16 * 16);
        pic.i_type = X264_TYPE_AUTO; pic.i_qpplus1 = 0; pic.i_pts = f;
        x264_encoder_encode(h, &nal, &nnal, &pic, &out);
    }
    x264_picture_clean(&pic);
    /* NOT closing h: this fork frees g_cost_mv[] in x264_encoder_close (S/encoder/encoder.c:2136-2145) */
    if (!g_cost_mv[qp] || !x264_cost_mv_fpel[qp][0]) abort();
}

int xo_lambda(int qp) { extern const int x264_lambda_tab[52]; return x264_lambda_tab[qp]; }

void xo_cost_mv_table(int qp, int16_t *out)
{
    ensure_costs(qp);
    memcpy(out, g_cost_mv[qp], (4 * 4 * 2048 + 1) * sizeof(int16_t));
}

/* ------------------------------------------------------------------------------------------------ */
static x264_pixel_function_t g_pixf;
static x264_mc_functions_t g_mc;
static x264_dct_function_t g_dctf;
static x264_quant_function_t g_quantf;
static int g_tables;
static void tables(void)
{
    if (g_tables) return;
    x264_pixel_init(0, &g_pixf);
    x264_mc_init(0, &g_mc);
    x264_dct_init(0, &g_dctf);
    x264_quant_init(NULL, 0, &g_quantf);
    g_tables = 1;
}

int xo_pixel_cmp(int metric, int i_pixel, const uint8_t *p1, int s1, const uint8_t *p2, int s2)
{
    tables();
    switch (metric) {
    case XO_SAD:  return g_pixf.sad[i_pixel]((uint8_t *)p1, s1, (uint8_t *)p2, s2);
    case XO_SSD:  return g_pixf.ssd[i_pixel]((uint8_t *)p1, s1, (uint8_t *)p2, s2);
    case XO_SATD: return g_pixf.satd[i_pixel]((uint8_t *)p1, s1, (uint8_t *)p2, s2);
    case XO_SA8D: return g_pixf.sa8d[i_pixel]((uint8_t *)p1, s1, (uint8_t *)p2, s2);
    }
    return -1;
}
int xo_pixel_var(int i_pixel, const uint8_t *pix, int stride) { tables(); return g_pixf.var[i_pixel]((uint8_t *)pix, stride); }
uint64_t xo_pixel_hadamard_ac(int i_pixel, const uint8_t *pix, int stride) { tables(); return g_pixf.hadamard_ac[i_pixel]((uint8_t *)pix, stride); }
int xo_pixel_ads(int i_pixel, const int enc_dc[4], const uint16_t *sums, int delta, const uint16_t *cost_mvx,
                 int16_t *mvs, int width, int thresh)
{
    tables();
    return g_pixf.ads[i_pixel]((int *)enc_dc, (uint16_t *)sums, delta, (uint16_t *)cost_mvx, mvs, width, thresh);
}

/* ------------------------------------------------------------------------------------------------ */
void xo_geometry(int width, int height, xo_geom *g)
{
    x264_frame_t *f;
    x264_t *h = get_h(width, height, X264_ME_ESA, 1, 0, 1, &f);
    memset(g, 0, sizeof(*g));
    g->width = width; g->height = height;
    g->mb_width = h->sps->i_mb_width; g->mb_height = h->sps->i_mb_height;
    g->stride = f->i_stride[0]; g->lines = f->i_lines[0];
    g->plane_size = g->stride * (g->lines + 2 * PADV);
    g->origin = g->stride * PADV + PADH;
    g->stride_lowres = f->i_stride_lowres; g->width_lowres = f->i_width_lowres; g->lines_lowres = f->i_lines_lowres;
    g->plane_size_lowres = g->stride_lowres * (g->lines_lowres + 2 * PADV);
    g->origin_lowres = g->stride_lowres * PADV + PADH;
}

/* copy a caller plane (pointer at pixel 0,0, padded all around) into/out of a reference frame plane */
static void put_plane(uint8_t *dst00, const uint8_t *src00, int stride, int lines)
{
    memcpy(dst00 - stride * PADV - PADH, src00 - stride * PADV - PADH, (size_t)stride * (lines + 2 * PADV));
}

void xo_frame_expand_border(const xo_geom *g, uint8_t *plane)
{
    x264_frame_t *f;
    x264_t *h = get_h(g->width, g->height, X264_ME_ESA, 1, 0, 0, &f);
    put_plane(f->plane[0], plane, g->stride, g->lines);
    x264_frame_expand_border_mod16(h, f);
    x264_frame_expand_border(h, f, 0, 1);
    put_plane(plane, f->plane[0], g->stride, g->lines);
}

void xo_frame_filter(const xo_geom *g, const uint8_t *plane, uint8_t *dsth, uint8_t *dstv, uint8_t *dstc,
                     uint16_t *integral, int b_sub8x8)
{
    x264_frame_t *f;
    x264_t *h = get_h(g->width, g->height, X264_ME_ESA, 1, b_sub8x8, 0, &f);
    size_t isz = (size_t)g->stride * (g->lines + 2 * PADV) * sizeof(uint16_t) << b_sub8x8;
    put_plane(f->plane[0], plane, g->stride, g->lines);
    if (integral) memcpy(f->buffer[3], integral - g->stride * PADV - PADH, isz);
    uint16_t *keep = f->integral;
    if (!integral) f->integral = NULL;
    x264_frame_filter(h, f, 0, 1);
    x264_frame_expand_border_filtered(h, f, 0, 1);
    f->integral = keep;
    put_plane(dsth, f->filtered[1], g->stride, g->lines);
    put_plane(dstv, f->filtered[2], g->stride, g->lines);
    put_plane(dstc, f->filtered[3], g->stride, g->lines);
    if (integral) memcpy(integral - g->stride * PADV - PADH, f->buffer[3], isz);
}

void xo_frame_init_lowres(const xo_geom *g, uint8_t *plane, uint8_t *l0, uint8_t *lh, uint8_t *lv, uint8_t *lc)
{
    x264_frame_t *f;
    x264_t *h = get_h(g->width, g->height, X264_ME_ESA, 1, 0, 1, &f);
    put_plane(f->plane[0], plane, g->stride, g->lines);
    x264_frame_init_lowres(h, f);
    put_plane(plane, f->plane[0], g->stride, g->lines);
    uint8_t *o[4] = { l0, lh, lv, lc };
    for (int i = 0; i < 4; i++)
        put_plane(o[i], f->lowres[i], g->stride_lowres, g->lines_lowres);
}

void xo_mc_luma(uint8_t *dst, int dst_stride, const uint8_t *const src[4], int src_stride, int mvx, int mvy, int w, int hgt)
{
    tables();
    g_mc.mc_luma(dst, dst_stride, (uint8_t **)src, src_stride, mvx, mvy, w, hgt);
}
void xo_mc_chroma(uint8_t *dst, int dst_stride, const uint8_t *src, int src_stride, int mvx, int mvy, int w, int hgt)
{
    tables();
    g_mc.mc_chroma(dst, dst_stride, (uint8_t *)src, src_stride, mvx, mvy, w, hgt);
}

void xo_pixel_avg(int i_pixel, uint8_t *dst, int dst_stride, const uint8_t *a, int a_stride, const uint8_t *b, int b_stride, int weight)
{
    tables();
    g_mc.avg[i_pixel](dst, dst_stride, (uint8_t *)a, a_stride, (uint8_t *)b, b_stride, weight);
}

/* ------------------------------------------------------------------------------------------------ */
static const int16_t *g_refine_mv; /* non-NULL: run x264_me_refine_qpel from (g_refine_mv, g_refine_cost) instead of the search */
static int g_refine_cost;
static void run_search_c(const xo_geom *g, const uint8_t *fenc_plane, const uint8_t *const fref[4], const uint16_t *integral,
                         const xo_chroma *ch, const xo_me_in *in, int subme, int mbcmp_satd, xo_me_out *out)
{
    /* the function tables depend on (me method, user subme>1): open the encoder the way the CLI would */
    int tables_subme = in->fpel_satd || mbcmp_satd ? 7 : 1;
    int tables_me = in->fpel_satd ? X264_ME_TESA : in->me_method;
    x264_frame_t *f;
    x264_t *h = get_h(g->width, g->height, tables_me, tables_subme, in->b_sub8x8, 0, &f);
    DECLARE_ALIGNED_16(uint8_t fenc[16 * 16]);
    x264_me_t m;
    DECLARE_ALIGNED_4(int16_t mvc[16][2]);
    ensure_costs(in->qp);
    memset(&m, 0, sizeof(m));
    for (int y = 0; y < x264_pixel_size[in->i_pixel].h; y++)
        memcpy(fenc + 16 * y, fenc_plane + (in->by + y) * g->stride + in->bx, x264_pixel_size[in->i_pixel].w);
    memcpy(mvc, in->mvc, sizeof(mvc));
    h->fenc = f;
    h->param.analyse.i_me_range = in->me_range;
    h->mb.i_me_method = in->me_method;
    h->mb.i_subpel_refine = subme;
    h->mb.b_chroma_me = ch != NULL;
    h->mb.i_qp = in->qp;
    for (int k = 0; k < 2; k++) {
        h->mb.mv_min_fpel[k] = in->mv_min_fpel[k]; h->mb.mv_max_fpel[k] = in->mv_max_fpel[k];
        h->mb.mv_min_spel[k] = in->mv_min_spel[k]; h->mb.mv_max_spel[k] = in->mv_max_spel[k];
    }
    m.i_pixel = in->i_pixel;
    m.p_cost_mv = g_cost_mv[in->qp] + 2 * 4 * 2048;
    m.i_stride[0] = g->stride;
    m.p_fenc[0] = fenc;
    for (int k = 0; k < 4; k++)
        m.p_fref[k] = fref[k] ? (uint8_t *)fref[k] + in->by * g->stride + in->bx : NULL;
    m.integral = integral ? (uint16_t *)integral + in->by * g->stride + in->bx : NULL;
    m.mvp[0] = in->mvp[0]; m.mvp[1] = in->mvp[1];
    DECLARE_ALIGNED_16(uint8_t fenc_c[2][16 * 8]);
    if (ch) { /* fenc chroma at FENC_STRIDE like h->mb.pic.p_fenc[1,2]; reference chroma planes as m->p_fref[4,5] */
        const uint8_t *fe[2] = { ch->fenc_u, ch->fenc_v }, *fr[2] = { ch->fref_u, ch->fref_v };
        for (int pl = 0; pl < 2; pl++) {
            for (int y = 0; y < x264_pixel_size[in->i_pixel].h / 2; y++)
                memcpy(fenc_c[pl] + 16 * y, fe[pl] + (in->by / 2 + y) * ch->stride_c + in->bx / 2, x264_pixel_size[in->i_pixel].w / 2);
            m.p_fenc[1 + pl] = fenc_c[pl];
            m.p_fref[4 + pl] = (uint8_t *)fr[pl] + (in->by / 2) * ch->stride_c + in->bx / 2;
        }
        m.i_stride[1] = ch->stride_c;
    }
    if (g_refine_mv) {
        m.mv[0] = g_refine_mv[0]; m.mv[1] = g_refine_mv[1]; m.cost = g_refine_cost; m.i_ref_cost = 0;
        x264_me_refine_qpel(h, &m);
        m.cost_mv = m.p_cost_mv[m.mv[0] - m.mvp[0]] + m.p_cost_mv[m.mv[1] - m.mvp[1]]; /* refine_subpel sets it too */
    } else
        x264_me_search_ref(h, &m, mvc, in->i_mvc, NULL);
    memset(out, 0, sizeof(*out));
    out->mv[0] = m.mv[0]; out->mv[1] = m.mv[1];
    out->cost = m.cost; out->cost_mv = m.cost_mv;
    /* the reference does not expose its internal full-pel state; report what can be derived */
    out->bmx = out->bmy = out->bcost = out->seed_mx = out->seed_my = out->seed_cost = -1;
}

static void run_search(const xo_geom *g, const uint8_t *fenc_plane, const uint8_t *const fref[4], const uint16_t *integral,
                       const xo_me_in *in, int subme, int mbcmp_satd, xo_me_out *out)
{
    run_search_c(g, fenc_plane, fref, integral, NULL, in, subme, mbcmp_satd, out);
}
void xo_me_search_subpel_chroma(const xo_geom *g, const uint8_t *fenc_plane, const uint8_t *const fref_planes[4], const uint16_t *integral,
                                const xo_chroma *ch, const xo_me_in *in, int subme, int mbcmp_satd, xo_me_out *out)
{
    run_search_c(g, fenc_plane, fref_planes, integral, ch, in, subme, mbcmp_satd, out);
}

void xo_me_refine_qpel(const xo_geom *g, const uint8_t *fenc_plane, const uint8_t *const fref_planes[4], const xo_chroma *ch, const xo_me_in *in,
                       int subme, int mbcmp_satd, const int16_t mv_in[2], int cost_in, xo_me_out *out)
{
    g_refine_mv = mv_in; g_refine_cost = cost_in;
    run_search_c(g, fenc_plane, fref_planes, NULL, ch, in, subme, mbcmp_satd, out);
    g_refine_mv = NULL;
}

void xo_me_search_fpel(const xo_geom *g, const uint8_t *fenc_plane, const uint8_t *fref_plane,
                       const uint16_t *integral, const xo_me_in *in, xo_me_out *out)
{
    const uint8_t *planes[4] = { fref_plane, NULL, NULL, NULL };
    run_search(g, fenc_plane, planes, integral, in, 1, 0, out);
}
void xo_me_search_subpel(const xo_geom *g, const uint8_t *fenc_plane, const uint8_t *const fref_planes[4],
                         const uint16_t *integral, const xo_me_in *in, int subme, int mbcmp_satd, xo_me_out *out)
{
    run_search(g, fenc_plane, fref_planes, integral, in, subme, mbcmp_satd, out);
}
void xo_me_search_fpel_batch(const xo_geom *g, const uint8_t *fenc_plane, const uint8_t *fref_plane,
                             const uint16_t *integral, const xo_me_in *in, int n, xo_me_out *out)
{
    for (int i = 0; i < n; i++)
        xo_me_search_fpel(g, fenc_plane, fref_plane, integral, in + i, out + i);
}

/* ------------------------------------------------------------------------------------------------ */
void xo_sub4x4_dct(int16_t dct[16], const uint8_t *p1, const uint8_t *p2) { tables(); g_dctf.sub4x4_dct((void *)dct, (uint8_t *)p1, (uint8_t *)p2); }
void xo_add4x4_idct(uint8_t *dst, int16_t dct[16]) { tables(); g_dctf.add4x4_idct(dst, (void *)dct); }
void xo_sub8x8_dct8(int16_t dct[64], const uint8_t *p1, const uint8_t *p2) { tables(); g_dctf.sub8x8_dct8((void *)dct, (uint8_t *)p1, (uint8_t *)p2); }
void xo_add8x8_idct8(uint8_t *dst, int16_t dct[64]) { tables(); g_dctf.add8x8_idct8(dst, (void *)dct); }
void xo_dct4x4dc(int16_t d[16]) { tables(); g_dctf.dct4x4dc((void *)d); }
void xo_idct4x4dc(int16_t d[16]) { tables(); g_dctf.idct4x4dc((void *)d); }
void xo_add_idct_dc(uint8_t *dst, const int16_t *dc, int n)
{
    tables();
    if (n == 4) g_dctf.add8x8_idct_dc(dst, (void *)dc);
    else g_dctf.add16x16_idct_dc(dst, (void *)dc);
}

/* quantiser tables straight out of x264_cqm_init (run by x264_encoder_open) */
static x264_t *g_cqm_h[2];
static x264_t *cqm_h(int cqm)
{
    if (!g_cqm_h[cqm]) {
        x264_param_t p;
        x264_param_default(&p);
        p.i_width = 32; p.i_height = 32; p.i_threads = 1; p.i_log_level = X264_LOG_NONE;
        p.rc.i_rc_method = X264_RC_CQP; p.rc.i_qp_constant = 26;
        p.analyse.b_transform_8x8 = 1;
        p.i_cqm_preset = cqm ? X264_CQM_JVT : X264_CQM_FLAT;
        g_cqm_h[cqm] = x264_encoder_open(&p);
        if (!g_cqm_h[cqm]) abort();
    }
    return g_cqm_h[cqm];
}
void xo_quant4_tables(int cqm, int list, int qp, uint16_t mf[16], uint16_t bias[16])
{
    x264_t *h = cqm_h(cqm);
    memcpy(mf, h->quant4_mf[list][qp], 32); memcpy(bias, h->quant4_bias[list][qp], 32);
}
void xo_quant8_tables(int cqm, int list, int qp, uint16_t mf[64], uint16_t bias[64])
{
    x264_t *h = cqm_h(cqm);
    memcpy(mf, h->quant8_mf[list][qp], 128); memcpy(bias, h->quant8_bias[list][qp], 128);
}
void xo_dequant4_table(int cqm, int list, int dequant_mf[6][16]) { memcpy(dequant_mf, cqm_h(cqm)->dequant4_mf[list], 6 * 16 * sizeof(int)); }
void xo_dequant8_table(int cqm, int list, int dequant_mf[6][64]) { memcpy(dequant_mf, cqm_h(cqm)->dequant8_mf[list], 6 * 64 * sizeof(int)); }
int xo_quant_4x4(int16_t dct[16], const uint16_t mf[16], const uint16_t bias[16]) { tables(); return g_quantf.quant_4x4((void *)dct, (uint16_t *)mf, (uint16_t *)bias); }
int xo_quant_8x8(int16_t dct[64], const uint16_t mf[64], const uint16_t bias[64]) { tables(); return g_quantf.quant_8x8((void *)dct, (uint16_t *)mf, (uint16_t *)bias); }
int xo_quant_4x4_dc(int16_t dct[16], int mf, int bias) { tables(); return g_quantf.quant_4x4_dc((void *)dct, mf, bias); }
int xo_quant_2x2_dc(int16_t dct[4], int mf, int bias) { tables(); return g_quantf.quant_2x2_dc((void *)dct, mf, bias); }
void xo_dequant_4x4(int16_t dct[16], const int dequant_mf[6][16], int qp) { tables(); g_quantf.dequant_4x4((void *)dct, (void *)dequant_mf, qp); }
void xo_dequant_8x8(int16_t dct[64], const int dequant_mf[6][64], int qp) { tables(); g_quantf.dequant_8x8((void *)dct, (void *)dequant_mf, qp); }
void xo_dequant_4x4_dc(int16_t dct[16], const int dequant_mf[6][16], int qp) { tables(); g_quantf.dequant_4x4_dc((void *)dct, (void *)dequant_mf, qp); }

/* ------------------------------------------------------------------------------------------------ */
void xo_zigzag_scan_4x4(int16_t level[16], const int16_t dct[16])
{
    x264_zigzag_function_t z; x264_zigzag_init(0, &z, 0);
    z.scan_4x4(level, (void *)dct);
}
void xo_zigzag_scan_8x8(int16_t level[64], const int16_t dct[64])
{
    x264_zigzag_function_t z; x264_zigzag_init(0, &z, 0);
    z.scan_8x8(level, (void *)dct);
}
int xo_decimate_score(const int16_t *dct, int i_max)
{
    tables();
    return i_max == 15 ? g_quantf.decimate_score15((int16_t *)dct) : i_max == 16 ? g_quantf.decimate_score16((int16_t *)dct)
                                                                                   : g_quantf.decimate_score64((int16_t *)dct);
}

/* drives the reference's own x264_macroblock_encode on a hand-loaded inter macroblock (prediction already in
 * p_fdec, b_skip_mc = 1) */
#include "encoder/macroblock.h"
static x264_t *g_res_h[2][2];
/* opens (once per cqm / 8x8dct pair) and hand-loads a handle: P slice, P_L0 16x16, fenc tiles and the prediction in p_fdec */
static x264_t *res_handle(const xo_resid_in *in, const uint8_t fenc_y[256], const uint8_t fenc_u[64], const uint8_t fenc_v[64],
                          const uint8_t rec_y[256], const uint8_t rec_u[64], const uint8_t rec_v[64])
{
    x264_t **ph = &g_res_h[!!in->cqm][!!in->b_transform_8x8], *h;
    if (!*ph) {
        x264_param_t p;
        x264_param_default(&p);
        p.i_width = 64; p.i_height = 64; p.i_threads = 1; p.i_log_level = X264_LOG_NONE;
        p.rc.i_rc_method = X264_RC_CQP; p.rc.i_qp_constant = 26;
        p.analyse.b_transform_8x8 = 1; /* tables for both paths */
        p.analyse.i_trellis = 0; p.analyse.i_noise_reduction = 0;
        p.i_cqm_preset = in->cqm ? X264_CQM_JVT : X264_CQM_FLAT;
        *ph = x264_encoder_open(&p);
        if (!*ph) abort();
    }
    h = *ph;
    h->sh.i_type = SLICE_TYPE_P;
    h->sh.b_mbaff = 0;
    h->param.analyse.b_dct_decimate = in->b_decimate;
    h->mb.i_type = P_L0; h->mb.i_partition = D_16x16;
    h->mb.b_lossless = 0; h->mb.b_trellis = 0; h->mb.b_noise_reduction = 0; h->mb.b_skip_mc = 1;
    h->mb.b_transform_8x8 = in->b_transform_8x8;
    h->mb.i_qp = in->qp; h->mb.i_chroma_qp = in->chroma_qp;
    h->mb.i_mb_xy = 0; h->mb.i_mb_x = h->mb.i_mb_y = 0;
    /* make the final P_SKIP test fail so i_type stays P_L0 */
    h->mb.cache.mv[0][x264_scan8[0]][0] = 4; h->mb.cache.mv[0][x264_scan8[0]][1] = 0;
    h->mb.cache.pskip_mv[0] = 0; h->mb.cache.pskip_mv[1] = 0;
    h->mb.cache.ref[0][x264_scan8[0]] = 0;
    memset(&h->dct, 0, sizeof(h->dct));
    memset(h->mb.cache.non_zero_count, 0, sizeof(h->mb.cache.non_zero_count));
    for (int y = 0; y < 16; y++) {
        memcpy(h->mb.pic.p_fenc[0] + FENC_STRIDE * y, fenc_y + 16 * y, 16);
        memcpy(h->mb.pic.p_fdec[0] + FDEC_STRIDE * y, rec_y + 16 * y, 16);
    }
    for (int y = 0; y < 8; y++) {
        memcpy(h->mb.pic.p_fenc[1] + FENC_STRIDE * y, fenc_u + 8 * y, 8);
        memcpy(h->mb.pic.p_fenc[2] + FENC_STRIDE * y, fenc_v + 8 * y, 8);
        memcpy(h->mb.pic.p_fdec[1] + FDEC_STRIDE * y, rec_u + 8 * y, 8);
        memcpy(h->mb.pic.p_fdec[2] + FDEC_STRIDE * y, rec_v + 8 * y, 8);
    }
    return h;
}

void xo_residual_inter_mb(const xo_resid_in *in, const uint8_t fenc_y[256], const uint8_t fenc_u[64], const uint8_t fenc_v[64],
                          uint8_t rec_y[256], uint8_t rec_u[64], uint8_t rec_v[64], xo_resid_out *out)
{
    x264_t *h = res_handle(in, fenc_y, fenc_u, fenc_v, rec_y, rec_u, rec_v);
    x264_macroblock_encode(h);
    memset(out, 0, sizeof(*out));
    memcpy(out->luma4x4, h->dct.luma4x4, sizeof(out->luma4x4));
    memcpy(out->luma8x8, h->dct.luma8x8, sizeof(out->luma8x8));
    memcpy(out->chroma_dc, h->dct.chroma_dc, sizeof(out->chroma_dc));
    for (int i = 0; i < 27; i++) out->nnz[i] = h->mb.cache.non_zero_count[x264_scan8[i]];
    out->cbp_luma = h->mb.i_cbp_luma; out->cbp_chroma = h->mb.i_cbp_chroma;
    for (int y = 0; y < 16; y++) memcpy(rec_y + 16 * y, h->mb.pic.p_fdec[0] + FDEC_STRIDE * y, 16);
    for (int y = 0; y < 8; y++) {
        memcpy(rec_u + 8 * y, h->mb.pic.p_fdec[1] + FDEC_STRIDE * y, 8);
        memcpy(rec_v + 8 * y, h->mb.pic.p_fdec[2] + FDEC_STRIDE * y, 8);
    }
}

/* one I_16x16 macroblock through the reference's own x264_macroblock_encode: neighbour pixels hand-loaded around p_fdec (the row above,
 * the left column and the corner are inside the fdec buffer, common/macroblock.c:990-1004), slice type picked so that b_decimate matches */
void xo_residual_intra16_mb(const xo_resid_in *in, int mode16, int mode_chroma, const uint8_t fenc_y[256], const uint8_t fenc_u[64],
                            const uint8_t fenc_v[64], const uint8_t nb_y[33], const uint8_t nb_u[17], const uint8_t nb_v[17],
                            uint8_t rec_y[256], uint8_t rec_u[64], uint8_t rec_v[64], xo_resid_out *out, int16_t luma_dc[16])
{
    xo_resid_in tmp = *in;
    tmp.b_transform_8x8 = 0;
    memset(rec_y, 0, 256); memset(rec_u, 0, 64); memset(rec_v, 0, 64);
    x264_t *h = res_handle(&tmp, fenc_y, fenc_u, fenc_v, rec_y, rec_u, rec_v);
    h->sh.i_type = in->b_decimate ? SLICE_TYPE_P : SLICE_TYPE_I; /* macroblock.c:193: B || (b_dct_decimate && P) */
    h->param.analyse.b_dct_decimate = in->b_decimate;
    h->mb.i_type = I_16x16;
    h->mb.i_intra16x16_pred_mode = mode16;
    h->mb.i_chroma_pred_mode = mode_chroma;
    const uint8_t *nb[3] = { nb_y, nb_u, nb_v };
    for (int p = 0; p < 3; p++) {
        const int n = p ? 8 : 16;
        uint8_t *d = h->mb.pic.p_fdec[p];
        d[-FDEC_STRIDE - 1] = nb[p][0];
        memcpy(d - FDEC_STRIDE, nb[p] + 1, n);
        for (int y = 0; y < n; y++) d[y * FDEC_STRIDE - 1] = nb[p][1 + n + y];
    }
    x264_macroblock_encode(h);
    memset(out, 0, sizeof(*out));
    memcpy(out->luma4x4, h->dct.luma4x4, sizeof(out->luma4x4));
    memcpy(out->chroma_dc, h->dct.chroma_dc, sizeof(out->chroma_dc));
    for (int i = 0; i < 27; i++) out->nnz[i] = h->mb.cache.non_zero_count[x264_scan8[i]];
    /* blocks the reference did not code keep stale levels in h->dct (it never reads them): report them as zero, like the port */
    for (int i = 0; i < 24; i++) if (!out->nnz[i]) memset(out->luma4x4[i], 0, sizeof(out->luma4x4[i]));
    for (int ch = 0; ch < 2; ch++) if (!out->nnz[25 + ch]) memset(out->chroma_dc[ch], 0, sizeof(out->chroma_dc[ch]));
    memset(luma_dc, 0, 16 * sizeof(int16_t));
    if (out->nnz[24]) memcpy(luma_dc, h->dct.luma16x16_dc, 16 * sizeof(int16_t));
    out->cbp_luma = h->mb.i_cbp_luma; out->cbp_chroma = h->mb.i_cbp_chroma;
    for (int y = 0; y < 16; y++) memcpy(rec_y + 16 * y, h->mb.pic.p_fdec[0] + FDEC_STRIDE * y, 16);
    for (int y = 0; y < 8; y++) {
        memcpy(rec_u + 8 * y, h->mb.pic.p_fdec[1] + FDEC_STRIDE * y, 8);
        memcpy(rec_v + 8 * y, h->mb.pic.p_fdec[2] + FDEC_STRIDE * y, 8);
    }
}

/* intra mode costs with the reference's own predictors (x264_predict_16x16_init / x264_predict_8x8c_init tables) and mbcmp functions,
 * on an FDEC_STRIDE tile loaded with the neighbour pixels; the candidate lists and the lambda terms are the few lines of glue around
 * them in x264_mb_analyse_intra / _intra_chroma (static there), followed literally */
static void load_tile(uint8_t *tile, const uint8_t *nb, int n)
{
    uint8_t *p = tile + FDEC_STRIDE + 16; /* block origin; row above and column left inside the tile */
    p[-FDEC_STRIDE - 1] = nb[0];
    memcpy(p - FDEC_STRIDE, nb + 1, n);
    for (int y = 0; y < n; y++) p[y * FDEC_STRIDE - 1] = nb[1 + n + y];
}
static void ref_predict(int chroma, int mode, const uint8_t *nb, uint8_t *pred)
{
    static x264_predict_t p16[7], p8c[7];
    static int init;
    DECLARE_ALIGNED_16(uint8_t tile[FDEC_STRIDE * 17 + 32]);
    if (!init) { x264_predict_16x16_init(0, p16); x264_predict_8x8c_init(0, p8c); init = 1; }
    const int n = chroma ? 8 : 16;
    memset(tile, 0xAA, sizeof(tile));
    load_tile(tile, nb, n);
    (chroma ? p8c : p16)[mode](tile + FDEC_STRIDE + 16);
    for (int y = 0; y < n; y++) memcpy(pred + n * y, tile + FDEC_STRIDE + 16 + y * FDEC_STRIDE, n);
}
void xo_predict_16x16(int mode, const uint8_t nb[33], uint8_t pred[256]) { ref_predict(0, mode, nb, pred); }
void xo_predict_8x8c(int mode, const uint8_t nb[17], uint8_t pred[64]) { ref_predict(1, mode, nb, pred); }

void xo_intra_mb_costs(const xo_intra_in *in, const uint8_t fenc_y[256], const uint8_t fenc_u[64], const uint8_t fenc_v[64],
                       const uint8_t nb_y[33], const uint8_t nb_u[17], const uint8_t nb_v[17], xo_intra_out *out)
{
    static x264_pixel_function_t pixf;
    static int init;
    if (!init) { x264_pixel_init(0, &pixf); init = 1; }
    x264_pixel_cmp_t *mbcmp = in->mbcmp_satd ? pixf.satd : pixf.sad;
    DECLARE_ALIGNED_16(uint8_t fe[16 * FENC_STRIDE]);
    DECLARE_ALIGNED_16(uint8_t fd[3][16 * 16]);
    int predict_mode[4], i_max;
    for (int i = 0; i < 7; i++) out->cost16[i] = out->cost_chroma[i] = -1;
    out->best16 = out->best_chroma = COST_MAX;
    out->mode16 = out->mode_chroma = 0;
    const unsigned nbr = in->neighbour;
    /* predict_16x16_mode_available, analyse.c:372-404 */
    if (nbr & MB_TOPLEFT) { predict_mode[0] = I_PRED_16x16_V; predict_mode[1] = I_PRED_16x16_H; predict_mode[2] = I_PRED_16x16_DC; predict_mode[3] = I_PRED_16x16_P; i_max = 4; }
    else if (nbr & MB_LEFT) { predict_mode[0] = I_PRED_16x16_DC_LEFT; predict_mode[1] = I_PRED_16x16_H; i_max = 2; }
    else if (nbr & MB_TOP) { predict_mode[0] = I_PRED_16x16_DC_TOP; predict_mode[1] = I_PRED_16x16_V; i_max = 2; }
    else { predict_mode[0] = I_PRED_16x16_DC_128; i_max = 1; }
    for (int y = 0; y < 16; y++) memcpy(fe + FENC_STRIDE * y, fenc_y + 16 * y, 16);
    for (int i = 0; i < i_max; i++) {
        int m = predict_mode[i];
        xo_predict_16x16(m, nb_y, fd[0]);
        int c = mbcmp[PIXEL_16x16](fd[0], 16, fe, FENC_STRIDE) + in->lambda * bs_size_ue(x264_mb_pred_mode16x16_fix[m]);
        out->cost16[m] = c;
        if (c < out->best16) { out->best16 = c; out->mode16 = m; }
    }
    if (in->b_slice_b) out->best16 += in->lambda * 9;
    /* predict_8x8chroma_mode_available, analyse.c:407-440 */
    if (nbr & MB_TOPLEFT) { predict_mode[0] = I_PRED_CHROMA_V; predict_mode[1] = I_PRED_CHROMA_H; predict_mode[2] = I_PRED_CHROMA_DC; predict_mode[3] = I_PRED_CHROMA_P; i_max = 4; }
    else if (nbr & MB_LEFT) { predict_mode[0] = I_PRED_CHROMA_DC_LEFT; predict_mode[1] = I_PRED_CHROMA_H; i_max = 2; }
    else if (nbr & MB_TOP) { predict_mode[0] = I_PRED_CHROMA_DC_TOP; predict_mode[1] = I_PRED_CHROMA_V; i_max = 2; }
    else { predict_mode[0] = I_PRED_CHROMA_DC_128; i_max = 1; }
    for (int i = 0; i < i_max; i++) {
        int m = predict_mode[i], c = in->lambda * bs_size_ue(x264_mb_pred_mode8x8c_fix[m]);
        for (int ch = 0; ch < 2; ch++) {
            const uint8_t *src = ch ? fenc_v : fenc_u;
            for (int y = 0; y < 8; y++) memcpy(fe + FENC_STRIDE * y, src + 8 * y, 8);
            xo_predict_8x8c(m, ch ? nb_v : nb_u, fd[1 + ch]);
            c += mbcmp[PIXEL_8x8](fd[1 + ch], 8, fe, FENC_STRIDE);
        }
        out->cost_chroma[m] = c;
        if (c < out->best_chroma) { out->best_chroma = c; out->mode_chroma = m; }
    }
}

/* the reference's own x264_macroblock_probe_skip on a hand-loaded macroblock, prediction already in p_fdec (b_bidir = 1) */
extern const int x264_lambda2_tab[52];
int xo_lambda2(int qp) { return x264_lambda2_tab[qp]; }
int xo_probe_skip_mb(const xo_resid_in *in, const uint8_t fenc_y[256], const uint8_t fenc_u[64], const uint8_t fenc_v[64],
                     const uint8_t pred_y[256], const uint8_t pred_u[64], const uint8_t pred_v[64])
{
    xo_resid_in tmp = *in;
    tmp.b_transform_8x8 = 0;
    x264_t *h = res_handle(&tmp, fenc_y, fenc_u, fenc_v, pred_y, pred_u, pred_v);
    return x264_macroblock_probe_skip(h, 1);
}

/* ------------------------------------------------------------------------------------------------ */
/* lowres lookahead through the reference's own x264_rc_analyse_slice -> x264_slicetype_frame_cost   */
#include "common/predict.h"
static x264_frame_t *g_la_frames[3];
static x264_t *g_la_h;
static int g_la_key[6];

static void ref_lowres_frame_cost(const xo_geom *g, const xo_lowres_in *in, const uint8_t *const fenc[4], const uint8_t *const fref0[4],
                                  const uint8_t *const fref1[4], int16_t (*mvs0)[2], int *costs0, int16_t (*mvs1)[2], int *costs1,
                                  const int16_t (*ref1_mvs)[2], uint16_t *intra_cost, xo_lowres_out *out, int b_vbv, const uint16_t *inv_qscale,
                                  int *row_satd)
{
    int key[6] = { g->width, g->height, in->me_method, in->mbcmp_satd, in->fpel_satd, in->b_weighted_bipred };
    if (!g_la_h || memcmp(key, g_la_key, sizeof(key))) {
        x264_param_t p;
        x264_param_default(&p);
        p.i_width = g->width; p.i_height = g->height; p.i_threads = 1; p.i_log_level = X264_LOG_NONE;
        p.rc.i_rc_method = X264_RC_CQP; p.rc.i_qp_constant = 26;
        p.analyse.i_me_method = in->fpel_satd ? X264_ME_TESA : in->me_method;
        p.analyse.i_subpel_refine = in->mbcmp_satd ? 7 : 1;
        p.analyse.b_weighted_bipred = in->b_weighted_bipred;
        p.i_bframe = 3; p.i_bframe_adaptive = X264_B_ADAPT_FAST;
        g_la_h = x264_encoder_open(&p);
        if (!g_la_h) abort();
        for (int i = 0; i < 3; i++) g_la_frames[i] = x264_frame_new(g_la_h);
        memcpy(g_la_key, key, sizeof(key));
    }
    x264_t *h = g_la_h;
    h->param.analyse.i_me_range = in->me_range;
    const int n_mb = g->mb_width * g->mb_height;
    x264_frame_t *f0 = g_la_frames[0], *f1 = g_la_frames[1], *fb = g_la_frames[2];
    /* the VBV form of x264_slicetype_frame_cost is selected by these two parameters alone (slicetype.c:300-309) */
    h->param.rc.i_vbv_buffer_size = b_vbv ? 1000 : 0;
    h->param.rc.i_aq_mode = inv_qscale ? X264_AQ_VARIANCE : X264_AQ_NONE;
    if (inv_qscale) {
        if (!fb->i_inv_qscale_factor) fb->i_inv_qscale_factor = x264_malloc(n_mb * sizeof(uint16_t));
        memcpy(fb->i_inv_qscale_factor, inv_qscale, n_mb * sizeof(uint16_t));
    }
    fb->i_row_satds[in->b - in->p0][in->p1 - in->b][0] = -1; /* "row sums not calculated yet" (slicetype.c:270) */
    const int b_bidir = in->b < in->p1;
    const uint8_t *const *src[3] = { fref0, fref1, fenc };
    x264_frame_t *dst[3] = { f0, f1, fb };
    for (int k = 0; k < 3; k++)
        for (int i = 0; i < 4; i++)
            if (src[k] && src[k][i]) put_plane(dst[k]->lowres[i], src[k][i], g->stride_lowres, g->lines_lowres);
    const int d0 = in->b - in->p0 - 1, d1 = in->p1 - in->b - 1;
    memset(fb->i_cost_est, -1, sizeof(fb->i_cost_est));
    if (in->b != in->p0) {
        memcpy(fb->lowres_mvs[0][d0], mvs0, n_mb * 4); memcpy(fb->lowres_mv_costs[0][d0], costs0, n_mb * 4);
        if (in->do_search[0]) fb->lowres_mvs[0][d0][0][0] = 0x7FFF;
    }
    if (b_bidir) {
        memcpy(fb->lowres_mvs[1][d1], mvs1, n_mb * 4); memcpy(fb->lowres_mv_costs[1][d1], costs1, n_mb * 4);
        if (in->do_search[1]) fb->lowres_mvs[1][d1][0][0] = 0x7FFF;
        memcpy(f1->lowres_mvs[0][in->p1 - in->p0 - 1], ref1_mvs, n_mb * 4);
    }
    memcpy(fb->i_intra_cost, intra_cost, n_mb * 2);
    fb->b_intra_calculated = in->b_intra_calculated;
    /* drive x264_rc_analyse_slice (slicetype.c:638-679) so that it evaluates exactly (p0,p1,b) */
    h->fenc = fb; h->fdec = g_la_frames[0]; /* fdec only receives copies of row satds */
    h->fref0[0] = f0; h->fref1[0] = f1;
    h->frames.current[0] = NULL;
    if (in->p0 == in->p1 && in->p0 == in->b) fb->i_type = X264_TYPE_I;
    else if (!b_bidir) {
        fb->i_type = X264_TYPE_P;
        /* p1 = 1 + number of leading B frames in h->frames.current: fake (p1-1) B entries */
        static x264_frame_t bf;
        bf.i_type = X264_TYPE_B;
        int i;
        for (i = 0; i < in->p1 - 1; i++) h->frames.current[i] = &bf;
        h->frames.current[i] = NULL;
    } else {
        fb->i_type = X264_TYPE_B;
        f0->i_poc = 0; f1->i_poc = 2 * in->p1; fb->i_poc = 2 * in->p1 - 2 * (in->p1 - in->b); /* p1=(poc1-poc0)/2, b as slicetype.c:664 */
        fb->i_poc = f1->i_poc - 2 * in->b; /* the reference computes b = (poc1 - poc_enc)/2 */
    }
    x264_rc_analyse_slice(h);
    h->frames.current[0] = NULL;
    int score = fb->i_cost_est[in->b - in->p0][in->p1 - in->b];
    (void)score;
    if (in->b != in->p0) { memcpy(mvs0, fb->lowres_mvs[0][d0], n_mb * 4); memcpy(costs0, fb->lowres_mv_costs[0][d0], n_mb * 4); }
    if (b_bidir) { memcpy(mvs1, fb->lowres_mvs[1][d1], n_mb * 4); memcpy(costs1, fb->lowres_mv_costs[1][d1], n_mb * 4); }
    memcpy(intra_cost, fb->i_intra_cost, n_mb * 2);
    memset(out, 0, sizeof(*out));
    out->score = fb->i_cost_est[in->b - in->p0][in->p1 - in->b]; /* NB: B scores arrive scaled by 100/(120+bias) (slicetype.c:338) */
    out->score_aq = fb->i_cost_est_aq[in->b - in->p0][in->p1 - in->b];
    out->intra_mbs = fb->i_intra_mbs[in->b - in->p0];
    out->intra_cost_sum = fb->i_cost_est[0][0];
    if (b_vbv && row_satd) memcpy(row_satd, fb->i_row_satds[in->b - in->p0][in->p1 - in->b], g->mb_height * sizeof(int));
    h->param.rc.i_vbv_buffer_size = 0; h->param.rc.i_aq_mode = X264_AQ_NONE;
}

void xo_lowres_frame_cost(const xo_geom *g, const xo_lowres_in *in, const uint8_t *const fenc[4], const uint8_t *const fref0[4],
                          const uint8_t *const fref1[4], int16_t (*mvs0)[2], int *costs0, int16_t (*mvs1)[2], int *costs1,
                          const int16_t (*ref1_mvs)[2], uint16_t *intra_cost, xo_lowres_out *out)
{
    ref_lowres_frame_cost(g, in, fenc, fref0, fref1, mvs0, costs0, mvs1, costs1, ref1_mvs, intra_cost, out, 0, NULL, NULL);
}

void xo_lowres_frame_cost_vbv(const xo_geom *g, const xo_lowres_in *in, const uint8_t *const fenc[4], const uint8_t *const fref0[4],
                              const uint8_t *const fref1[4], int16_t (*mvs0)[2], int *costs0, int16_t (*mvs1)[2], int *costs1,
                              const int16_t (*ref1_mvs)[2], uint16_t *intra_cost, xo_lowres_out *out, const uint16_t *inv_qscale, int *row_satd)
{
    ref_lowres_frame_cost(g, in, fenc, fref0, fref1, mvs0, costs0, mvs1, costs1, ref1_mvs, intra_cost, out, row_satd != NULL, inv_qscale, row_satd);
}

void xo_lowres_intra_pred(int mode, const uint8_t *l0, int stride, int bx, int by, uint8_t out[64])
{
    static x264_predict_t p8c[7];
    static x264_predict8x8_t p8[12];
    static int init;
    static x264_predict_8x8_filter_t filt;
    if (!init) { x264_predict_8x8c_init(0, p8c); x264_predict_8x8_init(0, p8, &filt); init = 1; }
    DECLARE_ALIGNED_16(uint8_t buf[9 * FDEC_STRIDE]);
    DECLARE_ALIGNED_16(uint8_t edge[33]);
    const uint8_t *src = l0 + by * stride + bx - 1;
    uint8_t *pix = &buf[8 + FDEC_STRIDE - 1];
    memcpy(pix - FDEC_STRIDE, src - stride, 17);
    for (int i = 0; i < 8; i++) pix[i * FDEC_STRIDE] = src[i * stride];
    pix++;
    if (mode < 4) p8c[mode](pix);
    else { x264_predict_8x8_filter(pix, edge, ALL_NEIGHBORS, ALL_NEIGHBORS); p8[mode - 1](pix, edge); }
    for (int y = 0; y < 8; y++) memcpy(out + 8 * y, pix + y * FDEC_STRIDE, 8);
}
int xo_lowres_intra_cost(const uint8_t *l0, int stride, int bx, int by, int mbcmp_satd)
{
    uint8_t pred[64], pp[8 * 16], fenc[8 * 16];
    int best = 1 << 30;
    for (int y = 0; y < 8; y++) memcpy(fenc + 16 * y, l0 + (by + y) * stride + bx, 8);
    for (int m = 0; m < 10; m++) {
        xo_lowres_intra_pred(m, l0, stride, bx, by, pred);
        for (int y = 0; y < 8; y++) memcpy(pp + 16 * y, pred + 8 * y, 8);
        int c = xo_pixel_cmp(mbcmp_satd ? XO_SATD : XO_SAD, XO_8x8, pp, 16, fenc, 16);
        if (c < best) best = c;
    }
    return best + 5;
}

/* ------------------------------------------------------------------------------------------------ */
/* deblocking through the reference's own x264_frame_deblock_row (S/common/frame.c:621-792)          */
void xo_frame_deblock(const xo_geom *g, const xo_deblock_in *d, uint8_t *py, uint8_t *pu, uint8_t *pv, int stride_c)
{
    x264_frame_t *f;
    x264_t *h = get_h(g->width, g->height, X264_ME_ESA, 1, 0, 1, &f); /* the i_bframe > 0 handle: its frames carry list-1 mv/ref arrays */
    static x264_t *fd_h[64];
    static x264_frame_t *fd_f[64];
    int k = 0;
    while (k < 63 && fd_h[k] && fd_h[k] != h) k++;
    if (!fd_h[k]) { fd_h[k] = h; fd_f[k] = x264_frame_new(h); } /* fdec is only attached by x264_encoder_encode: give it one */
    x264_frame_t *fd = fd_f[k];
    h->fdec = fd;
    /* x264_macroblock_slice_init (S/common/macroblock.c:777-781): per-macroblock state lives in the frame being decoded */
    h->mb.mv[0] = fd->mv[0]; h->mb.mv[1] = fd->mv[1]; h->mb.ref[0] = fd->ref[0]; h->mb.ref[1] = fd->ref[1]; h->mb.type = fd->mb_type;
    const int n = g->mb_width * g->mb_height;
    for (int y = 0; y < g->lines; y++) memcpy(fd->plane[0] + y * fd->i_stride[0], py + y * g->stride, 16 * g->mb_width);
    for (int y = 0; y < g->lines / 2; y++) {
        memcpy(fd->plane[1] + y * fd->i_stride[1], pu + y * stride_c, 8 * g->mb_width);
        memcpy(fd->plane[2] + y * fd->i_stride[2], pv + y * stride_c, 8 * g->mb_width);
    }
    const int save_inter = h->param.analyse.inter, save_off = h->param.analyse.i_chroma_qp_offset;
    h->sh.i_alpha_c0_offset = d->alpha_c0_offset; h->sh.i_beta_offset = d->beta_offset; h->sh.b_mbaff = 0;
    h->sh.i_type = d->b_slice_b ? SLICE_TYPE_B : SLICE_TYPE_P;
    h->param.analyse.i_chroma_qp_offset = d->chroma_qp_offset;
    h->chroma_qp_table = i_chroma_qp_table + 12 + d->chroma_qp_offset;
    h->param.analyse.inter = d->b_psub8x8 ? (save_inter | X264_ANALYSE_PSUB8x8) : (save_inter & ~X264_ANALYSE_PSUB8x8);
    x264_pps_t *pps = (x264_pps_t *)h->pps;
    const int save_cabac = pps->b_cabac, save_t8 = pps->b_transform_8x8_mode;
    if (d->b_cavlc_8x8dct) { pps->b_cabac = 0; pps->b_transform_8x8_mode = 1; } else pps->b_cabac = 1;
    memcpy(h->mb.type, d->type, n); memcpy(h->mb.qp, d->qp, n); memcpy(h->mb.mb_transform_size, d->transform8x8, n);
    memcpy(h->mb.non_zero_count, d->nnz, (size_t)n * 24);
    for (int l = 0; l < 2; l++) { memcpy(h->mb.ref[l], d->ref[l], (size_t)n * 4); memcpy(h->mb.mv[l], d->mv[l], (size_t)n * 16 * 4); }
    for (int mb_y = 0; mb_y < g->mb_height; mb_y++) x264_frame_deblock_row(h, mb_y);
    for (int y = 0; y < g->lines; y++) memcpy(py + y * g->stride, fd->plane[0] + y * fd->i_stride[0], 16 * g->mb_width);
    for (int y = 0; y < g->lines / 2; y++) {
        memcpy(pu + y * stride_c, fd->plane[1] + y * fd->i_stride[1], 8 * g->mb_width);
        memcpy(pv + y * stride_c, fd->plane[2] + y * fd->i_stride[2], 8 * g->mb_width);
    }
    h->param.analyse.inter = save_inter; h->param.analyse.i_chroma_qp_offset = save_off;
    h->chroma_qp_table = i_chroma_qp_table + 12 + save_off;
    pps->b_cabac = save_cabac; pps->b_transform_8x8_mode = save_t8;
}

/* ------------------------------------------------------------------------------------------------ */
/* whole-frame metrics through the reference's own functions                                          */
int64_t xo_frame_ssd(const uint8_t *p1, int s1, const uint8_t *p2, int s2, int width, int height)
{
    tables();
    return x264_pixel_ssd_wxh(&g_pixf, (uint8_t *)p1, s1, (uint8_t *)p2, s2, width, height);
}
float xo_frame_ssim(const uint8_t *p1, int s1, const uint8_t *p2, int s2, int width, int height)
{
    tables();
    void *buf = malloc(8 * (width / 4 + 3) * sizeof(int));
    float r = x264_pixel_ssim_wxh(&g_pixf, (uint8_t *)p1, s1, (uint8_t *)p2, s2, width, height, buf);
    free(buf);
    return r;
}
void xo_frame_ssim_sums(const uint8_t *p1, int s1, const uint8_t *p2, int s2, int width, int height, int (*sums)[4])
{
    tables();
    const int w4 = width >> 2, h4 = height >> 2;
    for (int by = 0; by < h4; by++)
        for (int bx = 0; bx + 1 < w4 + (w4 & 1); bx += 2) {
            int t[2][4];
            g_pixf.ssim_4x4x2_core(p1 + 4 * by * s1 + 4 * bx, s1, p2 + 4 * by * s2 + 4 * bx, s2, t);
            memcpy(sums[by * w4 + bx], t[0], 16);
            if (bx + 1 < w4) memcpy(sums[by * w4 + bx + 1], t[1], 16);
        }
}
void xo_frame_mb_energy(const xo_geom *g, const uint8_t *py, const uint8_t *pu, const uint8_t *pv, int stride_c, uint32_t *out)
{
    /* ac_energy_mb is static in ratecontrol.c: same three table calls; xo_frame_aq below goes through the real function */
    tables();
    for (int mb_y = 0; mb_y < g->mb_height; mb_y++)
        for (int mb_x = 0; mb_x < g->mb_width; mb_x++) {
            unsigned int var = g_pixf.var[PIXEL_16x16]((uint8_t *)py + 16 * (mb_x + mb_y * g->stride), g->stride);
            var += g_pixf.var[PIXEL_8x8]((uint8_t *)pu + 8 * (mb_x + mb_y * stride_c), stride_c);
            var += g_pixf.var[PIXEL_8x8]((uint8_t *)pv + 8 * (mb_x + mb_y * stride_c), stride_c);
            out[mb_x + mb_y * g->mb_width] = X264_MAX(var, 1);
        }
}
void xo_frame_mb_hadamard_ac(const xo_geom *g, const uint8_t *py, uint64_t *out)
{
    tables();
    for (int mb_y = 0; mb_y < g->mb_height; mb_y++)
        for (int mb_x = 0; mb_x < g->mb_width; mb_x++)
            out[mb_x + mb_y * g->mb_width] = g_pixf.hadamard_ac[PIXEL_16x16]((uint8_t *)py + 16 * (mb_x + mb_y * g->stride), g->stride);
}
void xo_aq_from_energy(const uint32_t *energy, int n, float aq_strength, float *qp_offset, uint16_t *inv_qscale)
{
    (void)energy; (void)n; (void)aq_strength; (void)qp_offset; (void)inv_qscale;
    abort(); /* the reference has no such entry point: use xo_frame_aq */
}
void xo_frame_aq(const xo_geom *g, const uint8_t *py, const uint8_t *pu, const uint8_t *pv, int stride_c, float aq_strength, float *qp_offset,
                 uint16_t *inv_qscale)
{
    x264_frame_t *f;
    x264_t *h = get_h(g->width, g->height, X264_ME_ESA, 1, 0, 1, &f); /* the lowres handle: frames carry i_inv_qscale_factor */
    for (int y = 0; y < g->lines; y++) memcpy(f->plane[0] + y * f->i_stride[0], py + y * g->stride, 16 * g->mb_width);
    for (int y = 0; y < g->lines / 2; y++) {
        memcpy(f->plane[1] + y * f->i_stride[1], pu + y * stride_c, 8 * g->mb_width);
        memcpy(f->plane[2] + y * f->i_stride[2], pv + y * stride_c, 8 * g->mb_width);
    }
    const int n_mb = g->mb_width * g->mb_height;
    /* CQP handles switch AQ off (encoder.c:429), so their frames lack the two arrays (frame.c:137-142): give them some */
    if (!f->f_qp_offset) f->f_qp_offset = x264_malloc(n_mb * sizeof(float));
    if (!f->i_inv_qscale_factor) f->i_inv_qscale_factor = x264_malloc(n_mb * sizeof(uint16_t));
    const float save = h->param.rc.f_aq_strength;
    h->param.rc.f_aq_strength = aq_strength;
    h->mb.b_interlaced = 0;
    x264_adaptive_quant_frame(h, f);
    h->param.rc.f_aq_strength = save;
    const int n = g->mb_width * g->mb_height;
    memcpy(qp_offset, f->f_qp_offset, n * sizeof(float));
    if (h->frames.b_have_lowres) memcpy(inv_qscale, f->i_inv_qscale_factor, n * sizeof(uint16_t));
}

/* ------------------------------------------------------------------------------------------------ */
int xo_me_refine_bidir_satd(const xo_geom *g, const uint8_t *fenc_plane, const uint8_t *const fref0[4], const uint8_t *const fref1[4],
                            const xo_me_in *in, const int16_t mvp0[2], const int16_t mvp1[2], int weight, int mbcmp_satd, int16_t mv0[2], int16_t mv1[2])
{
    x264_frame_t *f;
    x264_t *h = get_h(g->width, g->height, X264_ME_HEX, mbcmp_satd ? 7 : 1, 0, 0, &f);
    DECLARE_ALIGNED_16(uint8_t fenc[16 * 16]);
    x264_me_t m0, m1;
    ensure_costs(in->qp);
    memset(&m0, 0, sizeof(m0)); memset(&m1, 0, sizeof(m1));
    for (int y = 0; y < x264_pixel_size[in->i_pixel].h; y++)
        memcpy(fenc + 16 * y, fenc_plane + (in->by + y) * g->stride + in->bx, x264_pixel_size[in->i_pixel].w);
    for (int k = 0; k < 2; k++) { h->mb.mv_min_spel[k] = in->mv_min_spel[k]; h->mb.mv_max_spel[k] = in->mv_max_spel[k]; }
    m0.i_pixel = m1.i_pixel = in->i_pixel;
    m0.p_cost_mv = m1.p_cost_mv = g_cost_mv[in->qp] + 2 * 4 * 2048;
    m0.i_stride[0] = m1.i_stride[0] = g->stride;
    m0.p_fenc[0] = m1.p_fenc[0] = fenc;
    for (int k = 0; k < 4; k++) {
        m0.p_fref[k] = (uint8_t *)fref0[k] + in->by * g->stride + in->bx;
        m1.p_fref[k] = (uint8_t *)fref1[k] + in->by * g->stride + in->bx;
    }
    m0.mvp[0] = mvp0[0]; m0.mvp[1] = mvp0[1]; m1.mvp[0] = mvp1[0]; m1.mvp[1] = mvp1[1];
    m0.mv[0] = mv0[0]; m0.mv[1] = mv0[1]; m1.mv[0] = mv1[0]; m1.mv[1] = mv1[1];
    x264_me_refine_bidir_satd(h, &m0, &m1, weight);
    mv0[0] = m0.mv[0]; mv0[1] = m0.mv[1]; mv1[0] = m1.mv[0]; mv1[1] = m1.mv[1];
    return -1; /* the reference keeps the cost to itself */
}
