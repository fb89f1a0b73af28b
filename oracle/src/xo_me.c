/* xo_me.c — ORACLE (test infrastructure only): MV cost tables and the integer/sub-pel motion search of
 * S/encoder/me.c (x264_me_search_ref :156-631, refine_subpel :680-778), restated with whole padded planes
 * as inputs.  UMH is not restated (outside SURVEY.md §8).  The ESA branch is written as the plain raster
 * argmin that SURVEY.md App. D1 proved byte-identical to the SEA code; the TESA branch reproduces the
 * ADS/SAD threshold list exactly because its result depends on it. */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "xo.h"

static const int blk_w[7] = { 16, 16, 8, 8, 8, 4, 4 };
static const int blk_h[7] = { 16, 8, 16, 8, 4, 8, 4 };

static inline int clip3(int v, int lo, int hi) { return v < lo ? lo : v > hi ? hi : v; }

/* S/encoder/analyse.c:140-148: lambda = 2^(qp/6-2) rounded by hand */
int xo_lambda(int qp)
{
    static const uint8_t tab[52] = { 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 4, 4, 4,
                                     5, 6, 6, 7, 8, 9, 10, 11, 13, 14, 16, 18, 20, 23, 25, 29, 32, 36, 40, 45,
                                     51, 57, 64, 72, 81, 91 };
    return tab[qp];
}

/* S/encoder/analyse.c:40,192-203.  The reference evaluates, in this exact mixed float/double form,
 *   lambda * ( ((float)log(i+1)) / log(2) * 2 + 0.718f + !!i ) + .5f   truncated to int16.
 * Same compiler and -ffast-math as the reference build; equality with the reference's
 * own g_cost_mv[] is asserted for all 52 qps in tests/test_oracle_vs_ref.py. */
void xo_cost_mv_table(int qp, int16_t *out)
{
    int lambda = xo_lambda(qp);
    int16_t *c = out + 2 * 4 * 2048;
    for (int i = 0; i <= 2 * 4 * 2048; i++)
        c[-i] = c[i] = lambda * (((float)log((double)(i + 1))) / (log((double)2)) * 2 + 0.718f + !!i) + .5f;
}

/* per-process cache of the 52 tables */
static int16_t *g_tab[52];
static const int16_t *cost_table(int qp)
{
    if (!g_tab[qp]) {
        int16_t *t = malloc((4 * 4 * 2048 + 1) * sizeof(int16_t));
        xo_cost_mv_table(qp, t);
        g_tab[qp] = t;
    }
    return g_tab[qp] + 2 * 4 * 2048;
}

typedef struct {
    const xo_geom *g;
    const xo_me_in *in;
    uint8_t fenc[16 * 16];       /* stride-16 copy of the block, like h->mb.pic.p_fenc */
    const uint8_t *fref[4];      /* block-positioned pointers into full/h/v/c planes */
    const int16_t *cmx, *cmy;    /* p_cost_mv - mvp[k] */
    int bw, bh, stride;
    int fpel_metric, sub_metric; /* XO_SAD / XO_SATD */
    /* b_chroma_me (h->mb.b_chroma_me && i_pixel <= PIXEL_8x8): stride-16 copies of the fenc chroma blocks, block-positioned
     * reference chroma planes (m->p_fref[4], [5]) */
    int chroma_me, stride_c;
    uint8_t fenc_c[2][16 * 8];
    const uint8_t *fref_c[2];
} me_ctx;

static int fpel_cost(const me_ctx *c, int mx, int my)
{
    return xo_pixel_cmp(c->fpel_metric, c->in->i_pixel, c->fenc, 16, c->fref[0] + my * c->stride + mx, c->stride) +
           c->cmx[mx << 2] + c->cmy[my << 2];
}
#define TRY_FPEL(mx_, my_) do { int cst_ = fpel_cost(c, (mx_), (my_)); \
    if (cst_ < bcost) { bcost = cst_; bmx = (mx_); bmy = (my_); } } while (0)

/* qpel-position block compare through the 4 hpel planes (COST_MV_HPEL me.c:65-72, COST_MV_SAD/SATD :646-677) */
static int qpel_cost(const me_ctx *c, int metric, int mx, int my)
{
    uint8_t tmp[16 * 16];
    xo_mc_luma(tmp, 16, c->fref, c->stride, mx, my, c->bw, c->bh);
    return xo_pixel_cmp(metric, c->in->i_pixel, c->fenc, 16, tmp, 16) + c->cmx[mx] + c->cmy[my];
}

/* COST_MV_SATD (me.c:655-677): mbcmp of the luma block, plus — only while the candidate still beats bcost — the two chroma
 * blocks predicted by mc_chroma and compared with mbcmp[i_pixel+3] */
static int satd_cost(const me_ctx *c, int mx, int my, int bcost)
{
    int cost = qpel_cost(c, c->sub_metric, mx, my);
    if (c->chroma_me && cost < bcost) {
        uint8_t pix[8 * 8];
        for (int pl = 0; pl < 2; pl++) {
            xo_mc_chroma(pix, 8, c->fref_c[pl], c->stride_c, mx, my, c->bw / 2, c->bh / 2);
            cost += xo_pixel_cmp(c->sub_metric, c->in->i_pixel + 3, c->fenc_c[pl], 16, pix, 8);
            if (cost >= bcost) break;
        }
    }
    return cost;
}

static const int8_t hex_ring[8][2] = { { -1, -2 }, { -2, 0 }, { -1, 2 }, { 1, 2 }, { 2, 0 }, { 1, -2 }, { -1, -2 }, { -2, 0 } };
static const int8_t prev_of[8] = { 5, 0, 1, 2, 3, 4, 5, 0 }; /* (x-1)%6, me.c:46-47 */

typedef struct { int sad; int16_t mx, my; } cand_t;

/* extended input for sub-pel: subme level, mbcmp metric and the hpel planes */
typedef struct {
    int subme;      /* h->mb.i_subpel_refine */
    int mbcmp_satd; /* mbcmp == satd (user subme>1) */
    const xo_chroma *ch; /* non-NULL: h->mb.b_chroma_me */
} me_sub;

static void ctx_setup(me_ctx *c, const xo_geom *g, const uint8_t *fenc_plane, const uint8_t *const fref_planes[4], const xo_me_in *in,
                      const me_sub *sub, int stride, const int16_t *tab)
{
    c->g = g; c->in = in; c->stride = stride;
    c->bw = blk_w[in->i_pixel]; c->bh = blk_h[in->i_pixel];
    c->fpel_metric = in->fpel_satd ? XO_SATD : XO_SAD;
    c->sub_metric = sub->mbcmp_satd ? XO_SATD : XO_SAD;
    c->cmx = tab - in->mvp[0];
    c->cmy = tab - in->mvp[1];
    for (int y = 0; y < c->bh; y++)
        memcpy(c->fenc + 16 * y, fenc_plane + (in->by + y) * stride + in->bx, c->bw);
    for (int k = 0; k < 4; k++)
        c->fref[k] = fref_planes[k] ? fref_planes[k] + in->by * stride + in->bx : NULL;
    c->chroma_me = sub->ch && in->i_pixel <= 3; /* me.c:686 */
    if (c->chroma_me) {
        const xo_chroma *ch = sub->ch;
        const uint8_t *fe[2] = { ch->fenc_u, ch->fenc_v }, *fr[2] = { ch->fref_u, ch->fref_v };
        c->stride_c = ch->stride_c;
        for (int pl = 0; pl < 2; pl++) {
            for (int y = 0; y < c->bh / 2; y++)
                memcpy(c->fenc_c[pl] + 16 * y, fe[pl] + (in->by / 2 + y) * ch->stride_c + in->bx / 2, c->bw / 2);
            c->fref_c[pl] = fr[pl] + (in->by / 2) * ch->stride_c + in->bx / 2;
        }
    }
}

static void me_core(const xo_geom *g, const uint8_t *fenc_plane, const uint8_t *const fref_planes[4],
                    const uint16_t *integral, const xo_me_in *in, const me_sub *sub, xo_me_out *out)
{
    me_ctx cx, *c = &cx;
    const int stride = g->stride > 0 ? g->stride : -g->stride; /* negative: explicit stride (lowres) */
    int range = in->me_range;                                     /* UMH adapts it (me.c:354-399) */
    const int x_min = in->mv_min_fpel[0], y_min = in->mv_min_fpel[1];
    const int x_max = in->mv_max_fpel[0], y_max = in->mv_max_fpel[1];
    const int16_t *tab = cost_table(in->qp);
    int bmx, bmy, bcost, pmx, pmy;
    int bpred_mx = 0, bpred_my = 0, bpred_cost = XO_COST_MAX;

    ctx_setup(c, g, fenc_plane, fref_planes, in, sub, stride, tab);

    /* me.c:182-186 */
    bmx = clip3(in->mvp[0], x_min * 4, x_max * 4);
    bmy = clip3(in->mvp[1], y_min * 4, y_max * 4);
    pmx = (bmx + 2) >> 2;
    pmy = (bmy + 2) >> 2;
    bcost = XO_COST_MAX;

    if (sub->subme >= 3) { /* me.c:189-205: predictors compared at qpel precision with fpelcmp */
        uint32_t bmv = ((uint32_t)(uint16_t)bmx) | ((uint32_t)bmy << 16);
#define TRY_PRED(mx_, my_) do { int cst_ = qpel_cost(c, c->fpel_metric, (mx_), (my_)); \
        if (cst_ < bpred_cost) { bpred_cost = cst_; bpred_mx = (mx_); bpred_my = (my_); } } while (0)
        TRY_PRED(bmx, bmy);
        for (int i = 0; i < in->i_mvc; i++) {
            uint32_t v = ((uint32_t)(uint16_t)in->mvc[i][0]) | ((uint32_t)(uint16_t)in->mvc[i][1] << 16);
            if (v && (bmv - v)) {
                int mx = clip3(in->mvc[i][0], x_min * 4, x_max * 4);
                int my = clip3(in->mvc[i][1], y_min * 4, y_max * 4);
                TRY_PRED(mx, my);
            }
        }
        bmx = (bpred_mx + 2) >> 2;
        bmy = (bpred_my + 2) >> 2;
        TRY_FPEL(bmx, bmy);
    } else { /* me.c:207-227 */
        TRY_FPEL(pmx, pmy);
        bcost -= c->cmx[pmx << 2] + c->cmy[pmy << 2];
        for (int i = 0; i < in->i_mvc; i++) {
            int mx = (in->mvc[i][0] + 2) >> 2;
            int my = (in->mvc[i][1] + 2) >> 2;
            if ((mx | my) && ((mx - bmx) | (my - bmy))) {
                mx = clip3(mx, x_min, x_max);
                my = clip3(my, y_min, y_max);
                TRY_FPEL(mx, my);
            }
        }
    }
    TRY_FPEL(0, 0); /* me.c:229 */
    out->seed_mx = bmx; out->seed_my = bmy; out->seed_cost = bcost;

#define IN_RANGE(mx_, my_) ((mx_) >= x_min && (mx_) <= x_max && (my_) >= y_min && (my_) <= y_max)
    switch (in->me_method) {
    case XO_ME_DIA: { /* me.c:233-244 */
        int i = 0;
        do {
            int ox = bmx, oy = bmy;
            TRY_FPEL(ox, oy - 1); TRY_FPEL(ox, oy + 1); TRY_FPEL(ox - 1, oy); TRY_FPEL(ox + 1, oy);
            if (bmx == ox && bmy == oy) break;
            if (!IN_RANGE(bmx, bmy)) break;
        } while (++i < range);
        break;
    }
    case XO_ME_UMH: { /* me.c:306-447: uneven-cross multi-hexagon-grid search */
        static const int shift[7] = { 0, 1, 1, 2, 3, 3, 4 }; /* x264_pixel_size_shift */
#define SAD_THRESH(v) (bcost < ((v) >> shift[in->i_pixel]))
#define DIA1(cx_, cy_) do { omx = (cx_); omy = (cy_); TRY_FPEL(omx, omy - 1); TRY_FPEL(omx, omy + 1); TRY_FPEL(omx - 1, omy); TRY_FPEL(omx + 1, omy); } while (0)
#define X4(a0, b0, a1, b1, a2, b2, a3, b3) do { TRY_FPEL(omx + (a0), omy + (b0)); TRY_FPEL(omx + (a1), omy + (b1)); \
                                                 TRY_FPEL(omx + (a2), omy + (b2)); TRY_FPEL(omx + (a3), omy + (b3)); } while (0)
        /* CROSS (me.c:128-154): +-i along x for i = start, start+2, ... < x_max, then along y; the unchecked COST_MV_X4 form is
         * only used when every candidate is inside the limits, so one range-checked loop is the same thing */
#define CROSS(start_, xm_, ym_) do { \
            for (int i_ = (start_); i_ < (xm_); i_ += 2) { if (omx + i_ <= x_max) TRY_FPEL(omx + i_, omy); if (omx - i_ >= x_min) TRY_FPEL(omx - i_, omy); } \
            for (int i_ = (start_); i_ < (ym_); i_ += 2) { if (omy + i_ <= y_max) TRY_FPEL(omx, omy + i_); if (omy - i_ >= y_min) TRY_FPEL(omx, omy - i_); } } while (0)
        int omx, omy, ucost1, ucost2, cross_start = 1;
        ucost1 = bcost;
        DIA1(pmx, pmy);
        if (pmx | pmy) DIA1(0, 0);
        if (in->i_pixel == XO_4x4) goto me_hex2;
        ucost2 = bcost;
        if ((bmx | bmy) && ((bmx - pmx) | (bmy - pmy))) DIA1(bmx, bmy);
        if (bcost == ucost2) cross_start = 3;
        omx = bmx; omy = bmy;
        if (bcost == ucost2 && SAD_THRESH(2000)) { /* early termination */
            X4(0, -2, -1, -1, 1, -1, -2, 0);
            X4(2, 0, -1, 1, 1, 1, 0, 2);
            if (bcost == ucost1 && SAD_THRESH(500)) break;
            if (bcost == ucost2) {
                const int r = (range >> 1) | 1;
                CROSS(3, r, r);
                X4(-1, -2, 1, -2, -2, -1, 2, -1);
                X4(-2, 1, 2, 1, -1, 2, 1, 2);
                if (bcost == ucost2) break;
                cross_start = r + 2;
            }
        }
        if (in->i_mvc) { /* adaptive search range */
            static const int range_mul[4][4] = { { 3, 3, 4, 4 }, { 3, 4, 4, 4 }, { 4, 4, 4, 5 }, { 4, 4, 5, 6 } };
            int mvd, denom = 1;
            if (in->i_mvc == 1) {
                if (in->i_pixel == XO_16x16) mvd = 25;
                else mvd = abs(in->mvp[0] - in->mvc[0][0]) + abs(in->mvp[1] - in->mvc[0][1]);
            } else {
                denom = in->i_mvc - 1;
                mvd = 0;
                if (in->i_pixel != XO_16x16) { mvd = abs(in->mvp[0] - in->mvc[0][0]) + abs(in->mvp[1] - in->mvc[0][1]); denom++; }
                for (int k = 0; k < in->i_mvc - 1; k++) /* x264_predictor_difference, common.h:135-145 */
                    mvd += abs(in->mvc[k][0] - in->mvc[k + 1][0]) + abs(in->mvc[k][1] - in->mvc[k + 1][1]);
            }
            const int sad_ctx = SAD_THRESH(1000) ? 0 : SAD_THRESH(2000) ? 1 : SAD_THRESH(4000) ? 2 : 3;
            const int mvd_ctx = mvd < 10 * denom ? 0 : mvd < 20 * denom ? 1 : mvd < 40 * denom ? 2 : 3;
            range = range * range_mul[mvd_ctx][sad_ctx] / 4;
        }
        CROSS(cross_start, range, range / 2);
        X4(-2, -2, -2, 2, 2, -2, 2, 2);
        omx = bmx; omy = bmy;
        {
            static const int8_t hex4[16][2] = { { -4, 2 }, { -4, 1 }, { -4, 0 }, { -4, -1 }, { -4, -2 }, { 4, -2 }, { 4, -1 }, { 4, 0 }, { 4, 1 }, { 4, 2 },
                                                { 2, 3 }, { 0, 4 }, { -2, 3 }, { -2, -3 }, { 0, -4 }, { 2, -3 } };
            int i = 1;
            do { /* hexagon grid: the unchecked form is only used when all 16 points are inside the limits */
                for (int j = 0; j < 16; j++) {
                    const int mx = omx + hex4[j][0] * i, my = omy + hex4[j][1] * i;
                    if (IN_RANGE(mx, my)) TRY_FPEL(mx, my);
                }
            } while (++i <= range / 4);
        }
        if (bmy <= y_max) goto me_hex2;
        break;
#undef SAD_THRESH
#undef DIA1
#undef X4
#undef CROSS
    }
    case XO_ME_HEX: me_hex2: { /* me.c:246-305 (the de-duplicated form) */
        int costs[6], dir = -2;
        static const int8_t first[6][2] = { { -2, 0 }, { -1, 2 }, { 1, 2 }, { 2, 0 }, { 1, -2 }, { -1, -2 } };
        for (int k = 0; k < 6; k++)
            costs[k] = fpel_cost(c, bmx + first[k][0], bmy + first[k][1]);
        for (int k = 0; k < 6; k++)
            if (costs[k] < bcost) { bcost = costs[k]; dir = k; }
        if (dir != -2) {
            bmx += hex_ring[dir + 1][0];
            bmy += hex_ring[dir + 1][1];
            for (int i = 1; i < range / 2 && IN_RANGE(bmx, bmy); i++) {
                int odir = prev_of[dir + 1];
                for (int k = 0; k < 3; k++)
                    costs[k] = fpel_cost(c, bmx + hex_ring[odir + k][0], bmy + hex_ring[odir + k][1]);
                dir = -2;
                for (int k = 0; k < 3; k++)
                    if (costs[k] < bcost) { bcost = costs[k]; dir = odir - 1 + k; }
                if (dir == -2) break;
                bmx += hex_ring[dir + 1][0];
                bmy += hex_ring[dir + 1][1];
            }
        }
        { /* square refine, me.c:301-304 */
            int ox = bmx, oy = bmy;
            static const int8_t sq[8][2] = { { 0, -1 }, { 0, 1 }, { -1, 0 }, { 1, 0 }, { -1, -1 }, { -1, 1 }, { 1, -1 }, { 1, 1 } };
            for (int k = 0; k < 8; k++)
                TRY_FPEL(ox + sq[k][0], oy + sq[k][1]);
        }
        break;
    }
    case XO_ME_ESA:
    case XO_ME_TESA: { /* me.c:449-600 */
        const int min_x = bmx - range > x_min ? bmx - range : x_min;
        const int min_y = bmy - range > y_min ? bmy - range : y_min;
        const int max_x = bmx + range < x_max ? bmx + range : x_max;
        const int max_y = bmy + range < y_max ? bmy + range : y_max;
        const int width = (max_x - min_x + 3) & ~3;
        if (in->me_method == XO_ME_ESA) {
            /* ADS only removes candidates whose cost lower bound is >= bcost, so the survivors' scan in
             * raster order with strict '<' equals this loop. */
            for (int my = min_y; my <= max_y; my++)
                for (int mx = min_x; mx < min_x + width; mx++)
                    TRY_FPEL(mx, my);
        } else {
            /* me.c:469-489: DC of the fenc sub-blocks */
            int enc_dc[4];
            const int sad_size = in->i_pixel <= XO_8x8 ? XO_8x8 : XO_4x4;
            int delta = blk_w[sad_size];
            const uint16_t *sums_base = integral + in->by * stride + in->bx;
            static const uint8_t zero[8 * 16];
            uint16_t *row_cost = malloc((width + 16) * sizeof(uint16_t));
            int16_t *xs = malloc((width + 16) * sizeof(int16_t));
            cand_t *list = malloc(sizeof(cand_t) * (size_t)(width + 4) * (max_y - min_y + 2));
            int n = 0, limit;
            const int sad_thresh = range <= 16 ? 10 : range <= 24 ? 11 : 12;
            enc_dc[0] = xo_pixel_cmp(XO_SAD, sad_size, zero, 16, c->fenc, 16);
            enc_dc[1] = xo_pixel_cmp(XO_SAD, sad_size, zero, 16, c->fenc + delta, 16);
            enc_dc[2] = xo_pixel_cmp(XO_SAD, sad_size, zero, 16, c->fenc + delta * 16, 16);
            enc_dc[3] = xo_pixel_cmp(XO_SAD, sad_size, zero, 16, c->fenc + delta + delta * 16, 16);
            if (delta == 4)
                sums_base += stride * (g->lines + XO_PADV * 2);
            if (in->i_pixel == XO_16x16 || in->i_pixel == XO_8x16 || in->i_pixel == XO_4x8)
                delta *= stride;
            if (in->i_pixel == XO_8x16 || in->i_pixel == XO_4x8)
                enc_dc[1] = enc_dc[2];
            for (int i = 0; i < width; i++)
                row_cost[i] = (uint16_t)c->cmx[(min_x + i) << 2];

            int bsad = xo_pixel_cmp(XO_SAD, in->i_pixel, c->fenc, 16, c->fref[0] + bmy * stride + bmx, stride) +
                       c->cmx[bmx << 2] + c->cmy[bmy << 2];
            for (int my = min_y; my <= max_y; my++) {
                int ycost = c->cmy[my << 2];
                if (bsad <= ycost)
                    continue;
                bsad -= ycost;
                int xn = xo_pixel_ads(in->i_pixel, enc_dc, sums_base + min_x + my * stride, delta, row_cost, xs,
                                      width, bsad * 17 / 16);
                for (int i = 0; i < xn; i++) {
                    int mx = min_x + xs[i];
                    /* NB me.c:518,531: the reference indexes cost_fpel_mvx by xs[i] WITHOUT min_x here (while ADS
                     * got cost_fpel_mvx+min_x, :505): the SAD-stage x cost is that of mx = xs[i].  Reproduced. */
                    int sad = xo_pixel_cmp(XO_SAD, in->i_pixel, c->fenc, 16, c->fref[0] + mx + my * stride, stride) +
                              (uint16_t)c->cmx[xs[i] << 2];
                    if (sad < bsad * sad_thresh >> 3) {
                        if (sad < bsad) bsad = sad;
                        list[n].sad = sad + ycost;
                        list[n].mx = mx;
                        list[n].my = my;
                        n++;
                    }
                }
                bsad += ycost;
            }
            limit = range / 2;
            if (n > limit * 2) { /* me.c:543-558 */
                int i, j;
                bsad = bsad * (sad_thresh + 8) >> 4;
                for (i = 0; i < n && list[i].sad <= bsad; i++);
                for (j = i; j < n; j++)
                    if (list[j].sad <= bsad)
                        list[i++] = list[j];
                n = i;
            }
            if (n > limit) { /* me.c:559-576: partial selection sort, first index wins ties */
                for (int i = 0; i < limit; i++) {
                    int bj = i, bs = list[bj].sad;
                    for (int j = i + 1; j < n; j++)
                        if (list[j].sad < bs) { bs = list[j].sad; bj = j; }
                    if (bj > i) { cand_t t = list[i]; list[i] = list[bj]; list[bj] = t; }
                }
                n = limit;
            }
            for (int i = 0; i < n; i++)
                TRY_FPEL(list[i].mx, list[i].my);
            free(row_cost); free(xs); free(list);
        }
        break;
    }
    default:
        break;
    }
    out->bmx = bmx; out->bmy = bmy; out->bcost = bcost;

    /* me.c:603-620 */
    int mvx, mvy, cost;
    if (bpred_cost < bcost) { mvx = bpred_mx; mvy = bpred_my; cost = bpred_cost; }
    else { mvx = bmx << 2; mvy = bmy << 2; cost = bcost; }
    int cost_mv = c->cmx[mvx] + c->cmy[mvy];
    if (bmx == pmx && bmy == pmy && sub->subme < 3)
        cost += cost_mv;

    if (sub->subme >= 2) {
        /* refine_subpel(h, m, hpel, qpel, NULL, 0): me.c:680-778 */
        static const int8_t iters[10][4] = { { 0, 0, 0, 0 }, { 1, 1, 0, 0 }, { 0, 1, 1, 0 }, { 0, 2, 1, 0 }, { 0, 2, 1, 1 },
                                             { 0, 2, 1, 2 }, { 0, 0, 2, 2 }, { 0, 0, 2, 2 }, { 0, 0, 4, 10 }, { 0, 0, 4, 10 } };
        int hpel_iters = iters[sub->subme][2], qpel_iters = iters[sub->subme][3];
        int sbmx = mvx, sbmy = mvy, sbcost = cost, odir = -1, bdir;
        const int spel_y_max = in->mv_max_spel[1];
        if (hpel_iters && sub->subme < 3) {
            int mx = clip3(in->mvp[0], in->mv_min_spel[0], in->mv_max_spel[0]);
            int my = clip3(in->mvp[1], in->mv_min_spel[1], in->mv_max_spel[1]);
            if ((mx - sbmx) | (my - sbmy)) {
                int cst = qpel_cost(c, c->fpel_metric, mx, my);
                if (cst < sbcost) { sbcost = cst; sbmx = mx; sbmy = my; }
            }
        }
        for (int i = hpel_iters; i > 0; i--) { /* me.c:710-727: fpelcmp on the 4 half-pel neighbours */
            int ox = sbmx, oy = sbmy, cst;
            cst = qpel_cost(c, c->fpel_metric, ox, oy - 2); if (cst < sbcost) { sbcost = cst; sbmy = oy - 2; }
            cst = qpel_cost(c, c->fpel_metric, ox, oy + 2); if (cst < sbcost) { sbcost = cst; sbmy = oy + 2; }
            cst = qpel_cost(c, c->fpel_metric, ox - 2, oy); if (cst < sbcost) { sbcost = cst; sbmx = ox - 2; sbmy = oy; }
            cst = qpel_cost(c, c->fpel_metric, ox + 2, oy); if (cst < sbcost) { sbcost = cst; sbmx = ox + 2; sbmy = oy; }
            if (sbmx == ox && sbmy == oy) break;
        }
        /* !b_refine_qpel: me.c:729-736 */
        if (sbmy > spel_y_max) sbmy = spel_y_max;
        sbcost = satd_cost(c, sbmx, sbmy, XO_COST_MAX);
        bdir = -1;
        for (int i = qpel_iters; i > 0; i--) { /* me.c:755-767 */
            int ox = sbmx, oy = sbmy;
            static const int8_t d[4][2] = { { 0, -1 }, { 0, 1 }, { -1, 0 }, { 1, 0 } };
            odir = bdir;
            for (int k = 0; k < 4; k++)
                if ((k ^ 1) != odir) {
                    int cst = satd_cost(c, ox + d[k][0], oy + d[k][1], sbcost);
                    if (cst < sbcost) { sbcost = cst; sbmx = ox + d[k][0]; sbmy = oy + d[k][1]; bdir = k; }
                }
            if (sbmx == ox && sbmy == oy) break;
        }
        if (sbmy > spel_y_max) { /* me.c:770-775 */
            sbmy = spel_y_max;
            sbcost = satd_cost(c, sbmx, sbmy, XO_COST_MAX);
        }
        mvx = sbmx; mvy = sbmy; cost = sbcost;
        cost_mv = c->cmx[mvx] + c->cmy[mvy];
    } else if (mvy > in->mv_max_spel[1]) {
        mvy = in->mv_max_spel[1]; /* me.c:629-630 */
    }
    out->mv[0] = mvx; out->mv[1] = mvy; out->cost = cost; out->cost_mv = cost_mv;
}

/* x264_me_refine_qpel (me.c:633-643): refine_subpel( h, m, subpel_iterations[subme][0], [1], NULL, 1 ) from m->mv / m->cost.
 * The caller has already applied the "m->cost -= m->i_ref_cost" of :638-639 if it wants it. */
void xo_me_refine_qpel(const xo_geom *g, const uint8_t *fenc_plane, const uint8_t *const fref_planes[4], const xo_chroma *ch, const xo_me_in *in,
                       int subme, int mbcmp_satd, const int16_t mv_in[2], int cost_in, xo_me_out *out)
{
    static const int8_t iters[10][2] = { { 0, 0 }, { 1, 1 }, { 0, 1 }, { 0, 2 }, { 0, 2 }, { 0, 2 }, { 0, 0 }, { 0, 0 }, { 0, 0 }, { 0, 0 } };
    me_ctx cx, *c = &cx;
    me_sub sub = { subme, mbcmp_satd, ch };
    const int stride = g->stride > 0 ? g->stride : -g->stride;
    ctx_setup(c, g, fenc_plane, fref_planes, in, &sub, stride, cost_table(in->qp));
    int bmx = mv_in[0], bmy = mv_in[1], bcost = cost_in, bdir = -1;
    const int hpel_iters = iters[subme][0], qpel_iters = iters[subme][1], spel_y_max = in->mv_max_spel[1];
    if (hpel_iters && subme < 3) { /* me.c:699-705 */
        const int mx = clip3(in->mvp[0], in->mv_min_spel[0], in->mv_max_spel[0]), my = clip3(in->mvp[1], in->mv_min_spel[1], in->mv_max_spel[1]);
        if ((mx - bmx) | (my - bmy)) { const int cst = qpel_cost(c, c->fpel_metric, mx, my); if (cst < bcost) { bcost = cst; bmx = mx; bmy = my; } }
    }
    for (int i = hpel_iters; i > 0; i--) { /* me.c:708-727 */
        const int ox = bmx, oy = bmy;
        int cst;
        cst = qpel_cost(c, c->fpel_metric, ox, oy - 2); if (cst < bcost) { bcost = cst; bmy = oy - 2; }
        cst = qpel_cost(c, c->fpel_metric, ox, oy + 2); if (cst < bcost) { bcost = cst; bmy = oy + 2; }
        cst = qpel_cost(c, c->fpel_metric, ox - 2, oy); if (cst < bcost) { bcost = cst; bmx = ox - 2; bmy = oy; }
        cst = qpel_cost(c, c->fpel_metric, ox + 2, oy); if (cst < bcost) { bcost = cst; bmx = ox + 2; bmy = oy; }
        if (bmx == ox && bmy == oy) break;
    }
    for (int i = qpel_iters; i > 0; i--) { /* me.c:755-767 with b_refine_qpel: all four neighbours every time */
        static const int8_t d[4][2] = { { 0, -1 }, { 0, 1 }, { -1, 0 }, { 1, 0 } };
        const int ox = bmx, oy = bmy;
        for (int k = 0; k < 4; k++) {
            const int cst = satd_cost(c, ox + d[k][0], oy + d[k][1], bcost);
            if (cst < bcost) { bcost = cst; bmx = ox + d[k][0]; bmy = oy + d[k][1]; bdir = k; }
        }
        if (bmx == ox && bmy == oy) break;
    }
    (void)bdir;
    if (bmy > spel_y_max) { bmy = spel_y_max; bcost = satd_cost(c, bmx, bmy, XO_COST_MAX); } /* me.c:770-775 */
    memset(out, 0, sizeof(*out));
    out->mv[0] = bmx; out->mv[1] = bmy; out->cost = bcost; out->cost_mv = c->cmx[bmx] + c->cmy[bmy];
}

void xo_me_search_fpel(const xo_geom *g, const uint8_t *fenc_plane, const uint8_t *fref_plane,
                       const uint16_t *integral, const xo_me_in *in, xo_me_out *out)
{
    const uint8_t *planes[4] = { fref_plane, NULL, NULL, NULL };
    me_sub sub = { 1, 0, NULL };
    me_core(g, fenc_plane, planes, integral, in, &sub, out);
}

/* full search incl. sub-pel refinement; subme = h->mb.i_subpel_refine, mbcmp_satd = (user subme > 1) */
void xo_me_search_subpel(const xo_geom *g, const uint8_t *fenc_plane, const uint8_t *const fref_planes[4],
                         const uint16_t *integral, const xo_me_in *in, int subme, int mbcmp_satd, xo_me_out *out)
{
    me_sub sub = { subme, mbcmp_satd, NULL };
    me_core(g, fenc_plane, fref_planes, integral, in, &sub, out);
}

/* the same with chroma in the sub-pel cost (P-slices, subme >= 5, --chroma-me: S/encoder/analyse.c b_chroma_me) */
void xo_me_search_subpel_chroma(const xo_geom *g, const uint8_t *fenc_plane, const uint8_t *const fref_planes[4], const uint16_t *integral,
                                const xo_chroma *ch, const xo_me_in *in, int subme, int mbcmp_satd, xo_me_out *out)
{
    me_sub sub = { subme, mbcmp_satd, ch };
    me_core(g, fenc_plane, fref_planes, integral, in, &sub, out);
}

void xo_me_search_fpel_batch(const xo_geom *g, const uint8_t *fenc_plane, const uint8_t *fref_plane,
                             const uint16_t *integral, const xo_me_in *in, int n, xo_me_out *out)
{
    for (int i = 0; i < n; i++)
        xo_me_search_fpel(g, fenc_plane, fref_plane, integral, in + i, out + i);
}

/* same search on planes of an explicit stride (the half-resolution lookahead planes) */
void xo_me_search_subpel_strided(int stride, int lines_unused, const uint8_t *fenc_plane, const uint8_t *const fref_planes[4],
                                 const xo_me_in *in, int subme, int mbcmp_satd, xo_me_out *out)
{
    xo_geom g;
    memset(&g, 0, sizeof(g));
    g.stride = stride;
    (void)lines_unused;
    me_sub sub = { subme, mbcmp_satd, NULL };
    me_core(&g, fenc_plane, fref_planes, NULL, in, &sub, out);
}


/* x264_me_refine_bidir( h, m0, m1, i_weight, 0, 0, 0 ), me.c:843-927.  The 32 candidate pairs of a pass in the reference's order
 * (CHECK_BIDIR8 / CHECK_BIDIR2 expanded); `visited` is the reference's aliasing 8x8x8x8-bit map (indices & 7), reproduced as is.
 * The y cost tables are offset by mvp[1] clipped to the X limits — the reference does that (me.c:855,857) and so does this. */
int xo_me_refine_bidir_satd(const xo_geom *g, const uint8_t *fenc_plane, const uint8_t *const fref0[4], const uint8_t *const fref1[4],
                            const xo_me_in *in, const int16_t mvp0[2], const int16_t mvp1[2], int weight, int mbcmp_satd, int16_t mv0[2], int16_t mv1[2])
{
    static const int8_t cand[32][4] = {
        { 0, 0, 0, 1 }, { 0, 0, 0, -1 }, { 0, 0, 1, 0 }, { 0, 0, -1, 0 }, { 0, 1, 0, 0 }, { 0, -1, 0, 0 }, { 1, 0, 0, 0 }, { -1, 0, 0, 0 },
        { 0, 0, 1, 1 }, { 0, 0, -1, -1 }, { 0, 1, 1, 0 }, { 0, -1, -1, 0 }, { 1, 1, 0, 0 }, { -1, -1, 0, 0 }, { 1, 0, 0, 1 }, { -1, 0, 0, -1 },
        { 0, 1, 0, 1 }, { 0, -1, 0, -1 }, { 1, 0, 1, 0 }, { -1, 0, -1, 0 },
        { 0, 0, -1, 1 }, { 0, 0, 1, -1 }, { 0, -1, 1, 0 }, { 0, 1, -1, 0 }, { -1, 1, 0, 0 }, { 1, -1, 0, 0 }, { 1, 0, 0, -1 }, { -1, 0, 0, 1 },
        { 0, -1, 0, 1 }, { 0, 1, 0, -1 }, { -1, 0, 1, 0 }, { 1, 0, -1, 0 } };
    const int stride = g->stride, bw = blk_w[in->i_pixel], bh = blk_h[in->i_pixel], metric = mbcmp_satd ? XO_SATD : XO_SAD;
    const int16_t *tab = cost_table(in->qp);
    const int lo = in->mv_min_spel[0], hi = in->mv_max_spel[0];
    const int16_t *c0x = tab - clip3(mvp0[0], lo, hi), *c0y = tab - clip3(mvp0[1], lo, hi);
    const int16_t *c1x = tab - clip3(mvp1[0], lo, hi), *c1y = tab - clip3(mvp1[1], lo, hi);
    uint8_t fenc[16 * 16], visited[8][8][8];
    const uint8_t *p0[4], *p1[4];
    int bm0x = mv0[0], bm0y = mv0[1], bm1x = mv1[0], bm1y = mv1[1], om0x = bm0x, om0y = bm0y, om1x = bm1x, om1y = bm1y, bcost = XO_COST_MAX;
    if (bm0y > in->mv_max_spel[1] - 8 || bm1y > in->mv_max_spel[1] - 8) return bcost;
    for (int y = 0; y < bh; y++) memcpy(fenc + 16 * y, fenc_plane + (in->by + y) * stride + in->bx, bw);
    for (int k = 0; k < 4; k++) { p0[k] = fref0[k] + in->by * stride + in->bx; p1[k] = fref1[k] + in->by * stride + in->bx; }
    memset(visited, 0, sizeof(visited));
#define TRY_PAIR(m0x, m0y, m1x, m1y) do { \
        if (pass == 0 || !(visited[(m0x) & 7][(m0y) & 7][(m1x) & 7] & (1 << ((m1y) & 7)))) { \
            uint8_t a[16 * 16], b[16 * 16], pix[16 * 16]; \
            visited[(m0x) & 7][(m0y) & 7][(m1x) & 7] |= (1 << ((m1y) & 7)); \
            xo_mc_luma(a, 16, p0, stride, (m0x), (m0y), bw, bh); \
            xo_mc_luma(b, 16, p1, stride, (m1x), (m1y), bw, bh); \
            xo_pixel_avg(in->i_pixel, pix, 16, a, 16, b, 16, weight); \
            const int cost = xo_pixel_cmp(metric, in->i_pixel, fenc, 16, pix, 16) + c0x[(m0x)] + c0y[(m0y)] + c1x[(m1x)] + c1y[(m1y)]; \
            if (cost < bcost) { bcost = cost; bm0x = (m0x); bm0y = (m0y); bm1x = (m1x); bm1y = (m1y); } \
        } } while (0)
    int pass = 0;
    TRY_PAIR(om0x, om0y, om1x, om1y);
    for (pass = 0; pass < 8; pass++) {
        for (int k = 0; k < 32; k++) TRY_PAIR(om0x + cand[k][0], om0y + cand[k][1], om1x + cand[k][2], om1y + cand[k][3]);
        if (om0x == bm0x && om0y == bm0y && om1x == bm1x && om1y == bm1y) break;
        om0x = bm0x; om0y = bm0y; om1x = bm1x; om1y = bm1y;
    }
#undef TRY_PAIR
    mv0[0] = bm0x; mv0[1] = bm0y; mv1[0] = bm1x; mv1[1] = bm1y;
    return bcost;
}
