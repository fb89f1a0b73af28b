/* xo_lookahead.c — ORACLE (test infrastructure only): the half-resolution lookahead cost of one (p0,p1,b) frame
 * triple, i.e. x264_slicetype_frame_cost + x264_slicetype_mb_cost (S/encoder/slicetype.c:43-355): the default form
 * (interior blocks, :318-330) and the VBV form (every block, per-row sums, AQ-weighted costs, :300-316), with the 8x8 intra predictors it uses (S/common/predict.c:234-336, :499-748). */
#include <stdlib.h>
#include <string.h>
#include "xo.h"

void xo_me_search_subpel_strided(int stride, int lines_unused, const uint8_t *fenc_plane, const uint8_t *const fref_planes[4],
                                 const xo_me_in *in, int subme, int mbcmp_satd, xo_me_out *out);

static inline uint8_t clip_u8(int x) { return x < 0 ? 0 : x > 255 ? 255 : x; }
static inline int clip3(int v, int lo, int hi) { return v < lo ? lo : v > hi ? hi : v; }
static inline int median3(int a, int b, int c) { return a > b ? (b > c ? b : (a > c ? c : a)) : (a > c ? a : (b > c ? c : b)); }
#define F1(a, b) (((a) + (b) + 1) >> 1)
#define F2(a, b, c) (((a) + 2 * (b) + (c) + 2) >> 2)

/* the ten predictions of slicetype.c:205-229.  top[-1..15] = row above from x-1, left[0..7]; out: pred[8][8] */
static void intra_pred_8x8(int mode, const uint8_t *top /* top[-1] valid */, const uint8_t *left, uint8_t p[8][8])
{
    if (mode < 4) { /* predict_8x8c_{dc,h,v,p}, predict.c:234-336 */
        if (mode == 0) {
            int s0 = 0, s1 = 0, s2 = 0, s3 = 0;
            for (int i = 0; i < 4; i++) { s0 += top[i]; s1 += top[i + 4]; s2 += left[i]; s3 += left[i + 4]; }
            int dc[4] = { (s0 + s2 + 4) >> 3, (s1 + 2) >> 2, (s3 + 2) >> 2, (s1 + s3 + 4) >> 3 };
            for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) p[y][x] = dc[(y >> 2) * 2 + (x >> 2)];
        } else if (mode == 1) {
            for (int y = 0; y < 8; y++) memset(p[y], left[y], 8);
        } else if (mode == 2) {
            for (int y = 0; y < 8; y++) memcpy(p[y], top, 8);
        } else {
            int H = 0, V = 0;
            for (int i = 0; i < 4; i++) {
                H += (i + 1) * (top[4 + i] - top[2 - i]);
                V += (i + 1) * (left[4 + i] - (2 - i >= 0 ? left[2 - i] : top[-1]));
            }
            int a = 16 * (left[7] + top[7]), b = (17 * H + 16) >> 5, c = (17 * V + 16) >> 5, i00 = a - 3 * b - 3 * c + 16;
            for (int y = 0; y < 8; y++, i00 += c) { int pix = i00; for (int x = 0; x < 8; x++, pix += b) p[y][x] = clip_u8(pix >> 5); }
        }
        return;
    }
    /* x264_predict_8x8_filter with all neighbours (predict.c:499-540) */
    int lt = F2(top[0], top[-1], left[0]);
    int l[8], t[16];
    l[0] = F2(top[-1], left[0], left[1]);
    for (int y = 1; y < 7; y++) l[y] = F2(left[y - 1], left[y], left[y + 1]);
    l[7] = (left[6] + 3 * left[7] + 2) >> 2;
    t[0] = F2(top[-1], top[0], top[1]);
    for (int x = 1; x < 15; x++) t[x] = F2(top[x - 1], top[x], top[x + 1]);
    t[15] = (top[14] + 3 * top[15] + 2) >> 2;
#define P(x, y) p[y][x]
    switch (mode) { /* mode numbers: I_PRED_8x8_DDL=3 .. HU=8 mapped to 4..9 here */
    case 4: /* ddl, predict.c:605-623 */
        for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) { int k = x + y; P(x, y) = k == 14 ? F2(t[14], t[15], t[15]) : F2(t[k], t[k + 1], t[k + 2]); }
        break;
    case 5: { /* ddr, :624-645: diagonal d = x - y */
        int e[17]; /* e[0..7] = l7..l0, e[8] = lt, e[9..16] = t0..t7 */
        for (int i = 0; i < 8; i++) { e[i] = l[7 - i]; e[9 + i] = t[i]; }
        e[8] = lt;
        for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) { int k = 8 + x - y; P(x, y) = F2(e[k - 1], e[k], e[k + 1]); }
        break; }
    case 6: { /* vr, :646-674 */
        int e[17];
        for (int i = 0; i < 8; i++) { e[i] = l[7 - i]; e[9 + i] = t[i]; }
        e[8] = lt;
        for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) {
            int z = 2 * x - y;
            if (z >= 0) { int k = 8 + x - (y >> 1); P(x, y) = (z & 1) ? F2(e[k - 1], e[k], e[k + 1]) : F1(e[k], e[k + 1]); }
            else if (z == -1) P(x, y) = F2(l[0], lt, t[0]);
            else { int k = y - 2 * x - 1; /* z = -2 -> l1,l0,lt ... */ P(x, y) = F2(k >= 1 ? l[k] : l[k], k - 1 >= 0 ? l[k - 1] : lt, k - 2 >= 0 ? l[k - 2] : lt); }
        }
        break; }
    case 7: { /* hd, :675-703 */
        int e[17];
        for (int i = 0; i < 8; i++) { e[i] = l[7 - i]; e[9 + i] = t[i]; }
        e[8] = lt;
        for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) {
            int z = 2 * y - x;
            if (z >= 0) { int k = 8 - y + (x >> 1); P(x, y) = (z & 1) ? F2(e[k - 1], e[k], e[k + 1]) : F1(e[k - 1], e[k]); }
            else if (z == -1) P(x, y) = F2(l[0], lt, t[0]);
            else { int k = x - 2 * y - 1; P(x, y) = F2(t[k], t[k - 1], k - 2 >= 0 ? t[k - 2] : lt); }
        }
        break; }
    case 8: /* vl, :704-730 */
        for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) { int k = x + (y >> 1); P(x, y) = (y & 1) ? F2(t[k], t[k + 1], t[k + 2]) : F1(t[k], t[k + 1]); }
        break;
    default: /* hu, :731-748 */
        for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) {
            int z = x + 2 * y;
            if (z > 13) P(x, y) = l[7];
            else if (z == 13) P(x, y) = F2(l[6], l[7], l[7]);
            else { int k = y + (x >> 1); P(x, y) = (z & 1) ? F2(l[k], l[k + 1], l[k + 2]) : F1(l[k], l[k + 1]); }
        }
    }
#undef P
}

/* slicetype.c:192-233: min cost over the ten predictions + penalty 5 */
int xo_lowres_intra_cost(const uint8_t *l0, int stride, int bx, int by, int mbcmp_satd)
{
    const uint8_t *src = l0 + by * stride + bx;
    uint8_t top[17], left[8], fenc[8 * 16], pred[8][8], pp[8 * 16];
    memcpy(top, src - stride - 1, 17);
    for (int i = 0; i < 8; i++) { left[i] = src[i * stride - 1]; memcpy(fenc + 16 * i, src + i * stride, 8); }
    int best = 1 << 30;
    for (int m = 0; m < 10; m++) {
        intra_pred_8x8(m, top + 1, left, pred);
        for (int y = 0; y < 8; y++) memcpy(pp + 16 * y, pred[y], 8);
        int c = xo_pixel_cmp(mbcmp_satd ? XO_SATD : XO_SAD, XO_8x8, pp, 16, fenc, 16);
        if (c < best) best = c;
    }
    return best + 5;
}
void xo_lowres_intra_pred(int mode, const uint8_t *l0, int stride, int bx, int by, uint8_t out[64])
{
    const uint8_t *src = l0 + by * stride + bx;
    uint8_t top[17], left[8], pred[8][8];
    memcpy(top, src - stride - 1, 17);
    for (int i = 0; i < 8; i++) left[i] = src[i * stride - 1];
    intra_pred_8x8(mode, top + 1, left, pred);
    memcpy(out, pred, 64);
}

typedef struct {
    const xo_geom *g;
    const xo_lowres_in *in;
    const uint8_t *const *fenc, *const *fref[2];
    int stride;
} la_ctx;

/* TRY_BIDIR, slicetype.c:96-112 */
static int bidir_cost(const la_ctx *c, int bx, int by, const int *mv0, const int *mv1, int weight, int penalty)
{
    uint8_t a[8 * 16], b[8 * 16], fe[8 * 16];
    const uint8_t *p0[4], *p1[4];
    for (int k = 0; k < 4; k++) { p0[k] = c->fref[0][k] + by * c->stride + bx; p1[k] = c->fref[1][k] + by * c->stride + bx; }
    xo_mc_luma(a, 16, p0, c->stride, mv0[0], mv0[1], 8, 8);
    xo_mc_luma(b, 16, p1, c->stride, mv1[0], mv1[1], 8, 8);
    for (int y = 0; y < 8; y++) {
        memcpy(fe + 16 * y, c->fenc[0] + (by + y) * c->stride + bx, 8);
        for (int x = 0; x < 8; x++)
            a[16 * y + x] = weight == 32 ? (a[16 * y + x] + b[16 * y + x] + 1) >> 1
                                         : clip_u8((a[16 * y + x] * weight + b[16 * y + x] * (64 - weight) + 32) >> 6); /* mc.c:64-97 */
    }
    return penalty + xo_pixel_cmp(c->in->mbcmp_satd ? XO_SATD : XO_SAD, XO_8x8, fe, 16, a, 16);
}

/* b_vbv: h->param.rc.i_vbv_buffer_size != 0 (slicetype.c:300-316): every block is evaluated, row_satd[mb_height] receives the per-row
 * sums of the (AQ-weighted when inv_qscale != NULL, i.e. rc.i_aq_mode) block costs, out->score_aq the weighted interior sum */
static void lowres_frame_cost(const xo_geom *g, const xo_lowres_in *in, const uint8_t *const fenc[4], const uint8_t *const fref0[4],
                              const uint8_t *const fref1[4], int16_t (*mvs0)[2], int *costs0, int16_t (*mvs1)[2], int *costs1,
                              const int16_t (*ref1_mvs)[2], uint16_t *intra_cost, xo_lowres_out *out, int b_vbv, const uint16_t *inv_qscale,
                              int *row_satd)
{
    const int W = g->mb_width, H = g->mb_height, stride = g->stride_lowres;
    const int b_bidir = in->b < in->p1;
    int dsf = 128;
    if (in->p1 != in->p0) dsf = (((in->b - in->p0) << 8) + ((in->p1 - in->p0) >> 1)) / (in->p1 - in->p0);
    const int weight = in->b_weighted_bipred ? 64 - (dsf >> 2) : 32;
    la_ctx c = { g, in, fenc, { fref0, fref1 }, stride };
    int16_t (*mvs[2])[2] = { mvs0, mvs1 };
    int *costs[2] = { costs0, costs1 };
    memset(out, 0, sizeof(*out));
    const int small = W <= 2 || H <= 2;
    const int all = small || b_vbv;
    if (b_vbv && !small) memset(row_satd, 0, H * sizeof(int));
    for (int my = all ? H - 1 : H - 2; my >= (all ? 0 : 1); my--)
        for (int mx = all ? W - 1 : W - 2; mx >= (all ? 0 : 1); mx--) {
            const int xy = mx + my * W, bx = 8 * mx, by = 8 * my;
            int bcost = XO_COST_MAX;
            if (in->p0 != in->p1 || in->p0 != in->b) {
                xo_me_in mi;
                memset(&mi, 0, sizeof(mi));
                mi.mv_min_fpel[0] = -8 * mx - 4; mi.mv_max_fpel[0] = 8 * (W - mx - 1) + 4;
                mi.mv_min_fpel[1] = -8 * my - 4; mi.mv_max_fpel[1] = 8 * (H - my - 1) + 4;
                for (int k = 0; k < 2; k++) { mi.mv_min_spel[k] = 4 * (mi.mv_min_fpel[k] - 8); mi.mv_max_spel[k] = 4 * (mi.mv_max_fpel[k] + 8); }
                int m_mv[2][2] = { { 0, 0 }, { 0, 0 } }, m_cost[2] = { 0, 0 };
                if (b_bidir) { /* slicetype.c:121-142 */
                    const int16_t *mvr = ref1_mvs[xy];
                    int dmv[2][2], zero[2] = { 0, 0 };
                    for (int k = 0; k < 2; k++) {
                        dmv[0][k] = (mvr[k] * dsf + 128) >> 8;
                        dmv[1][k] = dmv[0][k] - mvr[k];
                        dmv[0][k] = clip3(dmv[0][k], mi.mv_min_spel[k], mi.mv_max_spel[k]);
                        dmv[1][k] = clip3(dmv[1][k], mi.mv_min_spel[k], mi.mv_max_spel[k]);
                    }
                    int v = bidir_cost(&c, bx, by, dmv[0], dmv[1], weight, 0);
                    if (bcost > v) bcost = v;
                    if (dmv[0][0] | dmv[0][1] | dmv[1][0] | dmv[1][1]) { v = bidir_cost(&c, bx, by, zero, zero, weight, 0); if (bcost > v) bcost = v; }
                }
                for (int l = 0; l < 1 + b_bidir; l++) {
                    if (in->do_search[l]) {
                        int16_t mvc[4][2] = { { 0 } };
                        int n = 0;
                        int16_t (*fm)[2] = mvs[l] + xy;
#define MVC(p) do { mvc[n][0] = (p)[0]; mvc[n][1] = (p)[1]; n++; } while (0)
                        if (mx < W - 1) MVC(fm[1]);
                        if (my < H - 1) {
                            MVC(fm[W]);
                            if (mx > 0) MVC(fm[W - 1]);
                            if (mx < W - 1) MVC(fm[W + 1]);
                        }
#undef MVC
                        mi.me_method = in->me_method < XO_ME_HEX ? in->me_method : XO_ME_HEX;
                        mi.me_range = in->me_range; mi.qp = 12; mi.i_pixel = XO_8x8; mi.bx = bx; mi.by = by;
                        mi.fpel_satd = 0; /* fpelcmp is SATD only for TESA, which the lookahead never uses... but mbcmp_init set it once: */
                        mi.fpel_satd = in->fpel_satd;
                        mi.mvp[0] = median3(mvc[0][0], mvc[1][0], mvc[2][0]); mi.mvp[1] = median3(mvc[0][1], mvc[1][1], mvc[2][1]);
                        mi.i_mvc = n;
                        memcpy(mi.mvc, mvc, sizeof(mvc));
                        xo_me_out mo;
                        xo_me_search_subpel_strided(stride, 0, fenc[0], l ? fref1 : fref0, &mi, 4, in->mbcmp_satd, &mo);
                        int cost = mo.cost - 2;
                        if (mo.mv[0] | mo.mv[1]) cost += 5;
                        fm[0][0] = mo.mv[0]; fm[0][1] = mo.mv[1];
                        costs[l][xy] = cost;
                    }
                    m_mv[l][0] = mvs[l][xy][0]; m_mv[l][1] = mvs[l][xy][1]; m_cost[l] = costs[l][xy];
                    if (m_cost[l] < bcost) bcost = m_cost[l];
                }
                if (b_bidir && (m_mv[0][0] | m_mv[0][1] | m_mv[1][0] | m_mv[1][1])) {
                    int v = bidir_cost(&c, bx, by, m_mv[0], m_mv[1], weight, 5);
                    if (bcost > v) bcost = v;
                }
            }
            if (!b_bidir) {
                int icost;
                if (!in->b_intra_calculated) { icost = xo_lowres_intra_cost(fenc[0], stride, bx, by, in->mbcmp_satd); intra_cost[xy] = icost; }
                else icost = intra_cost[xy];
                int b_intra = icost < bcost;
                if (b_intra) bcost = icost;
                if (mx > 0 && mx < W - 1 && my > 0 && my < H - 1) { out->intra_mbs += b_intra; out->intra_cost_sum += icost; }
            }
            if (b_vbv && !small) { /* slicetype.c:305-316 */
                const int aq = inv_qscale ? (bcost * inv_qscale[xy] + 128) >> 8 : bcost;
                row_satd[my] += aq;
                if (mx > 0 && mx < W - 1 && my > 0 && my < H - 1) { out->score += bcost; out->score_aq += aq; }
            } else {
                out->score += bcost;
                out->score_aq += (!small && inv_qscale) ? (bcost * inv_qscale[xy] + 128) >> 8 : small ? 0 : bcost; /* :318-330; tiny frames leave it 0 (:293-298) */
            }
        }
}

void xo_lowres_frame_cost(const xo_geom *g, const xo_lowres_in *in, const uint8_t *const fenc[4], const uint8_t *const fref0[4],
                          const uint8_t *const fref1[4], int16_t (*mvs0)[2], int *costs0, int16_t (*mvs1)[2], int *costs1,
                          const int16_t (*ref1_mvs)[2], uint16_t *intra_cost, xo_lowres_out *out)
{
    lowres_frame_cost(g, in, fenc, fref0, fref1, mvs0, costs0, mvs1, costs1, ref1_mvs, intra_cost, out, 0, NULL, NULL);
}

void xo_lowres_frame_cost_vbv(const xo_geom *g, const xo_lowres_in *in, const uint8_t *const fenc[4], const uint8_t *const fref0[4],
                              const uint8_t *const fref1[4], int16_t (*mvs0)[2], int *costs0, int16_t (*mvs1)[2], int *costs1,
                              const int16_t (*ref1_mvs)[2], uint16_t *intra_cost, xo_lowres_out *out, const uint16_t *inv_qscale, int *row_satd)
{
    lowres_frame_cost(g, in, fenc, fref0, fref1, mvs0, costs0, mvs1, costs1, ref1_mvs, intra_cost, out, row_satd != NULL, inv_qscale, row_satd);
}
