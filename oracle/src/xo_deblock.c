/* xo_deblock.c — ORACLE (test infrastructure only): in-loop deblocking filter of one progressive frame,
 * restating S/common/frame.c:376-800 (tables :377-421, edge filters :424-586, deblock_edge :588-618,
 * x264_frame_deblock_row :621-792).  MBAFF is not covered (the reference's b_interlaced paths). */
#include <stdlib.h>
#include <string.h>
#include "xo.h"

/* H.264 table 8-16 / 8-17 as indexed by the reference: entry [q + 12] for q in -12..63 (frame.c:377-421) */
static const uint8_t alpha_tab[52] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 4, 4, 5, 6, 7, 8, 9, 10, 12, 13, 15, 17, 20, 22, 25, 28, 32, 36, 40, 45,
                                       50, 56, 63, 71, 80, 90, 101, 113, 127, 144, 162, 182, 203, 226, 255, 255 };
static const uint8_t beta_tab[52] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10,
                                      11, 11, 12, 12, 13, 13, 14, 14, 15, 15, 16, 16, 17, 17, 18, 18 };
static const uint8_t tc0_tab[52][3] = { /* bS 1, 2, 3 */
    { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 0 },
    { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 1 }, { 0, 0, 1 }, { 0, 0, 1 }, { 0, 0, 1 }, { 0, 1, 1 },
    { 0, 1, 1 }, { 1, 1, 1 }, { 1, 1, 1 }, { 1, 1, 1 }, { 1, 1, 1 }, { 1, 1, 2 }, { 1, 1, 2 }, { 1, 1, 2 }, { 1, 1, 2 }, { 1, 2, 3 }, { 1, 2, 3 },
    { 2, 2, 3 }, { 2, 2, 4 }, { 2, 3, 4 }, { 2, 3, 4 }, { 3, 3, 5 }, { 3, 4, 6 }, { 3, 4, 6 }, { 4, 5, 7 }, { 4, 5, 8 }, { 4, 6, 9 }, { 5, 7, 10 },
    { 6, 8, 11 }, { 6, 8, 13 }, { 7, 10, 14 }, { 8, 11, 16 }, { 9, 12, 18 }, { 10, 13, 20 }, { 11, 15, 23 }, { 13, 17, 25 } };
/* S/common/macroblock.h:241-251 */
static const uint8_t chroma_qp_tab[52] = { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29,
                                           29, 30, 31, 32, 32, 33, 34, 34, 35, 35, 36, 36, 37, 37, 37, 38, 38, 38, 39, 39, 39, 39 };
static int idx52(int q) { return q < 0 ? 0 : q > 51 ? 51 : q; }
static int alpha_of(int q) { return q < 0 ? 0 : alpha_tab[idx52(q)]; }
static int beta_of(int q) { return q < 0 ? 0 : beta_tab[idx52(q)]; }
static int tc0_of(int q, int bs) { return bs == 0 ? -1 : q < 0 ? 0 : tc0_tab[idx52(q)][bs - 1]; }
static int chroma_qp(int qp, int off) { return chroma_qp_tab[idx52(qp + off)]; }
static int clip3(int v, int lo, int hi) { return v < lo ? lo : v > hi ? hi : v; }
static int clip_u8(int v) { return v < 0 ? 0 : v > 255 ? 255 : v; }

/* one line across a luma edge with bS < 4 (frame.c:424-467); xs = step across the edge */
static void luma_line(uint8_t *pix, int xs, int alpha, int beta, int tc0)
{
    const int p2 = pix[-3 * xs], p1 = pix[-2 * xs], p0 = pix[-xs], q0 = pix[0], q1 = pix[xs], q2 = pix[2 * xs];
    if (abs(p0 - q0) >= alpha || abs(p1 - p0) >= beta || abs(q1 - q0) >= beta) return;
    int tc = tc0;
    if (abs(p2 - p0) < beta) { pix[-2 * xs] = p1 + clip3(((p2 + ((p0 + q0 + 1) >> 1)) >> 1) - p1, -tc0, tc0); tc++; }
    if (abs(q2 - q0) < beta) { pix[xs] = q1 + clip3(((q2 + ((p0 + q0 + 1) >> 1)) >> 1) - q1, -tc0, tc0); tc++; }
    const int delta = clip3((((q0 - p0) << 2) + (p1 - q1) + 4) >> 3, -tc, tc);
    pix[-xs] = clip_u8(p0 + delta);
    pix[0] = clip_u8(q0 - delta);
}
/* frame.c:470-497 */
static void chroma_line(uint8_t *pix, int xs, int alpha, int beta, int tc)
{
    const int p1 = pix[-2 * xs], p0 = pix[-xs], q0 = pix[0], q1 = pix[xs];
    if (abs(p0 - q0) >= alpha || abs(p1 - p0) >= beta || abs(q1 - q0) >= beta) return;
    const int delta = clip3((((q0 - p0) << 2) + (p1 - q1) + 4) >> 3, -tc, tc);
    pix[-xs] = clip_u8(p0 + delta);
    pix[0] = clip_u8(q0 - delta);
}
/* frame.c:507-552 */
static void luma_intra_line(uint8_t *pix, int xs, int alpha, int beta)
{
    const int p2 = pix[-3 * xs], p1 = pix[-2 * xs], p0 = pix[-xs], q0 = pix[0], q1 = pix[xs], q2 = pix[2 * xs];
    if (abs(p0 - q0) >= alpha || abs(p1 - p0) >= beta || abs(q1 - q0) >= beta) return;
    if (abs(p0 - q0) < ((alpha >> 2) + 2)) {
        if (abs(p2 - p0) < beta) {
            const int p3 = pix[-4 * xs];
            pix[-xs] = (p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3;
            pix[-2 * xs] = (p2 + p1 + p0 + q0 + 2) >> 2;
            pix[-3 * xs] = (2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3;
        } else
            pix[-xs] = (2 * p1 + p0 + q1 + 2) >> 2;
        if (abs(q2 - q0) < beta) {
            const int q3 = pix[3 * xs];
            pix[0] = (p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3;
            pix[xs] = (p0 + q0 + q1 + q2 + 2) >> 2;
            pix[2 * xs] = (2 * q3 + 3 * q2 + q1 + q0 + p0 + 4) >> 3;
        } else
            pix[0] = (2 * q1 + q0 + p1 + 2) >> 2;
    } else {
        pix[-xs] = (2 * p1 + p0 + q1 + 2) >> 2;
        pix[0] = (2 * q1 + q0 + p1 + 2) >> 2;
    }
}
/* frame.c:562-580 */
static void chroma_intra_line(uint8_t *pix, int xs, int alpha, int beta)
{
    const int p1 = pix[-2 * xs], p0 = pix[-xs], q0 = pix[0], q1 = pix[xs];
    if (abs(p0 - q0) >= alpha || abs(p1 - p0) >= beta || abs(q1 - q0) >= beta) return;
    pix[-xs] = (2 * p1 + p0 + q1 + 2) >> 2;
    pix[0] = (2 * q1 + q0 + p1 + 2) >> 2;
}

/* nnz of 4x4 block (x,y) of macroblock mb as the deblocker sees it: with CAVLC + 8x8 transform an 8x8-transformed
 * macroblock reports per-8x8 "any coefficient" flags (munge_cavlc_nnz_row, frame.c:336-352) */
static int nnz_at(const xo_deblock_in *d, int mb, int x, int y)
{
    const uint8_t *n = d->nnz[mb];
    if (d->b_cavlc_8x8dct && d->transform8x8[mb]) {
        const int bx = x & 2, by = y & 2;
        return n[bx + by * 4] | n[bx + 1 + by * 4] | n[bx + (by + 1) * 4] | n[bx + 1 + (by + 1) * 4];
    }
    return n[x + y * 4];
}

void xo_frame_deblock(const xo_geom *g, const xo_deblock_in *d, uint8_t *py, uint8_t *pu, uint8_t *pv, int stride_c)
{
    const int W = g->mb_width, H = g->mb_height, stride = g->stride;
    const int s8 = 2 * W, s4 = 4 * W;
    const int qp_thresh = 15 - (d->alpha_c0_offset < d->beta_offset ? d->alpha_c0_offset : d->beta_offset) - (d->chroma_qp_offset > 0 ? d->chroma_qp_offset : 0);
    for (int mb_y = 0; mb_y < H; mb_y++)
        for (int mb_x = 0; mb_x < W; mb_x++) {
            const int mb = mb_y * W + mb_x, t8 = d->transform8x8[mb], qp = d->qp[mb];
            const int intra = d->type[mb] >= 0 && d->type[mb] <= 3;
            int edge_end = d->type[mb] == 6 /* P_SKIP */ ? 1 : 4;
            const int no_sub8x8 = d->type[mb] != 5 /* P_8x8 */ || !d->b_psub8x8;
            uint8_t *y0 = py + 16 * mb_y * stride + 16 * mb_x, *u0 = pu + 8 * mb_y * stride_c + 8 * mb_x, *v0 = pv + 8 * mb_y * stride_c + 8 * mb_x;
            if (qp <= qp_thresh) edge_end = 1;
            for (int dir = 0; dir < 2; dir++) {
                int edge = dir ? mb_y == 0 : mb_x == 0;
                if (edge) edge += t8; /* no neighbour: start at the first inner edge */
                for (; edge < edge_end; edge += t8 + 1) {
                    const int mbn = edge ? mb : dir == 0 ? mb - 1 : mb - W;
                    const int n_intra = d->type[mbn] >= 0 && d->type[mbn] <= 3;
                    const int qpn = d->qp[mbn];
                    const int q_luma = (qp + qpn + 1) >> 1;
                    const int q_chroma = (chroma_qp(qp, d->chroma_qp_offset) + chroma_qp(qpn, d->chroma_qp_offset) + 1) >> 1;
                    const int xs = dir == 0 ? 1 : stride, ys = dir == 0 ? stride : 1;
                    const int xsc = dir == 0 ? 1 : stride_c, ysc = dir == 0 ? stride_c : 1;
                    uint8_t *ly = y0 + 4 * edge * xs, *lu = u0 + 2 * edge * xsc, *lv = v0 + 2 * edge * xsc;
                    if (edge == 0 && (intra || n_intra)) { /* macroblock edge next to intra: bS 4 (frame.c:762-766) */
                        int a = alpha_of(q_luma + d->alpha_c0_offset), b = beta_of(q_luma + d->beta_offset);
                        if (a && b) for (int i = 0; i < 16; i++) luma_intra_line(ly + i * ys, xs, a, b);
                        a = alpha_of(q_chroma + d->alpha_c0_offset); b = beta_of(q_chroma + d->beta_offset);
                        if (a && b) for (int i = 0; i < 8; i++) { chroma_intra_line(lu + i * ysc, xsc, a, b); chroma_intra_line(lv + i * ysc, xsc, a, b); }
                        continue;
                    }
                    int bs[4] = { 0, 0, 0, 0 };
                    if (intra || n_intra) bs[0] = bs[1] = bs[2] = bs[3] = 3; /* inner edges of an intra macroblock */
                    else
                        for (int i = 0; i < 4; i++) { /* frame.c:706-741 */
                            const int x = dir == 0 ? edge : i, y = dir == 0 ? i : edge;
                            const int xn = dir == 0 ? (x - 1) & 3 : x, yn = dir == 0 ? y : (y - 1) & 3;
                            if (nnz_at(d, mb, x, y) || nnz_at(d, mbn, xn, yn)) bs[i] = 2;
                            else if (!(edge & no_sub8x8)) {
                                if ((i & no_sub8x8) && bs[i - 1] != 2) bs[i] = bs[i - 1];
                                else {
                                    const int mbx = mb_x - (edge == 0 && dir == 0), mby = mb_y - (edge == 0 && dir == 1); /* neighbour's coordinates */
                                    const int i8p = 2 * s8 * mb_y + 2 * mb_x + (x >> 1) + (y >> 1) * s8, i8q = 2 * s8 * mby + 2 * mbx + (xn >> 1) + (yn >> 1) * s8;
                                    const int i4p = 4 * s4 * mb_y + 4 * mb_x + x + y * s4, i4q = 4 * s4 * mby + 4 * mbx + xn + yn * s4;
                                    int diff = d->ref[0][i8p] != d->ref[0][i8q] || abs(d->mv[0][i4p][0] - d->mv[0][i4q][0]) >= 4 ||
                                               abs(d->mv[0][i4p][1] - d->mv[0][i4q][1]) >= 4;
                                    if (!diff && d->b_slice_b)
                                        diff = d->ref[1][i8p] != d->ref[1][i8q] || abs(d->mv[1][i4p][0] - d->mv[1][i4q][0]) >= 4 ||
                                               abs(d->mv[1][i4p][1] - d->mv[1][i4q][1]) >= 4;
                                    if (diff) bs[i] = 1;
                                }
                            }
                        }
                    if (!(bs[0] | bs[1] | bs[2] | bs[3])) continue;
                    { /* deblock_edge (frame.c:588-604), luma */
                        const int ia = q_luma + d->alpha_c0_offset, a = alpha_of(ia), b = beta_of(q_luma + d->beta_offset);
                        if (a && b)
                            for (int i = 0; i < 16; i++) {
                                const int tc0 = tc0_of(ia, bs[i >> 2]);
                                if (tc0 >= 0) luma_line(ly + i * ys, xs, a, b, tc0);
                            }
                    }
                    if (!(edge & 1)) {
                        const int ia = q_chroma + d->alpha_c0_offset, a = alpha_of(ia), b = beta_of(q_chroma + d->beta_offset);
                        if (a && b)
                            for (int i = 0; i < 8; i++) {
                                const int tc = tc0_of(ia, bs[i >> 1]) + 1;
                                if (tc > 0) { chroma_line(lu + i * ysc, xsc, a, b, tc); chroma_line(lv + i * ysc, xsc, a, b, tc); }
                            }
                    }
                }
            }
        }
}
