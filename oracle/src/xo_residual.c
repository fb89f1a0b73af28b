/* xo_residual.c — ORACLE (test infrastructure only): the residual path of one INTER macroblock, i.e. the
 * sequencing of transform / quant / zigzag / decimation / dequant / inverse transform done by
 * x264_macroblock_encode for inter MBs (S/encoder/macroblock.c:596-742) and x264_mb_encode_8x8_chroma
 * (:272-363), without trellis, noise reduction or lossless.  The primitives are the pinned xo_* functions. */
#include <string.h>
#include "xo.h"

/* classic zig-zag (frame) order as (row y, column x); the reference reads dct[x][y] because coefficients are
 * stored transposed (S/common/dct.c:488-560) */
static void zigzag_order(int n, uint8_t *flat)
{
    int y = 0, x = 0, up = 1;
    for (int i = 0; i < n * n; i++) {
        flat[i] = (uint8_t)(x * n + y);
        if (up) {
            if (x == n - 1) { y++; up = 0; }
            else if (y == 0) { x++; up = 0; }
            else { y--; x++; }
        } else {
            if (y == n - 1) { x++; up = 1; }
            else if (x == 0) { y++; up = 1; }
            else { y++; x--; }
        }
    }
}
void xo_zigzag_scan_4x4(int16_t level[16], const int16_t dct[16])
{
    uint8_t o[16];
    zigzag_order(4, o);
    for (int i = 0; i < 16; i++) level[i] = dct[o[i]];
}
void xo_zigzag_scan_8x8(int16_t level[64], const int16_t dct[64])
{
    uint8_t o[64];
    zigzag_order(8, o);
    for (int i = 0; i < 64; i++) level[i] = dct[o[i]];
}

/* S/common/quant.c:203-252: i_max = 15 (skip DC), 16 or 64 */
int xo_decimate_score(const int16_t *dct, int i_max)
{
    static const uint8_t t4[16] = { 3, 2, 2, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
    static const uint8_t t8[64] = { 3, 3, 3, 3, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1 };
    const uint8_t *tab = i_max == 64 ? t8 : t4;
    if (i_max == 15) dct++;
    int idx = i_max - 1, score = 0;
    while (idx >= 0 && dct[idx] == 0) idx--;
    while (idx >= 0) {
        if ((unsigned)(dct[idx--] + 1) > 2) return 9;
        int run = 0;
        while (idx >= 0 && dct[idx] == 0) { idx--; run++; }
        score += tab[run];
    }
    return score;
}

/* raster position of 4x4 block idx inside the MB (block_idx_x/y, S/common/macroblock.h:195-202) */
static const uint8_t bx4[16] = { 0, 1, 0, 1, 2, 3, 2, 3, 0, 1, 0, 1, 2, 3, 2, 3 };
static const uint8_t by4[16] = { 0, 0, 1, 1, 0, 0, 1, 1, 2, 2, 3, 3, 2, 2, 3, 3 };

static void to_tiles(uint8_t fe[16 * 16], uint8_t fd[32 * 16], const uint8_t *src, const uint8_t *pred, int n)
{
    for (int y = 0; y < n; y++) {
        memcpy(fe + 16 * y, src + n * y, n);
        memcpy(fd + 32 * y, pred + n * y, n);
    }
}

/* x264_mb_encode_8x8_chroma(h, b_inter, chroma_qp), macroblock.c:272-363; b_decimate as computed there (:275: only inter macroblocks decimate) */
static void encode_chroma(const xo_resid_in *in, int b_inter, int b_decimate, const uint8_t fenc_u[64], const uint8_t fenc_v[64],
                          uint8_t rec_u[64], uint8_t rec_v[64], xo_resid_out *out)
{
    uint8_t fe[16 * 16], fd[32 * 16];
    int cbp_chroma = 0;
    const int cqp = in->chroma_qp;
    uint16_t mf[16], bias[16];
    int dq[6][16];
    xo_quant4_tables(in->cqm, 2 + b_inter, cqp, mf, bias); /* CQM_4IC + b_inter */
    xo_dequant4_table(in->cqm, 2 + b_inter, dq);
    for (int ch = 0; ch < 2; ch++) {
        int16_t dct4[4][16], dc[4];
        int score = 0, nz_ac = 0;
        to_tiles(fe, fd, ch ? fenc_v : fenc_u, ch ? rec_v : rec_u, 8);
        for (int i = 0; i < 4; i++)
            xo_sub4x4_dct(dct4[i], fe + 4 * (i & 1) + 4 * (i >> 1) * 16, fd + 4 * (i & 1) + 4 * (i >> 1) * 32);
        { /* dct2x2dc, macroblock.c:72-85 */
            int d0 = dct4[0][0] + dct4[1][0], d1 = dct4[2][0] + dct4[3][0];
            int d2 = dct4[0][0] - dct4[1][0], d3 = dct4[2][0] - dct4[3][0];
            dc[0] = d0 + d1; dc[2] = d2 + d3; dc[1] = d0 - d1; dc[3] = d2 - d3; /* d[0][0], d[1][0], d[0][1], d[1][1] */
            for (int i = 0; i < 4; i++) dct4[i][0] = 0;
        }
        for (int i = 0; i < 4; i++) {
            int nz = xo_quant_4x4(dct4[i], mf, bias);
            out->nnz[16 + i + ch * 4] = nz;
            if (nz) {
                nz_ac = 1;
                xo_zigzag_scan_4x4(out->luma4x4[16 + i + ch * 4], dct4[i]);
                xo_dequant_4x4(dct4[i], (const int(*)[16])dq, cqp);
                if (b_decimate) score += xo_decimate_score(out->luma4x4[16 + i + ch * 4], 15);
            }
        }
        int nz_dc = xo_quant_2x2_dc(dc, mf[0] >> 1, bias[0] << 1);
        out->nnz[25 + ch] = nz_dc;
        /* IDCT_DEQUANT_START, macroblock.c:42-53 */
        int e0 = dc[0] + dc[1], e1 = dc[2] + dc[3], e2 = dc[0] - dc[1], e3 = dc[2] - dc[3];
        int dmf = dq[cqp % 6][0], qbits = cqp / 6 - 5;
        if (qbits > 0) { dmf <<= qbits; qbits = 0; }
        if ((b_decimate && score < 7) || !nz_ac) {
            for (int i = 0; i < 4; i++) out->nnz[16 + i + ch * 4] = 0;
            if (nz_dc) {
                int16_t o[4];
                for (int i = 0; i < 4; i++) out->chroma_dc[ch][i] = dc[(i & 1) * 2 + (i >> 1)]; /* zigzag_scan_2x2_dc: level[i] = dct[x][y] */
                o[0] = (int16_t)((e0 + e1) * dmf >> -qbits); o[1] = (int16_t)((e0 - e1) * dmf >> -qbits);
                o[2] = (int16_t)((e2 + e3) * dmf >> -qbits); o[3] = (int16_t)((e2 - e3) * dmf >> -qbits);
                xo_add_idct_dc(fd, o, 4);
            }
        } else {
            cbp_chroma = 1;
            if (nz_dc) {
                for (int i = 0; i < 4; i++) out->chroma_dc[ch][i] = dc[(i & 1) * 2 + (i >> 1)];
                dct4[0][0] = (int16_t)((e0 + e1) * dmf >> -qbits); dct4[1][0] = (int16_t)((e0 - e1) * dmf >> -qbits);
                dct4[2][0] = (int16_t)((e2 + e3) * dmf >> -qbits); dct4[3][0] = (int16_t)((e2 - e3) * dmf >> -qbits);
            }
            for (int i = 0; i < 4; i++) xo_add4x4_idct(fd + 4 * (i & 1) + 4 * (i >> 1) * 32, dct4[i]);
        }
        uint8_t *rec = ch ? rec_v : rec_u;
        for (int y = 0; y < 8; y++) memcpy(rec + 8 * y, fd + 32 * y, 8);
    }
    if (cbp_chroma) cbp_chroma = 2;
    else if (out->nnz[25] | out->nnz[26]) cbp_chroma = 1;
    out->cbp_chroma = cbp_chroma;
}

void xo_residual_inter_mb(const xo_resid_in *in, const uint8_t fenc_y[256], const uint8_t fenc_u[64], const uint8_t fenc_v[64],
                          uint8_t rec_y[256], uint8_t rec_u[64], uint8_t rec_v[64], xo_resid_out *out)
{
    uint8_t fe[16 * 16], fd[32 * 16];
    memset(out, 0, sizeof(*out));
    const int qp = in->qp, b_decimate = in->b_decimate;
    int decimate_mb = 0, cbp_luma = 0;

    /* ---- luma ---- */
    to_tiles(fe, fd, fenc_y, rec_y, 16);
    if (in->b_transform_8x8) { /* macroblock.c:627-677 */
        int16_t dct8[4][64];
        uint16_t mf[64], bias[64];
        int dq[6][64];
        xo_quant8_tables(in->cqm, 1, qp, mf, bias); /* CQM_8PY */
        xo_dequant8_table(in->cqm, 1, dq);
        for (int i = 0; i < 4; i++) {
            xo_sub8x8_dct8(dct8[i], fe + 8 * (i & 1) + 8 * (i >> 1) * 16, fd + 8 * (i & 1) + 8 * (i >> 1) * 32);
            if (xo_quant_8x8(dct8[i], mf, bias)) {
                xo_zigzag_scan_8x8(out->luma8x8[i], dct8[i]);
                if (b_decimate) {
                    int s = xo_decimate_score(out->luma8x8[i], 64);
                    decimate_mb += s;
                    if (s >= 4) cbp_luma |= 1 << i;
                } else
                    cbp_luma |= 1 << i;
            }
        }
        if (decimate_mb < 6 && b_decimate)
            cbp_luma = 0; /* nnz all zero (already) */
        else
            for (int i = 0; i < 4; i++)
                if (cbp_luma & (1 << i)) {
                    xo_dequant_8x8(dct8[i], (const int(*)[64])dq, qp);
                    xo_add8x8_idct8(fd + 8 * (i & 1) + 8 * (i >> 1) * 32, dct8[i]);
                    for (int k = 0; k < 4; k++) out->nnz[i * 4 + k] = 1;
                }
    } else { /* macroblock.c:678-742 */
        int16_t dct4[16][16];
        uint16_t mf[16], bias[16];
        int dq[6][16];
        xo_quant4_tables(in->cqm, 1, qp, mf, bias); /* CQM_4PY */
        xo_dequant4_table(in->cqm, 1, dq);
        for (int i = 0; i < 16; i++)
            xo_sub4x4_dct(dct4[i], fe + 4 * bx4[i] + 4 * by4[i] * 16, fd + 4 * bx4[i] + 4 * by4[i] * 32);
        for (int i8 = 0; i8 < 4; i8++) {
            int dec8 = 0, cbp = 0;
            for (int i4 = 0; i4 < 4; i4++) {
                int idx = i8 * 4 + i4;
                int nz = xo_quant_4x4(dct4[idx], mf, bias);
                out->nnz[idx] = nz;
                if (nz) {
                    xo_zigzag_scan_4x4(out->luma4x4[idx], dct4[idx]);
                    xo_dequant_4x4(dct4[idx], (const int(*)[16])dq, qp);
                    if (b_decimate && dec8 < 6) dec8 += xo_decimate_score(out->luma4x4[idx], 16);
                    cbp = 1;
                }
            }
            decimate_mb += dec8;
            if (b_decimate) {
                if (dec8 < 4) for (int k = 0; k < 4; k++) out->nnz[i8 * 4 + k] = 0;
                else cbp_luma |= 1 << i8;
            } else if (cbp)
                cbp_luma |= 1 << i8;
        }
        if (b_decimate && decimate_mb < 6) {
            cbp_luma = 0;
            memset(out->nnz, 0, 16);
        }
        for (int i8 = 0; i8 < 4; i8++)
            if (cbp_luma & (1 << i8))
                for (int i4 = 0; i4 < 4; i4++) { /* add8x8_idct: blocks with all-zero coefficients add nothing */
                    int idx = i8 * 4 + i4;
                    xo_add4x4_idct(fd + 4 * bx4[idx] + 4 * by4[idx] * 32, dct4[idx]);
                }
    }
    for (int y = 0; y < 16; y++) memcpy(rec_y + 16 * y, fd + 32 * y, 16);
    out->cbp_luma = cbp_luma;

    encode_chroma(in, 1, b_decimate, fenc_u, fenc_v, rec_u, rec_v, out);
}

/* ---------------------------------------------------------------------------------------------------------
 * One I_16x16 macroblock through x264_macroblock_encode (S/encoder/macroblock.c:512-530 predict + x264_mb_encode_i16x16 :184-270,
 * then :744-760 chroma predict + x264_mb_encode_8x8_chroma with b_inter = 0).  nb_* = corner, row above, left column (as for
 * xo_predict_*); in->b_decimate = (slice B) || (b_dct_decimate && slice P) (:193); in->b_transform_8x8 is ignored (:520).
 * out->luma4x4[0..15] hold the AC levels (DC slot zero), luma_dc the zigzagged 4x4 DC levels (h->dct.luma16x16_dc), nnz[24] its flag. */
void xo_residual_intra16_mb(const xo_resid_in *in, int mode16, int mode_chroma, const uint8_t fenc_y[256], const uint8_t fenc_u[64],
                            const uint8_t fenc_v[64], const uint8_t nb_y[33], const uint8_t nb_u[17], const uint8_t nb_v[17],
                            uint8_t rec_y[256], uint8_t rec_u[64], uint8_t rec_v[64], xo_resid_out *out, int16_t luma_dc[16])
{
    uint8_t fe[16 * 16], fd[32 * 16];
    int16_t dct4[16][16], dc[16];
    uint16_t mf[16], bias[16];
    int dq[6][16];
    const int qp = in->qp;
    memset(out, 0, sizeof(*out));
    memset(luma_dc, 0, 16 * sizeof(int16_t));
    xo_predict_16x16(mode16, nb_y, rec_y);
    to_tiles(fe, fd, fenc_y, rec_y, 16);
    xo_quant4_tables(in->cqm, 0, qp, mf, bias); /* CQM_4IY */
    xo_dequant4_table(in->cqm, 0, dq);
    int decimate_score = in->b_decimate ? 0 : 9, cbp_luma = 0;
    for (int i = 0; i < 16; i++) {
        xo_sub4x4_dct(dct4[i], fe + 4 * bx4[i] + 4 * by4[i] * 16, fd + 4 * bx4[i] + 4 * by4[i] * 32);
        dc[bx4[i] + 4 * by4[i]] = dct4[i][0]; /* block_idx_xy_1d, macroblock.c:219 */
        dct4[i][0] = 0;
        int nz = xo_quant_4x4(dct4[i], mf, bias);
        out->nnz[i] = nz;
        if (nz) {
            xo_zigzag_scan_4x4(out->luma4x4[i], dct4[i]);
            xo_dequant_4x4(dct4[i], (const int(*)[16])dq, qp);
            if (decimate_score < 6) decimate_score += xo_decimate_score(out->luma4x4[i], 15);
            cbp_luma = 0xf;
        }
    }
    if (decimate_score < 6) { cbp_luma = 0; memset(out->nnz, 0, 16); }
    xo_dct4x4dc(dc);
    int nz = xo_quant_4x4_dc(dc, mf[0] >> 1, bias[0] << 1);
    out->nnz[24] = nz;
    if (nz) {
        xo_zigzag_scan_4x4(luma_dc, dc);
        xo_idct4x4dc(dc);
        xo_dequant_4x4_dc(dc, (const int(*)[16])dq, qp);
        if (cbp_luma) for (int i = 0; i < 16; i++) dct4[i][0] = dc[bx4[i] + 4 * by4[i]];
    }
    if (cbp_luma) for (int i = 0; i < 16; i++) xo_add4x4_idct(fd + 4 * bx4[i] + 4 * by4[i] * 32, dct4[i]);
    else if (nz) xo_add_idct_dc(fd, dc, 16);
    for (int y = 0; y < 16; y++) memcpy(rec_y + 16 * y, fd + 32 * y, 16);
    out->cbp_luma = cbp_luma;
    xo_predict_8x8c(mode_chroma, nb_u, rec_u);
    xo_predict_8x8c(mode_chroma, nb_v, rec_v);
    encode_chroma(in, 0, 0, fenc_u, fenc_v, rec_u, rec_v, out);
    for (int i = 0; i < 24; i++) if (!out->nnz[i]) memset(out->luma4x4[i], 0, sizeof(out->luma4x4[i])); /* uncoded blocks read as zero */
}

/* ---------------------------------------------------------------------------------------------------------
 * x264_macroblock_probe_skip (S/encoder/macroblock.c:797-883) on a prediction the caller has already formed
 * (b_bidir = 1; the P-skip form only adds mc_luma / mc_chroma of the clipped pskip mv in front, :809-819, :851-856).
 * Returns 1 when the macroblock would quantise to nothing (skippable). */
int xo_lambda2(int qp)
{
    /* x264_lambda2_tab (S/encoder/analyse.c:151-160) restated as the closed form that generates it:
     * floor(0.9 * 2^((qp-12)/3) * 256); checked against every entry of the reference table in tests/test_oracle_vs_ref.py */
    static const double cbrt2[3] = { 1.0, 1.2599210498948732, 1.5874010519681994 };
    double v = 0.9 * 256.0 * cbrt2[qp % 3];
    int e = qp / 3 - 4;
    for (; e > 0; e--) v *= 2.0;
    for (; e < 0; e++) v *= 0.5;
    return (int)v;
}

int xo_probe_skip_mb(const xo_resid_in *in, const uint8_t fenc_y[256], const uint8_t fenc_u[64], const uint8_t fenc_v[64],
                     const uint8_t pred_y[256], const uint8_t pred_u[64], const uint8_t pred_v[64])
{
    uint8_t fe[16 * 16], fd[32 * 16];
    uint16_t mf[16], bias[16];
    int16_t dct[16], scan[16];
    int decimate_mb = 0;

    /* luma: 16 4x4 blocks in block order, early out once the running score reaches 6 (:822-840) */
    to_tiles(fe, fd, fenc_y, pred_y, 16);
    xo_quant4_tables(in->cqm, 1, in->qp, mf, bias); /* CQM_4PY */
    for (int i = 0; i < 16; i++) {
        xo_sub4x4_dct(dct, fe + 4 * bx4[i] + 4 * by4[i] * 16, fd + 4 * bx4[i] + 4 * by4[i] * 32);
        if (!xo_quant_4x4(dct, mf, bias)) continue;
        xo_zigzag_scan_4x4(scan, dct);
        decimate_mb += xo_decimate_score(scan, 16);
        if (decimate_mb >= 6) return 0;
    }

    /* chroma (:843-879): planes whose SSD is under the lambda2 threshold are not examined at all */
    const int cqp = in->chroma_qp;
    const int thresh = (xo_lambda2(cqp) + 32) >> 6;
    xo_quant4_tables(in->cqm, 3, cqp, mf, bias); /* CQM_4PC */
    for (int ch = 0; ch < 2; ch++) {
        const uint8_t *src = ch ? fenc_v : fenc_u, *prd = ch ? pred_v : pred_u;
        int ssd = 0;
        for (int k = 0; k < 64; k++) { int d = src[k] - prd[k]; ssd += d * d; }
        if (ssd < thresh) continue;
        int16_t dct4[4][16], dc[4];
        to_tiles(fe, fd, src, prd, 8);
        for (int i = 0; i < 4; i++)
            xo_sub4x4_dct(dct4[i], fe + 4 * (i & 1) + 4 * (i >> 1) * 16, fd + 4 * (i & 1) + 4 * (i >> 1) * 32);
        {
            int d0 = dct4[0][0] + dct4[1][0], d1 = dct4[2][0] + dct4[3][0];
            int d2 = dct4[0][0] - dct4[1][0], d3 = dct4[2][0] - dct4[3][0];
            dc[0] = d0 + d1; dc[2] = d2 + d3; dc[1] = d0 - d1; dc[3] = d2 - d3;
            for (int i = 0; i < 4; i++) dct4[i][0] = 0;
        }
        if (xo_quant_2x2_dc(dc, mf[0] >> 1, bias[0] << 1)) return 0;
        decimate_mb = 0;
        for (int i = 0; i < 4; i++) {
            if (!xo_quant_4x4(dct4[i], mf, bias)) continue;
            xo_zigzag_scan_4x4(scan, dct4[i]);
            decimate_mb += xo_decimate_score(scan, 15);
            if (decimate_mb >= 7) return 0;
        }
    }
    return 1;
}
