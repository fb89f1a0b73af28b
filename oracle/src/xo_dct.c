/* xo_dct.c — ORACLE (test infrastructure only): H.264 integer transforms, (de)quantisation and the
 * quantiser tables, after S/common/dct.c, S/common/quant.c and S/common/set.c:68-174.
 * Intermediates are stored to int16_t exactly where the reference stores to int16_t arrays. */
#include <string.h>
#include "xo.h"

static inline uint8_t clip_u8(int x) { return x < 0 ? 0 : x > 255 ? 255 : x; }
const char *xo_backend(void) { return "port"; }

/* one 4-point forward core butterfly (dct.c:131-154): out = {s03+s12, 2*d03+d12, s03-s12, d03-2*d12} */
static inline void fwd4(int a, int b, int c, int d, int o[4])
{
    int s03 = a + d, s12 = b + c, d03 = a - d, d12 = b - c;
    o[0] = s03 + s12; o[1] = 2 * d03 + d12; o[2] = s03 - s12; o[3] = d03 - 2 * d12;
}

/* S/common/dct.c:122-155.  Output layout is the reference's: dct[i][k] = transform along rows of the
 * column-transformed data, i.e. coefficients are stored transposed w.r.t. the textbook layout. */
void xo_sub4x4_dct(int16_t dct[16], const uint8_t *pix1, const uint8_t *pix2)
{
    int16_t tmp[4][4];
    for (int i = 0; i < 4; i++) {
        int d[4], o[4];
        for (int x = 0; x < 4; x++) d[x] = pix1[i * XO_FENC_STRIDE + x] - pix2[i * XO_FDEC_STRIDE + x];
        fwd4(d[0], d[1], d[2], d[3], o);
        for (int k = 0; k < 4; k++) tmp[k][i] = (int16_t)o[k];
    }
    for (int i = 0; i < 4; i++) {
        int o[4];
        fwd4(tmp[i][0], tmp[i][1], tmp[i][2], tmp[i][3], o);
        for (int k = 0; k < 4; k++) dct[i * 4 + k] = (int16_t)o[k];
    }
}

/* S/common/dct.c:174-216 */
void xo_add4x4_idct(uint8_t *dst, int16_t dct[16])
{
    int16_t tmp[4][4], d[4][4];
    for (int i = 0; i < 4; i++) {
        int s02 = dct[0 * 4 + i] + dct[2 * 4 + i], d02 = dct[0 * 4 + i] - dct[2 * 4 + i];
        int s13 = dct[1 * 4 + i] + (dct[3 * 4 + i] >> 1), d13 = (dct[1 * 4 + i] >> 1) - dct[3 * 4 + i];
        tmp[i][0] = (int16_t)(s02 + s13); tmp[i][1] = (int16_t)(d02 + d13);
        tmp[i][2] = (int16_t)(d02 - d13); tmp[i][3] = (int16_t)(s02 - s13);
    }
    for (int i = 0; i < 4; i++) {
        int s02 = tmp[0][i] + tmp[2][i], d02 = tmp[0][i] - tmp[2][i];
        int s13 = tmp[1][i] + (tmp[3][i] >> 1), d13 = (tmp[1][i] >> 1) - tmp[3][i];
        d[0][i] = (int16_t)((s02 + s13 + 32) >> 6); d[1][i] = (int16_t)((d02 + d13 + 32) >> 6);
        d[2][i] = (int16_t)((d02 - d13 + 32) >> 6); d[3][i] = (int16_t)((s02 - s13 + 32) >> 6);
    }
    for (int y = 0; y < 4; y++)
        for (int x = 0; x < 4; x++)
            dst[y * XO_FDEC_STRIDE + x] = clip_u8(dst[y * XO_FDEC_STRIDE + x] + d[y][x]);
}

/* S/common/dct.c:238-263 */
static inline void fwd8(const int s[8], int o[8])
{
    int s07 = s[0] + s[7], s16 = s[1] + s[6], s25 = s[2] + s[5], s34 = s[3] + s[4];
    int a0 = s07 + s34, a1 = s16 + s25, a2 = s07 - s34, a3 = s16 - s25;
    int d07 = s[0] - s[7], d16 = s[1] - s[6], d25 = s[2] - s[5], d34 = s[3] - s[4];
    int a4 = d16 + d25 + (d07 + (d07 >> 1));
    int a5 = d07 - d34 - (d25 + (d25 >> 1));
    int a6 = d07 + d34 - (d16 + (d16 >> 1));
    int a7 = d16 - d25 + (d34 + (d34 >> 1));
    o[0] = a0 + a1; o[1] = a4 + (a7 >> 2); o[2] = a2 + (a3 >> 1); o[3] = a5 + (a6 >> 2);
    o[4] = a0 - a1; o[5] = a6 - (a5 >> 2); o[6] = (a2 >> 1) - a3; o[7] = (a4 >> 2) - a7;
}

/* S/common/dct.c:265-285 */
void xo_sub8x8_dct8(int16_t dct[64], const uint8_t *pix1, const uint8_t *pix2)
{
    int16_t tmp[8][8];
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++)
            tmp[y][x] = pix1[y * XO_FENC_STRIDE + x] - pix2[y * XO_FDEC_STRIDE + x];
    for (int i = 0; i < 8; i++) { /* columns, in place */
        int s[8], o[8];
        for (int k = 0; k < 8; k++) s[k] = tmp[k][i];
        fwd8(s, o);
        for (int k = 0; k < 8; k++) tmp[k][i] = (int16_t)o[k];
    }
    for (int i = 0; i < 8; i++) { /* rows, written transposed */
        int s[8], o[8];
        for (int k = 0; k < 8; k++) s[k] = tmp[i][k];
        fwd8(s, o);
        for (int k = 0; k < 8; k++) dct[k * 8 + i] = (int16_t)o[k];
    }
}

/* S/common/dct.c:295-320 */
static inline void inv8(const int s[8], int o[8])
{
    int a0 = s[0] + s[4], a2 = s[0] - s[4], a4 = (s[2] >> 1) - s[6], a6 = (s[6] >> 1) + s[2];
    int b0 = a0 + a6, b2 = a2 + a4, b4 = a2 - a4, b6 = a0 - a6;
    int a1 = -s[3] + s[5] - s[7] - (s[7] >> 1);
    int a3 = s[1] + s[7] - s[3] - (s[3] >> 1);
    int a5 = -s[1] + s[7] + s[5] + (s[5] >> 1);
    int a7 = s[3] + s[5] + s[1] + (s[1] >> 1);
    int b1 = (a7 >> 2) + a1, b3 = a3 + (a5 >> 2), b5 = (a3 >> 2) - a5, b7 = a7 - (a1 >> 2);
    o[0] = b0 + b7; o[1] = b2 + b5; o[2] = b4 + b3; o[3] = b6 + b1;
    o[4] = b6 - b1; o[5] = b4 - b3; o[6] = b2 - b5; o[7] = b0 - b7;
}

/* S/common/dct.c:322-341 (modifies dct in place like the reference) */
void xo_add8x8_idct8(uint8_t *dst, int16_t dct[64])
{
    dct[0] += 32;
    for (int i = 0; i < 8; i++) {
        int s[8], o[8];
        for (int k = 0; k < 8; k++) s[k] = dct[k * 8 + i];
        inv8(s, o);
        for (int k = 0; k < 8; k++) dct[k * 8 + i] = (int16_t)o[k];
    }
    for (int i = 0; i < 8; i++) {
        int s[8], o[8];
        for (int k = 0; k < 8; k++) s[k] = dct[i * 8 + k];
        inv8(s, o);
        for (int k = 0; k < 8; k++)
            dst[i + k * XO_FDEC_STRIDE] = clip_u8(dst[i + k * XO_FDEC_STRIDE] + (o[k] >> 6));
    }
}

/* S/common/dct.c:39-105: 4x4 Hadamard of the luma DCs; forward rounds (x+1)>>1 */
static void hadamard_dc(int16_t d[16], int fwd)
{
    int16_t tmp[4][4];
    for (int i = 0; i < 4; i++) {
        int s01 = d[i * 4 + 0] + d[i * 4 + 1], d01 = d[i * 4 + 0] - d[i * 4 + 1];
        int s23 = d[i * 4 + 2] + d[i * 4 + 3], d23 = d[i * 4 + 2] - d[i * 4 + 3];
        tmp[0][i] = (int16_t)(s01 + s23); tmp[1][i] = (int16_t)(s01 - s23);
        tmp[2][i] = (int16_t)(d01 - d23); tmp[3][i] = (int16_t)(d01 + d23);
    }
    for (int i = 0; i < 4; i++) {
        int s01 = tmp[i][0] + tmp[i][1], d01 = tmp[i][0] - tmp[i][1];
        int s23 = tmp[i][2] + tmp[i][3], d23 = tmp[i][2] - tmp[i][3];
        int r = fwd ? 1 : 0, sh = fwd ? 1 : 0;
        d[i * 4 + 0] = (int16_t)((s01 + s23 + r) >> sh); d[i * 4 + 1] = (int16_t)((s01 - s23 + r) >> sh);
        d[i * 4 + 2] = (int16_t)((d01 - d23 + r) >> sh); d[i * 4 + 3] = (int16_t)((d01 + d23 + r) >> sh);
    }
}
void xo_dct4x4dc(int16_t d[16]) { hadamard_dc(d, 1); }
void xo_idct4x4dc(int16_t d[16]) { hadamard_dc(d, 0); }

/* S/common/dct.c:351-382: n=4 -> add8x8_idct_dc (2x2 dcs over an 8x8), n=16 -> add16x16_idct_dc */
void xo_add_idct_dc(uint8_t *dst, const int16_t *dc, int n)
{
    int per_row = n == 4 ? 2 : 4;
    for (int b = 0; b < n; b++) {
        uint8_t *p = dst + (b / per_row) * 4 * XO_FDEC_STRIDE + (b % per_row) * 4;
        int16_t v = (int16_t)((dc[b] + 32) >> 6);
        for (int y = 0; y < 4; y++)
            for (int x = 0; x < 4; x++)
                p[y * XO_FDEC_STRIDE + x] = clip_u8(p[y * XO_FDEC_STRIDE + x] + v);
    }
}

/* ---------------- quantiser tables: S/common/set.c:28-66 (constants), :68-174 ---------------- */
static const uint8_t dq4[6][3] = { { 10, 13, 16 }, { 11, 14, 18 }, { 13, 16, 20 }, { 14, 18, 23 }, { 16, 20, 25 }, { 18, 23, 29 } };
static const uint16_t q4[6][3] = { { 13107, 8066, 5243 }, { 11916, 7490, 4660 }, { 10082, 6554, 4194 },
                                   { 9362, 5825, 3647 },  { 8192, 5243, 3355 },  { 7282, 4559, 2893 } };
static const uint8_t q8scan[16] = { 0, 3, 4, 3, 3, 1, 5, 1, 4, 5, 2, 5, 3, 1, 5, 1 };
static const uint8_t dq8[6][6] = { { 20, 18, 32, 19, 25, 24 }, { 22, 19, 35, 21, 28, 26 }, { 26, 23, 42, 24, 33, 31 },
                                   { 28, 25, 45, 26, 35, 33 }, { 32, 28, 51, 30, 40, 38 }, { 36, 32, 58, 34, 46, 43 } };
static const uint16_t q8[6][6] = { { 13107, 11428, 20972, 12222, 16777, 15481 }, { 11916, 10826, 19174, 11058, 14980, 14290 },
                                   { 10082, 8943, 15978, 9675, 12710, 11985 },   { 9362, 8228, 14913, 8931, 11984, 11259 },
                                   { 8192, 7346, 13159, 7740, 10486, 9777 },     { 7282, 6428, 11570, 6830, 9118, 8640 } };
/* JVT matrices, S/common/set.h (x264_cqm_jvt4i/4p/8i/8p) */
static const uint8_t jvt4i[16] = { 6, 13, 20, 28, 13, 20, 28, 32, 20, 28, 32, 37, 28, 32, 37, 42 };
static const uint8_t jvt4p[16] = { 10, 14, 20, 24, 14, 20, 24, 27, 20, 24, 27, 30, 24, 27, 30, 34 };
static const uint8_t jvt8i[64] = { 6,  10, 13, 16, 18, 23, 25, 27, 10, 11, 16, 18, 23, 25, 27, 29, 13, 16, 18, 23, 25, 27,
                                   29, 31, 16, 18, 23, 25, 27, 29, 31, 33, 18, 23, 25, 27, 29, 31, 33, 36, 23, 25, 27, 29,
                                   31, 33, 36, 38, 25, 27, 29, 31, 33, 36, 38, 40, 27, 29, 31, 33, 36, 38, 40, 42 };
static const uint8_t jvt8p[64] = { 9,  13, 15, 17, 19, 21, 22, 24, 13, 13, 17, 19, 21, 22, 24, 25, 15, 17, 19, 21, 22, 24,
                                   25, 27, 17, 19, 21, 22, 24, 25, 27, 28, 19, 21, 22, 24, 25, 27, 28, 30, 21, 22, 24, 25,
                                   27, 28, 30, 32, 22, 24, 25, 27, 28, 30, 32, 33, 24, 25, 27, 28, 30, 32, 33, 35 };

static int scaling4(int cqm, int list, int i) { return cqm == 0 ? 16 : ((list & 1) ? jvt4p[i] : jvt4i[i]); }
static int scaling8(int cqm, int list, int i) { return cqm == 0 ? 16 : (list ? jvt8p[i] : jvt8i[i]); }
#define RDIV(n, d) (((n) + ((d) >> 1)) / (d))                                   /* set.c:25 DIV */
#define RSHIFT(x, s) ((s) < 0 ? (x) << -(s) : (s) == 0 ? (x) : ((x) + (1 << ((s)-1))) >> (s)) /* set.c:26 SHIFT */
static const int deadzone[4] = { 32 - 11, 32 - 21, 32 - 11, 32 - 21 }; /* set.c:77-79 with default dz 21 (inter) / 11 (intra) */

void xo_quant4_tables(int cqm, int list, int qp, uint16_t mf[16], uint16_t bias[16])
{
    for (int i = 0; i < 16; i++) {
        int k = (i & 1) + ((i >> 2) & 1);
        int base = RDIV(q4[qp % 6][k] * 16, scaling4(cqm, list, i));
        int j = RSHIFT(base, qp / 6 - 1);
        int b = RDIV(deadzone[list] << 10, j), cap = (1 << 15) / j;
        mf[i] = (uint16_t)j;
        bias[i] = (uint16_t)(b < cap ? b : cap);
    }
}

void xo_quant8_tables(int cqm, int list, int qp, uint16_t mf[64], uint16_t bias[64])
{
    for (int i = 0; i < 64; i++) {
        int k = q8scan[((i >> 1) & 12) | (i & 3)];
        int base = RDIV(q8[qp % 6][k] * 16, scaling8(cqm, list, i));
        int j = RSHIFT(base, qp / 6);
        int b = RDIV(deadzone[list] << 10, j), cap = (1 << 15) / j;
        mf[i] = (uint16_t)j;
        bias[i] = (uint16_t)(b < cap ? b : cap);
    }
}

void xo_dequant4_table(int cqm, int list, int dequant_mf[6][16])
{
    for (int q = 0; q < 6; q++)
        for (int i = 0; i < 16; i++)
            dequant_mf[q][i] = dq4[q][(i & 1) + ((i >> 2) & 1)] * scaling4(cqm, list, i);
}

void xo_dequant8_table(int cqm, int list, int dequant_mf[6][64])
{
    for (int q = 0; q < 6; q++)
        for (int i = 0; i < 64; i++)
            dequant_mf[q][i] = dq8[q][q8scan[((i >> 1) & 12) | (i & 3)]] * scaling8(cqm, list, i);
}

/* S/common/quant.c:33-74 */
static inline int quant_one(int16_t *c, int mf, int f)
{
    if (*c > 0) *c = (int16_t)((f + *c) * mf >> 16);
    else        *c = (int16_t)(-((f - *c) * mf >> 16));
    return *c;
}
int xo_quant_4x4(int16_t dct[16], const uint16_t mf[16], const uint16_t bias[16])
{
    int nz = 0;
    for (int i = 0; i < 16; i++) nz |= quant_one(&dct[i], mf[i], bias[i]);
    return !!nz;
}
int xo_quant_8x8(int16_t dct[64], const uint16_t mf[64], const uint16_t bias[64])
{
    int nz = 0;
    for (int i = 0; i < 64; i++) nz |= quant_one(&dct[i], mf[i], bias[i]);
    return !!nz;
}
int xo_quant_4x4_dc(int16_t dct[16], int mf, int bias)
{
    int nz = 0;
    for (int i = 0; i < 16; i++) nz |= quant_one(&dct[i], mf, bias);
    return !!nz;
}
int xo_quant_2x2_dc(int16_t dct[4], int mf, int bias)
{
    int nz = 0;
    for (int i = 0; i < 4; i++) nz |= quant_one(&dct[i], mf, bias);
    return !!nz;
}

/* S/common/quant.c:76-146 */
static void dequant_n(int16_t *dct, const int *mf, int n, int qbits)
{
    if (qbits >= 0)
        for (int i = 0; i < n; i++) dct[i] = (int16_t)((dct[i] * mf[i]) << qbits);
    else {
        int f = 1 << (-qbits - 1);
        for (int i = 0; i < n; i++) dct[i] = (int16_t)((dct[i] * mf[i] + f) >> (-qbits));
    }
}
void xo_dequant_4x4(int16_t dct[16], const int dequant_mf[6][16], int qp) { dequant_n(dct, dequant_mf[qp % 6], 16, qp / 6 - 4); }
void xo_dequant_8x8(int16_t dct[64], const int dequant_mf[6][64], int qp) { dequant_n(dct, dequant_mf[qp % 6], 64, qp / 6 - 6); }

/* S/common/quant.c:148-178 */
void xo_dequant_4x4_dc(int16_t dct[16], const int dequant_mf[6][16], int qp)
{
    int qbits = qp / 6 - 6;
    if (qbits >= 0) {
        int dmf = dequant_mf[qp % 6][0] << qbits;
        for (int i = 0; i < 16; i++) dct[i] = (int16_t)(dct[i] * dmf);
    } else {
        int dmf = dequant_mf[qp % 6][0], f = 1 << (-qbits - 1);
        for (int i = 0; i < 16; i++) dct[i] = (int16_t)((dct[i] * dmf + f) >> (-qbits));
    }
}
