/* xo_pixel.c — ORACLE (test infrastructure only): block metrics of S/common/pixel.c restated in
 * plain 32-bit integer arithmetic.  The reference packs two 16-bit lanes per uint32 (HADAMARD4/abs2,
 * pixel.c:164-181); for 8-bit input no lane can overflow (|H4 coef| <= 4080, |H8 coef| <= 16320, lane sums
 * < 65536), so the un-packed form below is the same function.  Pinned against the reference
 * itself by tests/test_oracle_vs_ref.py and the golden fixtures. */
#include <stdlib.h>
#include <string.h>
#include "xo.h"

static const int blk_w[7] = { 16, 16, 8, 8, 8, 4, 4 };
static const int blk_h[7] = { 16, 8, 16, 8, 4, 8, 4 };

/* S/common/pixel.c:40-65 */
static int sad_wxh(const uint8_t *a, int sa, const uint8_t *b, int sb, int w, int h)
{
    int sum = 0;
    for (int y = 0; y < h; y++, a += sa, b += sb)
        for (int x = 0; x < w; x++)
            sum += abs(a[x] - b[x]);
    return sum;
}

/* S/common/pixel.c:71-96 */
static int ssd_wxh(const uint8_t *a, int sa, const uint8_t *b, int sb, int w, int h)
{
    int sum = 0;
    for (int y = 0; y < h; y++, a += sa, b += sb)
        for (int x = 0; x < w; x++) {
            int d = a[x] - b[x];
            sum += d * d;
        }
    return sum;
}

/* un-normalised sum |H4 * D * H4^T| of one 4x4 difference block (pixel.c:187-210 without the >>1) */
static int hadamard4x4_abs_sum(const uint8_t *a, int sa, const uint8_t *b, int sb)
{
    int d[4][4], t[4][4], sum = 0;
    for (int y = 0; y < 4; y++)
        for (int x = 0; x < 4; x++)
            d[y][x] = a[y * sa + x] - b[y * sb + x];
    for (int y = 0; y < 4; y++) { /* rows */
        int s01 = d[y][0] + d[y][1], d01 = d[y][0] - d[y][1];
        int s23 = d[y][2] + d[y][3], d23 = d[y][2] - d[y][3];
        t[y][0] = s01 + s23; t[y][1] = s01 - s23; t[y][2] = d01 + d23; t[y][3] = d01 - d23;
    }
    for (int x = 0; x < 4; x++) { /* columns */
        int s01 = t[0][x] + t[1][x], d01 = t[0][x] - t[1][x];
        int s23 = t[2][x] + t[3][x], d23 = t[2][x] - t[3][x];
        sum += abs(s01 + s23) + abs(s01 - s23) + abs(d01 + d23) + abs(d01 - d23);
    }
    return sum;
}

/* S/common/pixel.c:187-253.  The reference halves once per 4x4 (satd_4x4) or once per 8x4 (satd_8x4):
 * sizes with w>=8 are tiled by 8x4 units, 4x8/4x4 by 4x4 units (PIXEL_SATD_C :233-253). */
static int satd_wxh(const uint8_t *a, int sa, const uint8_t *b, int sb, int w, int h)
{
    int sum = 0;
    if (w == 4) {
        for (int y = 0; y < h; y += 4)
            sum += hadamard4x4_abs_sum(a + y * sa, sa, b + y * sb, sb) >> 1;
        return sum;
    }
    for (int y = 0; y < h; y += 4)
        for (int x = 0; x < w; x += 8)
            sum += (hadamard4x4_abs_sum(a + y * sa + x, sa, b + y * sb + x, sb) +
                    hadamard4x4_abs_sum(a + y * sa + x + 4, sa, b + y * sb + x + 4, sb)) >> 1;
    return sum;
}

/* S/common/pixel.c:256-288: raw (un-rounded) 8x8 Hadamard magnitude sum */
static int sa8d_raw(const uint8_t *a, int sa, const uint8_t *b, int sb)
{
    int m[8][8], sum = 0;
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++)
            m[y][x] = a[y * sa + x] - b[y * sb + x];
    for (int pass = 0; pass < 2; pass++) {
        for (int i = 0; i < 8; i++) {
            int v[8];
            for (int k = 0; k < 8; k++) v[k] = pass ? m[k][i] : m[i][k];
            for (int step = 1; step < 8; step <<= 1) /* 3 butterfly stages = H8 up to output order */
                for (int k = 0; k < 8; k++)
                    if (!(k & step)) {
                        int p = v[k], q = v[k + step];
                        v[k] = p + q;
                        v[k + step] = p - q;
                    }
            for (int k = 0; k < 8; k++)
                if (pass) m[k][i] = v[k]; else m[i][k] = v[k];
        }
    }
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++)
            sum += abs(m[y][x]);
    return sum;
}

/* S/common/pixel.c:290-303 */
static int sa8d_wxh(const uint8_t *a, int sa, const uint8_t *b, int sb, int w, int h)
{
    int sum = 0;
    for (int y = 0; y < h; y += 8)
        for (int x = 0; x < w; x += 8)
            sum += sa8d_raw(a + y * sa + x, sa, b + y * sb + x, sb);
    return (sum + 2) >> 2;
}

int xo_pixel_cmp(int metric, int i_pixel, const uint8_t *p1, int s1, const uint8_t *p2, int s2)
{
    int w = blk_w[i_pixel], h = blk_h[i_pixel];
    switch (metric) {
    case XO_SAD:  return sad_wxh(p1, s1, p2, s2, w, h);
    case XO_SSD:  return ssd_wxh(p1, s1, p2, s2, w, h);
    case XO_SATD: return satd_wxh(p1, s1, p2, s2, w, h);
    case XO_SA8D: return sa8d_wxh(p1, s1, p2, s2, w, h); /* only 16x16 and 8x8 exist, pixel.c:607-608 */
    }
    return -1;
}

/* S/common/pixel.c:142-161 */
int xo_pixel_var(int i_pixel, const uint8_t *pix, int stride)
{
    int w = i_pixel == XO_16x16 ? 16 : 8, shift = i_pixel == XO_16x16 ? 8 : 6;
    uint32_t sum = 0, sqr = 0;
    for (int y = 0; y < w; y++, pix += stride)
        for (int x = 0; x < w; x++) {
            sum += pix[x];
            sqr += pix[x] * pix[x];
        }
    return (int)(sqr - (sum * sum >> shift));
}

/* S/common/pixel.c:306-358 */
static void hadamard_ac_8x8(const uint8_t *pix, int stride, int *sum4, int *sum8)
{
    int m[8][8], s4 = 0, s8 = 0, dc;
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++)
            m[y][x] = pix[y * stride + x];
    /* 4-point Hadamard on each 4x4 quadrant, both directions */
    for (int pass = 0; pass < 2; pass++)
        for (int i = 0; i < 8; i++)
            for (int half = 0; half < 8; half += 4) {
                int v[4];
                for (int k = 0; k < 4; k++) v[k] = pass ? m[half + k][i] : m[i][half + k];
                int s01 = v[0] + v[1], d01 = v[0] - v[1], s23 = v[2] + v[3], d23 = v[2] - v[3];
                v[0] = s01 + s23; v[1] = s01 - s23; v[2] = d01 + d23; v[3] = d01 - d23;
                for (int k = 0; k < 4; k++)
                    if (pass) m[half + k][i] = v[k]; else m[i][half + k] = v[k];
            }
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++)
            s4 += abs(m[y][x]);
    /* one more butterfly level across the quadrants -> 8x8 Hadamard magnitudes */
    for (int y = 0; y < 4; y++)
        for (int x = 0; x < 4; x++) {
            int a = m[y][x], b = m[y][x + 4], c = m[y + 4][x], d = m[y + 4][x + 4];
            s8 += abs(a + b + c + d) + abs(a - b + c - d) + abs(a + b - c - d) + abs(a - b - c + d);
        }
    dc = m[0][0] + m[0][4] + m[4][0] + m[4][4];
    *sum4 = s4 - dc;
    *sum8 = s8 - dc;
}

uint64_t xo_pixel_hadamard_ac(int i_pixel, const uint8_t *pix, int stride)
{
    int w = blk_w[i_pixel], h = blk_h[i_pixel];
    uint64_t sum = 0;
    for (int y = 0; y < h; y += 8)
        for (int x = 0; x < w; x += 8) {
            int s4, s8;
            hadamard_ac_8x8(pix + y * stride + x, stride, &s4, &s8);
            sum += ((uint64_t)(uint32_t)s8 << 32) + (uint32_t)s4;
        }
    return ((sum >> 34) << 32) + ((uint32_t)sum >> 1);
}

/* S/common/pixel.c:515-559 */
int xo_pixel_ads(int i_pixel, const int enc_dc[4], const uint16_t *sums, int delta, const uint16_t *cost_mvx,
                 int16_t *mvs, int width, int thresh)
{
    /* ads4 for 16x16; ads2 for 16x8, 8x16, 8x4, 4x8; ads1 for 8x8, 4x4 (pixel.c:591-594, 793-796) */
    int terms = i_pixel == XO_16x16 ? 4 : (i_pixel == XO_8x8 || i_pixel == XO_4x4) ? 1 : 2;
    int n = 0;
    for (int i = 0; i < width; i++) {
        int ads = abs(enc_dc[0] - sums[i]) + cost_mvx[i];
        if (terms == 2)
            ads += abs(enc_dc[1] - sums[i + delta]);
        else if (terms == 4)
            ads += abs(enc_dc[1] - sums[i + 8]) + abs(enc_dc[2] - sums[i + delta]) + abs(enc_dc[3] - sums[i + delta + 8]);
        if (ads < thresh)
            mvs[n++] = i;
    }
    return n;
}
