/* xo_metrics.c — ORACLE (test infrastructure only): whole-frame analysis metrics that have no inter-macroblock dependency.
 *   xo_frame_ssd            x264_pixel_ssd_wxh                S/common/pixel.c:98-136   (PSNR input, encoder.c:1034-1046)
 *   xo_frame_mb_energy      ac_energy_mb                      S/encoder/ratecontrol.c:171-191 (adaptive quantisation)
 *   xo_frame_mb_hadamard_ac x264_pixel_hadamard_ac_16x16      S/common/pixel.c:306-358  (psy-rd fenc energy)
 *   xo_frame_ssim           x264_pixel_ssim_wxh               S/common/pixel.c:435-509  (integer 4x4 sums + the float tail) */
#include <stdlib.h>
#include <string.h>
#include "xo.h"

int64_t xo_frame_ssd(const uint8_t *p1, int s1, const uint8_t *p2, int s2, int width, int height)
{
    /* the reference tiles the region with 16x16 / 8x16 / 8x8 blocks plus per-pixel remainders; the sum is the same */
    int64_t ssd = 0;
    for (int y = 0; y < height; y++)
        for (int x = 0; x < width; x++) {
            const int d = p1[y * s1 + x] - p2[y * s2 + x];
            ssd += d * d;
        }
    return ssd;
}

void xo_frame_mb_energy(const xo_geom *g, const uint8_t *py, const uint8_t *pu, const uint8_t *pv, int stride_c, uint32_t *out)
{
    for (int mb_y = 0; mb_y < g->mb_height; mb_y++)
        for (int mb_x = 0; mb_x < g->mb_width; mb_x++) {
            uint32_t var = (uint32_t)xo_pixel_var(XO_16x16, py + 16 * (mb_x + mb_y * g->stride), g->stride);
            var += (uint32_t)xo_pixel_var(XO_8x8, pu + 8 * (mb_x + mb_y * stride_c), stride_c);
            var += (uint32_t)xo_pixel_var(XO_8x8, pv + 8 * (mb_x + mb_y * stride_c), stride_c);
            out[mb_x + mb_y * g->mb_width] = var > 1 ? var : 1; /* X264_MAX(var,1) on unsigned int */
        }
}

void xo_frame_mb_hadamard_ac(const xo_geom *g, const uint8_t *py, uint64_t *out)
{
    for (int mb_y = 0; mb_y < g->mb_height; mb_y++)
        for (int mb_x = 0; mb_x < g->mb_width; mb_x++)
            out[mb_x + mb_y * g->mb_width] = xo_pixel_hadamard_ac(XO_16x16, py + 16 * (mb_x + mb_y * g->stride), g->stride);
}

/* ssim_4x4x2_core (pixel.c:435-460) for every 4x4 block: sums[y4][x4] = { s1, s2, ss, s12 } */
void xo_frame_ssim_sums(const uint8_t *p1, int s1, const uint8_t *p2, int s2, int width, int height, int (*sums)[4])
{
    const int w4 = width >> 2, h4 = height >> 2;
    for (int by = 0; by < h4; by++)
        for (int bx = 0; bx < w4; bx++) {
            uint32_t a1 = 0, a2 = 0, ss = 0, s12 = 0;
            for (int y = 0; y < 4; y++)
                for (int x = 0; x < 4; x++) {
                    const int a = p1[(4 * by + y) * s1 + 4 * bx + x], b = p2[(4 * by + y) * s2 + 4 * bx + x];
                    a1 += a; a2 += b; ss += a * a; ss += b * b; s12 += a * b;
                }
            int *o = sums[by * w4 + bx];
            o[0] = a1; o[1] = a2; o[2] = ss; o[3] = s12;
        }
}

/* ssim_end1 / ssim_end4 / x264_pixel_ssim_wxh (pixel.c:462-509): the float tail over the integer sums, same grouping and
 * accumulation order as the reference (windows of 2x2 blocks, four at a time) */
static float ssim_end1(int s1, int s2, int ss, int s12)
{
    static const int ssim_c1 = (int)(.01 * .01 * 255 * 255 * 64 + .5);
    static const int ssim_c2 = (int)(.03 * .03 * 255 * 255 * 64 * 63 + .5);
    int vars = ss * 64 - s1 * s1 - s2 * s2;
    int covar = s12 * 64 - s1 * s2;
    return (float)(2 * s1 * s2 + ssim_c1) * (float)(2 * covar + ssim_c2) / ((float)(s1 * s1 + s2 * s2 + ssim_c1) * (float)(vars + ssim_c2));
}
float xo_ssim_from_sums(const int (*sums)[4], int w4, int h4)
{
    float ssim = 0.0;
    for (int y = 1; y < h4; y++) {
        const int (*r0)[4] = sums + (size_t)y * w4, (*r1)[4] = sums + (size_t)(y - 1) * w4;
        for (int x = 0; x < w4 - 1; x += 4) {
            const int n = 4 < w4 - x - 1 ? 4 : w4 - x - 1;
            float part = 0.0;
            for (int i = 0; i < n; i++)
                part += ssim_end1(r0[x + i][0] + r0[x + i + 1][0] + r1[x + i][0] + r1[x + i + 1][0], r0[x + i][1] + r0[x + i + 1][1] + r1[x + i][1] + r1[x + i + 1][1],
                                  r0[x + i][2] + r0[x + i + 1][2] + r1[x + i][2] + r1[x + i + 1][2], r0[x + i][3] + r0[x + i + 1][3] + r1[x + i][3] + r1[x + i + 1][3]);
            ssim += part;
        }
    }
    return ssim;
}
float xo_frame_ssim(const uint8_t *p1, int s1, const uint8_t *p2, int s2, int width, int height)
{
    const int w4 = width >> 2, h4 = height >> 2;
    int (*sums)[4] = malloc((size_t)w4 * h4 * sizeof(*sums));
    xo_frame_ssim_sums(p1, s1, p2, s2, width, height, sums);
    const float r = xo_ssim_from_sums((const int (*)[4])sums, w4, h4);
    free(sums);
    return r;
}

/* x264_adaptive_quant_frame (ratecontrol.c:233-249): f_qp_offset and i_inv_qscale_factor from the macroblock energies.
 * The two tables are log2(1 + i/128) to five decimals and (2^((i+.5)/64) - 1) * 256 rounded (ratecontrol.c:193-219). */
#include <math.h>
static float log2_lut[128];
static uint8_t exp2_lut[64];
static void aq_tables(void)
{
    if (exp2_lut[63]) return;
    for (int i = 0; i < 128; i++) log2_lut[i] = (float)(floor(log2(1.0 + i / 128.0) * 1e5 + 0.5) / 1e5);
    for (int i = 0; i < 64; i++) exp2_lut[i] = (uint8_t)floor((pow(2.0, (i + 0.5) / 64.0) - 1.0) * 256.0 + 0.5);
}
static int exp2fix8(float x)
{
    int i, f;
    x += 8;
    if (x <= 0) return 0;
    if (x >= 16) return 0xffff;
    i = x;
    f = (x - i) * 64;
    return (exp2_lut[f] + 256) << i >> 8;
}
void xo_aq_from_energy(const uint32_t *energy, int n, float aq_strength, float *qp_offset, uint16_t *inv_qscale)
{
    aq_tables();
    const float strength = aq_strength * 1.0397;
    for (int k = 0; k < n; k++) {
        const uint32_t e = energy[k];
        const int lz = __builtin_clz(e);
        const float qp_adj = strength * (log2_lut[(e << lz >> 24) & 0x7f] - lz + 16.573f);
        qp_offset[k] = qp_adj;
        inv_qscale[k] = exp2fix8(qp_adj * (-1.f / 6.f));
    }
}

void xo_frame_aq(const xo_geom *g, const uint8_t *py, const uint8_t *pu, const uint8_t *pv, int stride_c, float aq_strength, float *qp_offset,
                 uint16_t *inv_qscale)
{
    const int n = g->mb_width * g->mb_height;
    uint32_t *e = malloc((size_t)n * sizeof(*e));
    xo_frame_mb_energy(g, py, pu, pv, stride_c, e);
    xo_aq_from_energy(e, n, aq_strength, qp_offset, inv_qscale);
    free(e);
}
