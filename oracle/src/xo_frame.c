/* xo_frame.c — ORACLE (test infrastructure only): frame geometry, border replication, the 6-tap half-pel
 * planes, the integral image and the half-resolution lookahead planes, restated whole-frame.
 * Follows S/common/frame.c and S/common/mc.c (line cites per function). */
#include <stdlib.h>
#include <string.h>
#include "xo.h"

#define ALIGN_UP(x, a) (((x) + ((a) - 1)) & ~((a) - 1))

static inline uint8_t clip_u8(int x) { return x < 0 ? 0 : x > 255 ? 255 : x; } /* S/common/common.h:104-107 */

/* S/common/frame.c:44-59,82-90 with cpu=0 (align=16), progressive */
void xo_geometry(int width, int height, xo_geom *g)
{
    memset(g, 0, sizeof(*g));
    g->width = width;
    g->height = height;
    g->mb_width = (width + 15) / 16;
    g->mb_height = (height + 15) / 16;
    int w16 = g->mb_width * 16;
    g->stride = ALIGN_UP(ALIGN_UP(w16 + 2 * XO_PADH, 16), 16);
    g->lines = g->mb_height * 16;
    g->plane_size = g->stride * (g->lines + 2 * XO_PADV);
    g->origin = g->stride * XO_PADV + XO_PADH;
    g->width_lowres = w16 / 2;
    g->stride_lowres = ALIGN_UP(g->width_lowres + 2 * XO_PADH, 16);
    g->lines_lowres = g->lines / 2;
    g->plane_size_lowres = g->stride_lowres * (g->lines_lowres + 2 * XO_PADV);
    g->origin_lowres = g->stride_lowres * XO_PADV + XO_PADH;
}

/* S/common/frame.c:218-238 */
static void expand_border(uint8_t *pix, int stride, int width, int height, int padh, int padv)
{
    for (int y = 0; y < height; y++) {
        uint8_t *row = pix + y * stride;
        memset(row - padh, row[0], padh);
        memset(row + width, row[width - 1], padh);
    }
    for (int y = 1; y <= padv; y++) {
        memcpy(pix - padh - y * stride, pix - padh, width + 2 * padh);
        memcpy(pix - padh + (height - 1 + y) * stride, pix - padh + (height - 1) * stride, width + 2 * padh);
    }
}

/* S/common/frame.c:304-331 (mod16 padding of the input picture) then :240-267 (32-px replication), luma */
void xo_frame_expand_border(const xo_geom *g, uint8_t *plane)
{
    int w16 = g->mb_width * 16;
    if (w16 > g->width)
        for (int y = 0; y < g->height; y++)
            memset(plane + y * g->stride + g->width, plane[y * g->stride + g->width - 1], w16 - g->width);
    for (int y = g->height; y < g->lines; y++)
        memcpy(plane + y * g->stride, plane + (g->height - 1) * g->stride, w16);
    expand_border(plane, g->stride, w16, g->lines, XO_PADH, XO_PADV);
}

static inline int tap6(int a, int b, int c, int d, int e, int f) { return a + f - 5 * (b + e) + 20 * (c + d); }

/* S/common/mc.c:133-155 driven as S/common/mc.c:404-426 does for (mb_y=0, b_end=1): rows [-8, lines+8),
 * columns [-8, width+8); then S/common/frame.c:269-295 replicates from column -4 / width+4 and row -8 /
 * lines+8 outward.  Intermediate "buf" is int16_t in the reference (fits: |v| <= 255*52). */
void xo_frame_filter(const xo_geom *g, const uint8_t *plane, uint8_t *dsth, uint8_t *dstv, uint8_t *dstc,
                     uint16_t *integral, int b_sub8x8)
{
    const int stride = g->stride, w16 = g->mb_width * 16, lines = g->lines;
    const int x0 = -8, x1 = w16 + 8; /* hpel_filter is called with offs=-8 cols and width+16 */
    int16_t *buf = malloc((w16 + 16 + 5 + 8) * sizeof(int16_t));
    for (int y = -8; y < lines + 8; y++) {
        const uint8_t *s = plane + y * stride;
        /* vertical filter for x in [x0-2, x1+3) -> buf[x - (x0-2)] */
        for (int x = x0 - 2; x < x1 + 3; x++) {
            int v = tap6(s[x - 2 * stride], s[x - stride], s[x], s[x + stride], s[x + 2 * stride], s[x + 3 * stride]);
            dstv[y * stride + x] = clip_u8((v + 16) >> 5);
            buf[x - (x0 - 2)] = (int16_t)v;
        }
        for (int x = x0; x < x1; x++) {
            const int16_t *b = buf + (x - (x0 - 2));
            dstc[y * stride + x] = clip_u8((tap6(b[-2], b[-1], b[0], b[1], b[2], b[3]) + 512) >> 10);
            dsth[y * stride + x] = clip_u8((tap6(s[x - 2], s[x - 1], s[x], s[x + 1], s[x + 2], s[x + 3]) + 16) >> 5);
        }
    }
    free(buf);
    /* frame.c:269-295: pix = filtered[i] + (0 - 8)*stride - 4 ; width = 16*mb_w + 8 ; height = lines+16 ;
     * padh = PADH-4, padv = PADV-8 */
    uint8_t *planes[3] = { dsth, dstv, dstc };
    for (int i = 0; i < 3; i++)
        expand_border(planes[i] - 8 * stride - 4, stride, w16 + 8, lines + 16, XO_PADH - 4, XO_PADV - 8);

    if (!integral)
        return;
    /* S/common/mc.c:428-461 with start=-PADV .. height = lines+8+PADV-9 = lines+PADV-1.
     * Row addressing is relative to `integral` == element (0,0); the h-pass of picture row y writes row y+1,
     * spanning columns [-PADH, stride-PADH-8 (or -4)).  uint16 wrap-around is intended. */
    memset(integral - XO_PADV * stride - XO_PADH, 0, stride * sizeof(uint16_t));
    for (int y = -XO_PADV; y < lines + XO_PADV - 1; y++) {
        const uint8_t *pix = plane + y * stride - XO_PADH;
        uint16_t *sum8 = integral + (y + 1) * stride - XO_PADH;
        if (b_sub8x8) {
            /* integral_init4h, mc.c:270-279 */
            int v = pix[0] + pix[1] + pix[2] + pix[3];
            for (int x = 0; x < stride - 4; x++) {
                sum8[x] = (uint16_t)(v + sum8[x - stride]);
                v += pix[x + 4] - pix[x];
            }
            sum8 -= 8 * stride;
            uint16_t *sum4 = sum8 + stride * (lines + XO_PADV * 2);
            if (y >= 8 - XO_PADV) {
                /* integral_init4v, mc.c:292-299 */
                for (int x = 0; x < stride - 8; x++)
                    sum4[x] = (uint16_t)(sum8[x + 4 * stride] - sum8[x]);
                for (int x = 0; x < stride - 8; x++)
                    sum8[x] = (uint16_t)(sum8[x + 8 * stride] + sum8[x + 8 * stride + 4] - sum8[x] - sum8[x + 4]);
            }
        } else {
            /* integral_init8h, mc.c:281-290 */
            int v = 0;
            for (int k = 0; k < 8; k++) v += pix[k];
            for (int x = 0; x < stride - 8; x++) {
                sum8[x] = (uint16_t)(v + sum8[x - stride]);
                v += pix[x + 8] - pix[x];
            }
            if (y >= 8 - XO_PADV) {
                /* integral_init8v, mc.c:301-306 */
                uint16_t *s = sum8 - 8 * stride;
                for (int x = 0; x < stride - 8; x++)
                    s[x] = (uint16_t)(s[x + 8 * stride] - s[x]);
            }
        }
    }
}

/* S/common/mc.c:306-357 + S/common/frame.c:297-302 */
void xo_frame_init_lowres(const xo_geom *g, uint8_t *plane, uint8_t *l0, uint8_t *lh, uint8_t *lv, uint8_t *lc)
{
    const int stride = g->stride, w16 = g->mb_width * 16, lines = g->lines, ls = g->stride_lowres;
    /* duplicate last column / row (mc.c:315-317) */
    for (int y = 0; y < lines; y++)
        plane[w16 + y * stride] = plane[w16 - 1 + y * stride];
    memcpy(plane + stride * lines, plane + stride * (lines - 1), w16);
#define AVG2(a, b) (((a) + (b) + 1) >> 1)
#define LOWRES_TAP(a, b, c, d) AVG2(AVG2(a, b), AVG2(c, d))
    for (int y = 0; y < g->lines_lowres; y++) {
        const uint8_t *r0 = plane + 2 * y * stride, *r1 = r0 + stride, *r2 = r1 + stride;
        for (int x = 0; x < g->width_lowres; x++) {
            l0[y * ls + x] = LOWRES_TAP(r0[2 * x], r1[2 * x], r0[2 * x + 1], r1[2 * x + 1]);
            lh[y * ls + x] = LOWRES_TAP(r0[2 * x + 1], r1[2 * x + 1], r0[2 * x + 2], r1[2 * x + 2]);
            lv[y * ls + x] = LOWRES_TAP(r1[2 * x], r2[2 * x], r1[2 * x + 1], r2[2 * x + 1]);
            lc[y * ls + x] = LOWRES_TAP(r1[2 * x + 1], r2[2 * x + 1], r1[2 * x + 2], r2[2 * x + 2]);
        }
    }
    /* frame.c:297-302: width passed is stride_lowres - 2*PADH (NOT width_lowres) */
    uint8_t *p[4] = { l0, lh, lv, lc };
    for (int i = 0; i < 4; i++)
        expand_border(p[i], ls, ls - 2 * XO_PADH, g->lines_lowres, XO_PADH, XO_PADV);
}

/* S/common/mc.c:157-202 (mc_luma; get_ref returns the same samples) */
void xo_mc_luma(uint8_t *dst, int dst_stride, const uint8_t *const src[4], int src_stride, int mvx, int mvy, int w, int h)
{
    static const int ref0[16] = { 0, 1, 1, 1, 0, 1, 1, 1, 2, 3, 3, 3, 0, 1, 1, 1 };
    static const int ref1[16] = { 0, 0, 0, 0, 2, 2, 3, 2, 2, 2, 3, 2, 2, 2, 3, 2 };
    int qidx = ((mvy & 3) << 2) + (mvx & 3);
    int offset = (mvy >> 2) * src_stride + (mvx >> 2);
    const uint8_t *s1 = src[ref0[qidx]] + offset + ((mvy & 3) == 3) * src_stride;
    if (qidx & 5) {
        const uint8_t *s2 = src[ref1[qidx]] + offset + ((mvx & 3) == 3);
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++)
                dst[y * dst_stride + x] = (s1[y * src_stride + x] + s2[y * src_stride + x] + 1) >> 1;
    } else {
        for (int y = 0; y < h; y++)
            memcpy(dst + y * dst_stride, s1 + y * src_stride, w);
    }
}

/* S/common/mc.c:205-236 */
void xo_mc_chroma(uint8_t *dst, int dst_stride, const uint8_t *src, int src_stride, int mvx, int mvy, int w, int h)
{
    int dx = mvx & 7, dy = mvy & 7;
    int cA = (8 - dx) * (8 - dy), cB = dx * (8 - dy), cC = (8 - dx) * dy, cD = dx * dy;
    src += (mvy >> 3) * src_stride + (mvx >> 3);
    for (int y = 0; y < h; y++, src += src_stride, dst += dst_stride)
        for (int x = 0; x < w; x++)
            dst[x] = (cA * src[x] + cB * src[x + 1] + cC * src[x + src_stride] + cD * src[x + src_stride + 1] + 32) >> 6;
}


/* mc.c:52-125 (PIXEL_AVG_C over pixel_avg_wxh / pixel_avg_weight_wxh) */
void xo_pixel_avg(int i_pixel, uint8_t *dst, int dst_stride, const uint8_t *a, int a_stride, const uint8_t *b, int b_stride, int weight)
{
    static const int ws[10] = { 16, 16, 8, 8, 8, 4, 4, 4, 2, 2 }, hs[10] = { 16, 8, 16, 8, 4, 8, 4, 2, 4, 2 };
    const int w = ws[i_pixel], h = hs[i_pixel];
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const int p = a[y * a_stride + x], q = b[y * b_stride + x];
            int v = weight == 32 ? (p + q + 1) >> 1 : (p * weight + q * (64 - weight) + 32) >> 6;
            dst[y * dst_stride + x] = v < 0 ? 0 : v > 255 ? 255 : v;
        }
}
