/* cuda_stub.c — TEST INFRASTRUCTURE ONLY.  A CPU stand-in for the few x264_cuda_* entry points integration/x264_b200_hooks.c calls,
 * served by the oracle restatement (oracle/src, xo_*).  It exists so that the HOST logic of the performance-mode integration (grid
 * bookkeeping, predictor stage, raster argmin, sub-pel stage, deferred end-of-frame pass, deferred PSNR/SSIM) can be checked for
 * byte-identical bitstreams in a container without a GPU (tests/test_integration_host.py, `-m "not gpu"`).  It is linked only into
 * oracle/_ref/x264_b200_stub; the product binary integration/_build/x264_b200 links libx264_cuda.so and nothing from oracle/. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "x264_cuda.h"
#include "src/xo.h"

struct x264_cuda_t { char err[256]; long long launches; };
struct x264_cuda_frame_t {
    xo_geom g; int flags;
    uint8_t *buf[4], *plane[4];
    uint8_t *cbuf[2], *chroma[2]; int stride_c, lines_c, width_c;
    uint16_t *ibuf, *integral;
};

int x264_cuda_open(x264_cuda_t **ctx, int device) { (void)device; *ctx = calloc(1, sizeof(**ctx)); return 0; }
void x264_cuda_close(x264_cuda_t *ctx) { free(ctx); }
int x264_cuda_synchronize(x264_cuda_t *ctx) { (void)ctx; return 0; }
const char *x264_cuda_error(const x264_cuda_t *ctx) { return ctx ? ctx->err : "stub"; }
long long x264_cuda_launch_count(const x264_cuda_t *ctx) { return ctx->launches; }
void *x264_cuda_host_alloc(size_t bytes) { return malloc(bytes); }
void x264_cuda_host_free(void *p) { free(p); }
void *x264_cuda_fence_record(x264_cuda_t *ctx) { (void)ctx; return (void *)1; }
int x264_cuda_fence_wait(x264_cuda_t *ctx, void *fence) { (void)ctx; (void)fence; return 0; }
int x264_cuda_set_cost_mv(x264_cuda_t *ctx, int qp, const int16_t *table) { (void)ctx; (void)qp; (void)table; return 0; }

x264_cuda_frame_t *x264_cuda_frame_new(x264_cuda_t *ctx, int width, int height, int flags)
{
    (void)ctx;
    x264_cuda_frame_t *f = calloc(1, sizeof(*f));
    xo_geometry(width, height, &f->g);
    f->flags = flags;
    for (int i = 0; i < 4; i++) { f->buf[i] = calloc(1, f->g.plane_size + 64); f->plane[i] = f->buf[i] + f->g.origin; }
    f->stride_c = f->g.stride / 2; f->lines_c = f->g.lines / 2; f->width_c = f->g.mb_width * 8;
    for (int i = 0; i < 2; i++) { f->cbuf[i] = calloc(1, (size_t)f->stride_c * (f->lines_c + 32) + 64); f->chroma[i] = f->cbuf[i] + f->stride_c * 16 + 16; }
    f->ibuf = calloc(2, (size_t)f->g.plane_size * 2 + 64);
    f->integral = f->ibuf + f->g.origin;
    return f;
}
void x264_cuda_frame_delete(x264_cuda_frame_t *f)
{
    if (!f) return;
    for (int i = 0; i < 4; i++) free(f->buf[i]);
    free(f->cbuf[0]); free(f->cbuf[1]); free(f->ibuf); free(f);
}
int x264_cuda_frame_upload(x264_cuda_t *ctx, x264_cuda_frame_t *f, const uint8_t *src, int src_stride, int cols, int rows)
{
    (void)ctx;
    for (int y = 0; y < rows; y++) memcpy(f->plane[0] + (size_t)y * f->g.stride, src + (size_t)y * src_stride, cols);
    return 0;
}
int x264_cuda_frame_upload_chroma(x264_cuda_t *ctx, x264_cuda_frame_t *f, int plane, const uint8_t *src, int src_stride, int cols, int rows)
{
    (void)ctx;
    uint8_t *d = f->chroma[plane - X264_CUDA_PLANE_CB];
    for (int y = 0; y < rows; y++) memcpy(d + (size_t)y * f->stride_c, src + (size_t)y * src_stride, cols);
    return 0;
}
int x264_cuda_frame_download(x264_cuda_t *ctx, const x264_cuda_frame_t *f, int plane, void *dst, int dst_stride)
{
    (void)ctx;
    if (plane < 4) {
        const uint8_t *s = f->plane[plane] - f->g.origin;
        const int cols = f->g.mb_width * 16 + 64;
        for (int y = 0; y < f->g.lines + 64; y++) memcpy((uint8_t *)dst + (size_t)y * dst_stride, s + (size_t)y * f->g.stride, cols);
    } else if (plane == X264_CUDA_PLANE_CB || plane == X264_CUDA_PLANE_CR) {
        const uint8_t *s = f->chroma[plane - X264_CUDA_PLANE_CB] - (f->stride_c * 16 + 16);
        for (int y = 0; y < f->lines_c + 32; y++) memcpy((uint8_t *)dst + (size_t)y * dst_stride, s + (size_t)y * f->stride_c, f->width_c + 32);
    } else if (plane == X264_CUDA_PLANE_INTEGRAL || plane == X264_CUDA_PLANE_INTEGRAL4) {
        const uint16_t *s = f->integral - f->g.origin + (plane == X264_CUDA_PLANE_INTEGRAL4 ? f->g.plane_size : 0);
        const int cols = f->g.mb_width * 16 + 64;
        for (int y = 0; y < f->g.lines + 64; y++) memcpy((uint16_t *)dst + (size_t)y * dst_stride, s + (size_t)y * f->g.stride, cols * 2);
    } else
        return -1;
    return 0;
}
static void expand_plane(uint8_t *p, int stride, int w, int h, int padh, int padv) /* S/common/frame.c:218-238 */
{
    for (int y = 0; y < h; y++) {
        memset(p + (size_t)y * stride - padh, p[(size_t)y * stride], padh);
        memset(p + (size_t)y * stride + w, p[(size_t)y * stride + w - 1], padh);
    }
    for (int y = 0; y < padv; y++) {
        memcpy(p - (size_t)(y + 1) * stride - padh, p - padh, w + 2 * padh);
        memcpy(p + (size_t)(h + y) * stride - padh, p + (size_t)(h - 1) * stride - padh, w + 2 * padh);
    }
}
int x264_cuda_frame_expand_border(x264_cuda_t *ctx, x264_cuda_frame_t *f)
{
    ctx->launches++;
    xo_frame_expand_border(&f->g, f->plane[0]);
    for (int i = 0; i < 2; i++) expand_plane(f->chroma[i], f->stride_c, f->width_c, f->lines_c, 16, 16);
    return 0;
}
int x264_cuda_frame_filter(x264_cuda_t *ctx, x264_cuda_frame_t *f)
{
    ctx->launches++;
    xo_frame_filter(&f->g, f->plane[0], f->plane[1], f->plane[2], f->plane[3], (f->flags & X264_CUDA_FRAME_INTEGRAL) ? f->integral : NULL,
                    !!(f->flags & X264_CUDA_FRAME_INTEGRAL4));
    return 0;
}
int x264_cuda_frame_deblock(x264_cuda_t *ctx, x264_cuda_frame_t *f, const x264_cuda_deblock_params_t *p, const int8_t *type, const int8_t *qp,
                            const int8_t *transform8x8, const uint8_t (*nnz)[24], const int8_t *ref0, const int16_t (*mv0)[2], const int8_t *ref1,
                            const int16_t (*mv1)[2])
{
    ctx->launches++;
    xo_deblock_in d = { p->alpha_c0_offset, p->beta_offset, p->chroma_qp_offset, p->b_slice_b, p->b_psub8x8, p->b_cavlc_8x8dct, type, qp, transform8x8, nnz,
                        { ref0, ref1 }, { mv0, mv1 } };
    xo_frame_deblock(&f->g, &d, f->plane[0], f->chroma[0], f->chroma[1], f->stride_c);
    return 0;
}
int x264_cuda_sad_grid_quad(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, const x264_cuda_frame_t *fref, int radius, const x264_cuda_grid_job_t *jobs,
                            int n_jobs, uint16_t *grid, int async)
{
    (void)async;
    ctx->launches++;
    const int GW = X264_CUDA_GRID_W(radius), GH = X264_CUDA_GRID_H(radius), s = fenc->g.stride;
    for (int n = 0; n < n_jobs; n++) {
        const x264_cuda_grid_job_t *j = &jobs[n];
        uint16_t *out = grid + (size_t)n * GW * GH * 4;
        for (int r = 0; r < GH; r++)
            for (int c = 0; c < GW; c++) {
                const int mx = j->cx - radius + c, my = j->cy - radius + r;
                uint16_t *o = out + ((size_t)r * GW + c) * 4;
                if (mx < j->mv_min_fpel[0] || mx > j->mv_max_fpel[0] + 3 || my < j->mv_min_fpel[1] || my > j->mv_max_fpel[1]) { o[0] = o[1] = o[2] = o[3] = 0xffff; continue; }
                for (int q = 0; q < 4; q++) {
                    const int x = j->mb_x * 16 + (q & 1) * 8, y = j->mb_y * 16 + (q >> 1) * 8;
                    o[q] = (uint16_t)xo_pixel_cmp(XO_SAD, XO_8x8, fenc->plane[0] + (size_t)y * s + x, s, fref->plane[0] + (size_t)(y + my) * s + x + mx, s);
                }
            }
    }
    return 0;
}
int x264_cuda_me_search(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, const x264_cuda_frame_t *fref, int me_range, const x264_cuda_me_job_t *jobs, int n_jobs,
                        x264_cuda_me_result_t *results)
{ /* seeded exhaustive search of single blocks (sub-8x8 partitions): plain raster argmin, S/encoder/me.c:459-465 */
    static int16_t *tab[52];
    static const int bw[7] = { 16, 16, 8, 8, 8, 4, 4 }, bh[7] = { 16, 8, 16, 8, 4, 8, 4 };
    ctx->launches++;
    const int s = fenc->g.stride;
    for (int n = 0; n < n_jobs; n++) {
        const x264_cuda_me_job_t *j = &jobs[n];
        if (!(j->flags & X264_CUDA_ME_SEEDED)) return -1;
        if (!tab[j->qp]) { tab[j->qp] = malloc((4 * 4 * 2048 + 1) * sizeof(int16_t)); xo_cost_mv_table(j->qp, tab[j->qp]); }
        const int16_t *cx = tab[j->qp] + 2 * 4 * 2048 - j->mvp[0], *cy = tab[j->qp] + 2 * 4 * 2048 - j->mvp[1];
        int bmx = j->seed_mv[0], bmy = j->seed_mv[1], bcost = j->seed_cost;
        const int min_x = bmx - me_range > j->mv_min_fpel[0] ? bmx - me_range : j->mv_min_fpel[0], min_y = bmy - me_range > j->mv_min_fpel[1] ? bmy - me_range : j->mv_min_fpel[1];
        const int max_x = bmx + me_range < j->mv_max_fpel[0] ? bmx + me_range : j->mv_max_fpel[0], max_y = bmy + me_range < j->mv_max_fpel[1] ? bmy + me_range : j->mv_max_fpel[1];
        const int width = (max_x - min_x + 3) & ~3;
        results[n].seed_mx = bmx; results[n].seed_my = bmy; results[n].seed_cost = bcost;
        for (int my = min_y; my <= max_y; my++)
            for (int mx = min_x; mx < min_x + width; mx++) {
                int sad = 0;
                for (int y = 0; y < bh[j->i_pixel]; y++)
                    for (int x = 0; x < bw[j->i_pixel]; x++)
                        sad += abs(fenc->plane[0][(size_t)(j->by + y) * s + j->bx + x] - fref->plane[0][(size_t)(j->by + y + my) * s + j->bx + x + mx]);
                const int c = sad + cx[mx << 2] + cy[my << 2];
                if (c < bcost) { bcost = c; bmx = mx; bmy = my; }
            }
        results[n].bmx = bmx; results[n].bmy = bmy; results[n].bcost = bcost;
    }
    return 0;
}
int x264_cuda_host_register(void *p, size_t bytes) { (void)p; (void)bytes; return 0; }
int x264_cuda_sad_grid_quad_direct(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, const x264_cuda_frame_t *fref, int radius, const x264_cuda_grid_job_t *jobs,
                                   int n_jobs, uint16_t *grid) { return x264_cuda_sad_grid_quad(ctx, fenc, fref, radius, jobs, n_jobs, grid, 0); }
int x264_cuda_grid_ring_reserve(x264_cuda_t *ctx, size_t bytes) { (void)ctx; (void)bytes; return 0; }
