/* x264_cuda_host.c — host-side C companions of the CUDA back-end: things the reference computes on the host
 * and hands to the device as data.  Compiled with the reference's own flags (-O4 -ffast-math). */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include "../../include/x264_cuda.h"

/* lambda = 2^(qp/6-2), the reference's hand-rounded table (S/encoder/analyse.c:140-148) */
int x264_cuda_host_lambda(int qp)
{
    static const uint8_t lambda_of_qp[52] = {
        1,  1,  1,  1,  1,  1,  1,  1,  1,  1,  1,  1,  1,  1,  1,  1,  2,  2,  2,  2,  3,  3,  3,  4,  4,  4,
        5,  6,  6,  7,  8,  9,  10, 11, 13, 14, 16, 18, 20, 23, 25, 29, 32, 36, 40, 45, 51, 57, 64, 72, 81, 91 };
    return qp < 0 ? 1 : qp > 51 ? 91 : lambda_of_qp[qp];
}

/* lambda2 = lambda^2 * .9 * 256 with lambda = 2^(qp/6-2) unrounded: x264_lambda2_tab (S/encoder/analyse.c:150-160) is
 * floor(0.9 * 256 * 2^((qp-12)/3)) for every qp, which is how it is produced here */
int x264_cuda_host_lambda2(int qp)
{
    static const double cube_root_of_2_pow[3] = { 1.0, 1.2599210498948732, 1.5874010519681994 };
    qp = qp < 0 ? 0 : qp > 51 ? 51 : qp;
    return (int)ldexp(0.9 * 256.0 * cube_root_of_2_pow[qp % 3], qp / 3 - 4);
}

/* p_cost_mv[qp] exactly as x264_mb_analyse_load_costs builds it (S/encoder/analyse.c:40,192-203): the
 * reference's log2f is a macro around double log(), narrowed to float before the division by log(2).
 * In a drop-in integration the reference's own table is uploaded instead (x264_cuda_set_cost_mv). */
void x264_cuda_host_cost_mv(int qp, int16_t *table)
{
    const int lambda = x264_cuda_host_lambda(qp);
    int16_t *centre = table + 2 * 4 * 2048;
    for (int i = 0; i <= 2 * 4 * 2048; i++) {
        float l2 = (float)log((double)(i + 1));
        centre[i] = centre[-i] = lambda * (l2 / (log((double)2)) * 2 + 0.718f + !!i) + .5f;
    }
}

/* ------------------------------------------------------------------------------------------------------------
 * Quantiser tables as x264_cqm_init builds them (S/common/set.c:28-66 constants, :68-174 derivation), for the
 * flat and JVT presets with the default deadzones (inter 21, intra 11: S/common/common.c:129-130).  In a drop-in
 * integration the tables of the live x264_t are uploaded instead (x264_cuda_set_quant_tables). */
static const uint8_t k_dequant4[6][3] = { { 10, 13, 16 }, { 11, 14, 18 }, { 13, 16, 20 }, { 14, 18, 23 }, { 16, 20, 25 }, { 18, 23, 29 } };
static const uint16_t k_quant4[6][3] = { { 13107, 8066, 5243 }, { 11916, 7490, 4660 }, { 10082, 6554, 4194 },
                                         { 9362, 5825, 3647 },  { 8192, 5243, 3355 },  { 7282, 4559, 2893 } };
static const uint8_t k_scan8[16] = { 0, 3, 4, 3, 3, 1, 5, 1, 4, 5, 2, 5, 3, 1, 5, 1 };
static const uint8_t k_dequant8[6][6] = { { 20, 18, 32, 19, 25, 24 }, { 22, 19, 35, 21, 28, 26 }, { 26, 23, 42, 24, 33, 31 },
                                          { 28, 25, 45, 26, 35, 33 }, { 32, 28, 51, 30, 40, 38 }, { 36, 32, 58, 34, 46, 43 } };
static const uint16_t k_quant8[6][6] = { { 13107, 11428, 20972, 12222, 16777, 15481 }, { 11916, 10826, 19174, 11058, 14980, 14290 },
                                         { 10082, 8943, 15978, 9675, 12710, 11985 },   { 9362, 8228, 14913, 8931, 11984, 11259 },
                                         { 8192, 7346, 13159, 7740, 10486, 9777 },     { 7282, 6428, 11570, 6830, 9118, 8640 } };
/* JVT default matrices (S/common/set.h:168-203) */
static const uint8_t k_jvt4i[16] = { 6, 13, 20, 28, 13, 20, 28, 32, 20, 28, 32, 37, 28, 32, 37, 42 };
static const uint8_t k_jvt4p[16] = { 10, 14, 20, 24, 14, 20, 24, 27, 20, 24, 27, 30, 24, 27, 30, 34 };
static const uint8_t k_jvt8i[64] = { 6,  10, 13, 16, 18, 23, 25, 27, 10, 11, 16, 18, 23, 25, 27, 29, 13, 16, 18, 23, 25, 27,
                                     29, 31, 16, 18, 23, 25, 27, 29, 31, 33, 18, 23, 25, 27, 29, 31, 33, 36, 23, 25, 27, 29,
                                     31, 33, 36, 38, 25, 27, 29, 31, 33, 36, 38, 40, 27, 29, 31, 33, 36, 38, 40, 42 };
static const uint8_t k_jvt8p[64] = { 9,  13, 15, 17, 19, 21, 22, 24, 13, 13, 17, 19, 21, 22, 24, 25, 15, 17, 19, 21, 22, 24,
                                     25, 27, 17, 19, 21, 22, 24, 25, 27, 28, 19, 21, 22, 24, 25, 27, 28, 30, 21, 22, 24, 25,
                                     27, 28, 30, 32, 22, 24, 25, 27, 28, 30, 32, 33, 24, 25, 27, 28, 30, 32, 33, 35 };

static int round_div(int n, int d) { return (n + (d >> 1)) / d; }
static int round_shift(int x, int s) { return s < 0 ? x << -s : s == 0 ? x : (x + (1 << (s - 1))) >> s; }

void x264_cuda_host_cqm_tables(int cqm_preset, uint16_t q4mf[4][52][16], uint16_t q4bias[4][52][16], int dq4[4][6][16],
                               uint16_t q8mf[2][52][64], uint16_t q8bias[2][52][64], int dq8[2][6][64])
{
    static const int deadzone[4] = { 32 - 11, 32 - 21, 32 - 11, 32 - 21 }; /* 4IY, 4PY, 4IC, 4PC (set.c:77-79) */
    for (int list = 0; list < 4; list++)
        for (int i = 0; i < 16; i++) {
            const int sl = cqm_preset ? ((list & 1) ? k_jvt4p[i] : k_jvt4i[i]) : 16;
            const int k = (i & 1) + ((i >> 2) & 1);
            for (int q = 0; q < 6; q++) dq4[list][q][i] = k_dequant4[q][k] * sl;
            for (int qp = 0; qp < 52; qp++) {
                const int j = round_shift(round_div(k_quant4[qp % 6][k] * 16, sl), qp / 6 - 1);
                const int b = round_div(deadzone[list] << 10, j), cap = (1 << 15) / j;
                q4mf[list][qp][i] = (uint16_t)j;
                q4bias[list][qp][i] = (uint16_t)(b < cap ? b : cap);
            }
        }
    for (int list = 0; list < 2; list++)
        for (int i = 0; i < 64; i++) {
            const int sl = cqm_preset ? (list ? k_jvt8p[i] : k_jvt8i[i]) : 16;
            const int k = k_scan8[((i >> 1) & 12) | (i & 3)];
            for (int q = 0; q < 6; q++) dq8[list][q][i] = k_dequant8[q][k] * sl;
            for (int qp = 0; qp < 52; qp++) {
                const int j = round_shift(round_div(k_quant8[qp % 6][k] * 16, sl), qp / 6);
                const int b = round_div(deadzone[list] << 10, j), cap = (1 << 15) / j;
                q8mf[list][qp][i] = (uint16_t)j;
                q8bias[list][qp][i] = (uint16_t)(b < cap ? b : cap);
            }
        }
}

/* ---------------------------------------------------------------------------------------------------------------------
 * Float tails of the frame metrics.  They run on the host with the reference's compiler flags because their results depend
 * on float evaluation order; the device supplies the integer inputs (csrc/metrics.cu). */

/* x264_pixel_ssim_wxh's accumulation (S/common/pixel.c:462-509) over the per-4x4 sums of x264_cuda_frame_ssim_sums: windows of
 * 2x2 blocks, evaluated four at a time per row pair, each group summed on its own before it joins the total. */
static float ssim_window(int s1, int s2, int ss, int s12)
{
    static const int c1 = (int)(.01 * .01 * 255 * 255 * 64 + .5);
    static const int c2 = (int)(.03 * .03 * 255 * 255 * 64 * 63 + .5);
    const int vars = ss * 64 - s1 * s1 - s2 * s2;
    const int covar = s12 * 64 - s1 * s2;
    return (float)(2 * s1 * s2 + c1) * (float)(2 * covar + c2) / ((float)(s1 * s1 + s2 * s2 + c1) * (float)(vars + c2));
}
float x264_cuda_host_ssim_end(const int (*sums)[4], int w4, int h4)
{
    float total = 0.0;
    for (int y = 1; y < h4; y++) {
        const int (*cur)[4] = sums + (size_t)y * w4, (*up)[4] = sums + (size_t)(y - 1) * w4;
        for (int x = 0; x < w4 - 1; x += 4) {
            const int n = w4 - x - 1 < 4 ? w4 - x - 1 : 4;
            float group = 0.0;
            for (int i = x; i < x + n; i++)
                group += ssim_window(cur[i][0] + cur[i + 1][0] + up[i][0] + up[i + 1][0], cur[i][1] + cur[i + 1][1] + up[i][1] + up[i + 1][1],
                                     cur[i][2] + cur[i + 1][2] + up[i][2] + up[i + 1][2], cur[i][3] + cur[i + 1][3] + up[i][3] + up[i + 1][3]);
            total += group;
        }
    }
    return total;
}

/* x264_adaptive_quant_frame (S/encoder/ratecontrol.c:233-249) from the macroblock energies of x264_cuda_frame_mb_energy:
 * qp_offset[mb] = strength*1.0397*(log2(energy) - 14.427...) via the reference's 7-bit log2 table, inv_qscale[mb] = 2^(-qp/6)
 * in .8 fixed point via its 6-bit exp2 table (x264_exp2fix8).  Both tables are five-decimal / integer roundings of the
 * functions and are rebuilt here rather than transcribed. */
static float aq_log2[128];
static uint8_t aq_exp2[64];
static pthread_once_t aq_once = PTHREAD_ONCE_INIT; /* several frame threads may arrive together */
static void aq_build(void)
{
    for (int i = 0; i < 128; i++) aq_log2[i] = (float)(floor(log2(1.0 + i / 128.0) * 1e5 + 0.5) / 1e5);
    for (int i = 0; i < 64; i++) aq_exp2[i] = (uint8_t)floor((pow(2.0, (i + 0.5) / 64.0) - 1.0) * 256.0 + 0.5);
}
static void aq_init(void) { pthread_once(&aq_once, aq_build); }
void x264_cuda_host_aq(const uint32_t *energy, int n_mb, float aq_strength, float *qp_offset, uint16_t *inv_qscale)
{
    aq_init();
    const float strength = aq_strength * 1.0397;
    for (int k = 0; k < n_mb; k++) {
        const uint32_t e = energy[k] ? energy[k] : 1; /* ac_energy_mb never returns 0 (ratecontrol.c:188: var clamped to >= 1); clz(0) is undefined */
        const int lz = __builtin_clz(e);
        const float adj = strength * (aq_log2[(e << lz >> 24) & 0x7f] - lz + 16.573f);
        qp_offset[k] = adj;
        if (inv_qscale) {
            float x = adj * (-1.f / 6.f) + 8;
            int v;
            if (x <= 0) v = 0;
            else if (x >= 16) v = 0xffff;
            else { const int i = x; const int f = (x - i) * 64; v = (aq_exp2[f] + 256) << i >> 8; }
            inv_qscale[k] = (uint16_t)v;
        }
    }
}

/* ---------------------------------------------------------------------------------------------------------------------
 * x264_me_search_ref, predictor stage + ESA loop (S/encoder/me.c:182-229, :449-492), replayed on a SAD grid produced by
 * x264_cuda_sad_grid.  Pure host arithmetic in the reference's order: strict '<', candidates in mvc order, raster scan. */
static inline int clip3_i(int v, int lo, int hi) { return v < lo ? lo : v > hi ? hi : v; }
int x264_cuda_host_esa_replay(const uint16_t *gp, int radius, int cx, int cy, const x264_cuda_me_job_t *job, int me_range, const int16_t *cost_table,
                              x264_cuda_me_result_t *res)
{
    const int GW = X264_CUDA_GRID_W(radius), GH = X264_CUDA_GRID_H(radius), gx0 = cx - radius, gy0 = cy - radius;
    const int16_t *cmx = cost_table + 2 * 4 * 2048 - job->mvp[0], *cmy = cost_table + 2 * 4 * 2048 - job->mvp[1];
    const int x_min = job->mv_min_fpel[0], y_min = job->mv_min_fpel[1], x_max = job->mv_max_fpel[0], y_max = job->mv_max_fpel[1];
#define GRID_SAD(mx, my, v) do { const int i_ = (mx) - gx0, j_ = (my) - gy0; if (i_ < 0 || i_ >= GW || j_ < 0 || j_ >= GH) return 1; \
                                 (v) = gp[j_ * GW + i_]; if ((v) == 0xffff) return 1; } while (0)
    int bmx = clip3_i(job->mvp[0], x_min * 4, x_max * 4), bmy = clip3_i(job->mvp[1], y_min * 4, y_max * 4);
    const int pmx = (bmx + 2) >> 2, pmy = (bmy + 2) >> 2;
    int bcost, v;
    bmx = pmx; bmy = pmy;
    GRID_SAD(pmx, pmy, v);
    bcost = v;                                   /* the prediction itself carries no mv cost (me.c:207) */
    const int n_mvc = job->i_mvc < 12 ? job->i_mvc : 12;
    for (int i = 0; i < n_mvc; i++) {            /* me.c:209-222 */
        const int mx0 = (job->mvc[i][0] + 2) >> 2, my0 = (job->mvc[i][1] + 2) >> 2;
        if ((mx0 | my0) && (mx0 != pmx || my0 != pmy)) {
            const int mx = clip3_i(mx0, x_min, x_max), my = clip3_i(my0, y_min, y_max);
            GRID_SAD(mx, my, v);
            const int c = v + cmx[mx << 2] + cmy[my << 2];
            if (c < bcost) { bcost = c; bmx = mx; bmy = my; }
        }
    }
    GRID_SAD(0, 0, v);                           /* me.c:224-225: COST_MV( 0, 0 ) */
    { const int c = v + cmx[0] + cmy[0]; if (c < bcost) { bcost = c; bmx = 0; bmy = 0; } }
    res->seed_mx = (int16_t)bmx; res->seed_my = (int16_t)bmy; res->seed_cost = bcost;
    /* me.c:451-457 and the plain-ESA loop :480-489 */
    const int min_x = bmx - me_range > x_min ? bmx - me_range : x_min, min_y = bmy - me_range > y_min ? bmy - me_range : y_min;
    const int max_x = bmx + me_range < x_max ? bmx + me_range : x_max, max_y = bmy + me_range < y_max ? bmy + me_range : y_max;
    const int width = (max_x - min_x + 3) & ~3;
    if (min_x < gx0 || min_x + width > gx0 + GW || min_y < gy0 || max_y >= gy0 + GH) return 1;
    for (int my = min_y; my <= max_y; my++) {
        const uint16_t *row = gp + (my - gy0) * GW - gx0;
        const int cy_ = cmy[my << 2];
        for (int mx = min_x; mx < min_x + width; mx++) {
            /* the reference also evaluates the up-to-3 columns beyond max_x that the width rounding adds (they are inside the
             * padded plane); the grid holds 0xffff for vectors beyond the MV limits: report "outside" rather than guess */
            if (row[mx] == 0xffff) return 1;
            const int c = row[mx] + cmx[mx << 2] + cy_;
            if (c < bcost) { bcost = c; bmx = mx; bmy = my; }
        }
    }
#undef GRID_SAD
    res->bmx = (int16_t)bmx; res->bmy = (int16_t)bmy; res->bcost = bcost;
    return 0;
}
