/* xo_intra.c — ORACLE (test infrastructure only): the parts of intra analysis that depend only on the source macroblock and the
 * reconstructed pixels of NEIGHBOURING macroblocks: the Intra16x16 mode costs of x264_mb_analyse_intra (S/encoder/analyse.c:612-664)
 * and the chroma mode costs of x264_mb_analyse_intra_chroma (:541-609), with the predictors of S/common/predict.c:40-170 (16x16) and
 * :172-336 (8x8 chroma).  The I4x4 / I8x8 stages need the reconstruction of earlier blocks of the SAME macroblock and are not here. */
#include <string.h>
#include "xo.h"

static inline uint8_t clip_u8(int x) { return x < 0 ? 0 : x > 255 ? 255 : x; }

/* neighbour vector layout: nb[0] = top-left, nb[1..n] = row above, nb[n+1..2n] = column to the left */
#define TL(nb) ((nb)[0])
#define TOP(nb, x) ((nb)[1 + (x)])
#define LEFT(nb, n, y) ((nb)[1 + (n) + (y)])
static int edge(const uint8_t *nb, int n, int is_left, int k) /* k = -1 is the corner for both edges */
{
    return k < 0 ? TL(nb) : is_left ? LEFT(nb, n, k) : TOP(nb, k);
}

/* mode numbering: enum intra16x16_pred_e (S/common/predict.h:48-58): V H DC P DC_LEFT DC_TOP DC_128 */
void xo_predict_16x16(int mode, const uint8_t nb[33], uint8_t pred[256])
{
    int s = 0;
    switch (mode) {
    case 0: for (int y = 0; y < 16; y++) for (int x = 0; x < 16; x++) pred[16 * y + x] = TOP(nb, x); return;
    case 1: for (int y = 0; y < 16; y++) memset(pred + 16 * y, LEFT(nb, 16, y), 16); return;
    case 2: for (int i = 0; i < 16; i++) s += TOP(nb, i) + LEFT(nb, 16, i); memset(pred, (s + 16) >> 5, 256); return;
    case 4: for (int i = 0; i < 16; i++) s += LEFT(nb, 16, i); memset(pred, (s + 8) >> 4, 256); return;
    case 5: for (int i = 0; i < 16; i++) s += TOP(nb, i); memset(pred, (s + 8) >> 4, 256); return;
    case 6: memset(pred, 128, 256); return;
    default: { /* plane, predict.c:134-167 */
        int H = 0, V = 0;
        for (int i = 0; i < 8; i++) {
            H += (i + 1) * (edge(nb, 16, 0, 8 + i) - edge(nb, 16, 0, 6 - i));
            V += (i + 1) * (edge(nb, 16, 1, 8 + i) - edge(nb, 16, 1, 6 - i));
        }
        const int a = 16 * (LEFT(nb, 16, 15) + TOP(nb, 15)), b = (5 * H + 32) >> 6, c = (5 * V + 32) >> 6;
        for (int y = 0; y < 16; y++)
            for (int x = 0; x < 16; x++) pred[16 * y + x] = clip_u8((a + b * (x - 7) + c * (y - 7) + 16) >> 5);
        return; }
    }
}

/* mode numbering: enum intra_chroma_pred_e (predict.h:31-41): DC H V P DC_LEFT DC_TOP DC_128 */
void xo_predict_8x8c(int mode, const uint8_t nb[17], uint8_t pred[64])
{
    int s0 = 0, s1 = 0, s2 = 0, s3 = 0; /* top-left half, top-right half, left-upper half, left-lower half */
    for (int i = 0; i < 4; i++) { s0 += TOP(nb, i); s1 += TOP(nb, i + 4); s2 += LEFT(nb, 8, i); s3 += LEFT(nb, 8, i + 4); }
    int dc[4]; /* quadrants: 0 1 / 2 3 */
    switch (mode) {
    case 0: dc[0] = (s0 + s2 + 4) >> 3; dc[1] = (s1 + 2) >> 2; dc[2] = (s3 + 2) >> 2; dc[3] = (s1 + s3 + 4) >> 3; break; /* predict.c:234-277 */
    case 4: dc[0] = dc[1] = (s2 + 2) >> 2; dc[2] = dc[3] = (s3 + 2) >> 2; break;                                          /* :184-212 */
    case 5: dc[0] = dc[2] = (s0 + 2) >> 2; dc[1] = dc[3] = (s1 + 2) >> 2; break;                                          /* :213-233 */
    case 6: dc[0] = dc[1] = dc[2] = dc[3] = 128; break;
    case 1: for (int y = 0; y < 8; y++) memset(pred + 8 * y, LEFT(nb, 8, y), 8); return;
    case 2: for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) pred[8 * y + x] = TOP(nb, x); return;
    default: { /* plane, :305-336 */
        int H = 0, V = 0;
        for (int i = 0; i < 4; i++) {
            H += (i + 1) * (edge(nb, 8, 0, 4 + i) - edge(nb, 8, 0, 2 - i));
            V += (i + 1) * (edge(nb, 8, 1, 4 + i) - edge(nb, 8, 1, 2 - i));
        }
        const int a = 16 * (LEFT(nb, 8, 7) + TOP(nb, 7)), b = (17 * H + 16) >> 5, c = (17 * V + 16) >> 5;
        for (int y = 0; y < 8; y++)
            for (int x = 0; x < 8; x++) pred[8 * y + x] = clip_u8((a + b * (x - 3) + c * (y - 3) + 16) >> 5);
        return; }
    }
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++) pred[8 * y + x] = (uint8_t)dc[(y >> 2) * 2 + (x >> 2)];
}

/* candidate list for a neighbour mask: predict_16x16_mode_available / predict_8x8chroma_mode_available (analyse.c:372-440).  Both lists
 * have the same shape once written in terms of (V, H, DC, P, DC_LEFT, DC_TOP, DC_128) of the respective enum. */
static int mode_list(int neighbour, int chroma, int modes[4])
{
    const int V = chroma ? 2 : 0, H = 1, DC = chroma ? 0 : 2, P = 3;
    if (neighbour & 8) { modes[0] = V; modes[1] = H; modes[2] = DC; modes[3] = P; return 4; } /* MB_TOPLEFT */
    if (neighbour & 1) { modes[0] = 4; modes[1] = H; return 2; }                               /* MB_LEFT */
    if (neighbour & 2) { modes[0] = 5; modes[1] = V; return 2; }                               /* MB_TOP */
    modes[0] = 6;
    return 1;
}

/* bs_size_ue of the mode number that gets written to the stream (x264_mb_pred_mode16x16_fix / 8x8c_fix: the DC variants code as DC) */
static int mode_bits(int mode, int chroma)
{
    static const int ue_bits[4] = { 1, 3, 3, 5 };
    if (mode > 3) mode = chroma ? 0 : 2;
    return ue_bits[mode];
}

void xo_intra_mb_costs(const xo_intra_in *in, const uint8_t fenc_y[256], const uint8_t fenc_u[64], const uint8_t fenc_v[64],
                       const uint8_t nb_y[33], const uint8_t nb_u[17], const uint8_t nb_v[17], xo_intra_out *out)
{
    const int metric = in->mbcmp_satd ? XO_SATD : XO_SAD;
    int modes[4], n;
    uint8_t pred[256], pu[64], pv[64];
    for (int i = 0; i < 7; i++) out->cost16[i] = out->cost_chroma[i] = -1;
    out->best16 = out->best_chroma = 1 << 28; /* COST_MAX */
    out->mode16 = out->mode_chroma = 0;

    n = mode_list(in->neighbour, 0, modes);
    for (int i = 0; i < n; i++) { /* analyse.c:641-657 (the merged x3 path :627-639 yields the same numbers) */
        xo_predict_16x16(modes[i], nb_y, pred);
        int cost = xo_pixel_cmp(metric, XO_16x16, pred, 16, fenc_y, 16) + in->lambda * mode_bits(modes[i], 0);
        out->cost16[modes[i]] = cost;
        if (cost < out->best16) { out->best16 = cost; out->mode16 = modes[i]; }
    }
    if (in->b_slice_b) out->best16 += in->lambda * 9; /* i_mb_b_cost_table[I_16x16], analyse.c:659-661 */

    n = mode_list(in->neighbour, 1, modes);
    for (int i = 0; i < n; i++) { /* analyse.c:583-606 */
        xo_predict_8x8c(modes[i], nb_u, pu);
        xo_predict_8x8c(modes[i], nb_v, pv);
        int cost = xo_pixel_cmp(metric, XO_8x8, pu, 8, fenc_u, 8) + xo_pixel_cmp(metric, XO_8x8, pv, 8, fenc_v, 8) +
                   in->lambda * mode_bits(modes[i], 1);
        out->cost_chroma[modes[i]] = cost;
        if (cost < out->best_chroma) { out->best_chroma = cost; out->mode_chroma = modes[i]; }
    }
}
