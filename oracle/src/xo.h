/* xo.h — ORACLE API.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, single-threaded restatement of the data-parallel core of x264-snapshot-20090216-2245
 * (S/ = /root/reference/x264-snapshot-20090216-2245/).  Every function cites the S/file:line it follows.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library; the product (libx264_cuda.so) never links or calls it.
 *
 * The SAME symbol set is exported by oracle/_ref/libref_harness.so, where each xo_* call is served by
 * the UNMODIFIED reference code compiled from S/ (see oracle/ref_harness.c).  Tests run both and
 * require identical bytes, which is how this restatement is pinned (the reference ships no golden
 * vectors: SURVEY.md §4 / §8c).
 */
#ifndef XO_H
#define XO_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* partition ids, S/common/pixel.h:30-42 */
enum { XO_16x16 = 0, XO_16x8, XO_8x16, XO_8x8, XO_8x4, XO_4x8, XO_4x4 };
/* metric ids for xo_pixel_cmp */
enum { XO_SAD = 0, XO_SSD = 1, XO_SATD = 2, XO_SA8D = 3 };
/* motion-search methods, S/x264.h:100-104 */
enum { XO_ME_DIA = 0, XO_ME_HEX = 1, XO_ME_UMH = 2, XO_ME_ESA = 3, XO_ME_TESA = 4 };

#define XO_PADH 32 /* S/common/frame.h:28-29 */
#define XO_PADV 32
#define XO_FENC_STRIDE 16 /* S/common/common.h:464-465 */
#define XO_FDEC_STRIDE 32
#define XO_COST_MAX (1 << 28) /* S/encoder/me.h:27 */

const char *xo_backend(void); /* "port" (restatement) or "reference" (harness over S/) */

/* ---------------- pixel metrics: S/common/pixel.c ---------------- */
int xo_pixel_cmp(int metric, int i_pixel, const uint8_t *pix1, int stride1, const uint8_t *pix2, int stride2);
int xo_pixel_var(int i_pixel /* XO_16x16 | XO_8x8 */, const uint8_t *pix, int stride);
uint64_t xo_pixel_hadamard_ac(int i_pixel /* 16x16..8x8 */, const uint8_t *pix, int stride);
/* S/common/pixel.c:515-559; i_pixel selects ads4/ads2/ads1 as x264_pixel_init does (:591-594, :793-796) */
int xo_pixel_ads(int i_pixel, const int enc_dc[4], const uint16_t *sums, int delta, const uint16_t *cost_mvx,
                 int16_t *mvs, int width, int thresh);

/* ---------------- cost tables: S/encoder/analyse.c:140-218 ---------------- */
int xo_lambda(int qp);
/* fills out[0 .. 4*4*2048] ; the centre (mvd 0) is out[2*4*2048] exactly like p_cost_mv */
void xo_cost_mv_table(int qp, int16_t *out);

/* ---------------- frame geometry: S/common/frame.c:29-152 (cpu=0 -> align 16) ---------------- */
typedef struct {
    int width, height;       /* picture size as given by the user */
    int mb_width, mb_height; /* in macroblocks */
    int stride, lines;       /* luma stride / mod16 lines */
    int plane_size;          /* stride*(lines+2*PADV) */
    int origin;              /* offset of pixel (0,0) inside a plane buffer: stride*PADV+PADH */
    int stride_lowres, width_lowres, lines_lowres, plane_size_lowres, origin_lowres;
} xo_geom;
void xo_geometry(int width, int height, xo_geom *g);

/* ---------------- whole-frame filters ---------------- */
/* S/common/frame.c:304-331 then :240-267 (luma only): pad the picture to mod16, replicate 32-px borders.
 * plane points at pixel (0,0). */
void xo_frame_expand_border(const xo_geom *g, uint8_t *plane);
/* S/common/mc.c:404-463 with (mb_y=0,b_end=1) + S/common/frame.c:269-295: hpel planes h,v,c and the
 * integral image(s).  Pointers address pixel (0,0) of their plane buffers; integral may be NULL.
 * integral has room for stride*(lines+2*PADV) uint16 (twice that when sub8x8). */
void xo_frame_filter(const xo_geom *g, const uint8_t *plane, uint8_t *dsth, uint8_t *dstv, uint8_t *dstc,
                     uint16_t *integral, int b_sub8x8);
/* S/common/mc.c:306-357 + S/common/frame.c:297-302; note: writes plane[width] column / lines row first,
 * exactly like the reference (:315-317), hence plane is not const. */
void xo_frame_init_lowres(const xo_geom *g, uint8_t *plane, uint8_t *l0, uint8_t *lh, uint8_t *lv, uint8_t *lc);

/* S/common/mc.c:157-202: qpel sample fetch from the 4 hpel planes into dst (stride dst_stride) */
void xo_mc_luma(uint8_t *dst, int dst_stride, const uint8_t *const src[4], int src_stride, int mvx, int mvy, int w, int h);
/* S/common/mc.c:205-236 */
void xo_mc_chroma(uint8_t *dst, int dst_stride, const uint8_t *src, int src_stride, int mvx, int mvy, int w, int h);
/* h->mc.avg[i_pixel] (S/common/mc.c:52-125): dst = (a + b + 1) >> 1 when weight == 32, else the implicit-weighted-bipred blend
 * clip((a*weight + b*(64-weight) + 32) >> 6); i_pixel: XO_16x16 .. XO_4x4 and 7..9 for 4x2, 2x4, 2x2 (mc.avg has ten entries) */
void xo_pixel_avg(int i_pixel, uint8_t *dst, int dst_stride, const uint8_t *a, int a_stride, const uint8_t *b, int b_stride, int weight);

/* ---------------- full-pel motion search: S/encoder/me.c:156-631 with i_subpel_refine = 1 ----------------
 * One call == one x264_me_search_ref() on a block, stopping before sub-pel refinement.
 * Planes are whole padded frames; (bx,by) is the block's pixel position. */
typedef struct {
    int me_method;         /* XO_ME_* (DIA, HEX, ESA, TESA supported) */
    int me_range;          /* h->param.analyse.i_me_range */
    int qp;                /* selects lambda / p_cost_mv */
    int fpel_satd;         /* 1: fpelcmp = SATD (subme>1 && TESA, S/encoder/encoder.c:608-618), 0: SAD */
    int i_pixel;           /* XO_16x16 .. XO_4x4 */
    int bx, by;            /* block position in pixels */
    int mv_min_fpel[2], mv_max_fpel[2]; /* h->mb.mv_{min,max}_fpel */
    int mv_min_spel[2], mv_max_spel[2]; /* h->mb.mv_{min,max}_spel (me.c:629, :699-700) */
    int16_t mvp[2];        /* qpel predictor */
    int i_mvc;             /* number of extra predictors */
    int16_t mvc[16][2];
    int b_sub8x8;          /* integral has the 4x4 plane */
} xo_me_in;
typedef struct {
    int16_t mv[2];         /* qpel units */
    int cost, cost_mv;
    /* full-pel state right after the search loop (before "-> qpel mv"): what a device kernel must reproduce */
    int bmx, bmy, bcost;
    /* the same, right BEFORE the ESA/TESA/DIA/HEX loop (after predictors and (0,0)): the seed */
    int seed_mx, seed_my, seed_cost;
} xo_me_out;
void xo_me_search_fpel(const xo_geom *g, const uint8_t *fenc_plane, const uint8_t *fref_plane,
                       const uint16_t *integral, const xo_me_in *in, xo_me_out *out);
/* n independent searches in one call (C loop; used to time the CPU path without per-call binding overhead) */
void xo_me_search_fpel_batch(const xo_geom *g, const uint8_t *fenc_plane, const uint8_t *fref_plane,
                             const uint16_t *integral, const xo_me_in *in, int n, xo_me_out *out);
/* the same search followed by refine_subpel (me.c:622-628): subme = h->mb.i_subpel_refine used by the
 * search, mbcmp_satd = whether mbcmp is SATD (user subme>1).  fref_planes = {full, h, v, c}. */
void xo_me_search_subpel(const xo_geom *g, const uint8_t *fenc_plane, const uint8_t *const fref_planes[4],
                         const uint16_t *integral, const xo_me_in *in, int subme, int mbcmp_satd, xo_me_out *out);

/* ---------------- transforms: S/common/dct.c, quant: S/common/quant.c, tables: S/common/set.c ---------------- */
void xo_sub4x4_dct(int16_t dct[16], const uint8_t *pix1 /*stride 16*/, const uint8_t *pix2 /*stride 32*/);
void xo_add4x4_idct(uint8_t *dst /*stride 32*/, int16_t dct[16]);
void xo_sub8x8_dct8(int16_t dct[64], const uint8_t *pix1, const uint8_t *pix2);
void xo_add8x8_idct8(uint8_t *dst, int16_t dct[64]);
void xo_dct4x4dc(int16_t d[16]);
void xo_idct4x4dc(int16_t d[16]);
void xo_add_idct_dc(uint8_t *dst, const int16_t *dc, int n /* 4: add8x8_idct_dc, 16: add16x16_idct_dc */);

/* cqm: 0 = flat16, 1 = JVT.  Tables as built by x264_cqm_init (S/common/set.c:68-174) with default deadzones 21/11.
 * list ids: 0 CQM_4IY 1 CQM_4PY 2 CQM_4IC 3 CQM_4PC (4x4), 0 CQM_8IY 1 CQM_8PY (8x8). */
void xo_quant4_tables(int cqm, int list, int qp, uint16_t mf[16], uint16_t bias[16]);
void xo_quant8_tables(int cqm, int list, int qp, uint16_t mf[64], uint16_t bias[64]);
void xo_dequant4_table(int cqm, int list, int dequant_mf[6][16]);
void xo_dequant8_table(int cqm, int list, int dequant_mf[6][64]);
int xo_quant_4x4(int16_t dct[16], const uint16_t mf[16], const uint16_t bias[16]);
int xo_quant_8x8(int16_t dct[64], const uint16_t mf[64], const uint16_t bias[64]);
int xo_quant_4x4_dc(int16_t dct[16], int mf, int bias);
int xo_quant_2x2_dc(int16_t dct[4], int mf, int bias);
void xo_dequant_4x4(int16_t dct[16], const int dequant_mf[6][16], int qp);
void xo_dequant_8x8(int16_t dct[64], const int dequant_mf[6][64], int qp);
void xo_dequant_4x4_dc(int16_t dct[16], const int dequant_mf[6][16], int qp);

/* zigzag (frame) scans, S/common/dct.c:488-560; decimation scores, S/common/quant.c:203-252 (i_max 15|16|64) */
void xo_zigzag_scan_4x4(int16_t level[16], const int16_t dct[16]);
void xo_zigzag_scan_8x8(int16_t level[64], const int16_t dct[64]);
int xo_decimate_score(const int16_t *dct, int i_max);

/* ---------------- residual path of one inter macroblock: S/encoder/macroblock.c:596-742 + :272-363 ----------------
 * fenc_* packed (16x16, 8x8, 8x8); rec_* hold the prediction on entry and the reconstruction on return. */
typedef struct {
    int qp, chroma_qp;
    int b_transform_8x8;
    int b_decimate; /* h->sh.i_type == SLICE_TYPE_B || h->param.analyse.b_dct_decimate */
    int cqm;        /* 0 flat, 1 JVT */
} xo_resid_in;
typedef struct {
    int16_t luma4x4[24][16]; /* h->dct.luma4x4: 16 luma + 8 chroma AC blocks, zigzag order; zero where nothing was coded */
    int16_t luma8x8[4][64];
    int16_t chroma_dc[2][4];
    uint8_t nnz[27];         /* non_zero_count[x264_scan8[i]]: 0..15 luma, 16..23 chroma AC, 24 luma DC, 25/26 chroma DC */
    uint8_t pad;
    int cbp_luma, cbp_chroma;
} xo_resid_out;
void xo_residual_inter_mb(const xo_resid_in *in, const uint8_t fenc_y[256], const uint8_t fenc_u[64], const uint8_t fenc_v[64],
                          uint8_t rec_y[256], uint8_t rec_u[64], uint8_t rec_v[64], xo_resid_out *out);

/* One I_16x16 macroblock through x264_macroblock_encode: predict_16x16[mode16] + x264_mb_encode_i16x16 (S/encoder/macroblock.c:512-530,
 * :184-270), predict_8x8c[mode_chroma] + x264_mb_encode_8x8_chroma(b_inter = 0) (:744-760, :272-363).  nb_y[33] / nb_u,nb_v[17] = corner,
 * row above, left column of the neighbouring reconstruction; in->b_decimate = slice B || (b_dct_decimate && slice P) (:193);
 * out->luma4x4[0..15] = AC levels (DC slot 0), luma_dc = h->dct.luma16x16_dc (zero unless out->nnz[24]); rec_* = reconstruction.
 * Levels of uncoded blocks are reported as zero. */
void xo_residual_intra16_mb(const xo_resid_in *in, int mode16, int mode_chroma, const uint8_t fenc_y[256], const uint8_t fenc_u[64],
                            const uint8_t fenc_v[64], const uint8_t nb_y[33], const uint8_t nb_u[17], const uint8_t nb_v[17],
                            uint8_t rec_y[256], uint8_t rec_u[64], uint8_t rec_v[64], xo_resid_out *out, int16_t luma_dc[16]);

/* x264_macroblock_probe_skip (S/encoder/macroblock.c:797-883) with the prediction supplied (b_bidir = 1); 1 = skippable.
 * in->b_transform_8x8 / b_decimate are ignored (the probe always uses the 4x4 transform and the decimation scores). */
int xo_probe_skip_mb(const xo_resid_in *in, const uint8_t fenc_y[256], const uint8_t fenc_u[64], const uint8_t fenc_v[64],
                     const uint8_t pred_y[256], const uint8_t pred_u[64], const uint8_t pred_v[64]);
int xo_lambda2(int qp); /* x264_lambda2_tab[qp], S/encoder/analyse.c:151-160 */

/* ---------------- intra analysis from neighbouring macroblocks: Intra16x16 and chroma 8x8 mode costs -----------------------------
 * x264_mb_analyse_intra's 16x16 stage (S/encoder/analyse.c:612-664) and x264_mb_analyse_intra_chroma (:541-609).
 * nb_*: [0] top-left, [1..n] row above, [n+1..2n] left column (n = 16 luma, 8 chroma) of the UNFILTERED reconstruction. */
typedef struct {
    int neighbour;  /* h->mb.i_neighbour: MB_LEFT 1, MB_TOP 2, MB_TOPRIGHT 4, MB_TOPLEFT 8 */
    int lambda;     /* a->i_lambda */
    int mbcmp_satd; /* h->pixf.mbcmp == satd (subme > 1) */
    int b_slice_b;
} xo_intra_in;
typedef struct {
    int cost16[7];      /* a->i_satd_i16x16_dir[mode] (mode = enum intra16x16_pred_e), -1 where the mode is not a candidate */
    int cost_chroma[7]; /* cost of chroma mode (enum intra_chroma_pred_e), -1 where not a candidate */
    int best16, best_chroma; /* a->i_satd_i16x16 (with the B-slice mb-type prefix), a->i_satd_i8x8chroma */
    int mode16, mode_chroma; /* a->i_predict16x16, a->i_predict8x8chroma */
} xo_intra_out;
void xo_predict_16x16(int mode, const uint8_t nb[33], uint8_t pred[256]);
void xo_predict_8x8c(int mode, const uint8_t nb[17], uint8_t pred[64]);
void xo_intra_mb_costs(const xo_intra_in *in, const uint8_t fenc_y[256], const uint8_t fenc_u[64], const uint8_t fenc_v[64],
                       const uint8_t nb_y[33], const uint8_t nb_u[17], const uint8_t nb_v[17], xo_intra_out *out);

/* chroma planes for b_chroma_me (pixel (0,0) pointers; borders expanded by 16 like x264_frame_expand_border does for planes 1,2) */
typedef struct { const uint8_t *fenc_u, *fenc_v, *fref_u, *fref_v; int stride_c; } xo_chroma;
void xo_me_search_subpel_chroma(const xo_geom *g, const uint8_t *fenc_plane, const uint8_t *const fref_planes[4], const uint16_t *integral,
                                const xo_chroma *ch, const xo_me_in *in, int subme, int mbcmp_satd, xo_me_out *out);

/* x264_me_refine_qpel (me.c:633-643) from (mv_in, cost_in); ch may be NULL (no chroma ME) */
void xo_me_refine_qpel(const xo_geom *g, const uint8_t *fenc_plane, const uint8_t *const fref_planes[4], const xo_chroma *ch, const xo_me_in *in,
                       int subme, int mbcmp_satd, const int16_t mv_in[2], int cost_in, xo_me_out *out);

/* x264_me_refine_bidir_satd (me.c:843-927): joint quarter-pel refinement of a list-0 / list-1 vector pair of one partition
 * (16x16, 16x8, 8x16 or 8x8 at (bx,by)) against the blended prediction.  mv0/mv1 are updated in place; returns the best cost seen
 * (the reference does not store it).  Uses in->i_pixel, bx, by, qp, mv_min_spel/mv_max_spel; mvp0/mvp1 are the two lists' predictors. */
int xo_me_refine_bidir_satd(const xo_geom *g, const uint8_t *fenc_plane, const uint8_t *const fref0[4], const uint8_t *const fref1[4],
                            const xo_me_in *in, const int16_t mvp0[2], const int16_t mvp1[2], int weight, int mbcmp_satd, int16_t mv0[2], int16_t mv1[2]);

/* ---------------- lowres lookahead: S/encoder/slicetype.c:43-355 ----------------
 * Planes are the four half-resolution planes (pixel 0,0 pointers, stride g->stride_lowres).  mvs/costs are the frame's
 * lowres_mvs[l][dist-1] / lowres_mv_costs[l][dist-1] arrays (mb_width*mb_height entries, updated in place when
 * do_search[l]); ref1_mvs = frames[p1]->lowres_mvs[0][p1-p0-1] (B evaluations only); intra_cost = i_intra_cost.
 * out->score is the plain sum before the B-frame scaling of slicetype.c:338-339. */
typedef struct {
    int p0, p1, b;
    int me_method;         /* user's --me; the lookahead uses min(HEX, me) (slicetype.c:38) */
    int me_range;          /* h->param.analyse.i_me_range */
    int mbcmp_satd;        /* user subme > 1 */
    int fpel_satd;         /* user subme > 1 && --me tesa */
    int b_weighted_bipred;
    int do_search[2];
    int b_intra_calculated;
} xo_lowres_in;
typedef struct { int score, score_aq, intra_mbs, intra_cost_sum; } xo_lowres_out;
void xo_lowres_frame_cost(const xo_geom *g, const xo_lowres_in *in, const uint8_t *const fenc[4], const uint8_t *const fref0[4],
                          const uint8_t *const fref1[4], int16_t (*mvs0)[2], int *costs0, int16_t (*mvs1)[2], int *costs1,
                          const int16_t (*ref1_mvs)[2], uint16_t *intra_cost, xo_lowres_out *out);
/* the VBV form (h->param.rc.i_vbv_buffer_size, slicetype.c:300-316): every block evaluated, row_satd[mb_height] = per-row sums of the
 * block costs weighted by inv_qscale (frames[b]->i_inv_qscale_factor; NULL = rc.i_aq_mode off), out->score_aq the weighted interior sum.
 * row_satd == NULL: the default (interior-only) form with the AQ weights (:318-330) */
void xo_lowres_frame_cost_vbv(const xo_geom *g, const xo_lowres_in *in, const uint8_t *const fenc[4], const uint8_t *const fref0[4],
                              const uint8_t *const fref1[4], int16_t (*mvs0)[2], int *costs0, int16_t (*mvs1)[2], int *costs1,
                              const int16_t (*ref1_mvs)[2], uint16_t *intra_cost, xo_lowres_out *out, const uint16_t *inv_qscale, int *row_satd);
/* one of the ten 8x8 predictions used by the lookahead (0..3: predict_8x8c DC,H,V,P; 4..9: predict_8x8 DDL,DDR,VR,HD,VL,HU
 * on the filtered edge), S/common/predict.c:234-336, :499-748 */
void xo_lowres_intra_pred(int mode, const uint8_t *l0, int stride, int bx, int by, uint8_t out[64]);
int xo_lowres_intra_cost(const uint8_t *l0, int stride, int bx, int by, int mbcmp_satd);

/* ---------------- in-loop deblocking of one progressive frame: S/common/frame.c:376-800 ----------------
 * Arrays use the reference's own layouts (S/common/common.h:420-436) with i_mb_stride = mb_width: type/qp/transform8x8
 * per macroblock; nnz[mb][24] (first 16 = luma 4x4 blocks, raster x+4y); ref[l] per 8x8 block on the frame-wide 8x8 grid
 * (stride 2*mb_width); mv[l] per 4x4 block on the frame-wide 4x4 grid (stride 4*mb_width).  Planes are filtered in place
 * (pixel 0,0 pointers; luma stride g->stride). */
typedef struct {
    int alpha_c0_offset, beta_offset;   /* sh.i_alpha_c0_offset, sh.i_beta_offset (= 2 x the --deblock arguments) */
    int chroma_qp_offset;               /* pps->i_chroma_qp_index_offset == param.analyse.i_chroma_qp_offset */
    int b_slice_b;                      /* sh.i_type == SLICE_TYPE_B */
    int b_psub8x8;                      /* param.analyse.inter & X264_ANALYSE_PSUB8x8 */
    int b_cavlc_8x8dct;                 /* !pps->b_cabac && pps->b_transform_8x8_mode */
    const int8_t *type, *qp, *transform8x8;
    const uint8_t (*nnz)[24];
    const int8_t *ref[2];
    const int16_t (*mv[2])[2];
} xo_deblock_in;
void xo_frame_deblock(const xo_geom *g, const xo_deblock_in *d, uint8_t *py, uint8_t *pu, uint8_t *pv, int stride_c);

/* ---------------- whole-frame analysis metrics (no inter-macroblock dependencies) ---------------- */
int64_t xo_frame_ssd(const uint8_t *p1, int s1, const uint8_t *p2, int s2, int width, int height);                 /* pixel.c:98-136 */
void xo_frame_mb_energy(const xo_geom *g, const uint8_t *py, const uint8_t *pu, const uint8_t *pv, int stride_c, uint32_t *out); /* ratecontrol.c:171-191 */
void xo_frame_mb_hadamard_ac(const xo_geom *g, const uint8_t *py, uint64_t *out);                                  /* pixel.c:306-358 */
void xo_frame_ssim_sums(const uint8_t *p1, int s1, const uint8_t *p2, int s2, int width, int height, int (*sums)[4]); /* pixel.c:435-460 */
float xo_frame_ssim(const uint8_t *p1, int s1, const uint8_t *p2, int s2, int width, int height);                 /* pixel.c:484-509 */
/* x264_adaptive_quant_frame (ratecontrol.c:233-249) given the energies: f_qp_offset[mb], i_inv_qscale_factor[mb] */
void xo_aq_from_energy(const uint32_t *energy, int n, float aq_strength, float *qp_offset, uint16_t *inv_qscale);
/* the same through the frame: energies + float formula in one go (reference: x264_adaptive_quant_frame itself) */
void xo_frame_aq(const xo_geom *g, const uint8_t *py, const uint8_t *pu, const uint8_t *pv, int stride_c, float aq_strength, float *qp_offset,
                 uint16_t *inv_qscale);

#ifdef __cplusplus
}
#endif
#endif
