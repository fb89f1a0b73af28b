/* x264_cuda.h — C ABI of the B200 (sm_100a) back-end for the data-parallel core of x264 (core build 66,
 * snapshot 20090216-2245; S/ = the reference tree).  Plain pointers and sizes only; no C++/torch types.
 *
 * Two layers:
 *  1. frame-batched entry points  x264_cuda_*      — the performance path.  Each one replaces a loop the
 *     reference runs per macroblock / per row with one (or a few) kernel launches over a whole frame.
 *  2. plugin-table fillers        x264_*_init_cuda — declared in x264_cuda_tables.h; they fill the reference's
 *     x264_pixel/mc/dct/quant function tables so the library is a drop-in behind *_init(cpu).
 *
 * Error convention (S/encoder/encoder.c:634-645, S/x264.c:759-762): functions return 0 on success, -1 on
 * failure; x264_cuda_error() gives the message a caller would pass to x264_log(h, X264_LOG_ERROR, ...).
 * There is NO CPU fallback: without a usable CUDA device every entry point fails.
 *
 * Threading (S/common/common.h:50, S/encoder/encoder.c:780): one x264_cuda_t per x264_t thread context; a
 * context owns one CUDA stream and is not re-entrant; different contexts may be used concurrently.
 * Device frames may be shared between contexts (one thread's fdec is the next thread's reference), but every call is
 * asynchronous on its own context's stream and NOTHING orders two contexts' streams: before another context reads a frame,
 * the producing context must have been waited for — x264_cuda_fence_wait(x264_cuda_fence_record(producer)) or
 * x264_cuda_synchronize(producer) — exactly where the reference waits on x264_frame_cond_wait (S/encoder/encoder.c:1176-1180).
 */
#ifndef X264_CUDA_H
#define X264_CUDA_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define X264_CUDA_API __attribute__((visibility("default")))

typedef struct x264_cuda_t x264_cuda_t;             /* device context */
typedef struct x264_cuda_frame_t x264_cuda_frame_t; /* device mirror of the planes of one x264_frame_t */

/* partition ids == enum in S/common/pixel.h:30-42 */
enum { X264_CUDA_PIXEL_16x16 = 0, X264_CUDA_PIXEL_16x8, X264_CUDA_PIXEL_8x16, X264_CUDA_PIXEL_8x8,
       X264_CUDA_PIXEL_8x4, X264_CUDA_PIXEL_4x8, X264_CUDA_PIXEL_4x4 };

/* ------------------------------------------------------------------ context ------------------------ */
/* hook: x264_encoder_open, next to x264_pixel_init..x264_quant_init (S/encoder/encoder.c:731-745) */
X264_CUDA_API int x264_cuda_open(x264_cuda_t **ctx, int device);
X264_CUDA_API void x264_cuda_close(x264_cuda_t *ctx);
X264_CUDA_API const char *x264_cuda_error(const x264_cuda_t *ctx); /* ctx may be NULL: last open() error */
/* run on a caller-provided cudaStream_t (e.g. the framework's current stream); NULL restores the own stream */
X264_CUDA_API int x264_cuda_set_stream(x264_cuda_t *ctx, void *cuda_stream);
X264_CUDA_API void *x264_cuda_get_stream(x264_cuda_t *ctx);
X264_CUDA_API int x264_cuda_synchronize(x264_cuda_t *ctx);
/* how the host thread waits for results inside the host-array entry points: 0 (default) spins in cudaStreamSynchronize — lowest latency;
 * 1 sleeps on a blocking-sync event — for many frame threads (S/encoder/encoder.c:1569-1608) sharing fewer host cores */
X264_CUDA_API int x264_cuda_set_blocking_wait(x264_cuda_t *ctx, int on);
/* number of kernel launches issued by this context so far (bench.py's gpu_launches) */
X264_CUDA_API long long x264_cuda_launch_count(const x264_cuda_t *ctx);
X264_CUDA_API int x264_cuda_sm_count(const x264_cuda_t *ctx);
/* roofline denominator for the SAD kernels: measured whole-GPU rate of VABSDIFF4.U8.ACC (thread-ops / second) */
X264_CUDA_API int x264_cuda_measure_int_pipe(x264_cuda_t *ctx, double *sad4_per_sec);

/* Page-locked host memory.  Every entry point accepts ordinary (pageable) host pointers; job lists, result arrays and picture
 * planes that live in memory from x264_cuda_host_alloc (a drop-in for x264_malloc, S/common/common.c:700-721) or registered
 * with x264_cuda_host_register are DMA'd directly instead of being copied through the context's staging buffer. */
X264_CUDA_API void *x264_cuda_host_alloc(size_t bytes);
X264_CUDA_API void x264_cuda_host_free(void *p);
X264_CUDA_API int x264_cuda_host_register(void *p, size_t bytes);
X264_CUDA_API int x264_cuda_host_unregister(void *p);

/* ------------------------------------------------------------------ frames ------------------------- */
/* Mirrors x264_frame_new (S/common/frame.c:29-152): a luma plane with PADH=PADV=32 borders, optionally the
 * three half-pel planes (filtered[1..3]), the integral image(s) and the four half-resolution planes.
 * Device layout: same padded geometry as the reference, stride rounded up to 128 bytes. */
#define X264_CUDA_FRAME_HPEL      1 /* filtered[1..3]: i_subpel_refine > 0 (frame.c:70-77) */
#define X264_CUDA_FRAME_INTEGRAL  2 /* me >= ESA (frame.c:99-104) */
#define X264_CUDA_FRAME_INTEGRAL4 4 /* + 4x4 sums: b_have_sub8x8_esa */
#define X264_CUDA_FRAME_LOWRES    8 /* b_have_lowres (frame.c:79-97) */
#define X264_CUDA_FRAME_CHROMA   16 /* plane[1], plane[2] (4:2:0), 16-px borders (frame.c:61-65) */

typedef struct x264_cuda_geom_t {
    int width, height;       /* picture size */
    int mb_width, mb_height; /* macroblocks */
    int stride, lines;       /* DEVICE luma stride (bytes), mod-16 lines */
    int stride_lowres, width_lowres, lines_lowres;
    int flags;
} x264_cuda_geom_t;

enum { X264_CUDA_PLANE_FULL = 0, X264_CUDA_PLANE_H = 1, X264_CUDA_PLANE_V = 2, X264_CUDA_PLANE_C = 3, /* filtered[0..3] */
       X264_CUDA_PLANE_LOWRES = 4, /* +0..3: lowres[0..3] */
       X264_CUDA_PLANE_INTEGRAL = 8, X264_CUDA_PLANE_INTEGRAL4 = 9, X264_CUDA_PLANE_CB = 10, X264_CUDA_PLANE_CR = 11 };

X264_CUDA_API x264_cuda_frame_t *x264_cuda_frame_new(x264_cuda_t *ctx, int width, int height, int flags);
X264_CUDA_API void x264_cuda_frame_delete(x264_cuda_frame_t *frame);
X264_CUDA_API void x264_cuda_frame_geometry(const x264_cuda_frame_t *frame, x264_cuda_geom_t *g);
/* device address of pixel (0,0) of a plane (NULL if the frame lacks it) — for zero-copy producers/consumers */
X264_CUDA_API void *x264_cuda_frame_plane(const x264_cuda_frame_t *frame, int plane);

/* host -> device copy of a cols x rows luma rectangle anchored at pixel (0,0); src points at pixel (0,0) of
 * the host plane (e.g. x264_frame_t.plane[0]).  No border handling. */
X264_CUDA_API int x264_cuda_frame_upload(x264_cuda_t *ctx, x264_cuda_frame_t *frame, const uint8_t *src, int src_stride,
                                         int cols, int rows);
/* chroma: plane = X264_CUDA_PLANE_CB / _CR, cols x rows in chroma samples */
X264_CUDA_API int x264_cuda_frame_upload_chroma(x264_cuda_t *ctx, x264_cuda_frame_t *frame, int plane, const uint8_t *src,
                                                int src_stride, int cols, int rows);
/* the same from a device-resident source (pitch-linear) */
X264_CUDA_API int x264_cuda_frame_upload_dev(x264_cuda_t *ctx, x264_cuda_frame_t *frame, const void *dsrc, int src_stride,
                                             int cols, int rows);
/* device -> host copy of a whole padded plane: dst points at the host buffer start (row -32, col -32);
 * elem size 1 (pixel planes) or 2 (integral).  dst_stride in elements. */
X264_CUDA_API int x264_cuda_frame_download(x264_cuda_t *ctx, const x264_cuda_frame_t *frame, int plane, void *dst,
                                           int dst_stride);

/* ≡ x264_frame_expand_border_mod16 + x264_frame_expand_border, luma (S/common/frame.c:304-331, :240-267) */
X264_CUDA_API int x264_cuda_frame_expand_border(x264_cuda_t *ctx, x264_cuda_frame_t *frame);
/* x264_frame_expand_border_mod16 only (S/common/frame.c:304-331): replicate the last column / row up to a multiple of 16 — all the
 * reference does to fenc (S/encoder/encoder.c:1413-1416); reference frames need the full x264_cuda_frame_expand_border */
X264_CUDA_API int x264_cuda_frame_expand_border_mod16(x264_cuda_t *ctx, x264_cuda_frame_t *frame);
/* ≡ x264_frame_filter(h, frame, 0, 1) + x264_frame_expand_border_filtered(h, frame, 0, 1): the three half-pel
 * planes and the integral image(s) of a whole (border-expanded) frame (S/common/mc.c:404-463,
 * S/common/frame.c:269-295).  Hook: x264_fdec_filter_row, S/encoder/encoder.c:1016-1023. */
X264_CUDA_API int x264_cuda_frame_filter(x264_cuda_t *ctx, x264_cuda_frame_t *frame);
/* ≡ x264_frame_init_lowres pixel work (S/common/mc.c:306-321 + frame.c:297-302). Hook: encoder.c:1418. */
X264_CUDA_API int x264_cuda_frame_init_lowres(x264_cuda_t *ctx, x264_cuda_frame_t *frame);

/* ------------------------------------------------------------------ MV cost tables ------------------ */
/* Upload the HOST-computed lambda*bits table of one qp: p_cost_mv as built by x264_mb_analyse_load_costs
 * (S/encoder/analyse.c:182-203), `table` = the malloc base, 4*4*2048+1 int16 entries (centre at +2*4*2048).
 * Float-derived on the host — never recomputed on the device.  A qp the caller never uploads falls back to the
 * library's own host-side builder below at first use (identical to the reference's table for all 52 qps in our tests). */
X264_CUDA_API int x264_cuda_set_cost_mv(x264_cuda_t *ctx, int qp, const int16_t *table);
/* host-side mirror of that table builder for standalone use (tests, bench): fills table[4*4*2048+1] */
X264_CUDA_API void x264_cuda_host_cost_mv(int qp, int16_t *table);
X264_CUDA_API int x264_cuda_host_lambda(int qp);
X264_CUDA_API int x264_cuda_host_lambda2(int qp); /* x264_lambda2_tab[qp], S/encoder/analyse.c:150-160 */

/* ------------------------------------------------------------------ motion search ------------------- */
/* One job == one x264_me_search_ref() call (S/encoder/me.c:156-631) up to and including the ESA/TESA loop,
 * for i_subpel_refine < 3 (full-pel predictor stage, me.c:207-229).  Jobs are independent; the reference's
 * sequential neighbour-predictor order lives in how the caller derives mvp/mvc. */
#define X264_CUDA_ME_MAX_MVC 12
#define X264_CUDA_ME_SEEDED 1 /* skip the predictor stage: (seed_mv, seed_cost) are bmx,bmy,bcost at me.c:229 */
#define X264_CUDA_ME_TESA   2 /* TESA candidate thresholds + final fpelcmp pass (me.c:491-578) */
#define X264_CUDA_ME_FPEL_SATD 4 /* fpelcmp is SATD (mbcmp_init, S/encoder/encoder.c:608-618) */
#define X264_CUDA_ME_MBCMP_SATD 8 /* mbcmp is SATD (user subme > 1) */
#define X264_CUDA_ME_CHROMA 32    /* h->mb.b_chroma_me (P slices, subme >= 5, --chroma-me): the sub-pel SATD cost of partitions >= 8x8 adds
                                   * mc_chroma + mbcmp of U and V (S/encoder/me.c:655-677); honoured by x264_cuda_me_search_small when fenc
                                   * and fref both carry chroma planes (X264_CUDA_FRAME_CHROMA, borders expanded), else the job is rejected
                                   * (cost -1) */
typedef struct x264_cuda_me_job_t {
    int16_t bx, by;              /* block position in luma pixels */
    uint8_t i_pixel;             /* X264_CUDA_PIXEL_* */
    uint8_t qp;                  /* h->mb.i_qp: selects the cost table */
    uint8_t i_mvc;               /* number of entries used in mvc[] */
    uint8_t flags;               /* X264_CUDA_ME_* */
    int16_t mvp[2];              /* m->mvp (qpel) */
    int16_t mv_min_fpel[2];      /* h->mb.mv_min_fpel */
    int16_t mv_max_fpel[2];      /* h->mb.mv_max_fpel */
    int16_t seed_mv[2];          /* only with X264_CUDA_ME_SEEDED */
    int32_t seed_cost;
    int16_t mvc[X264_CUDA_ME_MAX_MVC][2]; /* extra predictors (qpel) */
    int16_t mv_min_spel[2];      /* h->mb.mv_min_spel / mv_max_spel: used by x264_cuda_me_search_small (sub-pel stages) */
    int16_t mv_max_spel[2];
} x264_cuda_me_job_t;            /* 84 bytes */

typedef struct x264_cuda_me_result_t {
    int16_t bmx, bmy;            /* full-pel winner at me.c:601 */
    int32_t bcost;
    int16_t seed_mx, seed_my;    /* bmx,bmy,bcost entering the ESA loop (me.c:229) */
    int32_t seed_cost;
} x264_cuda_me_result_t;         /* 16 bytes */

/* host arrays in, host arrays out (H2D of jobs, launch, D2H of results, stream-synchronised on return) */
X264_CUDA_API int x264_cuda_me_search(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, const x264_cuda_frame_t *fref,
                                      int me_range, const x264_cuda_me_job_t *jobs, int n_jobs,
                                      x264_cuda_me_result_t *results);
/* device arrays in/out, asynchronous on the context's stream */
X264_CUDA_API int x264_cuda_me_search_dev(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, const x264_cuda_frame_t *fref,
                                          int me_range, const void *d_jobs, int n_jobs, void *d_results);
/* m->mv / m->cost / m->cost_mv from a result, i.e. me.c:603-630 for i_subpel_refine < 2 (host arithmetic) */
X264_CUDA_API void x264_cuda_me_finish(const x264_cuda_me_job_t *job, const x264_cuda_me_result_t *res,
                                       const int16_t *cost_table, int mv_max_spel_y, int16_t mv[2], int *cost, int *cost_mv);

/* ------------------------------------------------------------------ candidate grids + host replay ---- */
/* For callers that only learn a block's predictors sequentially (mvp depends on the neighbours' final vectors, S/encoder/analyse.c):
 * the device writes out the SAD of all nine inter partitions of a macroblock at every integer vector of a window around a centre
 * the caller guesses (e.g. the lowres lookahead's vector), and x264_cuda_host_esa_replay then runs x264_me_search_ref's predictor
 * stage and ESA loop (S/encoder/me.c:182-229, :449-492) on that grid with the exact mvp — same vectors, same costs.
 *   grid[job][part 0..8][j][i] (uint16) = x264_pixel_sad_<part>( fenc block, fref block at mv (cx - radius + i, cy - radius + j) ),
 *   i < X264_CUDA_GRID_W(radius), j < X264_CUDA_GRID_H(radius); 0xffff outside [mv_min_fpel, mv_max_fpel] (plus the 3 columns
 *   beyond mv_max_fpel[0] that the ESA loop's width rounding tests, me.c:456-457) or for masked partitions.
 * Partition order as in x264_cuda_me_mb_job_t: 16x16 | 16x8 top,bottom | 8x16 left,right | 8x8 TL,TR,BL,BR. */
#define X264_CUDA_GRID_W(radius) ((2 * (radius) + 1 + 3) & ~3)
#define X264_CUDA_GRID_H(radius) (2 * (radius) + 1)
typedef struct x264_cuda_grid_job_t {
    int16_t mb_x, mb_y;
    int16_t cx, cy;                          /* grid centre, full-pel */
    int16_t mv_min_fpel[2], mv_max_fpel[2];  /* h->mb.mv_{min,max}_fpel */
    uint16_t part_mask;                      /* bit p: partition p wanted */
    uint16_t reserved;
} x264_cuda_grid_job_t;                      /* 20 bytes */
X264_CUDA_API int x264_cuda_sad_grid(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, const x264_cuda_frame_t *fref, int radius,
                                     const x264_cuda_grid_job_t *jobs, int n_jobs, uint16_t *grid);
X264_CUDA_API int x264_cuda_sad_grid_dev(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, const x264_cuda_frame_t *fref, int radius,
                                         const void *d_jobs, int n_jobs, void *d_grid);
/* Quadrant grids — the form the live encoder uses (integration/x264_b200_hooks.c).  Every partition SAD is a plain sum of the SADs of
 * the macroblock's four 8x8 quadrants (PIXEL_SAD_C, S/common/pixel.c:40-56), so only those are written, interleaved per position:
 *   grid[job][j][i][q] (uint16), q = 0 TL | 1 TR | 2 BL | 3 BR, same window, limits and 0xffff marking as above; part_mask is ignored.
 * 8 bytes per position instead of 18.  async != 0: the call returns once the work is queued on the context's stream; `jobs` and `grid`
 * must then be page-locked (x264_cuda_host_alloc) and stay untouched until a fence recorded after the call has been waited for. */
#define X264_CUDA_GRID_QUAD_BYTES(radius) ((size_t)X264_CUDA_GRID_W(radius) * X264_CUDA_GRID_H(radius) * 4 * sizeof(uint16_t))
X264_CUDA_API int x264_cuda_sad_grid_quad(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, const x264_cuda_frame_t *fref, int radius,
                                          const x264_cuda_grid_job_t *jobs, int n_jobs, uint16_t *grid, int async);
X264_CUDA_API int x264_cuda_sad_grid_quad_dev(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, const x264_cuda_frame_t *fref, int radius,
                                              const void *d_jobs, int n_jobs, void *d_grid);
/* the latency path: no copies — the kernel reads `jobs` from and writes `grid` to page-locked host memory directly, on a high-priority
 * stream of its own, and the call returns when the grids are in host memory.  The frames must be complete (nothing pending on them). */
X264_CUDA_API int x264_cuda_sad_grid_quad_direct(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, const x264_cuda_frame_t *fref, int radius,
                                                 const x264_cuda_grid_job_t *jobs, int n_jobs, uint16_t *grid);
/* size the device ring the asynchronous calls carve their job copies and grids from (default: 64 MB or four calls' worth); a caller
 * that knows its working set reserves it once, so that the ring never has to drain the stream in the middle of a frame */
X264_CUDA_API int x264_cuda_grid_ring_reserve(x264_cuda_t *ctx, size_t bytes);
/* completion markers for asynchronous calls: record returns a fence covering everything queued on the context so far (NULL on failure);
 * wait blocks until it has completed and releases it */
X264_CUDA_API void *x264_cuda_fence_record(x264_cuda_t *ctx);
X264_CUDA_API int x264_cuda_fence_wait(x264_cuda_t *ctx, void *fence);
/* Host side, no device involved: the full-pel part of x264_me_search_ref for --me esa on one partition's grid plane
 * (grid_part = grid + (job_index * 9 + part) * GW * GH).  job carries mvp, mvc[], i_mvc, qp limits exactly as for
 * x264_cuda_me_search (bx/by are not used); cost_table = p_cost_mv of the qp (x264_cuda_host_cost_mv).  Returns 0 and fills res
 * (bmx, bmy, bcost, seed_*), or 1 when a vector the search must test lies outside the grid (then enlarge the radius or fall back
 * to x264_cuda_me_search for that block). */
X264_CUDA_API int x264_cuda_host_esa_replay(const uint16_t *grid_part, int radius, int cx, int cy, const x264_cuda_me_job_t *job, int me_range,
                                            const int16_t *cost_table, x264_cuda_me_result_t *res);

/* ------------------------------------------------------------------ iterative + sub-pel search -------- */
/* One job == one complete x264_me_search_ref() (S/encoder/me.c:156-631) for the small iterative methods, or the tail
 * of one for ESA/TESA:
 *   method X264_CUDA_ME_METHOD_DIA / _HEX : predictor stage (full-pel for subme < 3, quarter-pel via get_ref for
 *       subme >= 3: me.c:188-229), the search loop (me.c:233-305), "-> qpel mv" (:603-620), refine_subpel (:622-628,
 *       :680-778; b_chroma_me = 0);
 *   method X264_CUDA_ME_METHOD_UMH        : the uneven-cross multi-hexagon-grid search (me.c:306-447: cross, early terminations,
 *       range adapted to the agreement of mvp and mvc[], hexagon grid) followed by the hexagon refinement; same tail;
 *   method X264_CUDA_ME_METHOD_TESA       : the complete --me tesa search: predictor stage, ADS threshold on the
 *       integral image (pixf.ads, S/common/pixel.c:515-559), SAD threshold list, keep merange/2, fpelcmp on the keepers
 *       (me.c:491-578), then the tail as above.  fref needs X264_CUDA_FRAME_INTEGRAL (+_INTEGRAL4 for sub-8x8 blocks);
 *   method X264_CUDA_ME_METHOD_REFINE_QPEL: x264_me_refine_qpel (me.c:633-643) — refine_subpel with b_refine_qpel = 1 from job.seed_mv
 *       (= m->mv, QUARTER-pel) and job.seed_cost (= m->cost after the caller's i_ref_cost adjustment); bmx/bmy in the result are
 *       not meaningful;
 *   method X264_CUDA_ME_METHOD_SEEDED     : job.seed_mv/seed_cost carry the full-pel winner of x264_cuda_me_search
 *       (ESA/TESA, subme < 3 predictor stage) and only ":603-631" runs.
 * fref must have the half-pel planes (X264_CUDA_FRAME_HPEL, x264_cuda_frame_filter) when subme >= 2. */
enum { X264_CUDA_ME_METHOD_DIA = 0, X264_CUDA_ME_METHOD_HEX = 1, X264_CUDA_ME_METHOD_UMH = 2, X264_CUDA_ME_METHOD_TESA = 4, X264_CUDA_ME_METHOD_SEEDED = 8,
       X264_CUDA_ME_METHOD_REFINE_QPEL = 16 };
typedef struct x264_cuda_me_final_t {
    int16_t mv[2];               /* m->mv (qpel) */
    int32_t cost;                /* m->cost */
    int32_t cost_mv;             /* m->cost_mv */
    int16_t bmx, bmy;            /* full-pel winner before the sub-pel stages */
} x264_cuda_me_final_t;          /* 16 bytes */
X264_CUDA_API int x264_cuda_me_search_small(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, const x264_cuda_frame_t *fref,
                                            int method, int me_range, int subme, const x264_cuda_me_job_t *jobs, int n_jobs,
                                            x264_cuda_me_final_t *results);
X264_CUDA_API int x264_cuda_me_search_small_dev(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, const x264_cuda_frame_t *fref,
                                                int method, int me_range, int subme, const void *d_jobs, int n_jobs,
                                                void *d_results);

/* Function-level block metrics over PACKED operands (checkasm-style, S/tools/checkasm.c:222-295): block i of pix1 and
 * pix2 is a 16x16 byte tile (stride 16) whose top-left w x h corner is compared.  metric: 0 SAD, 1 SSD, 2 SATD, 3 SA8D
 * (16x16 and 8x8 only, like the reference table). */
X264_CUDA_API int x264_cuda_block_cmp(x264_cuda_t *ctx, int metric, int i_pixel, int n, const uint8_t *pix1, const uint8_t *pix2,
                                      int *out);

/* ------------------------------------------------------------------ lowres lookahead ------------------ */
/* One call == one x264_slicetype_frame_cost(h, a, frames, p0, p1, b) evaluation (S/encoder/slicetype.c:256-355).  Default
 * form (:318-330): every interior 8x8 block of the half-resolution frame b gets
 * x264_slicetype_mb_cost (:43-248) — bidirectional direct-like try, one DIA/HEX + subme-4 search per list with the
 * reverse-raster neighbour predictors, bidirectional retry, intra (ten 8x8 predictions) — and the block costs are
 * summed.  Blocks run as an anti-diagonal wavefront (x+2y descending) inside one launch; the per-frame
 * state the reference keeps in x264_frame_t (lowres_mvs, lowres_mv_costs, i_intra_cost) lives on the device with the
 * frame.  Call order and do_search/b_intra_calculated bookkeeping stay with the host, as in the reference.
 * fenc = frames[b], fref0 = frames[p0], fref1 = frames[p1] (may equal fref0 when b == p1); all three need
 * X264_CUDA_FRAME_LOWRES + x264_cuda_frame_init_lowres. */
#define X264_CUDA_LOWRES_WEIGHTED_BIPRED 16 /* param.analyse.b_weighted_bipred; other flag bits: X264_CUDA_ME_MBCMP_SATD / _FPEL_SATD */
#define X264_CUDA_LOWRES_VBV 32             /* h->param.rc.i_vbv_buffer_size != 0: the all-blocks form of slicetype.c:300-316 (the frame-edge
                                             * blocks are searched too, and i_row_satds[b-p0][p1-b][] is produced); only through the _rc entries */
typedef struct x264_cuda_lowres_params_t {
    int p0, p1, b;
    int me_method;          /* the user's --me (X264_ME_*: 0 dia, 1 hex, ...); the lookahead uses min(HEX, me) */
    int me_range;
    int flags;
    int do_search[2];       /* slicetype.c:279-282 */
    int b_intra_calculated; /* frames[b]->b_intra_calculated */
} x264_cuda_lowres_params_t;
typedef struct x264_cuda_lowres_result_t {
    int score;              /* sum of block costs BEFORE the B-frame scaling of slicetype.c:338-339 */
    int intra_mbs;          /* i_intra_mbs[b-p0]   (b == p1 only) */
    int intra_cost_sum;     /* i_cost_est[0][0]    (b == p1 only) */
    int score_aq;           /* i_cost_est_aq[b-p0][p1-b]: the interior sum of (cost * i_inv_qscale_factor + 128) >> 8 (slicetype.c:307-315, :324-329);
                             * equals score without an inv_qscale array, 0 for frames of at most 2 macroblocks in a dimension (:293-298) */
} x264_cuda_lowres_result_t;
/* (re)allocates the lookahead state of a frame: n_dist = i_bframe + 1 distances per list, zero-filled like frame.c:92-98 */
X264_CUDA_API int x264_cuda_frame_lookahead_alloc(x264_cuda_t *ctx, x264_cuda_frame_t *frame, int n_dist);
/* host access to lowres_mvs[list][dist] (int16[mb][2]) / lowres_mv_costs[list][dist] (int[mb]) / i_intra_cost (uint16[mb]);
 * NULL pointers are skipped */
X264_CUDA_API int x264_cuda_frame_lookahead_get(x264_cuda_t *ctx, const x264_cuda_frame_t *frame, int list, int dist, int16_t *mvs,
                                                int *costs, uint16_t *intra_cost);
X264_CUDA_API int x264_cuda_frame_lookahead_set(x264_cuda_t *ctx, x264_cuda_frame_t *frame, int list, int dist, const int16_t *mvs,
                                                const int *costs, const uint16_t *intra_cost);
X264_CUDA_API int x264_cuda_lowres_frame_cost(x264_cuda_t *ctx, x264_cuda_frame_t *fenc, const x264_cuda_frame_t *fref0,
                                              const x264_cuda_frame_t *fref1, const x264_cuda_lowres_params_t *params,
                                              x264_cuda_lowres_result_t *result);
/* Several evaluations in ONE launch (up to 64): the slicetype decision asks for many (p0,p1,b) costs of a lookahead window
 * (S/encoder/slicetype.c:357-470), and each wavefront alone occupies only a few dozen warps.  The evaluations of a batch must be
 * independent: no two may search the same (frame, list, distance) state, and none may read (ref1's list-0 vectors) what another
 * one of the batch searches.  E.g. all P costs cost(i-1, i, i) of a window, or all B costs between two fixed P frames whose own
 * vectors are already cached.  results[i] as for the single call. */
X264_CUDA_API int x264_cuda_lowres_frame_cost_batch(x264_cuda_t *ctx, int n_evals, x264_cuda_frame_t *const *fencs,
                                                    const x264_cuda_frame_t *const *fref0s, const x264_cuda_frame_t *const *fref1s,
                                                    const x264_cuda_lowres_params_t *params, x264_cuda_lowres_result_t *results);
/* The rate-control forms: inv_qscale = frames[b]->i_inv_qscale_factor (uint16[mb_width*mb_height], host; NULL when rc.i_aq_mode is off)
 * weights every block cost into result->score_aq; with X264_CUDA_LOWRES_VBV in params->flags the evaluation covers EVERY block in
 * the reference's reverse-raster dependency order and row_satd[mb_height] (host) receives frames[b]->i_row_satds[b-p0][p1-b][] —
 * what x264_rc_analyse_slice hands to the VBV row predictor (slicetype.c:638-679, ratecontrol.c).  A batch is either all-VBV or all-default. */
X264_CUDA_API int x264_cuda_lowres_frame_cost_rc(x264_cuda_t *ctx, x264_cuda_frame_t *fenc, const x264_cuda_frame_t *fref0,
                                                 const x264_cuda_frame_t *fref1, const x264_cuda_lowres_params_t *params,
                                                 const uint16_t *inv_qscale, x264_cuda_lowres_result_t *result, int *row_satd);
X264_CUDA_API int x264_cuda_lowres_frame_cost_batch_rc(x264_cuda_t *ctx, int n_evals, x264_cuda_frame_t *const *fencs,
                                                       const x264_cuda_frame_t *const *fref0s, const x264_cuda_frame_t *const *fref1s,
                                                       const x264_cuda_lowres_params_t *params, const uint16_t *const *inv_qscales,
                                                       x264_cuda_lowres_result_t *results, int *const *row_satds);

/* ------------------------------------------------------------------ in-loop deblocking ---------------- */
/* x264_frame_deblock (S/common/frame.c:794-799 -> x264_frame_deblock_row :621-792) of one progressive frame, in place on the
 * device frame's luma and chroma planes (X264_CUDA_FRAME_CHROMA).  The per-macroblock arrays are the reference's own, in its
 * layouts (S/common/common.h:420-436, i_mb_stride == mb_width): h->mb.type, h->mb.qp, h->mb.mb_transform_size,
 * h->mb.non_zero_count ([mb][16+4+4]), h->mb.ref[l] (8x8 grid, stride 2*mb_width), h->mb.mv[l] (4x4 grid, stride 4*mb_width);
 * list-1 arrays may be NULL unless b_slice_b.  MBAFF (sh.b_mbaff) is not supported.  Called where the reference calls
 * x264_frame_deblock_row for the last row of a frame (S/encoder/encoder.c:1009-1014, single-thread schedule). */
typedef struct x264_cuda_deblock_params_t {
    int alpha_c0_offset, beta_offset;   /* h->sh.i_alpha_c0_offset, h->sh.i_beta_offset */
    int chroma_qp_offset;               /* h->pps->i_chroma_qp_index_offset */
    int b_slice_b;                      /* h->sh.i_type == SLICE_TYPE_B */
    int b_psub8x8;                      /* h->param.analyse.inter & X264_ANALYSE_PSUB8x8 */
    int b_cavlc_8x8dct;                 /* !h->pps->b_cabac && h->pps->b_transform_8x8_mode (nnz munging, frame.c:334-373) */
} x264_cuda_deblock_params_t;
X264_CUDA_API int x264_cuda_frame_deblock(x264_cuda_t *ctx, x264_cuda_frame_t *fdec, const x264_cuda_deblock_params_t *params,
                                          const int8_t *type, const int8_t *qp, const int8_t *transform8x8, const uint8_t (*nnz)[24],
                                          const int8_t *ref0, const int16_t (*mv0)[2], const int8_t *ref1, const int16_t (*mv1)[2]);
X264_CUDA_API int x264_cuda_frame_deblock_dev(x264_cuda_t *ctx, x264_cuda_frame_t *fdec, const x264_cuda_deblock_params_t *params,
                                              const int8_t *d_type, const int8_t *d_qp, const int8_t *d_transform8x8, const uint8_t *d_nnz,
                                              const int8_t *d_ref0, const int16_t *d_mv0, const int8_t *d_ref1, const int16_t *d_mv1);

/* ------------------------------------------------------------------ whole-frame analysis metrics ------ */
/* per-frame / per-macroblock measures without inter-macroblock dependencies.  The device computes the
 * integer parts; the float tails stay on the host (x264_cuda_host_*) because their results depend on float evaluation order.
 * plane: X264_CUDA_PLANE_FULL / _CB / _CR (or a half-pel plane id). */
/* x264_pixel_ssd_wxh (S/common/pixel.c:98-136) over the width x height pixels at (x0,y0) of one plane of two frames: PSNR input
 * (S/encoder/encoder.c:1034-1046; the reference's per-row slabs add up to one whole-frame call) */
X264_CUDA_API int x264_cuda_frame_ssd(x264_cuda_t *ctx, const x264_cuda_frame_t *a, const x264_cuda_frame_t *b, int plane, int x0, int y0, int width,
                                      int height, int64_t *ssd);
/* ssim_4x4x2_core (pixel.c:435-460) for every 4x4 block of the region at (x0,y0): sums[(height/4)*(width/4)][4] = s1, s2, ss, s12;
 * x264_cuda_host_ssim_end then gives x264_pixel_ssim_wxh's value (pixel.c:462-509).  The encoder evaluates SSIM in row slabs
 * starting at x0 = 2 (encoder.c:1047-1056): pass the same (2, min_y, i_width-2, max_y-min_y) per slab for the identical float. */
X264_CUDA_API int x264_cuda_frame_ssim_sums(x264_cuda_t *ctx, const x264_cuda_frame_t *a, const x264_cuda_frame_t *b, int plane, int x0, int y0, int width,
                                            int height, int (*sums)[4]);
X264_CUDA_API float x264_cuda_host_ssim_end(const int (*sums)[4], int w4, int h4);
/* ac_energy_mb (S/encoder/ratecontrol.c:171-191) of every macroblock (frame needs X264_CUDA_FRAME_CHROMA): energy[mb_width*mb_height];
 * x264_cuda_host_aq then is x264_adaptive_quant_frame (:233-249): f_qp_offset[mb] and (optional) i_inv_qscale_factor[mb] */
X264_CUDA_API int x264_cuda_frame_mb_energy(x264_cuda_t *ctx, const x264_cuda_frame_t *frame, uint32_t *energy);
X264_CUDA_API void x264_cuda_host_aq(const uint32_t *energy, int n_mb, float aq_strength, float *qp_offset, uint16_t *inv_qscale);
/* x264_pixel_hadamard_ac_16x16 (pixel.c:306-358) of every macroblock of the luma plane (psy-rd source energy): out[mb] */
X264_CUDA_API int x264_cuda_frame_mb_hadamard_ac(x264_cuda_t *ctx, const x264_cuda_frame_t *frame, uint64_t *out);

/* ------------------------------------------------------------------ motion compensation -------------- */
/* Frame-batched x264_mb_mc_0xywh (S/common/macroblock.c:462-486): for each job the w x h luma block at (bx,by) is
 * predicted from fref at quarter-pel mv (mc_luma, S/common/mc.c:160-179) and, when both frames carry chroma planes, the
 * two w/2 x h/2 chroma blocks by 1/8-pel bilinear interpolation (mc_chroma, mc.c:205-236); results are written into fdec.
 * fref must be border-expanded and filtered (X264_CUDA_FRAME_HPEL); chroma borders are the caller's to upload. */
typedef struct x264_cuda_mc_job_t {
    int16_t bx, by;   /* luma position */
    int16_t mvx, mvy; /* quarter-pel */
    uint8_t w, h;     /* luma size: 4, 8 or 16 */
    uint8_t reserved[2];
} x264_cuda_mc_job_t; /* 12 bytes */
X264_CUDA_API int x264_cuda_mc_blocks(x264_cuda_t *ctx, const x264_cuda_frame_t *fref, x264_cuda_frame_t *fdec,
                                      const x264_cuda_mc_job_t *jobs, int n_jobs);
X264_CUDA_API int x264_cuda_mc_blocks_dev(x264_cuda_t *ctx, const x264_cuda_frame_t *fref, x264_cuda_frame_t *fdec,
                                          const void *d_jobs, int n_jobs);
/* Bi-predicted blocks, x264_mb_mc_01xywh (S/common/macroblock.c:508-546): the two lists' predictions (luma qpel fetch and chroma
 * mc_chroma from fref0 with mv0, from fref1 with mv1) blended by h->mc.avg (S/common/mc.c:52-125): weight 32 = rounded average,
 * otherwise implicit weighted bi-prediction with weight = h->mb.bipred_weight[ref0][ref1]. */
typedef struct x264_cuda_mc_bi_job_t {
    int16_t bx, by;       /* luma position */
    int16_t mv0[2], mv1[2];
    uint8_t w, h;         /* luma size: 4, 8 or 16 */
    uint8_t weight;       /* 32: plain average */
    uint8_t reserved;
} x264_cuda_mc_bi_job_t; /* 16 bytes */
X264_CUDA_API int x264_cuda_mc_blocks_bi(x264_cuda_t *ctx, const x264_cuda_frame_t *fref0, const x264_cuda_frame_t *fref1, x264_cuda_frame_t *fdec,
                                         const x264_cuda_mc_bi_job_t *jobs, int n_jobs);
X264_CUDA_API int x264_cuda_mc_blocks_bi_dev(x264_cuda_t *ctx, const x264_cuda_frame_t *fref0, const x264_cuda_frame_t *fref1,
                                             x264_cuda_frame_t *fdec, const void *d_jobs, int n_jobs);

/* ------------------------------------------------------------------ transform / quantisation --------- */
/* Quantiser tables exactly as x264_cqm_init leaves them in x264_t (S/common/set.c:68-174, S/common/common.h:294-304):
 * quant4_mf[list] -> uint16[52][16], quant4_bias likewise, dequant4_mf[list] -> int[6][16] (lists CQM_4IY,4PY,4IC,4PC),
 * quant8_mf[list] -> uint16[52][64], quant8_bias, dequant8_mf[list] -> int[6][64] (lists CQM_8IY, 8PY; may be NULL). */
X264_CUDA_API int x264_cuda_set_quant_tables(x264_cuda_t *ctx, const uint16_t *const quant4_mf[4], const uint16_t *const quant4_bias[4],
                                             const int *const dequant4_mf[4], const uint16_t *const quant8_mf[2],
                                             const uint16_t *const quant8_bias[2], const int *const dequant8_mf[2]);
/* host-side mirror of x264_cqm_init for standalone use: cqm_preset 0 flat / 1 JVT, default deadzones 21/11.
 * Output arrays are laid out as above (caller-allocated, contiguous over lists). */
X264_CUDA_API void x264_cuda_host_cqm_tables(int cqm_preset, uint16_t q4mf[4][52][16], uint16_t q4bias[4][52][16], int dq4[4][6][16],
                                             uint16_t q8mf[2][52][64], uint16_t q8bias[2][52][64], int dq8[2][6][64]);
/* convenience: build with x264_cuda_host_cqm_tables and upload */
X264_CUDA_API int x264_cuda_set_quant_preset(x264_cuda_t *ctx, int cqm_preset);

/* Function-level batches over PACKED blocks (host arrays; the checkasm-style entry points, S/tools/checkasm.c:469-688,
 * :976-1291).  kind 0: 4x4 blocks (16 samples / coefficients each), kind 1: 8x8 (64).  For block i:
 *   dct   = sub4x4_dct | sub8x8_dct8 (fenc_i - pred_i)                  (S/common/dct.c:122-155, :265-285)   -> dct_out
 *   nz    = quant_4x4 | quant_8x8 (dct, mf[cat_i][qp_i], bias[..])      (S/common/quant.c:33-58)             -> level_out, nz_out
 *   recon = pred_i + idct(dequant(level))                                (quant.c:76-146, dct.c:174-216, :322-341) -> recon_out
 * cat_i = CQM list id.  Any output pointer may be NULL. */
X264_CUDA_API int x264_cuda_block_residual(x264_cuda_t *ctx, int kind, int n, const uint8_t *fenc, const uint8_t *pred,
                                           const uint8_t *qp, const uint8_t *cat, int16_t *dct_out, int16_t *level_out,
                                           uint8_t *nz_out, uint8_t *recon_out);
/* DC chains on packed blocks of 16 (kind 0: dct4x4dc -> quant_4x4_dc -> idct4x4dc -> dequant_4x4_dc, the luma-DC path of
 * x264_mb_encode_i16x16, S/encoder/macroblock.c:246-262) : fwd_out / level_out / deq_out, nz_out. */
X264_CUDA_API int x264_cuda_block_dc(x264_cuda_t *ctx, int n, const int16_t *dc_in, const uint8_t *qp, const uint8_t *cat,
                                     int16_t *fwd_out, int16_t *level_out, uint8_t *nz_out, int16_t *deq_out);

/* Frame-batched residual coding of INTER macroblocks: the non-trellis, non-lossless inter branch of x264_macroblock_encode
 * (S/encoder/macroblock.c:596-742) + x264_mb_encode_8x8_chroma (:272-363).  fdec holds the motion-compensated
 * prediction of the listed macroblocks on entry (luma + chroma planes) and their reconstruction on return;
 * coefficient output mirrors h->dct / non_zero_count / cbp of each macroblock. */
#define X264_CUDA_RESID_8x8DCT   1 /* h->mb.b_transform_8x8 */
#define X264_CUDA_RESID_DECIMATE 2 /* B slice or param.analyse.b_dct_decimate */
typedef struct x264_cuda_resid_job_t {
    int16_t mb_x, mb_y;
    uint8_t qp, chroma_qp; /* h->mb.i_qp, h->mb.i_chroma_qp */
    uint8_t flags, reserved;
} x264_cuda_resid_job_t;     /* 8 bytes */
typedef struct x264_cuda_mb_coeffs_t {
    int16_t luma[256];       /* h->dct.luma4x4[0..15] (zigzag, block order) or h->dct.luma8x8[0..3] with 8x8dct; 0 where uncoded */
    int16_t chroma_ac[8][16];/* h->dct.luma4x4[16..23] */
    int16_t chroma_dc[2][4]; /* h->dct.chroma_dc */
    uint8_t nnz[27];         /* non_zero_count[x264_scan8[0..26]] */
    uint8_t cbp_luma, cbp_chroma, reserved[3];
} x264_cuda_mb_coeffs_t;     /* 816 bytes */
X264_CUDA_API int x264_cuda_residual_inter(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, x264_cuda_frame_t *fdec,
                                           const x264_cuda_resid_job_t *jobs, int n_jobs, x264_cuda_mb_coeffs_t *coeffs);
X264_CUDA_API int x264_cuda_residual_inter_dev(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, x264_cuda_frame_t *fdec,
                                               const void *d_jobs, int n_jobs, void *d_coeffs);

/* Frame-batched residual coding of I_16x16 macroblocks: what x264_macroblock_encode does for h->mb.i_type == I_16x16 (S/encoder/macroblock.c:512-530
 * predict_16x16[mode] + x264_mb_encode_i16x16 :184-270; :744-760 predict_8x8c[mode] + x264_mb_encode_8x8_chroma with b_inter = 0, :272-363), non-trellis,
 * non-lossless.  The prediction reads the UNFILTERED reconstruction of the left / top / top-left neighbours in fdec, so the listed macroblocks run
 * as a wavefront inside one launch: a macroblock must be listed AFTER any of those three neighbours that is also in the list (raster order always
 * qualifies); neighbours that are not listed must already be final in fdec.  On return fdec holds the reconstruction of the listed macroblocks
 * (luma + chroma).  Modes are the reference's enums: mode16 = enum intra16x16_pred_e (V H DC P DC_LEFT DC_TOP DC_128), mode_chroma =
 * enum intra_chroma_pred_e (DC H V P DC_LEFT DC_TOP DC_128) — e.g. what x264_cuda_intra_mb_costs picked.  flags: X264_CUDA_RESID_DECIMATE =
 * slice B || (b_dct_decimate && slice P) (:193).  Coefficient record: c.luma[16*i + 1..15] = AC levels of block i (slot 0 zero), luma_dc =
 * h->dct.luma16x16_dc, c.nnz[24] its flag; levels of blocks whose nnz is 0 read as zero.
 * The host-array entry takes the list in any dependency-respecting order and works through it by anti-diagonals (x + y ascending: a whole
 * diagonal is independent), returning results in list order; the _dev entry takes its tickets in LIST order, so a device-resident list
 * should itself be sorted by x + y — in raster order only a few macroblock rows are ever in flight. */
typedef struct x264_cuda_intra16_job_t {
    int16_t mb_x, mb_y;
    uint8_t qp, chroma_qp;     /* h->mb.i_qp, h->mb.i_chroma_qp */
    uint8_t mode16, mode_chroma;
    uint8_t flags, reserved[3];
} x264_cuda_intra16_job_t;     /* 12 bytes */
typedef struct x264_cuda_mb_coeffs_i16_t {
    x264_cuda_mb_coeffs_t c;
    int16_t luma_dc[16];       /* h->dct.luma16x16_dc */
} x264_cuda_mb_coeffs_i16_t;   /* 848 bytes */
X264_CUDA_API int x264_cuda_residual_intra16(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, x264_cuda_frame_t *fdec,
                                             const x264_cuda_intra16_job_t *jobs, int n_jobs, x264_cuda_mb_coeffs_i16_t *coeffs);
X264_CUDA_API int x264_cuda_residual_intra16_dev(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, x264_cuda_frame_t *fdec,
                                                 const void *d_jobs, int n_jobs, void *d_coeffs);

/* x264_macroblock_probe_skip (S/encoder/macroblock.c:797-883), frame-batched: skip[i] = 1 when macroblock i quantises to nothing
 * against its skip prediction — luma decimation total < 6 (:822-840) and, for each chroma plane whose SSD reaches
 * (x264_lambda2_tab[chroma_qp] + 32) >> 6, a zero 2x2 DC and an AC decimation total < 7 (:843-879).  Called per macroblock by
 * x264_macroblock_analyse for P (analyse.c:2207 and :1111, with the pskip mv) and B (analyse.c:2489, b_bidir = 1) slices.
 *   default               : the prediction is mc_luma(16x16) / mc_chroma(8x8) of (mvx, mvy) from fref (:809-819, :851-856); mvx/mvy are
 *                           h->mb.cache.pskip_mv ALREADY clipped to h->mb.mv_min/mv_max (:812-813).  fref needs HPEL | CHROMA.
 *   SKIP_PRED_IN_FDEC     : b_bidir = 1 — the prediction of that macroblock is already in fdec (e.g. from x264_cuda_mc_blocks_bi).
 *   SKIP_STORE_PRED       : also write the motion-compensated prediction into fdec, the side effect h->mb.b_skip_mc = 1 relies on
 *                           (the reference stops writing at its first early exit; here luma and chroma are always both written).
 * fref may be NULL when every job has PRED_IN_FDEC, fdec may be NULL when no job has either flag. */
#define X264_CUDA_SKIP_PRED_IN_FDEC 1
#define X264_CUDA_SKIP_STORE_PRED   2
typedef struct x264_cuda_skip_job_t {
    int16_t mb_x, mb_y;
    int16_t mvx, mvy;      /* quarter-pel */
    uint8_t qp, chroma_qp; /* h->mb.i_qp, h->mb.i_chroma_qp */
    uint8_t flags, reserved;
} x264_cuda_skip_job_t;    /* 12 bytes */
X264_CUDA_API int x264_cuda_probe_skip(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, const x264_cuda_frame_t *fref, x264_cuda_frame_t *fdec,
                                       const x264_cuda_skip_job_t *jobs, int n_jobs, uint8_t *skip);
/* device-resident job list / result bytes; any_mc / any_fdec say whether some job needs fref / fdec (validated up front) */
X264_CUDA_API int x264_cuda_probe_skip_dev(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, const x264_cuda_frame_t *fref,
                                           x264_cuda_frame_t *fdec, const void *d_jobs, int n_jobs, int any_mc, int any_fdec, void *d_skip);

/* ------------------------------------------------------------------ intra analysis from neighbouring macroblocks ------ */
/* The two stages of intra analysis whose inputs are the source macroblock and reconstructed pixels of NEIGHBOURING macroblocks only:
 *   - Intra16x16: every candidate of predict_16x16_mode_available (S/encoder/analyse.c:372-404) predicted (S/common/predict.c:40-170)
 *     and costed  mbcmp[PIXEL_16x16] + lambda * bs_size_ue(mode)  — x264_mb_analyse_intra, analyse.c:612-664;
 *   - chroma 8x8: every candidate of predict_8x8chroma_mode_available (:407-440), predictors predict.c:172-336, cost
 *     mbcmp[PIXEL_8x8](U) + mbcmp[PIXEL_8x8](V) + lambda * bs_size_ue(mode) — x264_mb_analyse_intra_chroma, analyse.c:541-609.
 * These are also the numbers the reference's optional table entries intra_mbcmp_x3_16x16 / intra_satd_x3_8x8c return
 * (S/common/pixel.h:97-100; asm-only there, checked against predict + mbcmp by tools/checkasm.c:380-407).
 * fdec holds the UNFILTERED reconstruction of the neighbours (row above, column left, corner of each listed macroblock: what
 * x264_macroblock_cache_load puts around p_fdec: the intra_border_backup row and the previous macroblock's last column, S/common/macroblock.c:839-848, :1015-1021); the caller lists macroblocks whose neighbours are
 * final — a wavefront diagonal, or every macroblock when the neighbourhood comes from another source.  The I4x4 / I8x8 stages need
 * reconstructed blocks of the same macroblock and stay on the host. */
#define X264_CUDA_INTRA_SATD    1 /* h->pixf.mbcmp == satd (subme > 1); else SAD */
#define X264_CUDA_INTRA_SLICE_B 2 /* adds the B-slice mb-type prefix lambda * i_mb_b_cost_table[I_16x16] to best16 (analyse.c:659-661) */
typedef struct x264_cuda_intra_job_t {
    int16_t mb_x, mb_y;
    uint8_t neighbour;     /* h->mb.i_neighbour: MB_LEFT 1 | MB_TOP 2 | MB_TOPRIGHT 4 | MB_TOPLEFT 8 (S/common/macroblock.h:28-34) */
    uint8_t flags;
    uint16_t lambda;       /* a->i_lambda = x264_lambda_tab[qp] (x264_cuda_host_lambda) */
} x264_cuda_intra_job_t;   /* 8 bytes */
typedef struct x264_cuda_intra_result_t {
    int32_t cost16[7];     /* a->i_satd_i16x16_dir[mode], mode = enum intra16x16_pred_e (V H DC P DC_LEFT DC_TOP DC_128); -1: not a candidate */
    int32_t cost_chroma[7];/* per enum intra_chroma_pred_e (DC H V P DC_LEFT DC_TOP DC_128); the reference stores these by list position in
                            * a->i_satd_i8x8chroma_dir[i]; -1: not a candidate */
    int32_t best16;        /* a->i_satd_i16x16 */
    int32_t best_chroma;   /* a->i_satd_i8x8chroma */
    uint8_t mode16;        /* a->i_predict16x16 */
    uint8_t mode_chroma;   /* a->i_predict8x8chroma */
    uint8_t reserved[2];
} x264_cuda_intra_result_t;/* 68 bytes */
X264_CUDA_API int x264_cuda_intra_mb_costs(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, const x264_cuda_frame_t *fdec,
                                           const x264_cuda_intra_job_t *jobs, int n_jobs, x264_cuda_intra_result_t *results);
X264_CUDA_API int x264_cuda_intra_mb_costs_dev(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, const x264_cuda_frame_t *fdec,
                                               const void *d_jobs, int n_jobs, void *d_results);

/* ------------------------------------------------------------------ bidirectional refinement --------- */
/* x264_me_refine_bidir_satd (S/encoder/me.c:843-927): joint quarter-pel refinement of the list-0 / list-1 vectors of one B partition
 * (16x16, 16x8, 8x16, 8x8) against the blended prediction, called by x264_mb_analyse_inter_b* refinement (analyse.c:2085-2105).
 * Both references need X264_CUDA_FRAME_HPEL.  cost: best blended cost + mv bits seen (the reference does not return it); -1 = input
 * vectors outside the cost table; the job's vectors are returned unchanged when the reference's early exit applies (:874-876). */
typedef struct x264_cuda_bidir_job_t {
    int16_t bx, by;                  /* luma position of the partition */
    uint8_t i_pixel;                 /* X264_CUDA_PIXEL_16x16 .. _8x8 */
    uint8_t qp;                      /* selects p_cost_mv */
    uint8_t weight;                  /* h->mb.bipred_weight[ref0][ref1]; 32 = plain average */
    uint8_t flags;                   /* X264_CUDA_ME_MBCMP_SATD */
    int16_t mv0[2], mv1[2];          /* m0->mv, m1->mv (quarter-pel) */
    int16_t mvp0[2], mvp1[2];        /* m0->mvp, m1->mvp */
    int16_t mv_min_spel[2], mv_max_spel[2];
} x264_cuda_bidir_job_t;             /* 32 bytes */
typedef struct x264_cuda_bidir_result_t { int16_t mv0[2], mv1[2]; int32_t cost; } x264_cuda_bidir_result_t; /* 12 bytes */
X264_CUDA_API int x264_cuda_me_refine_bidir(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, const x264_cuda_frame_t *fref0,
                                            const x264_cuda_frame_t *fref1, const x264_cuda_bidir_job_t *jobs, int n_jobs,
                                            x264_cuda_bidir_result_t *results);
X264_CUDA_API int x264_cuda_me_refine_bidir_dev(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, const x264_cuda_frame_t *fref0,
                                                const x264_cuda_frame_t *fref1, const void *d_jobs, int n_jobs, void *d_results);

/* ------------------------------------------------------------------ macroblock-batched motion search - */
/* One job == ALL inter partition searches of one macroblock against one reference: 16x16, 2x16x8, 2x8x16,
 * 4x8x8 (the x264_me_search_ref calls of x264_mb_analyse_inter_p16x16/p8x8/p16x8/p8x16, S/encoder/analyse.c:
 * 1077-1371), each with its own mvp/mvc, hence its own seed and its own +-me_range window.  The kernel walks the
 * union of the windows once, evaluates the four 8x8 SADs of the macroblock per candidate position and derives
 * all nine partition costs from them (PIXEL_SAD_C is a plain sum: S/common/pixel.c:40-56), so nine searches
 * cost about one.  Results are identical to nine x264_cuda_me_search jobs. */
#define X264_CUDA_ME_MB_PARTS 9 /* 0: 16x16 | 1,2: 16x8 top,bottom | 3,4: 8x16 left,right | 5..8: 8x8 TL,TR,BL,BR */
#define X264_CUDA_ME_MB_MVC 4
/* The 16x16 search gets up to nine extra predictors from x264_mb_predict_mv_ref16x16 (S/common/macroblock.c:376-449); the 16x8 / 8x16
 * searches get at most three and the 8x8 searches one to four, growing in block order (S/encoder/analyse.c:1229-1256, :1278, :1328).
 * Partition 0 may therefore carry up to 4 + 7: predictor 4 + k is stored in the fourth slot of partition 1 + k (k = 0..6; use the
 * accessor below), and a job that does so must keep i_mvc[1 + k] <= 3 for every slot it borrows.  The last 8x8 block keeps its four. */
#define X264_CUDA_ME_MB_MVC16_EXTRA 7
#define X264_CUDA_ME_MB_MVC16(job, k) ((k) < X264_CUDA_ME_MB_MVC ? (job)->mvc[0][k] : (job)->mvc[(k) - X264_CUDA_ME_MB_MVC + 1][X264_CUDA_ME_MB_MVC - 1])
typedef struct x264_cuda_me_mb_job_t {
    int16_t mb_x, mb_y;              /* macroblock coordinates */
    uint16_t part_mask;              /* bit p: search partition p */
    uint8_t qp, flags;               /* flags: X264_CUDA_ME_SEEDED */
    int16_t mv_min_fpel[2], mv_max_fpel[2];
    uint8_t i_mvc[X264_CUDA_ME_MB_PARTS]; /* <= 4; i_mvc[0] <= 11 (see X264_CUDA_ME_MB_MVC16) */
    uint8_t reserved[3];
    int16_t mvp[X264_CUDA_ME_MB_PARTS][2];
    int16_t mvc[X264_CUDA_ME_MB_PARTS][X264_CUDA_ME_MB_MVC][2];
    int16_t seed_mv[X264_CUDA_ME_MB_PARTS][2]; /* only with X264_CUDA_ME_SEEDED */
    int32_t seed_cost[X264_CUDA_ME_MB_PARTS];
} x264_cuda_me_mb_job_t;             /* 280 bytes */
typedef struct x264_cuda_me_mb_result_t {
    x264_cuda_me_result_t part[X264_CUDA_ME_MB_PARTS];
} x264_cuda_me_mb_result_t;          /* 144 bytes */

X264_CUDA_API int x264_cuda_me_search_mb(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, const x264_cuda_frame_t *fref,
                                         int me_range, const x264_cuda_me_mb_job_t *jobs, int n_jobs,
                                         x264_cuda_me_mb_result_t *results);
X264_CUDA_API int x264_cuda_me_search_mb_dev(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, const x264_cuda_frame_t *fref,
                                             int me_range, const void *d_jobs, int n_jobs, void *d_results);

#ifdef __cplusplus
}
#endif
#endif
