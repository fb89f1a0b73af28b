/* x264_cuda_tables.h — the reference's plugin surface, filled by the CUDA back-end.
 *
 * x264 (core 66) reaches its DSP code only through four structs of C function pointers embedded by value in x264_t
 * (S/common/common.h:625-629).  The structs below are layout-identical mirrors of
 *     x264_pixel_function_t  S/common/pixel.h:63-103        x264_dct_function_t   S/common/dct.h:89-114
 *     x264_mc_functions_t    S/common/mc.h:31-77            x264_quant_function_t S/common/quant.h:26-44
 * so a maintainer can pass &h->pixf etc. straight in (INTEGRATION.md shows the four call sites).  Like
 * x264_pixel_altivec_init (S/common/pixel.c:781-786) the *_init_cuda functions OVERRIDE entries of a table that
 * x264_*_init(cpu=0) has already filled with the C bodies; members this back-end does not implement (ssim (unused), ssim_end4 (float),
 * intra_satd_x3_4x4, intra_sa8d_x3_8x8, copy, plane_copy, prefetch, memcpy, integral_init*, denoise, decimate, coeff_last/level_run,
 * zigzag) keep their C bodies.
 *
 * Every overridden entry runs on the GPU: operands are staged to the device, one kernel computes the result with the
 * same device functions the frame-batched entry points use, and the result is copied back.  They are correct drop-ins
 * for any caller (tools/checkasm.c included) but cost one PCIe round trip per call; the performance path is the
 * frame-batched API of x264_cuda.h.  There is no CPU fallback: if the device cannot be opened the init functions
 * return -1 and leave the table untouched, and x264_encoder_open should fail (S/encoder/encoder.c:634-645).
 */
#ifndef X264_CUDA_TABLES_H
#define X264_CUDA_TABLES_H
#include "x264_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

/* the cpu-flag bit a maintainer would add next to X264_CPU_* (S/x264.h:49-65: next free bit) */
#define X264_CPU_CUDA 0x020000

typedef int (*x264_cuda_pixel_cmp_t)(uint8_t *, int, uint8_t *, int);
typedef void (*x264_cuda_pixel_cmp_x3_t)(uint8_t *, uint8_t *, uint8_t *, uint8_t *, int, int[3]);
typedef void (*x264_cuda_pixel_cmp_x4_t)(uint8_t *, uint8_t *, uint8_t *, uint8_t *, uint8_t *, int, int[4]);

typedef struct x264_cuda_pixel_function_t { /* == x264_pixel_function_t */
    x264_cuda_pixel_cmp_t sad[7];
    x264_cuda_pixel_cmp_t ssd[7];
    x264_cuda_pixel_cmp_t satd[7];
    x264_cuda_pixel_cmp_t ssim[7];
    x264_cuda_pixel_cmp_t sa8d[4];
    x264_cuda_pixel_cmp_t mbcmp[7];
    x264_cuda_pixel_cmp_t mbcmp_unaligned[7];
    x264_cuda_pixel_cmp_t fpelcmp[7];
    x264_cuda_pixel_cmp_x3_t fpelcmp_x3[7];
    x264_cuda_pixel_cmp_x4_t fpelcmp_x4[7];
    x264_cuda_pixel_cmp_t sad_aligned[7];
    int (*var[4])(uint8_t *pix, int stride);
    uint64_t (*hadamard_ac[4])(uint8_t *pix, int stride);
    void (*ssim_4x4x2_core)(const uint8_t *pix1, int stride1, const uint8_t *pix2, int stride2, int sums[2][4]);
    float (*ssim_end4)(int sum0[5][4], int sum1[5][4], int width);
    x264_cuda_pixel_cmp_x3_t sad_x3[7];
    x264_cuda_pixel_cmp_x4_t sad_x4[7];
    x264_cuda_pixel_cmp_x3_t satd_x3[7];
    x264_cuda_pixel_cmp_x4_t satd_x4[7];
    int (*ads[7])(int enc_dc[4], uint16_t *sums, int delta, uint16_t *cost_mvx, int16_t *mvs, int width, int thresh);
    void (*intra_mbcmp_x3_16x16)(uint8_t *fenc, uint8_t *fdec, int res[3]);
    void (*intra_satd_x3_16x16)(uint8_t *fenc, uint8_t *fdec, int res[3]);
    void (*intra_sad_x3_16x16)(uint8_t *fenc, uint8_t *fdec, int res[3]);
    void (*intra_satd_x3_8x8c)(uint8_t *fenc, uint8_t *fdec, int res[3]);
    void (*intra_satd_x3_4x4)(uint8_t *fenc, uint8_t *fdec, int res[3]);
    void (*intra_sa8d_x3_8x8)(uint8_t *fenc, uint8_t edge[33], int res[3]);
} x264_cuda_pixel_function_t;

typedef struct x264_cuda_dct_function_t { /* == x264_dct_function_t; pix1 stride 16, pix2 / p_dst stride 32 */
    void (*sub4x4_dct)(int16_t dct[4][4], uint8_t *pix1, uint8_t *pix2);
    void (*add4x4_idct)(uint8_t *p_dst, int16_t dct[4][4]);
    void (*sub8x8_dct)(int16_t dct[4][4][4], uint8_t *pix1, uint8_t *pix2);
    void (*add8x8_idct)(uint8_t *p_dst, int16_t dct[4][4][4]);
    void (*add8x8_idct_dc)(uint8_t *p_dst, int16_t dct[2][2]);
    void (*sub16x16_dct)(int16_t dct[16][4][4], uint8_t *pix1, uint8_t *pix2);
    void (*add16x16_idct)(uint8_t *p_dst, int16_t dct[16][4][4]);
    void (*add16x16_idct_dc)(uint8_t *p_dst, int16_t dct[4][4]);
    void (*sub8x8_dct8)(int16_t dct[8][8], uint8_t *pix1, uint8_t *pix2);
    void (*add8x8_idct8)(uint8_t *p_dst, int16_t dct[8][8]);
    void (*sub16x16_dct8)(int16_t dct[4][8][8], uint8_t *pix1, uint8_t *pix2);
    void (*add16x16_idct8)(uint8_t *p_dst, int16_t dct[4][8][8]);
    void (*dct4x4dc)(int16_t d[4][4]);
    void (*idct4x4dc)(int16_t d[4][4]);
} x264_cuda_dct_function_t;

typedef struct x264_cuda_run_level_t { int last; int16_t level[16]; uint8_t run[16]; } x264_cuda_run_level_t; /* S/common/bs.h */

typedef struct x264_cuda_quant_function_t { /* == x264_quant_function_t */
    int (*quant_8x8)(int16_t dct[8][8], uint16_t mf[64], uint16_t bias[64]);
    int (*quant_4x4)(int16_t dct[4][4], uint16_t mf[16], uint16_t bias[16]);
    int (*quant_4x4_dc)(int16_t dct[4][4], int mf, int bias);
    int (*quant_2x2_dc)(int16_t dct[2][2], int mf, int bias);
    void (*dequant_8x8)(int16_t dct[8][8], int dequant_mf[6][8][8], int i_qp);
    void (*dequant_4x4)(int16_t dct[4][4], int dequant_mf[6][4][4], int i_qp);
    void (*dequant_4x4_dc)(int16_t dct[4][4], int dequant_mf[6][4][4], int i_qp);
    void (*denoise_dct)(int16_t *dct, uint32_t *sum, uint16_t *offset, int size);
    int (*decimate_score15)(int16_t *dct);
    int (*decimate_score16)(int16_t *dct);
    int (*decimate_score64)(int16_t *dct);
    int (*coeff_last[6])(int16_t *dct);
    int (*coeff_level_run[5])(int16_t *dct, x264_cuda_run_level_t *runlevel);
} x264_cuda_quant_function_t;

typedef struct x264_cuda_mc_functions_t { /* == x264_mc_functions_t */
    void (*mc_luma)(uint8_t *dst, int i_dst, uint8_t **src, int i_src, int mvx, int mvy, int i_width, int i_height);
    uint8_t *(*get_ref)(uint8_t *dst, int *i_dst, uint8_t **src, int i_src, int mvx, int mvy, int i_width, int i_height);
    void (*mc_chroma)(uint8_t *dst, int i_dst, uint8_t *src, int i_src, int mvx, int mvy, int i_width, int i_height);
    void (*avg[10])(uint8_t *dst, int, uint8_t *src1, int, uint8_t *src2, int, int i_weight);
    void (*copy[7])(uint8_t *dst, int, uint8_t *src, int, int i_height);
    void (*copy_16x16_unaligned)(uint8_t *dst, int, uint8_t *src, int, int i_height);
    void (*plane_copy)(uint8_t *dst, int i_dst, uint8_t *src, int i_src, int w, int h);
    void (*hpel_filter)(uint8_t *dsth, uint8_t *dstv, uint8_t *dstc, uint8_t *src, int i_stride, int i_width, int i_height, int16_t *buf);
    void (*prefetch_fenc)(uint8_t *pix_y, int stride_y, uint8_t *pix_uv, int stride_uv, int mb_x);
    void (*prefetch_ref)(uint8_t *pix, int stride, int parity);
    void *(*memcpy_aligned)(void *dst, const void *src, size_t n);
    void (*memzero_aligned)(void *dst, int n);
    void (*integral_init4h)(uint16_t *sum, uint8_t *pix, int stride);
    void (*integral_init8h)(uint16_t *sum, uint8_t *pix, int stride);
    void (*integral_init4v)(uint16_t *sum8, uint16_t *sum4, int stride);
    void (*integral_init8v)(uint16_t *sum8, int stride);
    void (*frame_init_lowres_core)(uint8_t *src0, uint8_t *dst0, uint8_t *dsth, uint8_t *dstv, uint8_t *dstc, int src_stride,
                                   int dst_stride, int width, int height);
} x264_cuda_mc_functions_t;

/* Each returns 0, or -1 when no CUDA device is usable (table left as it was).  The per-call entries share one
 * process-wide device context (device = $X264_CUDA_DEVICE or 0), serialised by a mutex: re-entrant as the reference
 * requires (S/common/common.h:50), though not concurrent. */
X264_CUDA_API int x264_pixel_init_cuda(x264_cuda_pixel_function_t *pixf); /* sad, sad_aligned, ssd, satd, sa8d, sad_x3/x4, satd_x3/x4, ads, var, hadamard_ac,
                                                                           * ssim_4x4x2_core, intra_{mbcmp,satd,sad}_x3_16x16, intra_satd_x3_8x8c */
X264_CUDA_API int x264_dct_init_cuda(x264_cuda_dct_function_t *dctf);     /* all 14 entries */
X264_CUDA_API int x264_quant_init_cuda(x264_cuda_quant_function_t *pf);   /* quant_*, dequant_* */
X264_CUDA_API int x264_mc_init_cuda(x264_cuda_mc_functions_t *pf);        /* mc_luma, get_ref, mc_chroma, avg[10], hpel_filter, frame_init_lowres_core */
X264_CUDA_API long long x264_cuda_tables_launches(void);                  /* kernels launched by table entries so far (diagnostic) */
/* The table signatures cannot report failure (S/common/pixel.h:26-28).  A device error inside an entry is recorded once (sticky), every
 * later entry returns zeros without touching the device, and this returns the message (NULL while healthy): the encoder-side hook
 * checks it once per frame and makes x264_encoder_encode() return -1 after x264_log(h, X264_LOG_ERROR, ...) — the reference's own
 * convention (S/x264.c:759-762).  Entries may be called concurrently from any number of host threads (one context per thread). */
X264_CUDA_API const char *x264_cuda_tables_error(void);
X264_CUDA_API void x264_cuda_tables_shutdown(void);                       /* closes every thread's table context (no entry may be running) */

#ifdef __cplusplus
}
#endif
#endif
