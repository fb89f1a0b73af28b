// mc.cu — frame-batched motion compensation: the luma qpel fetch of mc_luma/get_ref (S/common/mc.c:157-202) and the
// 1/8-pel bilinear chroma of mc_chroma (mc.c:205-236) for lists of blocks, writing the prediction straight into a
// device frame (the fdec the residual kernel then turns into the reconstruction) — x264_mb_mc_0xywh's data flow
// (S/common/macroblock.c:462-486) without the host round trip.  HBM/L2-bound gather-copy kernels.
#include "pixel_dev.cuh"

namespace {

struct McPlanes {
    const uint8_t *ref[4]; const uint8_t *ref_cb, *ref_cr;
    uint8_t *dst, *dst_cb, *dst_cr;
    int stride, stride_c;
};

// Four horizontally adjacent pixels of mc_chroma (mc.c:205-236) at once: s = top-left source byte of the first pixel (any alignment),
// coef = cA | cB << 8 | cC << 16 | cD << 24 (each <= 64).  Every output pixel is ONE dp4a over (s[x], s[x+1], t[x], t[x+1]) — the same sum
// of four products, + 32, >> 6 as the reference's expression.
__device__ __forceinline__ uint32_t chroma4(const uint8_t *s, int stride, uint32_t coef)
{
    uint32_t lo[2], hi[2]; // per source row: bytes 0..3 and bytes 1..4
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const uintptr_t a = (uintptr_t)(s + (ptrdiff_t)r * stride);
        const uint32_t *w = (const uint32_t *)(a & ~(uintptr_t)3);
        const int sh = (int)(a & 3) * 8;
        const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1);
        lo[r] = __funnelshift_r(w0, w1, sh);
        hi[r] = sh == 24 ? w1 : __funnelshift_r(w0, w1, sh + 8);
    }
    uint32_t out = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint32_t sel = (uint32_t)j | (uint32_t)(4 + j) << 4;
        const uint32_t quad = __byte_perm(__byte_perm(lo[0], hi[0], sel), __byte_perm(lo[1], hi[1], sel), 0x5410);
        out |= (__dp4a(quad, coef, 32u) >> 6) << (8 * j);
    }
    return out;
}

__global__ void __launch_bounds__(128) mc_blocks_kernel(McPlanes pl, const x264_cuda_mc_job_t *__restrict__ jobs, int n_jobs, int do_chroma)
{
    const int lane = threadIdx.x & 31;
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (j >= n_jobs) return;
    const x264_cuda_mc_job_t job = jobs[j];
    const int w = job.w, h = job.h;
    { // luma: rows of 4-pixel words; a 16x16 block is 64 words = 2 per lane
        const size_t off = (size_t)job.by * pl.stride + job.bx;
        const uint8_t *const planes[4] = { pl.ref[0] + off, pl.ref[1] + off, pl.ref[2] + off, pl.ref[3] + off };
        const QpelSrc src = qpel_src(planes, pl.stride, job.mvx, job.mvy);
        const int wpr = w >> 2;
        for (int i = lane; i < wpr * h; i += 32) {
            const int y = i / wpr, x = (i - y * wpr) * 4;
            *(uint32_t *)(pl.dst + off + (size_t)y * pl.stride + x) = qpel_row4(src, (ptrdiff_t)y * pl.stride + x);
        }
    }
    if (do_chroma) { // mc.c:205-236 with the luma mv (chroma units are 1/8 pel); block is w/2 x h/2
        const int d8x = job.mvx & 7, d8y = job.mvy & 7;
        const int cA = (8 - d8x) * (8 - d8y), cB = d8x * (8 - d8y), cC = (8 - d8x) * d8y, cD = d8x * d8y;
        const int cw = w >> 1, ch = h >> 1;
        const size_t off = (size_t)(job.by >> 1) * pl.stride_c + (job.bx >> 1);
        const ptrdiff_t so = (ptrdiff_t)(job.mvy >> 3) * pl.stride_c + (job.mvx >> 3);
        if (cw >= 4 && !((job.bx >> 1) & 3)) { // word-aligned destination rows: four pixels per lane and step (chroma4), 32-bit stores
            const uint32_t coef = (uint32_t)cA | (uint32_t)cB << 8 | (uint32_t)cC << 16 | (uint32_t)cD << 24;
            const int two = cw >> 3, per = (cw >> 2) * ch; // 4-pixel units per row - 1 (cw is 4 or 8), units per plane
            for (int i = lane; i < 2 * per; i += 32) {
                const int p = i >= per, k = i - p * per, y = k >> two, x = (k & two) * 4;
                const ptrdiff_t o = (ptrdiff_t)y * pl.stride_c + x;
                *(uint32_t *)((p ? pl.dst_cr : pl.dst_cb) + off + o) = chroma4((p ? pl.ref_cr : pl.ref_cb) + off + so + o, pl.stride_c, coef);
            }
        } else
        for (int i = lane; i < 2 * cw * ch; i += 32) {
            const int p = i / (cw * ch), k = i - p * cw * ch, y = k / cw, x = k - y * cw;
            const uint8_t *s = (p ? pl.ref_cr : pl.ref_cb) + off + so + (ptrdiff_t)y * pl.stride_c + x;
            uint8_t *d = (p ? pl.dst_cr : pl.dst_cb) + off + (size_t)y * pl.stride_c + x;
            *d = (uint8_t)((cA * s[0] + cB * s[1] + cC * s[pl.stride_c] + cD * s[pl.stride_c + 1] + 32) >> 6);
        }
    }
}

// x264_mb_mc_01xywh (S/common/macroblock.c:508-546): both lists' predictions blended by h->mc.avg — the rounded average when
// weight == 32, else implicit weighted bi-prediction clip((a*w + b*(64-w) + 32) >> 6) (mc.c:52-125)
__device__ __forceinline__ uint32_t blend4(uint32_t a, uint32_t b, int weight)
{
    if (weight == 32) return __vavgu4(a, b);
    uint32_t r = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int p = (a >> (8 * k)) & 255, q = (b >> (8 * k)) & 255;
        r |= (uint32_t)clip_u8((p * weight + q * (64 - weight) + 32) >> 6) << (8 * k);
    }
    return r;
}

struct McBiPlanes {
    const uint8_t *ref[2][4]; const uint8_t *ref_cb[2], *ref_cr[2];
    uint8_t *dst, *dst_cb, *dst_cr;
    int stride, stride_c;
};

__global__ void __launch_bounds__(128) mc_blocks_bi_kernel(McBiPlanes pl, const x264_cuda_mc_bi_job_t *__restrict__ jobs, int n_jobs, int do_chroma)
{
    const int lane = threadIdx.x & 31;
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (j >= n_jobs) return;
    const x264_cuda_mc_bi_job_t job = jobs[j];
    const int w = job.w, h = job.h, weight = job.weight;
    {
        const size_t off = (size_t)job.by * pl.stride + job.bx;
        const uint8_t *const p0[4] = { pl.ref[0][0] + off, pl.ref[0][1] + off, pl.ref[0][2] + off, pl.ref[0][3] + off };
        const uint8_t *const p1[4] = { pl.ref[1][0] + off, pl.ref[1][1] + off, pl.ref[1][2] + off, pl.ref[1][3] + off };
        const QpelSrc s0 = qpel_src(p0, pl.stride, job.mv0[0], job.mv0[1]), s1 = qpel_src(p1, pl.stride, job.mv1[0], job.mv1[1]);
        const int wpr = w >> 2;
        for (int i = lane; i < wpr * h; i += 32) {
            const int y = i / wpr, x = (i - y * wpr) * 4;
            const ptrdiff_t o = (ptrdiff_t)y * pl.stride + x;
            *(uint32_t *)(pl.dst + off + o) = blend4(qpel_row4(s0, o), qpel_row4(s1, o), weight);
        }
    }
    if (do_chroma) {
        const int cw = w >> 1, ch = h >> 1;
        const size_t off = (size_t)(job.by >> 1) * pl.stride_c + (job.bx >> 1);
        int cA[2], cB[2], cC[2], cD[2];
        ptrdiff_t so[2];
#pragma unroll
        for (int l = 0; l < 2; l++) {
            const int mvx = l ? job.mv1[0] : job.mv0[0], mvy = l ? job.mv1[1] : job.mv0[1];
            const int d8x = mvx & 7, d8y = mvy & 7;
            cA[l] = (8 - d8x) * (8 - d8y); cB[l] = d8x * (8 - d8y); cC[l] = (8 - d8x) * d8y; cD[l] = d8x * d8y;
            so[l] = (ptrdiff_t)(mvy >> 3) * pl.stride_c + (mvx >> 3);
        }
        if (cw >= 4 && !((job.bx >> 1) & 3)) { // four pixels per lane and step: both lists' chroma4, blended per byte like the luma words
            uint32_t coef[2];
#pragma unroll
            for (int l = 0; l < 2; l++) coef[l] = (uint32_t)cA[l] | (uint32_t)cB[l] << 8 | (uint32_t)cC[l] << 16 | (uint32_t)cD[l] << 24;
            const int two = cw >> 3, per = (cw >> 2) * ch;
            for (int i = lane; i < 2 * per; i += 32) {
                const int p = i >= per, k = i - p * per, y = k >> two, x = (k & two) * 4;
                const ptrdiff_t o = (ptrdiff_t)y * pl.stride_c + x;
                const uint32_t a = chroma4((p ? pl.ref_cr[0] : pl.ref_cb[0]) + off + so[0] + o, pl.stride_c, coef[0]);
                const uint32_t b = chroma4((p ? pl.ref_cr[1] : pl.ref_cb[1]) + off + so[1] + o, pl.stride_c, coef[1]);
                *(uint32_t *)((p ? pl.dst_cr : pl.dst_cb) + off + o) = blend4(a, b, weight);
            }
        } else
        for (int i = lane; i < 2 * cw * ch; i += 32) {
            const int p = i / (cw * ch), k = i - p * cw * ch, y = k / cw, x = k - y * cw;
            int v[2];
#pragma unroll
            for (int l = 0; l < 2; l++) {
                const uint8_t *s = (p ? pl.ref_cr[l] : pl.ref_cb[l]) + off + so[l] + (ptrdiff_t)y * pl.stride_c + x;
                v[l] = (cA[l] * s[0] + cB[l] * s[1] + cC[l] * s[pl.stride_c] + cD[l] * s[pl.stride_c + 1] + 32) >> 6;
            }
            uint8_t *d = (p ? pl.dst_cr : pl.dst_cb) + off + (size_t)y * pl.stride_c + x;
            *d = (uint8_t)(weight == 32 ? (v[0] + v[1] + 1) >> 1 : clip_u8((v[0] * weight + v[1] * (64 - weight) + 32) >> 6));
        }
    }
}

} // namespace

extern "C" int x264_cuda_mc_blocks_bi_dev(x264_cuda_t *ctx, const x264_cuda_frame_t *fref0, const x264_cuda_frame_t *fref1, x264_cuda_frame_t *fdec,
                                          const void *d_jobs, int n_jobs)
{
    x264_cuda_enter(ctx);
    if (n_jobs <= 0) return 0;
    if (!(fref0->g.flags & fref1->g.flags & X264_CUDA_FRAME_HPEL) || fref0->g.stride != fdec->g.stride || fref1->g.stride != fdec->g.stride) {
        snprintf(ctx->err, 256, "x264_cuda_mc_blocks_bi: both references need the half-pel planes and the same geometry as fdec");
        return -1;
    }
    const int do_chroma = fref0->buf_chroma && fref1->buf_chroma && fdec->buf_chroma;
    McBiPlanes pl;
    for (int k = 0; k < 4; k++) { pl.ref[0][k] = fref0->plane[k]; pl.ref[1][k] = fref1->plane[k]; }
    pl.ref_cb[0] = fref0->chroma[0]; pl.ref_cr[0] = fref0->chroma[1]; pl.ref_cb[1] = fref1->chroma[0]; pl.ref_cr[1] = fref1->chroma[1];
    pl.dst = fdec->plane[0]; pl.dst_cb = fdec->chroma[0]; pl.dst_cr = fdec->chroma[1]; pl.stride = fdec->g.stride; pl.stride_c = fdec->stride_c;
    mc_blocks_bi_kernel<<<(n_jobs + 3) / 4, 128, 0, ctx->stream>>>(pl, (const x264_cuda_mc_bi_job_t *)d_jobs, n_jobs, do_chroma);
    LAUNCH_CHECK(ctx, "mc_blocks_bi_kernel");
    return 0;
}

extern "C" int x264_cuda_mc_blocks_bi(x264_cuda_t *ctx, const x264_cuda_frame_t *fref0, const x264_cuda_frame_t *fref1, x264_cuda_frame_t *fdec,
                                      const x264_cuda_mc_bi_job_t *jobs, int n_jobs)
{
    x264_cuda_enter(ctx);
    if (n_jobs <= 0) return 0;
    const size_t jb = (size_t)n_jobs * sizeof(x264_cuda_mc_bi_job_t);
    if (x264_cuda_stage(ctx, jb, jb)) return -1;
    if (x264_cuda_jobs_in(ctx, ctx->d_stage, jobs, ctx->h_stage, jb)) return -1;
    if (x264_cuda_mc_blocks_bi_dev(ctx, fref0, fref1, fdec, ctx->d_stage, n_jobs)) return -1;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int x264_cuda_mc_blocks_dev(x264_cuda_t *ctx, const x264_cuda_frame_t *fref, x264_cuda_frame_t *fdec, const void *d_jobs, int n_jobs)
{
    x264_cuda_enter(ctx);
    if (n_jobs <= 0) return 0;
    if (!(fref->g.flags & X264_CUDA_FRAME_HPEL) || fref->g.stride != fdec->g.stride) {
        snprintf(ctx->err, 256, "x264_cuda_mc_blocks: fref needs the half-pel planes and the same geometry as fdec");
        return -1;
    }
    const int do_chroma = fref->buf_chroma && fdec->buf_chroma;
    McPlanes pl = { { fref->plane[0], fref->plane[1], fref->plane[2], fref->plane[3] }, fref->chroma[0], fref->chroma[1],
                    fdec->plane[0], fdec->chroma[0], fdec->chroma[1], fref->g.stride, fref->stride_c };
    mc_blocks_kernel<<<(n_jobs + 3) / 4, 128, 0, ctx->stream>>>(pl, (const x264_cuda_mc_job_t *)d_jobs, n_jobs, do_chroma);
    LAUNCH_CHECK(ctx, "mc_blocks_kernel");
    return 0;
}

extern "C" int x264_cuda_mc_blocks(x264_cuda_t *ctx, const x264_cuda_frame_t *fref, x264_cuda_frame_t *fdec, const x264_cuda_mc_job_t *jobs,
                                   int n_jobs)
{
    x264_cuda_enter(ctx);
    if (n_jobs <= 0) return 0;
    const size_t jb = (size_t)n_jobs * sizeof(x264_cuda_mc_job_t);
    if (x264_cuda_stage(ctx, jb, jb)) return -1;
    if (x264_cuda_jobs_in(ctx, ctx->d_stage, jobs, ctx->h_stage, jb)) return -1;
    if (x264_cuda_mc_blocks_dev(ctx, fref, fdec, ctx->d_stage, n_jobs)) return -1;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
