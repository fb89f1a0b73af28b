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
        for (int i = lane; i < 2 * cw * ch; i += 32) {
            const int p = i / (cw * ch), k = i - p * cw * ch, y = k / cw, x = k - y * cw;
            const uint8_t *s = (p ? pl.ref_cr : pl.ref_cb) + off + so + (ptrdiff_t)y * pl.stride_c + x;
            uint8_t *d = (p ? pl.dst_cr : pl.dst_cb) + off + (size_t)y * pl.stride_c + x;
            *d = (uint8_t)((cA * s[0] + cB * s[1] + cC * s[pl.stride_c] + cD * s[pl.stride_c + 1] + 32) >> 6);
        }
    }
}

} // namespace

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
