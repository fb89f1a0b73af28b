// metrics.cu — whole-frame analysis metrics with no inter-macroblock dependency: the integer parts run
// on the device, the float tails (log2 / SSIM ratio) stay on the host (host/x264_cuda_host.c) because the reference's
// results depend on float evaluation order.  All HBM-bound: every input byte is read once.
//   x264_cuda_frame_ssd            x264_pixel_ssd_wxh            S/common/pixel.c:98-136
//   x264_cuda_frame_mb_energy      ac_energy_mb                  S/encoder/ratecontrol.c:171-191
//   x264_cuda_frame_mb_hadamard_ac x264_pixel_hadamard_ac_16x16  S/common/pixel.c:306-358
//   x264_cuda_frame_ssim_sums      ssim_4x4x2_core               S/common/pixel.c:435-460
#include "pixel_dev.cuh"

namespace {

__device__ __forceinline__ uint32_t sq4(uint32_t a, uint32_t b, uint32_t acc)
{
    const uint32_t d = __vabsdiffu4(a, b);
    return __dp4a(d, d, acc); // four byte products, unsigned
}

// SSD over a width x height region; words of 4 pixels, tail bytes masked
__global__ void __launch_bounds__(256) ssd_kernel(const uint8_t *__restrict__ p1, int s1, const uint8_t *__restrict__ p2, int s2, int width, int height,
                                                  unsigned long long *__restrict__ out)
{
    const int wpr = (width + 3) >> 2;
    const long long n = (long long)wpr * height;
    unsigned long long total = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / wpr), x = (int)(i - (long long)y * wpr) * 4;
        uint32_t a = ldg4(p1 + (size_t)y * s1 + x), b = ldg4(p2 + (size_t)y * s2 + x); // any alignment (regions may start at odd x)
        if (x + 4 > width) { const uint32_t m = 0xffffffffu >> (8 * (x + 4 - width)); a &= m; b &= m; }
        total += sq4(a, b, 0);
    }
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
    if ((threadIdx.x & 31) == 0 && total) atomicAdd(out, total);
}

// one warp per macroblock: lanes 0-15 the luma rows, 16-23 the U rows, 24-31 the V rows
__global__ void __launch_bounds__(128) mb_energy_kernel(const uint8_t *__restrict__ py, int stride, const uint8_t *__restrict__ pu,
                                                        const uint8_t *__restrict__ pv, int stride_c, int W, int n_mb, uint32_t *__restrict__ out)
{
    const int lane = threadIdx.x & 31, mb = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (mb >= n_mb) return;
    const int mb_x = mb % W, mb_y = mb / W;
    uint32_t sum = 0, sqr = 0;
    if (lane < 16) {
        const uint4 v = __ldg((const uint4 *)(py + (size_t)(16 * mb_y + lane) * stride + 16 * mb_x));
        const uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
        for (int k = 0; k < 4; k++) { sum = __dp4a(w[k], 0x01010101u, sum); sqr = __dp4a(w[k], w[k], sqr); }
    } else {
        const uint8_t *p = (lane < 24 ? pu : pv) + (size_t)(8 * mb_y + (lane & 7)) * stride_c + 8 * mb_x;
        const uint2 v = __ldg((const uint2 *)p);
        sum = __dp4a(v.x, 0x01010101u, __dp4a(v.y, 0x01010101u, 0u));
        sqr = __dp4a(v.x, v.x, __dp4a(v.y, v.y, 0u));
    }
    // segmented sums: 16 luma lanes, 8 + 8 chroma lanes
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) { sum += __shfl_xor_sync(0xffffffffu, sum, o); sqr += __shfl_xor_sync(0xffffffffu, sqr, o); }
    const uint32_t s8 = __shfl_xor_sync(0xffffffffu, sum, 8), q8 = __shfl_xor_sync(0xffffffffu, sqr, 8);
    if (lane < 16) { sum += s8; sqr += q8; }
    const uint32_t var = lane < 16 ? sqr - (sum * sum >> 8) : sqr - (sum * sum >> 6); // PIXEL_VAR_C, pixel.c:142-161
    const uint32_t vy = __shfl_sync(0xffffffffu, var, 0), vu = __shfl_sync(0xffffffffu, var, 16), vv = __shfl_sync(0xffffffffu, var, 24);
    if (lane == 0) out[mb] = max(vy + vu + vv, 1u);
}

// four threads per macroblock (its 8x8 quadrants), HADAMARD_AC(16,16) pixel.c:343-355
__global__ void __launch_bounds__(128) mb_hadamard_ac_kernel(const uint8_t *__restrict__ py, int stride, int W, int n_mb, unsigned long long *__restrict__ out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x, mb = t >> 2, q = t & 3;
    unsigned long long v = 0;
    if (mb < n_mb) {
        const int mb_x = mb % W, mb_y = mb / W;
        v = hadamard_ac_8x8(py + (size_t)(16 * mb_y + 8 * (q >> 1)) * stride + 16 * mb_x + 8 * (q & 1), stride);
    }
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    if (mb < n_mb && q == 0) out[mb] = ((v >> 34) << 32) + ((uint32_t)v >> 1);
}

// one thread per 4x4 block
__global__ void __launch_bounds__(128) ssim_sums_kernel(const uint8_t *__restrict__ p1, int s1, const uint8_t *__restrict__ p2, int s2, int w4, int h4,
                                                        int4 *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w4 * h4) return;
    const int by = i / w4, bx = i - by * w4;
    uint32_t a1 = 0, a2 = 0, ss = 0, s12 = 0;
#pragma unroll
    for (int y = 0; y < 4; y++) {
        const uint32_t a = ldg4(p1 + (size_t)(4 * by + y) * s1 + 4 * bx), b = ldg4(p2 + (size_t)(4 * by + y) * s2 + 4 * bx);
        a1 = __dp4a(a, 0x01010101u, a1); a2 = __dp4a(b, 0x01010101u, a2);
        ss = __dp4a(a, a, ss); ss = __dp4a(b, b, ss); s12 = __dp4a(a, b, s12);
    }
    out[i] = make_int4((int)a1, (int)a2, (int)ss, (int)s12);
}

const uint8_t *plane00(const x264_cuda_frame_t *f, int plane, int *stride)
{
    if (plane == X264_CUDA_PLANE_CB || plane == X264_CUDA_PLANE_CR) { *stride = f->stride_c; return f->buf_chroma ? f->chroma[plane - X264_CUDA_PLANE_CB] : nullptr; }
    if (plane >= 0 && plane < 4) { *stride = f->g.stride; return f->plane[plane]; }
    return nullptr;
}

} // namespace

extern "C" int x264_cuda_frame_ssd(x264_cuda_t *ctx, const x264_cuda_frame_t *a, const x264_cuda_frame_t *b, int plane, int x0, int y0, int width,
                                   int height, int64_t *ssd)
{
    x264_cuda_enter(ctx);
    int s1 = 0, s2 = 0;
    const uint8_t *p1 = plane00(a, plane, &s1), *p2 = plane00(b, plane, &s2);
    if (p1 && p2) { p1 += (ptrdiff_t)y0 * s1 + x0; p2 += (ptrdiff_t)y0 * s2 + x0; }
    if (!p1 || !p2 || width < 1 || height < 1) { snprintf(ctx->err, 256, "x264_cuda_frame_ssd: frames lack plane %d or empty region", plane); return -1; }
    if (x264_cuda_stage(ctx, 256, 256)) return -1;
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_stage, 0, 8, ctx->stream));
    const long long n = (long long)((width + 3) >> 2) * height;
    const int blocks = (int)min((n + 255) / 256, (long long)ctx->sm_count * 8);
    ssd_kernel<<<blocks, 256, 0, ctx->stream>>>(p1, s1, p2, s2, width, height, (unsigned long long *)ctx->d_stage);
    LAUNCH_CHECK(ctx, "ssd_kernel");
    return x264_cuda_results_out(ctx, ssd, ctx->d_stage, ctx->h_stage, 8);
}

extern "C" int x264_cuda_frame_mb_energy(x264_cuda_t *ctx, const x264_cuda_frame_t *f, uint32_t *energy)
{
    x264_cuda_enter(ctx);
    if (!f->buf_chroma) { snprintf(ctx->err, 256, "x264_cuda_frame_mb_energy: frame needs X264_CUDA_FRAME_CHROMA"); return -1; }
    const int n = f->g.mb_width * f->g.mb_height;
    if (x264_cuda_stage(ctx, (size_t)n * 4, (size_t)n * 4)) return -1;
    mb_energy_kernel<<<(n + 3) / 4, 128, 0, ctx->stream>>>(f->plane[0], f->g.stride, f->chroma[0], f->chroma[1], f->stride_c, f->g.mb_width, n, (uint32_t *)ctx->d_stage);
    LAUNCH_CHECK(ctx, "mb_energy_kernel");
    return x264_cuda_results_out(ctx, energy, ctx->d_stage, ctx->h_stage, (size_t)n * 4);
}

extern "C" int x264_cuda_frame_mb_hadamard_ac(x264_cuda_t *ctx, const x264_cuda_frame_t *f, uint64_t *out)
{
    x264_cuda_enter(ctx);
    const int n = f->g.mb_width * f->g.mb_height;
    if (x264_cuda_stage(ctx, (size_t)n * 8, (size_t)n * 8)) return -1;
    mb_hadamard_ac_kernel<<<(4 * n + 127) / 128, 128, 0, ctx->stream>>>(f->plane[0], f->g.stride, f->g.mb_width, n, (unsigned long long *)ctx->d_stage);
    LAUNCH_CHECK(ctx, "mb_hadamard_ac_kernel");
    return x264_cuda_results_out(ctx, out, ctx->d_stage, ctx->h_stage, (size_t)n * 8);
}

extern "C" int x264_cuda_frame_ssim_sums(x264_cuda_t *ctx, const x264_cuda_frame_t *a, const x264_cuda_frame_t *b, int plane, int x0, int y0, int width,
                                         int height, int (*sums)[4])
{
    x264_cuda_enter(ctx);
    int s1 = 0, s2 = 0;
    const uint8_t *p1 = plane00(a, plane, &s1), *p2 = plane00(b, plane, &s2);
    if (p1 && p2) { p1 += (ptrdiff_t)y0 * s1 + x0; p2 += (ptrdiff_t)y0 * s2 + x0; }
    const int w4 = width >> 2, h4 = height >> 2;
    if (!p1 || !p2 || w4 < 1 || h4 < 1) { snprintf(ctx->err, 256, "x264_cuda_frame_ssim_sums: frames lack plane %d or region smaller than 4x4", plane); return -1; }
    const size_t bytes = (size_t)w4 * h4 * 16;
    if (x264_cuda_stage(ctx, bytes, bytes)) return -1;
    ssim_sums_kernel<<<(w4 * h4 + 127) / 128, 128, 0, ctx->stream>>>(p1, s1, p2, s2, w4, h4, (int4 *)ctx->d_stage);
    LAUNCH_CHECK(ctx, "ssim_sums_kernel");
    return x264_cuda_results_out(ctx, sums, ctx->d_stage, ctx->h_stage, bytes);
}
