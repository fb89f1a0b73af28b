// microbench.cu — measures the roofline denominator that MEASURED_PEAKS.json lacks: the sustained issue rate
// of VABSDIFF4.U8.ACC (the 4-byte SAD-accumulate every SAD kernel here is made of) on the integer ALU pipe.
#include "common.cuh"

namespace {
__global__ void __launch_bounds__(256) sad4_rate_kernel(uint32_t *out, int iters, uint32_t seed)
{
    uint32_t a[8], acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = seed * (threadIdx.x + 1) + i * 0x01030507u; acc[i] = i; }
    uint32_t b = seed ^ (blockIdx.x * 0x9e3779b9u);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++) acc[i] = sad4_acc(a[i], b, acc[i]);
            b += 0x01010101u; // keeps the compiler from hoisting; one extra ALU op per 8 SADs (accounted below)
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += acc[i];
    if (s == 0x12345678u) out[0] = s; // practically never; defeats dead-code elimination
}
} // namespace

/* thread-level VABSDIFF4.ACC operations per second, whole GPU, at the clocks the GPU runs under this load */
extern "C" int x264_cuda_measure_int_pipe(x264_cuda_t *ctx, double *sad4_per_sec)
{
    x264_cuda_enter(ctx);
    uint32_t *d = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&d, 256));
    cudaEvent_t e0, e1;
    CUDA_TRY(ctx, cudaEventCreate(&e0));
    CUDA_TRY(ctx, cudaEventCreate(&e1));
    const int blocks = ctx->sm_count * 8, iters = 4096;
    double best = 0;
    for (int rep = 0; rep < 4; rep++) {
        CUDA_TRY(ctx, cudaEventRecord(e0, ctx->stream));
        sad4_rate_kernel<<<blocks, 256, 0, ctx->stream>>>(d, iters, 0x1234567u + rep);
        ctx->launches++;
        CUDA_TRY(ctx, cudaEventRecord(e1, ctx->stream));
        CUDA_TRY(ctx, cudaEventSynchronize(e1));
        float ms = 0;
        CUDA_TRY(ctx, cudaEventElapsedTime(&ms, e0, e1));
        const double ops = (double)blocks * 256 * iters * 64;
        if (rep > 0 && ops / (ms * 1e-3) > best) best = ops / (ms * 1e-3);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    *sad4_per_sec = best;
    return 0;
}
