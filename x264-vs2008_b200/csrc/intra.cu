// intra.cu — the intra-analysis stages that depend only on the source macroblock and on reconstructed pixels of NEIGHBOURING
// macroblocks: the Intra16x16 candidates of x264_mb_analyse_intra (S/encoder/analyse.c:612-664) and the chroma 8x8 candidates of
// x264_mb_analyse_intra_chroma (:541-609), predictors of S/common/predict.c:40-170 and :172-336.  (I4x4 / I8x8 need the
// reconstruction of earlier blocks of the same macroblock: host)
//
// ONE WARP PER MACROBLOCK.  Luma: lane = candidate (0..3) x 8x4 unit (0..7), i.e. all four candidates of a macroblock with a
// complete neighbourhood are costed in one pass; chroma: lane = candidate x (plane, upper/lower half).  The neighbour pixels
// (33 + 2 x 17 bytes) are staged in shared memory once per warp.  Latency/HBM class: 384 B of source + ~100 B of edges in, 68 B out.
#include "intra_dev.cuh"

namespace {

struct IntraPlanes { const uint8_t *fe_y, *fe_u, *fe_v, *fd_y, *fd_u, *fd_v; int stride, stride_c; };

// candidate i of the list for a neighbour mask (predict_16x16_mode_available / predict_8x8chroma_mode_available, analyse.c:372-440):
// returns the mode number in the respective enum and the predictor kind (0 V, 1 H, 2 DC-like, 3 plane); n = list length
__device__ __forceinline__ int candidate(int neighbour, bool chroma, int i, int &n, int &kind)
{
    const int mV = chroma ? 2 : 0, mDC = chroma ? 0 : 2;
    if (neighbour & 8) { n = 4; kind = i == 0 ? 0 : i == 1 ? 1 : i == 2 ? 2 : 3; return i == 0 ? mV : i == 1 ? 1 : i == 2 ? mDC : 3; }
    if (neighbour & 1) { n = 2; kind = i == 0 ? 2 : 1; return i == 0 ? 4 : 1; }
    if (neighbour & 2) { n = 2; kind = i == 0 ? 2 : 0; return i == 0 ? 5 : mV; }
    n = 1; kind = 2;
    return 6;
}
// bs_size_ue(x264_mb_pred_mode16x16_fix[mode]) / (..8x8c_fix[mode]): the DC variants are coded as DC
__device__ __forceinline__ int mode_bits(int mode, bool chroma)
{
    if (mode > 3) mode = chroma ? 0 : 2;
    return mode == 0 ? 1 : mode == 3 ? 5 : 3;
}

__global__ void __launch_bounds__(128) intra_mb_costs_kernel(IntraPlanes pl, const x264_cuda_intra_job_t *__restrict__ jobs, int n_jobs,
                                                             x264_cuda_intra_result_t *__restrict__ results)
{
    __shared__ __align__(4) IntraEdges s_edges[4];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int jb = blockIdx.x * 4 + wid;
    if (jb >= n_jobs) return;
    const x264_cuda_intra_job_t job = jobs[jb];
    IntraEdges &E = s_edges[wid];
    const unsigned FULL = 0xffffffffu;
    const bool satd = job.flags & X264_CUDA_INTRA_SATD;
    const int lambda = job.lambda;

    // ---- stage the neighbour pixels (the frames are padded, so the reads are in bounds even where the mask says "not available")
    {
        const uint8_t *py = pl.fd_y + (size_t)job.mb_y * 16 * pl.stride + job.mb_x * 16;
        if (lane < 16) { E.y[4 + lane] = py[-(ptrdiff_t)pl.stride + lane]; E.y[20 + lane] = py[(ptrdiff_t)lane * pl.stride - 1]; }
        if (lane == 16) E.y[3] = py[-(ptrdiff_t)pl.stride - 1];
        const size_t co = (size_t)job.mb_y * 8 * pl.stride_c + job.mb_x * 8;
        if (lane >= 16 && lane < 32) {
            const int k = lane & 7;
            const uint8_t *pc = ((lane & 8) ? pl.fd_v : pl.fd_u) + co;
            uint8_t *e = (lane & 8) ? E.v : E.u;
            e[4 + k] = pc[-(ptrdiff_t)pl.stride_c + k]; e[12 + k] = pc[(ptrdiff_t)k * pl.stride_c - 1];
            if (k == 0) e[3] = pc[-(ptrdiff_t)pl.stride_c - 1];
        }
    }
    __syncwarp();

    x264_cuda_intra_result_t res;
#pragma unroll
    for (int i = 0; i < 7; i++) res.cost16[i] = res.cost_chroma[i] = -1;
    res.best16 = res.best_chroma = 1 << 28; // COST_MAX
    res.mode16 = res.mode_chroma = 0;
    res.reserved[0] = res.reserved[1] = 0;

    // ---- Intra16x16: lane = candidate (lane >> 3) x unit (lane & 7)
    {
        const int ci = lane >> 3, u = lane & 7, x0 = (u & 1) * 8, y0 = (u >> 1) * 4;
        int n, kind;
        const int mode = candidate(job.neighbour, false, ci, n, kind);
        int cost = 0;
        if (ci < n) {
            int dcq[4] = { 128, 0, 0, 0 };
            if (kind == 2 && mode != 6) {
                int st = 0, sl = 0;
#pragma unroll
                for (int i = 0; i < 16; i++) { st += E.y[4 + i]; sl += E.y[20 + i]; }
                dcq[0] = mode == 2 ? (st + sl + 16) >> 5 : mode == 4 ? (sl + 8) >> 4 : (st + 8) >> 4; // predict.c:52-96
            }
            uint2 f[4], r[4];
            const uint8_t *fe = pl.fe_y + ((size_t)job.mb_y * 16 + y0) * pl.stride + job.mb_x * 16 + x0;
#pragma unroll
            for (int k = 0; k < 4; k++) f[k] = ldg8(fe + (size_t)k * pl.stride);
            predict_unit<16>(E.y, kind, dcq, x0, y0, r);
            cost = unit_metric(satd, f, r);
        }
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) cost += __shfl_xor_sync(FULL, cost, o);
        cost += lambda * mode_bits(mode, false);
#pragma unroll
        for (int i = 0; i < 4; i++) { // list order, strict '<' keeps the first minimum (COPY2_IF_LT)
            const int c = __shfl_sync(FULL, cost, i * 8), m = __shfl_sync(FULL, mode, i * 8);
            if (i < n) {
#pragma unroll
                for (int k = 0; k < 7; k++) if (k == m) res.cost16[k] = c;
                if (c < res.best16) { res.best16 = c; res.mode16 = (uint8_t)m; }
            }
        }
        if (job.flags & X264_CUDA_INTRA_SLICE_B) res.best16 += lambda * 9; // i_mb_b_cost_table[I_16x16], analyse.c:659-661
    }

    // ---- chroma: lane = candidate (lane >> 2) x (plane, half); lanes 16..31 idle
    {
        const int ci = (lane >> 2) & 3, pv = (lane >> 1) & 1, y0 = (lane & 1) * 4;
        int n, kind;
        const int mode = candidate(job.neighbour, true, ci, n, kind);
        int cost = 0;
        if (lane < 16 && ci < n) {
            const uint8_t *e = pv ? E.v : E.u;
            int dcq[4] = { 128, 128, 128, 128 };
            if (kind == 2 && mode != 6) {
                int s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
                for (int i = 0; i < 4; i++) { s0 += e[4 + i]; s1 += e[8 + i]; s2 += e[12 + i]; s3 += e[16 + i]; }
                if (mode == 0) { dcq[0] = (s0 + s2 + 4) >> 3; dcq[1] = (s1 + 2) >> 2; dcq[2] = (s3 + 2) >> 2; dcq[3] = (s1 + s3 + 4) >> 3; } // predict.c:234-277
                else if (mode == 4) { dcq[0] = dcq[1] = (s2 + 2) >> 2; dcq[2] = dcq[3] = (s3 + 2) >> 2; }                                // :184-212
                else { dcq[0] = dcq[2] = (s0 + 2) >> 2; dcq[1] = dcq[3] = (s1 + 2) >> 2; }                                               // :213-233
            }
            uint2 f[4], r[4];
            const uint8_t *fe = (pv ? pl.fe_v : pl.fe_u) + ((size_t)job.mb_y * 8 + y0) * pl.stride_c + job.mb_x * 8;
#pragma unroll
            for (int k = 0; k < 4; k++) f[k] = ldg8(fe + (size_t)k * pl.stride_c);
            predict_unit<8>(e, kind, dcq, 0, y0, r);
            cost = unit_metric(satd, f, r);
        }
        cost += __shfl_xor_sync(FULL, cost, 1);
        cost += __shfl_xor_sync(FULL, cost, 2);
        cost += lambda * mode_bits(mode, true);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int c = __shfl_sync(FULL, cost, i * 4), m = __shfl_sync(FULL, mode, i * 4);
            if (i < n) {
#pragma unroll
                for (int k = 0; k < 7; k++) if (k == m) res.cost_chroma[k] = c;
                if (c < res.best_chroma) { res.best_chroma = c; res.mode_chroma = (uint8_t)m; }
            }
        }
    }
    if (lane == 0) results[jb] = res;
}
} // namespace

extern "C" int x264_cuda_intra_mb_costs_dev(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, const x264_cuda_frame_t *fdec, const void *d_jobs,
                                            int n_jobs, void *d_results)
{
    x264_cuda_enter(ctx);
    if (n_jobs <= 0) return 0;
    if (!fenc->buf_chroma || !fdec->buf_chroma || fenc->g.stride != fdec->g.stride) {
        snprintf(ctx->err, 256, "x264_cuda_intra_mb_costs: frames need X264_CUDA_FRAME_CHROMA and equal geometry");
        return -1;
    }
    IntraPlanes pl = { fenc->plane[0], fenc->chroma[0], fenc->chroma[1], fdec->plane[0], fdec->chroma[0], fdec->chroma[1], fenc->g.stride,
                       fenc->stride_c };
    intra_mb_costs_kernel<<<(n_jobs + 3) / 4, 128, 0, ctx->stream>>>(pl, (const x264_cuda_intra_job_t *)d_jobs, n_jobs,
                                                                     (x264_cuda_intra_result_t *)d_results);
    LAUNCH_CHECK(ctx, "intra_mb_costs_kernel");
    return 0;
}

extern "C" int x264_cuda_intra_mb_costs(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, const x264_cuda_frame_t *fdec,
                                        const x264_cuda_intra_job_t *jobs, int n_jobs, x264_cuda_intra_result_t *results)
{
    x264_cuda_enter(ctx);
    if (n_jobs <= 0) return 0;
    const size_t jb = (size_t)n_jobs * sizeof(x264_cuda_intra_job_t), rb = (size_t)n_jobs * sizeof(x264_cuda_intra_result_t);
    const size_t jb_al = (jb + 255) & ~(size_t)255;
    if (x264_cuda_stage(ctx, jb_al + rb, jb_al + rb)) return -1;
    uint8_t *hs = (uint8_t *)ctx->h_stage, *ds = (uint8_t *)ctx->d_stage;
    if (x264_cuda_jobs_in(ctx, ds, jobs, hs, jb)) return -1;
    if (x264_cuda_intra_mb_costs_dev(ctx, fenc, fdec, ds, n_jobs, ds + jb_al)) return -1;
    if (x264_cuda_results_out(ctx, results, ds + jb_al, hs + jb_al, rb)) return -1;
    return 0;
}
