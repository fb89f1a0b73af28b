// sad_grid.cu — candidate grids: the SAD of all nine inter partitions of a macroblock at every integer vector of a window,
// written out (north_star: "kernels that evaluate whole per-macroblock candidate grids ... the lambda-weighted MV-bit cost and
// the choice among candidates remain in the reference's sequential neighbour-predictor order").  The host replays
// x264_me_search_ref's predictor stage and ESA loop on a grid with the exact, sequentially known mvp (host/x264_cuda_host.c:
// x264_cuda_host_esa_replay); the winner-only kernels (me_search_mb.cu) are the path when predictors are known up front.
//
// One warp per macroblock, same scan as me_search_mb.cu: lane = column, 16-row register ring of the reference strip, the
// macroblock's 64 VABSDIFF4 per position give the four 8x8 quadrant SADs and from them the nine partition SADs, which are
// stored (uint16; 0xffff outside the MV limits) instead of reduced.  18 B written per position: with the full grid written
// the kernel streams ~160 MB per 1080p frame.
#include "common.cuh"

namespace {

struct GridGeo { const uint8_t *fenc; const uint8_t *fref; int stride; };

__device__ __forceinline__ void grid_load_row16(uint32_t (&dst)[4], const uint8_t *p, int sh)
{
    uint32_t w[5];
#pragma unroll
    for (int i = 0; i < 5; i++) w[i] = __ldg((const uint32_t *)p + i);
#pragma unroll
    for (int i = 0; i < 4; i++) dst[i] = __funnelshift_r(w[i], w[i + 1], sh);
}

// QUAD = false: nine partition planes per job, grid[job][part][GH][GW].
// QUAD = true : the four 8x8 quadrant SADs (TL, TR, BL, BR) interleaved per position, grid[job][GH][GW][4] — every partition SAD is a
//               sum of these (PIXEL_SAD_C is a plain sum, S/common/pixel.c:40-56), 8 B instead of 18 B per position and one 8-byte
//               store per lane (a warp row = 256 contiguous bytes).
template <bool QUAD>
__global__ void __launch_bounds__(128, QUAD ? 3 : 2) // the nine-plane form keeps nine running sums per position: two CTAs per SM, no spills
sad_grid_kernel(GridGeo geo, const x264_cuda_grid_job_t *__restrict__ jobs, int n_jobs, int radius, uint16_t *__restrict__ grid)
{
    __shared__ __align__(16) uint32_t s_F[4][16][4];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int jb = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (jb >= n_jobs) return;
    const x264_cuda_grid_job_t job = jobs[jb];
    const int GW = X264_CUDA_GRID_W(radius), GH = X264_CUDA_GRID_H(radius), stride = geo.stride;
    const uint8_t *fe = geo.fenc + (size_t)job.mb_y * 16 * stride + job.mb_x * 16;
    const uint8_t *ref0 = geo.fref + (size_t)job.mb_y * 16 * stride + job.mb_x * 16;
    for (int i = lane; i < 64; i += 32) s_F[wid][i >> 2][i & 3] = __ldg((const uint32_t *)(fe + (size_t)(i >> 2) * stride) + (i & 3));
    __syncwarp();
    const uint4 *F4 = (const uint4 *)&s_F[wid][0][0];
    uint16_t *out = grid + (size_t)jb * (QUAD ? 4 : 9) * GH * GW;
    const size_t pstride = (size_t)GH * GW;
    const int ux0 = job.cx - radius, uy0 = job.cy - radius;
    for (int c0 = 0; c0 < GW; c0 += 32) {
        // narrow tail chunks are folded to 16x2 / 8x4 / 4x8 (columns x row segments) so that all 32 lanes stay busy
        const int rem = GW - c0;
        const int cw = rem > 16 ? 32 : rem > 8 ? 16 : rem > 4 ? 8 : 4, segs = 32 / cw;
        const int lcol = lane & (cw - 1), seg = lane / cw;
        const int col = c0 + lcol;
        const bool col_ok = col < GW;
        const int mx = ux0 + min(col, GW - 1);
        const int seg_rows = (GH + segs - 1) / segs, rbeg = seg * seg_rows;
        // columns outside the MV limits are not even loaded from (they may lie outside the padded plane).  The ESA loop rounds its
        // width up to a multiple of 4 and so tests up to 3 columns beyond mv_max_fpel[0] (me.c:456-457): those are kept.
        const bool x_ok = col_ok && mx >= job.mv_min_fpel[0] && mx <= job.mv_max_fpel[0] + 3;
        const int mxc = clip3i(mx, job.mv_min_fpel[0], job.mv_max_fpel[0] + 3);
        const uint8_t *a = ref0 + (ptrdiff_t)(uy0 + rbeg) * stride + mxc;
        const int sh = ((uintptr_t)a & 3) * 8;
        const uint8_t *pr = (const uint8_t *)((uintptr_t)a & ~(uintptr_t)3);
        // rows are clamped into the MV limits for loading; their results are overwritten with 0xffff
        const int ylo = job.mv_min_fpel[1] - uy0 - rbeg, yhi = job.mv_max_fpel[1] - uy0 - rbeg; // valid r range (relative to rbeg)
        auto row_ptr = [&](int r) { return pr + (ptrdiff_t)clip3i(r, ylo, yhi + 15) * stride; };
        uint32_t R[16][4];
#pragma unroll
        for (int y = 0; y < 15; y++) grid_load_row16(R[y], row_ptr(y), sh);
        for (int base = 0; base < seg_rows; base += 16) {
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const int r = base + j;
                if (r >= seg_rows) break; // warp-uniform
                grid_load_row16(R[(j + 15) % 16], row_ptr(r + 15), sh);
                uint32_t tl = 0, tr = 0, bl = 0, br = 0, tl2 = 0, tr2 = 0, bl2 = 0, br2 = 0;
#pragma unroll
                for (int y = 0; y < 8; y++) {
                    const uint4 f = F4[y], g = F4[y + 8];
                    tl = sad4_acc(f.x, R[(j + y) % 16][0], tl); tr = sad4_acc(f.z, R[(j + y) % 16][2], tr);
                    bl = sad4_acc(g.x, R[(j + y + 8) % 16][0], bl); br = sad4_acc(g.z, R[(j + y + 8) % 16][2], br);
                    tl2 = sad4_acc(f.y, R[(j + y) % 16][1], tl2); tr2 = sad4_acc(f.w, R[(j + y) % 16][3], tr2);
                    bl2 = sad4_acc(g.y, R[(j + y + 8) % 16][1], bl2); br2 = sad4_acc(g.w, R[(j + y + 8) % 16][3], br2);
                }
                tl += tl2; tr += tr2; bl += bl2; br += br2;
                const int row = rbeg + r;
                if (col_ok && row < GH) {
                    const bool ok = x_ok && r >= ylo && r <= yhi;
                    if (QUAD) {
                        uint2 v2 = ok ? make_uint2(tl | (tr << 16), bl | (br << 16)) : make_uint2(0xffffffffu, 0xffffffffu);
                        *(uint2 *)(out + ((size_t)row * GW + col) * 4) = v2;
                        continue;
                    }
                    const uint32_t top = tl + tr, bot = bl + br, lft = tl + bl, rgt = tr + br, all = top + bot;
                    const uint32_t v[9] = { all, top, bot, lft, rgt, tl, tr, bl, br };
                    uint16_t *o = out + (size_t)row * GW + col;
#pragma unroll
                    for (int p = 0; p < 9; p++) o[p * pstride] = (ok && (job.part_mask >> p & 1)) ? (uint16_t)v[p] : (uint16_t)0xffff;
                }
            }
        }
    }
}

} // namespace

static int sad_grid_launch(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, const x264_cuda_frame_t *fref, int radius, const void *d_jobs,
                           int n_jobs, void *d_grid, bool quad)
{
    x264_cuda_enter(ctx);
    if (n_jobs <= 0) return 0;
    if (fenc->g.stride != fref->g.stride || fenc->g.lines != fref->g.lines || radius < 1 || radius > 64) {
        snprintf(ctx->err, 256, "x264_cuda_sad_grid: fenc/fref geometry mismatch or radius not in 1..64");
        return -1;
    }
    GridGeo geo = { fenc->plane[0], fref->plane[0], fenc->g.stride };
    if (quad)
        sad_grid_kernel<true><<<(n_jobs + 3) / 4, 128, 0, ctx->stream>>>(geo, (const x264_cuda_grid_job_t *)d_jobs, n_jobs, radius, (uint16_t *)d_grid);
    else
        sad_grid_kernel<false><<<(n_jobs + 3) / 4, 128, 0, ctx->stream>>>(geo, (const x264_cuda_grid_job_t *)d_jobs, n_jobs, radius, (uint16_t *)d_grid);
    LAUNCH_CHECK(ctx, "sad_grid_kernel");
    return 0;
}
extern "C" int x264_cuda_sad_grid_dev(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, const x264_cuda_frame_t *fref, int radius, const void *d_jobs,
                                      int n_jobs, void *d_grid)
{
    return sad_grid_launch(ctx, fenc, fref, radius, d_jobs, n_jobs, d_grid, false);
}
extern "C" int x264_cuda_sad_grid_quad_dev(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, const x264_cuda_frame_t *fref, int radius, const void *d_jobs,
                                           int n_jobs, void *d_grid)
{
    return sad_grid_launch(ctx, fenc, fref, radius, d_jobs, n_jobs, d_grid, true);
}

extern "C" int x264_cuda_sad_grid(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, const x264_cuda_frame_t *fref, int radius,
                                  const x264_cuda_grid_job_t *jobs, int n_jobs, uint16_t *grid)
{
    x264_cuda_enter(ctx);
    if (n_jobs <= 0) return 0;
    const size_t jb = (size_t)n_jobs * sizeof(x264_cuda_grid_job_t), jb_al = (jb + 255) & ~(size_t)255;
    const size_t gb = (size_t)n_jobs * 9 * X264_CUDA_GRID_W(radius) * X264_CUDA_GRID_H(radius) * sizeof(uint16_t);
    if (x264_cuda_stage(ctx, jb_al + gb, jb_al + 256)) return -1; // the grid goes straight to the caller's memory (page-lock it: it is large)
    uint8_t *hs = (uint8_t *)ctx->h_stage, *ds = (uint8_t *)ctx->d_stage;
    if (x264_cuda_jobs_in(ctx, ds, jobs, hs, jb)) return -1;
    if (x264_cuda_sad_grid_dev(ctx, fenc, fref, radius, ds, n_jobs, ds + jb_al)) return -1;
    CUDA_TRY(ctx, cudaMemcpyAsync(grid, ds + jb_al, gb, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

// Quadrant grids through host memory.  `async` != 0: nothing is waited for — jobs must stay valid and `grid` must be page-locked
// (x264_cuda_host_alloc) until x264_cuda_fence_wait() on a fence recorded after this call returns; the device-side job copy and
// grid live in a per-call slice of the context's grid ring (x264_cuda_grid_ring), so several calls may be in flight.
extern "C" int x264_cuda_sad_grid_quad(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, const x264_cuda_frame_t *fref, int radius,
                                       const x264_cuda_grid_job_t *jobs, int n_jobs, uint16_t *grid, int async)
{
    x264_cuda_enter(ctx);
    if (n_jobs <= 0) return 0;
    const size_t jb = (size_t)n_jobs * sizeof(x264_cuda_grid_job_t), jb_al = (jb + 255) & ~(size_t)255;
    const size_t gb = (size_t)n_jobs * X264_CUDA_GRID_QUAD_BYTES(radius);
    uint8_t *ds;
    if (async) {
        ds = (uint8_t *)x264_cuda_grid_ring(ctx, jb_al + gb);
        if (!ds) return -1;
        CUDA_TRY(ctx, cudaMemcpyAsync(ds, jobs, jb, cudaMemcpyHostToDevice, ctx->stream));
    } else {
        if (x264_cuda_stage(ctx, jb_al + gb, jb_al + 256)) return -1;
        ds = (uint8_t *)ctx->d_stage;
        if (x264_cuda_jobs_in(ctx, ds, jobs, ctx->h_stage, jb)) return -1;
    }
    if (sad_grid_launch(ctx, fenc, fref, radius, ds, n_jobs, ds + jb_al, true)) return -1;
    CUDA_TRY(ctx, cudaMemcpyAsync(grid, ds + jb_al, gb, cudaMemcpyDeviceToHost, ctx->stream));
    if (!async) CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

// One-shot grids with NO copies: the kernel reads the jobs from and writes the grids to page-locked host memory itself (mapped into the
// device's address space), on a high-priority stream of its own.  This is the latency path of the live encoder — a macroblock whose
// guessed centre was off needs its grid NOW, and must neither queue behind the asynchronous chunk traffic of the context's main stream
// nor behind other encoder threads' copies on the shared copy engines.  20 KB per macroblock cross PCIe as 256-byte posted writes.
extern "C" int x264_cuda_sad_grid_quad_direct(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, const x264_cuda_frame_t *fref, int radius,
                                              const x264_cuda_grid_job_t *jobs, int n_jobs, uint16_t *grid)
{
    x264_cuda_enter(ctx);
    if (n_jobs <= 0) return 0;
    if (fenc->g.stride != fref->g.stride || fenc->g.lines != fref->g.lines || radius < 1 || radius > 64) {
        snprintf(ctx->err, 256, "x264_cuda_sad_grid: fenc/fref geometry mismatch or radius not in 1..64");
        return -1;
    }
    if (!ctx->aux_stream) {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi); // hi is the numerically lowest = highest priority
        CUDA_TRY(ctx, cudaStreamCreateWithPriority(&ctx->aux_stream, cudaStreamNonBlocking, hi));
    }
    void *d_jobs = nullptr, *d_grid = nullptr;
    if (cudaHostGetDevicePointer(&d_jobs, (void *)jobs, 0) != cudaSuccess || cudaHostGetDevicePointer(&d_grid, (void *)grid, 0) != cudaSuccess) {
        cudaGetLastError();
        snprintf(ctx->err, 256, "x264_cuda_sad_grid_quad_direct: jobs and grid must be page-locked (x264_cuda_host_alloc)");
        return -1;
    }
    GridGeo geo = { fenc->plane[0], fref->plane[0], fenc->g.stride };
    sad_grid_kernel<true><<<(n_jobs + 3) / 4, 128, 0, ctx->aux_stream>>>(geo, (const x264_cuda_grid_job_t *)d_jobs, n_jobs, radius, (uint16_t *)d_grid);
    LAUNCH_CHECK(ctx, "sad_grid_kernel (direct)");
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->aux_stream));
    return 0;
}
