// me_search.cu — batched full-pel motion search: the ESA branch of x264_me_search_ref
// (S/encoder/me.c:156-229 predictor stage, :449-600 exhaustive window) for thousands of independent blocks.
//
// Mapping: ONE WARP PER JOB.  Lane l owns candidate column mx = min_x + l (the reference window is at most
// (2*merange+3)&~3 columns wide; wider windows are walked in chunks of 32 columns).  The warp walks the rows my
// top to bottom.  Each lane keeps the BH reference rows its block currently covers in a register ring, so moving
// one row down costs one new row of loads (BW/4+1 aligned words, funnel-shifted to the lane's byte phase) and
// BH*BW/4 VABSDIFF4.U8.ACC instructions — the whole search is integer-ALU-pipe work, which is its roofline.
// The encode block (fenc) lives in registers too (same for all lanes).  Reference rows are read straight
// through L1 (neighbouring lanes/rows/jobs hit the same 128-byte lines); no shared memory, no block barrier.
//
// Exactness: ESA == plain raster-order argmin with strict '<', seeded by the predictor stage.
// Each lane scans its column top-down with strict '<' (first row wins), then the warp takes the minimum of
// (cost, my, mx) lexicographically == first candidate in raster order among the minima; the seed wins ties.
#include "common.cuh"

namespace {

struct Geo { const uint8_t *fenc; const uint8_t *fref; int stride; };

#define ME_WARPS_PER_BLOCK 4
#define ME_ROW_CHUNK 64 // rows of partial SADs kept in shared memory between the two column passes of 16-wide blocks

template <int NW>
__device__ __forceinline__ void load_row(uint32_t (&dst)[NW], const uint8_t *p, int sh)
{
    uint32_t w[NW + 1];
#pragma unroll
    for (int i = 0; i <= NW; i++) w[i] = __ldg((const uint32_t *)p + i);
#pragma unroll
    for (int i = 0; i < NW; i++) dst[i] = __funnelshift_r(w[i], w[i + 1], sh);
}

__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// One lane walks its candidate column top-down over `nrows` rows.  mode 0: complete block (cost -> best key);
// mode 1: left half of a 16-wide block (SAD -> partial[]); mode 2: right half (adds partial[], cost -> best key).
// The ring R keeps the BH reference rows under the block; slot s holds the row r with r % BH == s, which is
// static after unrolling by BH.  key = cost<<12 | row: min() == strict '<' with the earlier row winning ties.
template <int NW, int BH>
__device__ __noinline__ void scan_column(const uint8_t *fe, const uint8_t *p, int sh, int stride, int nrows, int mode,
                                         int cost_x, const int16_t *cmy_row, uint16_t *partial, int row0, uint32_t &best_key)
{
    uint32_t F[BH][NW], R[BH][NW];
#pragma unroll
    for (int y = 0; y < BH; y++)
#pragma unroll
        for (int w = 0; w < NW; w++) F[y][w] = __ldg((const uint32_t *)(fe + (size_t)y * stride) + w);
#pragma unroll
    for (int y = 0; y < BH - 1; y++) load_row<NW>(R[y], p + (size_t)y * stride, sh);
    int cy = cmy_row[0];
    for (int base = 0; base < nrows; base += BH) {
#pragma unroll
        for (int j = 0; j < BH; j++) {
            const int r = base + j;
            if (r >= nrows) break; // warp-uniform
            load_row<NW>(R[(j + BH - 1) % BH], p + (size_t)(r + BH - 1) * stride, sh);
            const int cy_next = cmy_row[4 * (r + 1)]; // one row ahead: hides the L1 latency (the table has slack)
            uint32_t acc[4] = { 0, 0, 0, 0 };
#pragma unroll
            for (int y = 0; y < BH; y++)
#pragma unroll
                for (int w = 0; w < NW; w++) {
                    const int k = (y * NW + w) & 3;
                    acc[k] = sad4_acc(F[y][w], R[(j + y) % BH][w], acc[k]);
                }
            uint32_t sad = (acc[0] + acc[1]) + (acc[2] + acc[3]);
            if (mode == 1) {
                partial[r * 32] = (uint16_t)sad;
            } else {
                if (mode == 2) sad += partial[r * 32];
                const uint32_t key = ((sad + cost_x + cy) << 12) | (uint32_t)(row0 + r);
                best_key = min(best_key, key);
            }
            cy = cy_next;
        }
    }
}

// SAD of the BH x (4*NW*npass) fenc block against the reference block at byte address `a`
template <int NW, int BH>
__device__ __noinline__ int sad_block_at(const uint8_t *fe, const uint8_t *a, int stride, int npass)
{
    uint32_t acc0 = 0, acc1 = 0;
    for (int h = 0; h < npass; h++, fe += 4 * NW, a += 4 * NW) {
        const int sh = ((uintptr_t)a & 3) * 8;
        const uint8_t *p = (const uint8_t *)((uintptr_t)a & ~(uintptr_t)3);
#pragma unroll
        for (int y = 0; y < BH; y++) {
            uint32_t r[NW];
            load_row<NW>(r, p + (size_t)y * stride, sh);
#pragma unroll
            for (int w = 0; w < NW; w++) {
                const uint32_t f = __ldg((const uint32_t *)(fe + (size_t)y * stride) + w);
                if ((y + w) & 1) acc1 = sad4_acc(f, r[w], acc1);
                else acc0 = sad4_acc(f, r[w], acc0);
            }
        }
    }
    return (int)(acc0 + acc1);
}

// NW words per column pass; 16-wide blocks (npass == 2) take two 8-wide passes, which halves the registers and
// lets 16x16/8x16 (and 16x8/8x8) share one instantiation — the hot loops of a frame's job mix fit the I-cache.
// variant -> (NW, BH) instantiation shared by several partition shapes
__device__ __forceinline__ void scan_dispatch(int variant, const uint8_t *fe, const uint8_t *p, int sh, int stride, int nrows,
                                              int mode, int cost_x, const int16_t *cmy_row, uint16_t *partial, int row0,
                                              uint32_t &key)
{
    switch (variant) {
    case 0: scan_column<2, 16>(fe, p, sh, stride, nrows, mode, cost_x, cmy_row, partial, row0, key); break;
    case 1: scan_column<2, 8>(fe, p, sh, stride, nrows, mode, cost_x, cmy_row, partial, row0, key); break;
    case 2: scan_column<2, 4>(fe, p, sh, stride, nrows, mode, cost_x, cmy_row, partial, row0, key); break;
    case 3: scan_column<1, 8>(fe, p, sh, stride, nrows, mode, cost_x, cmy_row, partial, row0, key); break;
    default: scan_column<1, 4>(fe, p, sh, stride, nrows, mode, cost_x, cmy_row, partial, row0, key); break;
    }
}
__device__ __forceinline__ int sad_dispatch(int variant, const uint8_t *fe, const uint8_t *a, int stride, int npass)
{
    switch (variant) {
    case 0: return sad_block_at<2, 16>(fe, a, stride, npass);
    case 1: return sad_block_at<2, 8>(fe, a, stride, npass);
    case 2: return sad_block_at<2, 4>(fe, a, stride, npass);
    case 3: return sad_block_at<1, 8>(fe, a, stride, npass);
    default: return sad_block_at<1, 4>(fe, a, stride, npass);
    }
}

__device__ void search_job(const Geo &geo, const x264_cuda_me_job_t &job, const int16_t *tab, int me_range,
                           x264_cuda_me_result_t *res, int lane, uint16_t *partial)
{
    // i_pixel -> {variant, passes}: 16x16 8x16 share <2,16>, 16x8 8x8 share <2,8>
    const int ip = min((int)job.i_pixel, 6);
    const int variant = ip == X264_CUDA_PIXEL_16x16 || ip == X264_CUDA_PIXEL_8x16 ? 0
                      : ip == X264_CUDA_PIXEL_16x8 || ip == X264_CUDA_PIXEL_8x8 ? 1
                      : ip == X264_CUDA_PIXEL_8x4 ? 2 : ip == X264_CUDA_PIXEL_4x8 ? 3 : 4;
    const int npass = ip <= X264_CUDA_PIXEL_16x8 ? 2 : 1;
    const int NW = variant <= 2 ? 2 : 1;
    const int BH = variant == 0 ? 16 : (variant == 1 || variant == 3) ? 8 : 4;
    const int BW = 4 * NW * npass;
    const int stride = geo.stride;
    const int16_t *cmx = tab - job.mvp[0]; // p_cost_mvx, me.c:179-180
    const int16_t *cmy = tab - job.mvp[1];
    const int x_min = job.mv_min_fpel[0], y_min = job.mv_min_fpel[1];
    const int x_max = job.mv_max_fpel[0], y_max = job.mv_max_fpel[1];

    // every table index this job can form must stay inside p_cost_mv's +-2*4*2048 (analyse.c:196-203)
    {
        const int lim = 2 * 4 * 2048 - 8;
        const int ax = max(abs(4 * x_min - job.mvp[0]), abs(4 * x_max - job.mvp[0]));
        const int ay = max(abs(4 * y_min - job.mvp[1]), abs(4 * y_max - job.mvp[1]));
        if (ax > lim || ay > lim || x_min > 0 || x_max < 0 || y_min > 0 || y_max < 0) {
            if (lane == 0) { x264_cuda_me_result_t r = { 0, 0, -1, 0, 0, -1 }; *res = r; } // rejected job
            return;
        }
    }
    const uint8_t *fe = geo.fenc + (size_t)job.by * stride + job.bx;   // m->p_fenc[0] (in place, plane stride)
    const uint8_t *ref0 = geo.fref + (size_t)job.by * stride + job.bx; // m->p_fref[0]

    // ---- predictor stage, me.c:182-229 (i_subpel_refine < 3): lane i evaluates candidate i
    int bmx, bmy, bcost;
    if (job.flags & X264_CUDA_ME_SEEDED) {
        bmx = clip3i(job.seed_mv[0], x_min, x_max); bmy = clip3i(job.seed_mv[1], y_min, y_max); bcost = job.seed_cost;
    } else {
        const int pmx = (clip3i(job.mvp[0], x_min * 4, x_max * 4) + 2) >> 2;
        const int pmy = (clip3i(job.mvp[1], y_min * 4, y_max * 4) + 2) >> 2;
        const int n_mvc = min((int)job.i_mvc, X264_CUDA_ME_MAX_MVC);
        int cx = 0, cy = 0, valid = 0;
        if (lane == 0) { cx = pmx; cy = pmy; valid = 1; }
        else if (lane <= n_mvc) {
            int mx = (job.mvc[lane - 1][0] + 2) >> 2, my = (job.mvc[lane - 1][1] + 2) >> 2;
            valid = (mx | my) != 0; // zero predictors are covered by the final (0,0) test (me.c:220)
            cx = clip3i(mx, x_min, x_max); cy = clip3i(my, y_min, y_max);
        } else if (lane == n_mvc + 1) { cx = 0; cy = 0; valid = 1; }
        int cost = COST_MAX + 1;
        if (valid) {
            cost = sad_dispatch(variant, fe, ref0 + (ptrdiff_t)cy * stride + cx, stride, npass);
            if (lane != 0) cost += cmx[cx << 2] + cmy[cy << 2]; // me.c:217: the mvp's own MV cost is removed again
        }
        // sequential strict '<' over the list == min over (cost, list index).  Skipping "same as current best"
        // candidates (me.c:222) never changes the outcome: such a candidate cannot be strictly better.
        const unsigned key = __reduce_min_sync(0xffffffffu, (unsigned)cost);
        const unsigned who = __reduce_min_sync(0xffffffffu, (unsigned)cost == key ? (unsigned)lane : 0xffu);
        bcost = (int)key;
        bmx = __shfl_sync(0xffffffffu, cx, who);
        bmy = __shfl_sync(0xffffffffu, cy, who);
    }
    const int seed_mx = bmx, seed_my = bmy, seed_cost = bcost;

    // ---- exhaustive window, me.c:451-457
    const int min_x = max(bmx - me_range, x_min), min_y = max(bmy - me_range, y_min);
    const int max_x = min(bmx + me_range, x_max), max_y = min(bmy + me_range, y_max);
    const int width = (max_x - min_x + 3) & ~3;
    const int rows = max_y - min_y + 1;

    unsigned long long best = ~0ull; // (cost<<12|row) << 12 | col
    for (int c0 = 0; c0 < width; c0 += 32) {
        const int col = c0 + lane;
        const bool active = col < width;
        const int mx = min_x + (active ? col : 0);
        const int cost_x = cmx[mx << 2];
        for (int row0 = 0; row0 < rows; row0 += ME_ROW_CHUNK) {
            const int nrows = min(ME_ROW_CHUNK, rows - row0);
            const uint8_t *a = ref0 + (ptrdiff_t)(min_y + row0) * stride + mx;
            // pull the warp's whole window tile into L1 up front: later row loads are L1 hits consumed a full
            // row-step after issue.  The tile is (nrows+BH-1) rows x (32+BW+3) bytes, i.e. at most 2 lines a row.
            {
                const uint8_t *t0 = ref0 + (ptrdiff_t)(min_y + row0) * stride + min_x + c0;
                for (int r = lane; r < nrows + BH - 1; r += 32) {
                    prefetch_l1(t0 + (size_t)r * stride);
                    prefetch_l1(t0 + (size_t)r * stride + 32 + BW);
                }
            }
            const int sh = ((uintptr_t)a & 3) * 8;
            const uint8_t *p = (const uint8_t *)((uintptr_t)a & ~(uintptr_t)3);
            const int16_t *cmy_row = cmy + 4 * (min_y + row0);
            uint32_t key = 0xffffffffu;
            if (npass == 1) {
                scan_dispatch(variant, fe, p, sh, stride, nrows, 0, cost_x, cmy_row, partial + lane, row0, key);
            } else {
                scan_dispatch(variant, fe, p, sh, stride, nrows, 1, cost_x, cmy_row, partial + lane, row0, key);
                __syncwarp();
                scan_dispatch(variant, fe + 4 * NW, p + 4 * NW, sh, stride, nrows, 2, cost_x, cmy_row, partial + lane, row0, key);
                __syncwarp();
            }
            if (active) best = min(best, ((unsigned long long)key << 12) | (unsigned)col);
        }
    }
    // warp argmin of (cost, my, mx): first candidate in raster order among the minima
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
    const int w_cost = (int)(best >> 24);
    if (best != ~0ull && w_cost < bcost) {
        bcost = w_cost;
        bmy = min_y + (int)((best >> 12) & 0xfff);
        bmx = min_x + (int)(best & 0xfff);
    }
    if (lane == 0) {
        x264_cuda_me_result_t r;
        r.bmx = (int16_t)bmx; r.bmy = (int16_t)bmy; r.bcost = bcost;
        r.seed_mx = (int16_t)seed_mx; r.seed_my = (int16_t)seed_my; r.seed_cost = seed_cost;
        *res = r;
    }
}

__global__ void __launch_bounds__(ME_WARPS_PER_BLOCK * 32, 4)
me_search_kernel(Geo geo, const x264_cuda_me_job_t *__restrict__ jobs, int n_jobs, const int16_t *const *__restrict__ cost_tabs,
                 int me_range, x264_cuda_me_result_t *__restrict__ results)
{
    __shared__ uint16_t s_partial[ME_WARPS_PER_BLOCK][ME_ROW_CHUNK * 32];
    __shared__ x264_cuda_me_job_t s_job[ME_WARPS_PER_BLOCK];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
    for (int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < n_jobs; j += warps_per_grid) {
        // stage the 76-byte job through shared memory (19 words, one per lane)
        __syncwarp();
        if (lane < (int)(sizeof(x264_cuda_me_job_t) / 4)) ((uint32_t *)&s_job[wid])[lane] = __ldg((const uint32_t *)(jobs + j) + lane);
        __syncwarp();
        const x264_cuda_me_job_t &job = s_job[wid];
        const int16_t *tab = cost_tabs[job.qp > 51 ? 51 : job.qp] + 2 * 4 * 2048;
        uint16_t *partial = s_partial[wid];
        search_job(geo, job, tab, me_range, results + j, lane, partial);
    }
}

} // namespace

extern "C" int x264_cuda_me_search_dev(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, const x264_cuda_frame_t *fref,
                                       int me_range, const void *d_jobs, int n_jobs, void *d_results)
{
    x264_cuda_enter(ctx);
    if (n_jobs <= 0) return 0;
    if (fenc->g.stride != fref->g.stride || fenc->g.lines != fref->g.lines) {
        snprintf(ctx->err, 256, "x264_cuda_me_search: fenc/fref geometry mismatch");
        return -1;
    }
    if (me_range < 1 || me_range > 1024) {
        snprintf(ctx->err, 256, "x264_cuda_me_search: me_range %d out of range", me_range);
        return -1;
    }
    const int16_t *const *d_tabs;
    if (x264_cuda_cost_tables(ctx, &d_tabs)) return -1;
    Geo geo = { fenc->plane[0], fref->plane[0], fenc->g.stride };
    const int warps_per_block = ME_WARPS_PER_BLOCK;
    int blocks = (n_jobs + warps_per_block - 1) / warps_per_block;
    me_search_kernel<<<blocks, warps_per_block * 32, 0, ctx->stream>>>(geo, (const x264_cuda_me_job_t *)d_jobs, n_jobs, d_tabs,
                                                                        me_range, (x264_cuda_me_result_t *)d_results);
    LAUNCH_CHECK(ctx, "me_search_kernel");
    return 0;
}

extern "C" int x264_cuda_me_search(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, const x264_cuda_frame_t *fref,
                                   int me_range, const x264_cuda_me_job_t *jobs, int n_jobs, x264_cuda_me_result_t *results)
{
    x264_cuda_enter(ctx);
    if (n_jobs <= 0) return 0;
    const size_t jb = (size_t)n_jobs * sizeof(x264_cuda_me_job_t), rb = (size_t)n_jobs * sizeof(x264_cuda_me_result_t);
    const size_t jb_al = (jb + 255) & ~(size_t)255;
    if (x264_cuda_stage(ctx, jb_al + rb, jb_al + rb)) return -1;
    uint8_t *hs = (uint8_t *)ctx->h_stage, *ds = (uint8_t *)ctx->d_stage;
    if (x264_cuda_jobs_in(ctx, ds, jobs, hs, jb)) return -1;
    if (x264_cuda_me_search_dev(ctx, fenc, fref, me_range, ds, n_jobs, ds + jb_al)) return -1;
    if (x264_cuda_results_out(ctx, results, ds + jb_al, hs + jb_al, rb)) return -1;
    return 0;
}

// me.c:603-630 for i_subpel_refine < 2 — pure host arithmetic on the result
extern "C" void x264_cuda_me_finish(const x264_cuda_me_job_t *job, const x264_cuda_me_result_t *res, const int16_t *cost_table,
                                    int mv_max_spel_y, int16_t mv[2], int *cost, int *cost_mv)
{
    const int16_t *c = cost_table + 2 * 4 * 2048;
    int x_min = job->mv_min_fpel[0] * 4, x_max = job->mv_max_fpel[0] * 4;
    int y_min = job->mv_min_fpel[1] * 4, y_max = job->mv_max_fpel[1] * 4;
    int pmx = ((job->mvp[0] < x_min ? x_min : job->mvp[0] > x_max ? x_max : job->mvp[0]) + 2) >> 2;
    int pmy = ((job->mvp[1] < y_min ? y_min : job->mvp[1] > y_max ? y_max : job->mvp[1]) + 2) >> 2;
    int mvx = res->bmx << 2, mvy = res->bmy << 2;
    *cost = res->bcost;
    *cost_mv = c[mvx - job->mvp[0]] + c[mvy - job->mvp[1]];
    if (res->bmx == pmx && res->bmy == pmy) *cost += *cost_mv;
    if (mvy > mv_max_spel_y) mvy = mv_max_spel_y;
    mv[0] = (int16_t)mvx; mv[1] = (int16_t)mvy;
}
