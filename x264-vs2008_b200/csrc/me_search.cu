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

template <int NW>
__device__ __forceinline__ void load_row(uint32_t (&dst)[NW], const uint32_t *p, int sh)
{
    uint32_t w[NW + 1];
#pragma unroll
    for (int i = 0; i <= NW; i++) w[i] = __ldg(p + i);
#pragma unroll
    for (int i = 0; i < NW; i++) dst[i] = __funnelshift_r(w[i], w[i + 1], sh);
}

// SAD of the register block F (BH x NW words) against the block at byte address `a` (any alignment)
template <int BW, int BH>
__device__ __forceinline__ int sad_block_at(const uint32_t (&F)[BH][BW / 4], const uint8_t *a, int stride)
{
    constexpr int NW = BW / 4;
    const int sh = ((uintptr_t)a & 3) * 8;
    const uint32_t *p = (const uint32_t *)((uintptr_t)a & ~(uintptr_t)3);
    uint32_t acc0 = 0, acc1 = 0;
#pragma unroll
    for (int y = 0; y < BH; y++) {
        uint32_t r[NW];
        load_row<NW>(r, (const uint32_t *)((const uint8_t *)p + (size_t)y * stride), sh);
#pragma unroll
        for (int w = 0; w < NW; w++) {
            if ((y + w) & 1) acc1 = sad4_acc(F[y][w], r[w], acc1);
            else acc0 = sad4_acc(F[y][w], r[w], acc0);
        }
    }
    return (int)(acc0 + acc1);
}

__device__ __forceinline__ int tab_at(const int16_t *tab, int i)
{
    // p_cost_mv is defined for |i| <= 2*4*2048 (analyse.c:196-203); x264's mv range keeps indices inside
    return tab[max(-2 * 4 * 2048, min(2 * 4 * 2048, i))];
}

template <int BW, int BH>
__device__ void search_job(const Geo &geo, const x264_cuda_me_job_t &job, const int16_t *tab, int me_range,
                           x264_cuda_me_result_t *res, int lane)
{
    constexpr int NW = BW / 4;
    const int stride = geo.stride;
    const int16_t *cmx = tab - job.mvp[0]; // p_cost_mvx, me.c:179-180
    const int16_t *cmy = tab - job.mvp[1];
    const int x_min = job.mv_min_fpel[0], y_min = job.mv_min_fpel[1];
    const int x_max = job.mv_max_fpel[0], y_max = job.mv_max_fpel[1];

    // ---- fenc block -> registers (uniform across the warp)
    uint32_t F[BH][NW];
    {
        const uint8_t *fe = geo.fenc + (size_t)job.by * stride + job.bx;
#pragma unroll
        for (int y = 0; y < BH; y++)
#pragma unroll
            for (int w = 0; w < NW; w++) F[y][w] = __ldg((const uint32_t *)(fe + (size_t)y * stride) + w);
    }
    const uint8_t *ref0 = geo.fref + (size_t)job.by * stride + job.bx; // m->p_fref[0]

    // ---- predictor stage, me.c:182-229 (i_subpel_refine < 3): lane i evaluates candidate i
    int bmx, bmy, bcost;
    const int pmx = (clip3i(job.mvp[0], x_min * 4, x_max * 4) + 2) >> 2;
    const int pmy = (clip3i(job.mvp[1], y_min * 4, y_max * 4) + 2) >> 2;
    if (job.flags & X264_CUDA_ME_SEEDED) {
        bmx = job.seed_mv[0]; bmy = job.seed_mv[1]; bcost = job.seed_cost;
    } else {
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
            cost = sad_block_at<BW, BH>(F, ref0 + (ptrdiff_t)cy * stride + cx, stride);
            if (lane != 0) cost += tab_at(cmx, cx << 2) + tab_at(cmy, cy << 2); // me.c:217: mvp cost is removed again
        }
        // sequential strict '<' over the list == min over (cost, list index).  Skipping "same as current best"
        // candidates (me.c:222) never changes the outcome: such a candidate cannot be strictly better.
        unsigned key = __reduce_min_sync(0xffffffffu, (unsigned)cost);
        unsigned who = __reduce_min_sync(0xffffffffu, (unsigned)cost == key ? (unsigned)lane : 0xffu);
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

    unsigned long long best_key = ~0ull;
    for (int c0 = 0; c0 < width; c0 += 32) {
        const int col = c0 + lane;
        const bool active = col < width;
        const int mx = min_x + (active ? col : 0);
        const uint8_t *a = ref0 + (ptrdiff_t)min_y * stride + mx;
        const int sh = ((uintptr_t)a & 3) * 8;
        const uint8_t *p = (const uint8_t *)((uintptr_t)a & ~(uintptr_t)3);
        const int cost_x = tab_at(cmx, mx << 2);
        uint32_t R[BH][NW]; // ring: slot s holds reference row r with r % BH == s
#pragma unroll
        for (int y = 0; y < BH - 1; y++) load_row<NW>(R[y], (const uint32_t *)(p + (size_t)y * stride), sh);
        int lane_best = 0x7fffffff, lane_my = 0;
        for (int base = 0; base < rows; base += BH) {
#pragma unroll
            for (int j = 0; j < BH; j++) {
                const int my_idx = base + j;
                if (my_idx >= rows) break; // warp-uniform
                load_row<NW>(R[(j + BH - 1) % BH], (const uint32_t *)(p + (size_t)(my_idx + BH - 1) * stride), sh);
                uint32_t acc[4] = { 0, 0, 0, 0 };
#pragma unroll
                for (int y = 0; y < BH; y++)
#pragma unroll
                    for (int w = 0; w < NW; w++) {
                        const int k = (y * NW + w) & 3;
                        acc[k] = sad4_acc(F[y][w], R[(j + y) % BH][w], acc[k]);
                    }
                const int cost = (int)(acc[0] + acc[1] + acc[2] + acc[3]) + cost_x + tab_at(cmy, (min_y + my_idx) << 2);
                if (cost < lane_best) { lane_best = cost; lane_my = my_idx; }
            }
        }
        if (active) {
            unsigned long long key = ((unsigned long long)(unsigned)lane_best << 24) | ((unsigned)lane_my << 12) | (unsigned)col;
            best_key = min(best_key, key);
        }
    }
    // warp argmin of (cost, my, mx)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, best_key, o);
        best_key = min(best_key, other);
    }
    const int w_cost = (int)(best_key >> 24);
    if (best_key != ~0ull && w_cost < bcost) {
        bcost = w_cost;
        bmy = min_y + (int)((best_key >> 12) & 0xfff);
        bmx = min_x + (int)(best_key & 0xfff);
    }
    if (lane == 0) {
        x264_cuda_me_result_t r;
        r.bmx = (int16_t)bmx; r.bmy = (int16_t)bmy; r.bcost = bcost;
        r.seed_mx = (int16_t)seed_mx; r.seed_my = (int16_t)seed_my; r.seed_cost = seed_cost;
        *res = r;
    }
}

__global__ void __launch_bounds__(128) me_search_kernel(Geo geo, const x264_cuda_me_job_t *__restrict__ jobs, int n_jobs,
                                                        const int16_t *const *__restrict__ cost_tabs, int me_range,
                                                        x264_cuda_me_result_t *__restrict__ results)
{
    const int lane = threadIdx.x & 31;
    const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
    for (int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < n_jobs; j += warps_per_grid) {
        const x264_cuda_me_job_t job = jobs[j];
        const int16_t *tab = cost_tabs[job.qp] + 2 * 4 * 2048;
        switch (job.i_pixel) {
        case X264_CUDA_PIXEL_16x16: search_job<16, 16>(geo, job, tab, me_range, results + j, lane); break;
        case X264_CUDA_PIXEL_16x8:  search_job<16, 8>(geo, job, tab, me_range, results + j, lane); break;
        case X264_CUDA_PIXEL_8x16:  search_job<8, 16>(geo, job, tab, me_range, results + j, lane); break;
        case X264_CUDA_PIXEL_8x8:   search_job<8, 8>(geo, job, tab, me_range, results + j, lane); break;
        case X264_CUDA_PIXEL_8x4:   search_job<8, 4>(geo, job, tab, me_range, results + j, lane); break;
        case X264_CUDA_PIXEL_4x8:   search_job<4, 8>(geo, job, tab, me_range, results + j, lane); break;
        default:                    search_job<4, 4>(geo, job, tab, me_range, results + j, lane); break;
        }
    }
}

} // namespace

extern "C" int x264_cuda_me_search_dev(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, const x264_cuda_frame_t *fref,
                                       int me_range, const void *d_jobs, int n_jobs, void *d_results)
{
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
    const int warps_per_block = 4;
    int blocks = (n_jobs + warps_per_block - 1) / warps_per_block;
    me_search_kernel<<<blocks, warps_per_block * 32, 0, ctx->stream>>>(geo, (const x264_cuda_me_job_t *)d_jobs, n_jobs, d_tabs,
                                                                        me_range, (x264_cuda_me_result_t *)d_results);
    LAUNCH_CHECK(ctx, "me_search_kernel");
    return 0;
}

extern "C" int x264_cuda_me_search(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, const x264_cuda_frame_t *fref,
                                   int me_range, const x264_cuda_me_job_t *jobs, int n_jobs, x264_cuda_me_result_t *results)
{
    if (n_jobs <= 0) return 0;
    const size_t jb = (size_t)n_jobs * sizeof(x264_cuda_me_job_t), rb = (size_t)n_jobs * sizeof(x264_cuda_me_result_t);
    const size_t jb_al = (jb + 255) & ~(size_t)255;
    if (x264_cuda_stage(ctx, jb_al + rb, jb_al + rb)) return -1;
    uint8_t *hs = (uint8_t *)ctx->h_stage, *ds = (uint8_t *)ctx->d_stage;
    memcpy(hs, jobs, jb); // pageable caller memory -> pinned staging, so the copies below are truly asynchronous
    CUDA_TRY(ctx, cudaMemcpyAsync(ds, hs, jb, cudaMemcpyHostToDevice, ctx->stream));
    if (x264_cuda_me_search_dev(ctx, fenc, fref, me_range, ds, n_jobs, ds + jb_al)) return -1;
    CUDA_TRY(ctx, cudaMemcpyAsync(hs + jb_al, ds + jb_al, rb, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(results, hs + jb_al, rb);
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
