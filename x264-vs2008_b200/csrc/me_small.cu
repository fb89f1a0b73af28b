// me_small.cu — the iterative searches (DIA, HEX), the "-> qpel mv" step and refine_subpel of x264_me_search_ref
// (S/encoder/me.c:156-305, :603-631, :680-778), batched: ONE WARP PER SEARCH.
//
// Every step of these searches evaluates a handful of candidate vectors (4 for a diamond, 3 or 6 for a hexagon, up to
// 14 predictors).  A step is one "round": the 32 lanes are split into groups of U lanes, each group owns one candidate
// and each lane of a group one 8x4 (or 4x4) unit of the block — the granularity at which the reference's own SATD is
// defined — so a 16x16 SATD candidate is spread over 8 lanes and four candidates fill the warp.  Group sums travel by
// shuffles; the selection among candidates is the reference's sequential strict-'<' scan, done uniformly by all lanes.
// Sub-pel samples follow get_ref (mc.c:181-202): one of the four half-pel planes or the byte average of two.
#include "pixel_dev.cuh"

namespace {

struct SearchCtx {
    const uint8_t *fe;            // fenc block origin (plane stride)
    const uint8_t *planes[4];     // fref full/h/v/c planes at the block origin
    int stride;
    int bw, bh, U, units;         // block size, lanes per candidate (pow2), units in the block
    const int16_t *cmx, *cmy;     // p_cost_mv - mvp
    bool fpel_satd, mbcmp_satd;
};

// cost of THIS lane's candidate (qpel mv), identical on all lanes of the group; invalid candidates return COST_MAX+1
__device__ __forceinline__ int eval_round(const SearchCtx &c, bool satd, int lane, bool valid, int mx, int my, bool add_mv_cost = true)
{
    const int u = lane & (c.U - 1);
    int v = 0;
    if (valid && u < c.units) {
        int ux, uy;
        unit_pos(c.bw, u, ux, uy);
        const QpelSrc src = qpel_src(c.planes, c.stride, mx, my);
        v = unit_cost(satd, c.bw, c.fe, c.stride, src, c.stride, ux, uy);
    }
    for (int o = c.U >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (!valid) return COST_MAX + 1;
    return add_mv_cost ? v + c.cmx[mx] + c.cmy[my] : v;
}
__device__ __forceinline__ int cand_cost(const SearchCtx &c, int cost, int k) { return __shfl_sync(0xffffffffu, cost, k * c.U); }

__constant__ int8_t c_hex2[8][2] = { { -1, -2 }, { -2, 0 }, { -1, 2 }, { 1, 2 }, { 2, 0 }, { 1, -2 }, { -1, -2 }, { -2, 0 } };
__constant__ int8_t c_mod6m1[8] = { 5, 0, 1, 2, 3, 4, 5, 0 };
__constant__ int8_t c_square1[8][2] = { { 0, -1 }, { 0, 1 }, { -1, 0 }, { 1, 0 }, { -1, -1 }, { -1, 1 }, { 1, -1 }, { 1, 1 } }; // order of me.c:303-304
__constant__ int8_t c_subpel_iters[10][4] = { { 0, 0, 0, 0 }, { 1, 1, 0, 0 }, { 0, 1, 1, 0 }, { 0, 2, 1, 0 }, { 0, 2, 1, 1 },
                                              { 0, 2, 1, 2 }, { 0, 0, 2, 2 }, { 0, 0, 2, 2 }, { 0, 0, 4, 10 }, { 0, 0, 4, 10 } };

// evaluate up to n full-pel candidates (offsets relative to (ox,oy), from a constant table) and fold them into
// (bcost,bmx,bmy) in table order with strict '<' — COST_MV_X4 / COST_MV_X3_DIR + COPYn_IF_LT of me.c:74-104
template <typename Tab>
__device__ __forceinline__ int scan_fpel(const SearchCtx &c, int lane, const Tab &tab, int first, int n, int ox, int oy, int &bcost,
                                         int &bmx, int &bmy)
{
    const int per = 32 / c.U;
    int best_k = -1;
    for (int k0 = 0; k0 < n; k0 += per) {
        const int k = k0 + lane / c.U;
        const bool valid = k < n;
        const int mx = ox + (valid ? tab[first + k][0] : 0), my = oy + (valid ? tab[first + k][1] : 0);
        const int cost = eval_round(c, c.fpel_satd, lane, valid, mx << 2, my << 2);
        for (int j = 0; j < per && k0 + j < n; j++) {
            const int cj = cand_cost(c, cost, j);
            if (cj < bcost) { bcost = cj; bmx = ox + tab[first + k0 + j][0]; bmy = oy + tab[first + k0 + j][1]; best_k = k0 + j; }
        }
    }
    return best_k;
}

__device__ void warp_search(const SearchCtx &c, const x264_cuda_me_job_t &job, int method, int me_range, int subme, int lane,
                            x264_cuda_me_final_t *out)
{
    const int x_min = job.mv_min_fpel[0], y_min = job.mv_min_fpel[1], x_max = job.mv_max_fpel[0], y_max = job.mv_max_fpel[1];
    const int per = 32 / c.U;
    int bmx, bmy, bcost, pmx, pmy;
    int bpred_mx = 0, bpred_my = 0, bpred_cost = COST_MAX;
    bmx = clip3i(job.mvp[0], x_min * 4, x_max * 4);
    bmy = clip3i(job.mvp[1], y_min * 4, y_max * 4);
    pmx = (bmx + 2) >> 2; pmy = (bmy + 2) >> 2;
    bcost = COST_MAX;
    const int n_mvc = min((int)job.i_mvc, X264_CUDA_ME_MAX_MVC);

    if (method == X264_CUDA_ME_METHOD_SEEDED) {
        bmx = job.seed_mv[0]; bmy = job.seed_mv[1]; bcost = job.seed_cost;
    } else {
        if (subme >= 3) { // me.c:189-205: predictors at quarter-pel precision
            const uint32_t bmv = ((uint32_t)bmx & 0xffff) | ((uint32_t)bmy << 16);
            const int n = 1 + n_mvc;
            for (int k0 = 0; k0 < n; k0 += per) {
                const int k = k0 + lane / c.U;
                bool valid = k < n;
                int mx = bmx, my = bmy;
                if (valid && k > 0) {
                    const int vx = job.mvc[k - 1][0], vy = job.mvc[k - 1][1];
                    const uint32_t v = ((uint32_t)vx & 0xffff) | ((uint32_t)vy << 16);
                    valid = v != 0 && v != bmv;
                    mx = clip3i(vx, x_min * 4, x_max * 4); my = clip3i(vy, y_min * 4, y_max * 4);
                }
                const int cost = eval_round(c, c.fpel_satd, lane, valid, mx, my);
                for (int j = 0; j < per && k0 + j < n; j++) {
                    const int cj = cand_cost(c, cost, j);
                    if (cj < bpred_cost) { bpred_cost = cj; bpred_mx = __shfl_sync(0xffffffffu, mx, j * c.U); bpred_my = __shfl_sync(0xffffffffu, my, j * c.U); }
                }
            }
            bmx = (bpred_mx + 2) >> 2; bmy = (bpred_my + 2) >> 2;
            const int cost = eval_round(c, c.fpel_satd, lane, true, bmx << 2, bmy << 2);
            if (cost < bcost) bcost = cost; // COST_MV(bmx,bmy) with bcost == COST_MAX
        } else { // me.c:207-227
            const int n = 1 + n_mvc;
            int cur_x = pmx, cur_y = pmy;
            for (int k0 = 0; k0 < n; k0 += per) {
                const int k = k0 + lane / c.U;
                bool valid = k < n;
                int mx = pmx, my = pmy;
                if (valid && k > 0) {
                    mx = (job.mvc[k - 1][0] + 2) >> 2; my = (job.mvc[k - 1][1] + 2) >> 2;
                    valid = (mx | my) != 0; // the "same as current best" skip (me.c:222) cannot change the outcome
                    mx = clip3i(mx, x_min, x_max); my = clip3i(my, y_min, y_max);
                }
                const int cost = eval_round(c, c.fpel_satd, lane, valid, mx << 2, my << 2, !(k == 0));
                for (int j = 0; j < per && k0 + j < n; j++) {
                    const int cj = cand_cost(c, cost, j);
                    if (cj < bcost) { bcost = cj; cur_x = __shfl_sync(0xffffffffu, mx, j * c.U); cur_y = __shfl_sync(0xffffffffu, my, j * c.U); }
                }
            }
            bmx = cur_x; bmy = cur_y;
        }
        { // COST_MV(0,0), me.c:229
            const int cost = eval_round(c, c.fpel_satd, lane, true, 0, 0);
            if (cost < bcost) { bcost = cost; bmx = 0; bmy = 0; }
        }
    }

#define IN_RANGE(mx_, my_) ((mx_) >= x_min && (mx_) <= x_max && (my_) >= y_min && (my_) <= y_max)
    if (method == X264_CUDA_ME_METHOD_DIA) { // me.c:233-244
        int i = 0;
        do {
            const int ox = bmx, oy = bmy;
            scan_fpel(c, lane, c_square1, 0, 4, ox, oy, bcost, bmx, bmy);
            if (bmx == ox && bmy == oy) break;
            if (!IN_RANGE(bmx, bmy)) break;
        } while (++i < me_range);
    } else if (method == X264_CUDA_ME_METHOD_HEX) { // me.c:266-304
        int ox = bmx, oy = bmy;
        int dir = scan_fpel(c, lane, c_hex2, 1, 6, ox, oy, bcost, ox, oy); // hex2[1..6] == (-2,0),(-1,2),(1,2),(2,0),(1,-2),(-1,-2)
        if (dir >= 0) {
            bmx += c_hex2[dir + 1][0]; bmy += c_hex2[dir + 1][1];
            for (int i = 1; i < me_range / 2 && IN_RANGE(bmx, bmy); i++) {
                const int odir = c_mod6m1[dir + 1];
                ox = bmx; oy = bmy;
                const int k = scan_fpel(c, lane, c_hex2, odir, 3, bmx, bmy, bcost, ox, oy);
                if (k < 0) break;
                dir = odir - 1 + k;
                bmx += c_hex2[dir + 1][0]; bmy += c_hex2[dir + 1][1];
            }
        }
        ox = bmx; oy = bmy; // square refine, me.c:301-304
        scan_fpel(c, lane, c_square1, 0, 8, ox, oy, bcost, bmx, bmy);
    }
    const int fbmx = bmx, fbmy = bmy;

    // ---- "-> qpel mv", me.c:603-620
    int mvx, mvy, cost;
    if (bpred_cost < bcost) { mvx = bpred_mx; mvy = bpred_my; cost = bpred_cost; }
    else { mvx = bmx << 2; mvy = bmy << 2; cost = bcost; }
    int cost_mv = c.cmx[mvx] + c.cmy[mvy];
    if (bmx == pmx && bmy == pmy && subme < 3) cost += cost_mv;

    // ---- refine_subpel(h, m, hpel, qpel, NULL, 0), me.c:680-778 (b_chroma_me == 0)
    if (subme >= 2) {
        const int hpel_iters = c_subpel_iters[subme][2], qpel_iters = c_subpel_iters[subme][3];
        int sx = mvx, sy = mvy, sc = cost;
        const int spel_ymax = job.mv_max_spel[1];
        if (hpel_iters && subme < 3) {
            const int mx = clip3i(job.mvp[0], job.mv_min_spel[0], job.mv_max_spel[0]);
            const int my = clip3i(job.mvp[1], job.mv_min_spel[1], job.mv_max_spel[1]);
            if ((mx - sx) | (my - sy)) {
                const int v = eval_round(c, c.fpel_satd, lane, true, mx, my);
                if (v < sc) { sc = v; sx = mx; sy = my; }
            }
        }
        for (int i = hpel_iters; i > 0; i--) { // half-pel diamond with fpelcmp, me.c:708-727
            const int ox = sx, oy = sy;
            const int k = lane / c.U; // 4 candidates; when U == 8 they fill the warp, otherwise extra groups idle
            const int dx = k == 2 ? -2 : k == 3 ? 2 : 0, dy = k == 0 ? -2 : k == 1 ? 2 : 0;
            const int v = eval_round(c, c.fpel_satd, lane, k < 4, ox + dx, oy + dy);
            const int v0 = cand_cost(c, v, 0), v1 = cand_cost(c, v, 1), v2 = cand_cost(c, v, 2), v3 = cand_cost(c, v, 3);
            if (v0 < sc) { sc = v0; sy = oy - 2; }
            if (v1 < sc) { sc = v1; sy = oy + 2; }
            if (v2 < sc) { sc = v2; sx = ox - 2; sy = oy; }
            if (v3 < sc) { sc = v3; sx = ox + 2; sy = oy; }
            if (sx == ox && sy == oy) break;
        }
        if (sy > spel_ymax) sy = spel_ymax; // !b_refine_qpel, me.c:729-736
        sc = eval_round(c, c.mbcmp_satd, lane, true, sx, sy);
        int bdir = -1;
        for (int i = qpel_iters; i > 0; i--) { // quarter-pel diamond with mbcmp, me.c:755-767
            const int odir = bdir, ox = sx, oy = sy;
            const int k = lane / c.U;
            const int dx = k == 2 ? -1 : k == 3 ? 1 : 0, dy = k == 0 ? -1 : k == 1 ? 1 : 0;
            const bool valid = k < 4 && (k ^ 1) != odir;
            const int v = eval_round(c, c.mbcmp_satd, lane, valid, ox + dx, oy + dy);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int vj = cand_cost(c, v, j);
                if ((j ^ 1) != odir && vj < sc) {
                    sc = vj; bdir = j;
                    sx = ox + (j == 2 ? -1 : j == 3 ? 1 : 0); sy = oy + (j == 0 ? -1 : j == 1 ? 1 : 0);
                }
            }
            if (sx == ox && sy == oy) break;
        }
        if (sy > spel_ymax) { sy = spel_ymax; sc = eval_round(c, c.mbcmp_satd, lane, true, sx, sy); } // me.c:770-775
        mvx = sx; mvy = sy; cost = sc;
        cost_mv = c.cmx[mvx] + c.cmy[mvy];
    } else if (mvy > job.mv_max_spel[1]) {
        mvy = job.mv_max_spel[1]; // me.c:629-630
    }
    if (lane == 0) {
        x264_cuda_me_final_t r;
        r.mv[0] = (int16_t)mvx; r.mv[1] = (int16_t)mvy; r.cost = cost; r.cost_mv = cost_mv; r.bmx = (int16_t)fbmx; r.bmy = (int16_t)fbmy;
        *out = r;
    }
}

struct Planes { const uint8_t *fenc; const uint8_t *ref[4]; int stride; };

__global__ void __launch_bounds__(128) me_small_kernel(Planes pl, const x264_cuda_me_job_t *__restrict__ jobs, int n_jobs,
                                                       const int16_t *const *__restrict__ cost_tabs, int method, int me_range, int subme,
                                                       x264_cuda_me_final_t *__restrict__ results)
{
    __shared__ x264_cuda_me_job_t s_job[4];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (j >= n_jobs) return;
    if (lane < (int)(sizeof(x264_cuda_me_job_t) / 4)) ((uint32_t *)&s_job[wid])[lane] = __ldg((const uint32_t *)(jobs + j) + lane);
    __syncwarp();
    const x264_cuda_me_job_t &job = s_job[wid];
    const int16_t *tab = cost_tabs[job.qp > 51 ? 51 : job.qp] + 2 * 4 * 2048;
    // reject jobs whose cost-table indices could leave +-2*4*2048 or whose limits are inconsistent
    {
        const int lim = 2 * 4 * 2048 - 16;
        const int ax = max(abs(4 * job.mv_min_fpel[0] - job.mvp[0]), abs(4 * job.mv_max_fpel[0] - job.mvp[0])) + 40;
        const int ay = max(abs(4 * job.mv_min_fpel[1] - job.mvp[1]), abs(4 * job.mv_max_fpel[1] - job.mvp[1])) + 40;
        if (ax > lim || ay > lim || job.mv_min_fpel[0] > 0 || job.mv_max_fpel[0] < 0 || job.mv_min_fpel[1] > 0 || job.mv_max_fpel[1] < 0) {
            if (lane == 0) { x264_cuda_me_final_t r = { { 0, 0 }, -1, -1, 0, 0 }; results[j] = r; }
            return;
        }
    }
    const int ip = min((int)job.i_pixel, 6);
    SearchCtx c;
    c.stride = pl.stride;
    c.bw = ip <= X264_CUDA_PIXEL_16x8 ? 16 : ip <= X264_CUDA_PIXEL_8x4 ? 8 : 4;
    c.bh = (ip == X264_CUDA_PIXEL_16x16 || ip == X264_CUDA_PIXEL_8x16) ? 16 : (ip == X264_CUDA_PIXEL_8x4 || ip == X264_CUDA_PIXEL_4x4) ? 4 : 8;
    c.units = unit_count(c.bw, c.bh);
    c.U = c.units; // 8,4,4,2,1,2,1: already powers of two
    const size_t off = (size_t)job.by * pl.stride + job.bx;
    c.fe = pl.fenc + off;
#pragma unroll
    for (int k = 0; k < 4; k++) c.planes[k] = pl.ref[k] ? pl.ref[k] + off : pl.ref[0] + off;
    c.cmx = tab - job.mvp[0]; c.cmy = tab - job.mvp[1];
    c.fpel_satd = job.flags & X264_CUDA_ME_FPEL_SATD; c.mbcmp_satd = job.flags & X264_CUDA_ME_MBCMP_SATD;
    warp_search(c, job, method, me_range, subme, lane, results + j);
}

// ---- function-level block metrics over packed 16x16 tiles: one thread per block
__global__ void __launch_bounds__(128) block_cmp_kernel(int metric, int bw, int bh, int n, const uint8_t *__restrict__ p1,
                                                        const uint8_t *__restrict__ p2, int *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t *a = p1 + (size_t)i * 256, *b = p2 + (size_t)i * 256;
    int sum = 0;
    if (metric == 0 || metric == 1) {
        for (int y = 0; y < bh; y++)
            for (int x = 0; x < bw; x += 4) {
                const uint32_t wa = *(const uint32_t *)(a + y * 16 + x), wb = *(const uint32_t *)(b + y * 16 + x);
                if (metric == 0) sum = (int)sad4_acc(wa, wb, (uint32_t)sum);
                else
                    for (int k = 0; k < 4; k++) { const int d = (int)((wa >> (8 * k)) & 255) - (int)((wb >> (8 * k)) & 255); sum += d * d; }
            }
    } else if (metric == 2) {
        const uint8_t *const pl[4] = { b, b, b, b };
        const QpelSrc src = qpel_src(pl, 16, 0, 0);
        for (int u = 0; u < unit_count(bw, bh); u++) {
            int ux, uy;
            unit_pos(bw, u, ux, uy);
            sum += unit_cost(true, bw, a, 16, src, 16, ux, uy);
        }
    } else { // SA8D: (sum of raw 8x8 sums + 2) >> 2, pixel.c:290-303
        for (int y0 = 0; y0 < bh; y0 += 8)
            for (int x0 = 0; x0 < bw; x0 += 8) {
                uint2 f[8], r[8];
                for (int y = 0; y < 8; y++) {
                    f[y] = *(const uint2 *)(a + (y0 + y) * 16 + x0);
                    r[y] = *(const uint2 *)(b + (y0 + y) * 16 + x0);
                }
                sum += sa8d_8x8_rows(f, r);
            }
        sum = (sum + 2) >> 2;
    }
    out[i] = sum;
}

} // namespace

extern "C" int x264_cuda_me_search_small_dev(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, const x264_cuda_frame_t *fref, int method,
                                             int me_range, int subme, const void *d_jobs, int n_jobs, void *d_results)
{
    if (n_jobs <= 0) return 0;
    if (fenc->g.stride != fref->g.stride || fenc->g.lines != fref->g.lines) {
        snprintf(ctx->err, 256, "x264_cuda_me_search_small: fenc/fref geometry mismatch");
        return -1;
    }
    if (subme < 0 || subme > 9 || me_range < 1 || (method != X264_CUDA_ME_METHOD_DIA && method != X264_CUDA_ME_METHOD_HEX &&
                                                   method != X264_CUDA_ME_METHOD_SEEDED)) {
        snprintf(ctx->err, 256, "x264_cuda_me_search_small: bad method %d / subme %d / me_range %d", method, subme, me_range);
        return -1;
    }
    if (subme >= 2 && !(fref->g.flags & X264_CUDA_FRAME_HPEL)) {
        snprintf(ctx->err, 256, "x264_cuda_me_search_small: subme %d needs the half-pel planes (X264_CUDA_FRAME_HPEL)", subme);
        return -1;
    }
    const int16_t *const *d_tabs;
    if (x264_cuda_cost_tables(ctx, &d_tabs)) return -1;
    Planes pl = { fenc->plane[0], { fref->plane[0], fref->plane[1], fref->plane[2], fref->plane[3] }, fenc->g.stride };
    me_small_kernel<<<(n_jobs + 3) / 4, 128, 0, ctx->stream>>>(pl, (const x264_cuda_me_job_t *)d_jobs, n_jobs, d_tabs, method, me_range, subme,
                                                                (x264_cuda_me_final_t *)d_results);
    LAUNCH_CHECK(ctx, "me_small_kernel");
    return 0;
}

extern "C" int x264_cuda_me_search_small(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, const x264_cuda_frame_t *fref, int method,
                                         int me_range, int subme, const x264_cuda_me_job_t *jobs, int n_jobs, x264_cuda_me_final_t *results)
{
    if (n_jobs <= 0) return 0;
    const size_t jb = (size_t)n_jobs * sizeof(x264_cuda_me_job_t), rb = (size_t)n_jobs * sizeof(x264_cuda_me_final_t);
    const size_t jb_al = (jb + 255) & ~(size_t)255;
    if (x264_cuda_stage(ctx, jb_al + rb, jb_al + rb)) return -1;
    uint8_t *hs = (uint8_t *)ctx->h_stage, *ds = (uint8_t *)ctx->d_stage;
    memcpy(hs, jobs, jb);
    CUDA_TRY(ctx, cudaMemcpyAsync(ds, hs, jb, cudaMemcpyHostToDevice, ctx->stream));
    if (x264_cuda_me_search_small_dev(ctx, fenc, fref, method, me_range, subme, ds, n_jobs, ds + jb_al)) return -1;
    CUDA_TRY(ctx, cudaMemcpyAsync(hs + jb_al, ds + jb_al, rb, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(results, hs + jb_al, rb);
    return 0;
}

extern "C" int x264_cuda_block_cmp(x264_cuda_t *ctx, int metric, int i_pixel, int n, const uint8_t *pix1, const uint8_t *pix2, int *out)
{
    if (n <= 0) return 0;
    static const int bws[7] = { 16, 16, 8, 8, 8, 4, 4 }, bhs[7] = { 16, 8, 16, 8, 4, 8, 4 };
    if (metric < 0 || metric > 3 || i_pixel < 0 || i_pixel > 6 || (metric == 3 && i_pixel != 0 && i_pixel != 3)) {
        snprintf(ctx->err, 256, "x264_cuda_block_cmp: no such table entry (metric %d, i_pixel %d)", metric, i_pixel); // NULL in the C table too
        return -1;
    }
    const size_t tb = (size_t)n * 256, ob = (size_t)n * sizeof(int);
    if (x264_cuda_stage(ctx, 2 * tb + ob, 2 * tb + ob)) return -1;
    uint8_t *hs = (uint8_t *)ctx->h_stage, *ds = (uint8_t *)ctx->d_stage;
    memcpy(hs, pix1, tb); memcpy(hs + tb, pix2, tb);
    CUDA_TRY(ctx, cudaMemcpyAsync(ds, hs, 2 * tb, cudaMemcpyHostToDevice, ctx->stream));
    block_cmp_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(metric, bws[i_pixel], bhs[i_pixel], n, ds, ds + tb, (int *)(ds + 2 * tb));
    LAUNCH_CHECK(ctx, "block_cmp_kernel");
    CUDA_TRY(ctx, cudaMemcpyAsync(hs + 2 * tb, ds + 2 * tb, ob, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(out, hs + 2 * tb, ob);
    return 0;
}
