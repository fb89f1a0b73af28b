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
    // TESA only
    const uint16_t *sums8, *sums4; // integral planes at the block origin (8x8 and 4x4 box sums)
    int lines_pad;                 // unused
    int2 *list; int list_cap;      // candidate list scratch: (sad, mx | my << 16)
    int i_pixel;
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

// candidate-list accessors: the list is written and read by DIFFERENT lanes of the warp between __syncwarp()s; volatile
// keeps every access a real memory operation (no register caching across the barriers)
__device__ __forceinline__ int2 list_get(const int2 *l, int i) { const volatile int *p = (const volatile int *)(l + i); return make_int2(p[0], p[1]); }
__device__ __forceinline__ void list_put(int2 *l, int i, int2 v) { volatile int *p = (volatile int *)(l + i); p[0] = v.x; p[1] = v.y; }

// ---- TESA candidate selection, me.c:491-578: ADS threshold, SAD threshold, keep the best few, fpelcmp on those.
// Kept __noinline__ on purpose: when inlined into warp_search, ptxas (12.9, -O1 and up) produced code that dropped the
// second ADS term for 4x8 blocks (caught by tests/test_gpu_me_small.py::test_tesa; -Xptxas -O0 and the out-of-line
// form are both correct).
__device__ __noinline__ void tesa_search(const SearchCtx &c, const x264_cuda_me_job_t &job, int me_range, int lane, int &bcost, int &bmx, int &bmy)
{
    const int x_min = job.mv_min_fpel[0], y_min = job.mv_min_fpel[1], x_max = job.mv_max_fpel[0], y_max = job.mv_max_fpel[1];
    const int min_x = max(bmx - me_range, x_min), min_y = max(bmy - me_range, y_min);
    const int max_x = min(bmx + me_range, x_max), max_y = min(bmy + me_range, y_max);
    const int width = (max_x - min_x + 3) & ~3;
    const int stride = c.stride, ip = c.i_pixel;
    // enc_dc: pixel sums of the fenc sub-blocks (sad_x4 against zero, me.c:481-489)
    const int small = ip > X264_CUDA_PIXEL_8x8;         // 4x4 sums
    const int d = small ? 4 : 8;
    const int terms = ip == X264_CUDA_PIXEL_16x16 ? 4 : (ip == X264_CUDA_PIXEL_8x8 || ip == X264_CUDA_PIXEL_4x4) ? 1 : 2;
    const bool vertical2 = ip == X264_CUDA_PIXEL_8x16 || ip == X264_CUDA_PIXEL_4x8; // second term below, not beside
    int dc0, dc1, dc2, dc3;
    {
        // lane l < 4 sums sub-block l (only the ones inside the block are ever used)
        int v = 0;
        const int sx = (lane & 1) * d, sy = ((lane >> 1) & 1) * d;
        if (lane < 4 && sx < c.bw && sy < c.bh)
            for (int y = 0; y < d; y++)
                for (int x = 0; x < d; x += 4) v = (int)sad4_acc(__ldg((const uint32_t *)(c.fe + (size_t)(sy + y) * stride + sx + x)), 0u, (uint32_t)v);
        __syncwarp();
        dc0 = __shfl_sync(0xffffffffu, v, 0);
        const int v1 = __shfl_sync(0xffffffffu, v, 1);
        dc2 = __shfl_sync(0xffffffffu, v, 2);
        dc3 = __shfl_sync(0xffffffffu, v, 3);
        dc1 = vertical2 ? dc2 : v1; // me.c:488-489
    }
    const uint16_t *sums_base = small ? c.sums4 : c.sums8;
    const int delta = (ip == X264_CUDA_PIXEL_16x16 || vertical2) ? d * stride : d;
    const int n_extra = terms - 1;                                  // 0, 1 or 3 box sums besides the first
    const int off1 = terms == 4 ? 8 : delta, off2 = delta, off3 = delta + 8;
    const int sad_thresh = me_range <= 16 ? 10 : me_range <= 24 ? 11 : 12;
    const uint8_t *ref0 = c.planes[0];
    int nmv = 0;
    // bsad = plain SAD at the seed + its MV bits (me.c:497-498)
    int bsad = eval_round(c, false, lane, true, bmx << 2, bmy << 2);
    for (int my = min_y; my <= max_y; my++) {
        const int ycost = c.cmy[my << 2];
        if (bsad <= ycost) continue;
        bsad -= ycost;
        const int thresh = bsad * 17 / 16;
        for (int c0 = 0; c0 < width; c0 += 32) {
            const int i = c0 + lane;
            bool surv = false;
            if (i < width) {
                const uint16_t *sp = sums_base + (ptrdiff_t)my * stride + min_x + i;
                // pixf.ads[i_pixel]: ads4 / ads2 / ads1 (pixel.c:515-559); offsets of the extra box sums precomputed per job
                const int s0 = __ldg(sp), s1 = __ldg(sp + off1), s2 = __ldg(sp + off2), s3 = __ldg(sp + off3);
                int ads = max(dc0 - s0, s0 - dc0) + (int)(uint16_t)c.cmx[(min_x + i) << 2];
                if (n_extra >= 1) ads += max(dc1 - s1, s1 - dc1);
                if (n_extra == 3) ads += max(dc2 - s2, s2 - dc2) + max(dc3 - s3, s3 - dc3);
                surv = ads < thresh;
            }
            if (!__any_sync(0xffffffffu, surv)) continue;
            // plain SAD of every survivor (one lane each), + the x cost indexed WITHOUT min_x as the reference does
            // (me.c:518,531: cost_fpel_mvx[xs[i]])
            int sad = 0x3fffffff;
            if (surv) {
                const uint8_t *a = ref0 + (ptrdiff_t)my * stride + min_x + i;
                uint32_t acc = 0;
                for (int y = 0; y < c.bh; y++)
                    for (int x = 0; x < c.bw; x += 4)
                        acc = sad4_acc(__ldg((const uint32_t *)(c.fe + (size_t)y * stride + x)), ldg4(a + (size_t)y * stride + x), acc);
                sad = (int)acc + (int)(uint16_t)c.cmx[i << 2];
            }

            // bsad seen by lane j = min(bsad, sads of the survivors before it): exclusive prefix-min over the lanes
            int pm = sad;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, pm, o);
                if (lane >= o) pm = min(pm, t);
            }
            int excl = __shfl_up_sync(0xffffffffu, pm, 1);
            excl = lane == 0 ? bsad : min(excl, bsad);
            const bool acc_ = surv && sad < ((excl * sad_thresh) >> 3);
            const unsigned m = __ballot_sync(0xffffffffu, acc_);
            if (acc_) {
                const int pos = nmv + __popc(m & ((1u << lane) - 1));
                if (pos < c.list_cap) list_put(c.list, pos, make_int2(sad + ycost, ((min_x + i) & 0xffff) | (my << 16)));
            }
            nmv += __popc(m);
            bsad = min(bsad, __shfl_sync(0xffffffffu, pm, 31));
        }
        bsad += ycost;
    }
    __threadfence_block();
    __syncwarp();
    nmv = min(nmv, c.list_cap);
    const int limit = me_range / 2;
    if (nmv > limit * 2) { // stable filter by sad <= bsad*(sad_thresh+8)>>4, me.c:543-558
        const int cut = bsad * (sad_thresh + 8) >> 4;
        int w = 0;
        for (int b0 = 0; b0 < nmv; b0 += 32) {
            const int j = b0 + lane;
            int2 e = make_int2(0, 0);
            bool keep = false;
            if (j < nmv) { e = list_get(c.list, j); keep = e.x <= cut; }
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            __syncwarp();
            if (keep) list_put(c.list, w + __popc(m & ((1u << lane) - 1)), e);
            w += __popc(m);
            __syncwarp();
        }
        nmv = w;
    }
    if (nmv > limit) { // partial selection sort, first index wins ties, me.c:559-576
        for (int i = 0; i < limit; i++) {
            unsigned long long best = ~0ull;
            for (int j = i + lane; j < nmv; j += 32) best = min(best, ((unsigned long long)(unsigned)list_get(c.list, j).x << 32) | (unsigned)j);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
            const int bj = (int)(best & 0xffffffffu);
            if (lane == 0 && bj > i) { const int2 t = list_get(c.list, i); list_put(c.list, i, list_get(c.list, bj)); list_put(c.list, bj, t); }
            __threadfence_block();
            __syncwarp();
        }
        nmv = limit;
    }
    // COST_MV with fpelcmp on the keepers, in list order (me.c:577-578)
    const int per = 32 / c.U;
    for (int k0 = 0; k0 < nmv; k0 += per) {
        const int k = k0 + lane / c.U;
        const bool valid = k < nmv;
        int mx = 0, my = 0;
        if (valid) { const int2 e = list_get(c.list, k); mx = (int)(int16_t)(e.y & 0xffff); my = e.y >> 16; }
        const int cost = eval_round(c, c.fpel_satd, lane, valid, mx << 2, my << 2);
        for (int j = 0; j < per && k0 + j < nmv; j++) {
            const int cj = cand_cost(c, cost, j);
            if (cj < bcost) { bcost = cj; bmx = __shfl_sync(0xffffffffu, mx, j * c.U); bmy = __shfl_sync(0xffffffffu, my, j * c.U); }
        }
    }
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
    else if (method == X264_CUDA_ME_METHOD_TESA)
        tesa_search(c, job, me_range, lane, bcost, bmx, bmy);
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

struct Planes { const uint8_t *fenc; const uint8_t *ref[4]; int stride; const uint16_t *sums8, *sums4; int2 *scratch; int list_cap; };

__global__ void __launch_bounds__(128) me_small_kernel(Planes pl, const x264_cuda_me_job_t *__restrict__ jobs, int n_jobs,
                                                       const int16_t *const *__restrict__ cost_tabs, int method, int me_range, int subme,
                                                       x264_cuda_me_final_t *__restrict__ results)
{
    __shared__ x264_cuda_me_job_t s_job[4];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int warps_per_grid = (gridDim.x * blockDim.x) >> 5, gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  for (int j = gw; j < n_jobs; j += warps_per_grid) {
    __syncwarp();
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
            continue;
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
    c.i_pixel = ip;
    c.sums8 = pl.sums8 ? pl.sums8 + off : nullptr; c.sums4 = pl.sums4 ? pl.sums4 + off : nullptr;
    c.list = pl.scratch ? pl.scratch + (size_t)gw * pl.list_cap : nullptr; c.list_cap = pl.list_cap; c.lines_pad = 0;
    if (method == X264_CUDA_ME_METHOD_TESA && (!c.list || !(ip > X264_CUDA_PIXEL_8x8 ? c.sums4 : c.sums8))) {
        if (lane == 0) { x264_cuda_me_final_t r = { { 0, 0 }, -1, -1, 0, 0 }; results[j] = r; } // no integral plane for this block size
        continue;
    }
    warp_search(c, job, method, me_range, subme, lane, results + j);
  }
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
    if (subme < 0 || subme > 9 || me_range < 1 || me_range > 64 ||
        (method != X264_CUDA_ME_METHOD_DIA && method != X264_CUDA_ME_METHOD_HEX && method != X264_CUDA_ME_METHOD_SEEDED &&
         method != X264_CUDA_ME_METHOD_TESA)) {
        snprintf(ctx->err, 256, "x264_cuda_me_search_small: bad method %d / subme %d / me_range %d", method, subme, me_range);
        return -1;
    }
    if (subme >= 2 && !(fref->g.flags & X264_CUDA_FRAME_HPEL)) {
        snprintf(ctx->err, 256, "x264_cuda_me_search_small: subme %d needs the half-pel planes (X264_CUDA_FRAME_HPEL)", subme);
        return -1;
    }
    const int16_t *const *d_tabs;
    if (x264_cuda_cost_tables(ctx, &d_tabs)) return -1;
    Planes pl = { fenc->plane[0], { fref->plane[0], fref->plane[1], fref->plane[2], fref->plane[3] }, fenc->g.stride, nullptr, nullptr, nullptr, 0 };
    int blocks = (n_jobs + 3) / 4;
    if (method == X264_CUDA_ME_METHOD_TESA) {
        if (!fref->integral) {
            snprintf(ctx->err, 256, "x264_cuda_me_search_small: TESA needs the integral image (X264_CUDA_FRAME_INTEGRAL + x264_cuda_frame_filter)");
            return -1;
        }
        pl.sums8 = fref->integral;
        pl.sums4 = (fref->g.flags & X264_CUDA_FRAME_INTEGRAL4) ? fref->integral + fref->plane_size : nullptr;
        // candidate list scratch: one (2R+1) x (2R+4) list per resident warp (persistent grid-stride warps)
        pl.list_cap = (2 * me_range + 1) * ((2 * me_range + 4) & ~3);
        blocks = min(blocks, ctx->sm_count * 8);
        const size_t need = (size_t)blocks * 4 * pl.list_cap * sizeof(int2);
        if (need > ctx->d_scratch_size) {
            CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
            cudaFree(ctx->d_scratch); ctx->d_scratch = nullptr; ctx->d_scratch_size = 0;
            CUDA_TRY(ctx, cudaMalloc(&ctx->d_scratch, need));
            ctx->d_scratch_size = need;
        }
        pl.scratch = (int2 *)ctx->d_scratch;
    }
    me_small_kernel<<<blocks, 128, 0, ctx->stream>>>(pl, (const x264_cuda_me_job_t *)d_jobs, n_jobs, d_tabs, method, me_range, subme,
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
