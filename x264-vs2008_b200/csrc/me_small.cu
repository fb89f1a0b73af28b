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
#include <atomic>

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
    // b_chroma_me (me.c:686): chroma planes of fenc / fref at the block's chroma origin
    const uint8_t *fe_c[2], *ref_c[2];
    int stride_c;
    bool chroma_me;
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

// The chroma term of COST_MV_SATD (me.c:655-677): mc_chroma (mc.c:205-236) of the bw/2 x bh/2 blocks of U and V at the luma
// qpel mv, compared with mbcmp[i_pixel+3].  Chroma units (8x4 for 16-wide partitions, 4x4 for 8-wide ones) are spread over
// the lanes of the candidate's group like the luma units; there are never more of them than luma units.  The reference
// adds U and V only while the candidate still beats bcost — the decision is the same as with the full sum (costs are >= 0).
__device__ __forceinline__ int chroma_unit_cost(const SearchCtx &c, bool satd, int u, int mx, int my)
{
    const bool wide = c.bw == 16;                 // chroma block 8 wide, else 4
    const int per_plane = c.bh >> 3;              // units per plane: bh/2 rows in units of 4 rows
    if (u >= 2 * per_plane) return 0;
    const int pl = u / per_plane, cy0 = (u - pl * per_plane) * 4;
    const int d8x = mx & 7, d8y = my & 7;
    const int cA = (8 - d8x) * (8 - d8y), cB = d8x * (8 - d8y), cC = (8 - d8x) * d8y, cD = d8x * d8y;
    const uint8_t *s = c.ref_c[pl] + (ptrdiff_t)((my >> 3) + cy0) * c.stride_c + (mx >> 3);
    const uint8_t *f = c.fe_c[pl] + (ptrdiff_t)cy0 * c.stride_c;
    const int w = wide ? 8 : 4;
    uint32_t pr[4][2] = { { 0, 0 }, { 0, 0 }, { 0, 0 }, { 0, 0 } };
    int top[9];
#pragma unroll
    for (int x = 0; x < 9; x++) top[x] = x <= w ? s[x] : 0;
#pragma unroll
    for (int y = 0; y < 4; y++) {
        int bot[9];
#pragma unroll
        for (int x = 0; x < 9; x++) bot[x] = x <= w ? s[(ptrdiff_t)(y + 1) * c.stride_c + x] : 0;
#pragma unroll
        for (int x = 0; x < 8; x++)
            if (x < w) pr[y][x >> 2] |= (uint32_t)((cA * top[x] + cB * top[x + 1] + cC * bot[x] + cD * bot[x + 1] + 32) >> 6) << (8 * (x & 3));
#pragma unroll
        for (int x = 0; x < 9; x++) top[x] = bot[x];
    }
    if (wide) {
        uint2 fr[4], rr[4];
#pragma unroll
        for (int y = 0; y < 4; y++) { fr[y] = ldg8(f + (ptrdiff_t)y * c.stride_c); rr[y] = make_uint2(pr[y][0], pr[y][1]); }
        if (satd) return satd_8x4_rows(fr, rr);
        uint32_t acc = 0;
#pragma unroll
        for (int y = 0; y < 4; y++) { acc = sad4_acc(fr[y].x, rr[y].x, acc); acc = sad4_acc(fr[y].y, rr[y].y, acc); }
        return (int)acc;
    }
    uint32_t fr[4], rr[4];
#pragma unroll
    for (int y = 0; y < 4; y++) { fr[y] = ldg4(f + (ptrdiff_t)y * c.stride_c); rr[y] = pr[y][0]; }
    if (satd) return satd_4x4_rows(fr, rr);
    uint32_t acc = 0;
#pragma unroll
    for (int y = 0; y < 4; y++) acc = sad4_acc(fr[y], rr[y], acc);
    return (int)acc;
}

// COST_MV_SATD: mbcmp of the luma block (+ mv cost) + the chroma term when b_chroma_me
__device__ __forceinline__ int eval_round_satd(const SearchCtx &c, int lane, bool valid, int mx, int my)
{
    int v = eval_round(c, c.mbcmp_satd, lane, valid, mx, my);
    if (c.chroma_me) { // warp-uniform
        int cv = valid ? chroma_unit_cost(c, c.mbcmp_satd, lane & (c.U - 1), mx, my) : 0;
        for (int o = c.U >> 1; o > 0; o >>= 1) cv += __shfl_xor_sync(0xffffffffu, cv, o);
        if (valid) v += cv;
    }
    return v;
}

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

// the same for candidates produced by a generator: gen(k, mx, my) -> "candidate k exists", in list order
template <typename Gen>
__device__ __forceinline__ void scan_gen(const SearchCtx &c, int lane, int n, Gen gen, int &bcost, int &bmx, int &bmy)
{
    const int per = 32 / c.U;
    for (int k0 = 0; k0 < n; k0 += per) {
        const int k = k0 + lane / c.U;
        int mx = 0, my = 0;
        const bool valid = k < n && gen(k, mx, my);
        const int cost = eval_round(c, c.fpel_satd, lane, valid, mx << 2, my << 2); // COST_MAX + 1 for absent candidates
        for (int j = 0; j < per && k0 + j < n; j++) {
            const int cj = cand_cost(c, cost, j);
            const int jx = __shfl_sync(0xffffffffu, mx, j * c.U), jy = __shfl_sync(0xffffffffu, my, j * c.U);
            if (cj < bcost) { bcost = cj; bmx = jx; bmy = jy; }
        }
    }
}

__constant__ int8_t c_umh_oct[8][2] = { { 0, -2 }, { -1, -1 }, { 1, -1 }, { -2, 0 }, { 2, 0 }, { -1, 1 }, { 1, 1 }, { 0, 2 } };     // me.c:334-335
__constant__ int8_t c_umh_star[8][2] = { { -1, -2 }, { 1, -2 }, { -2, -1 }, { 2, -1 }, { -2, 1 }, { 2, 1 }, { -1, 2 }, { 1, 2 } };  // me.c:342-343
__constant__ int8_t c_umh_diag2[4][2] = { { -2, -2 }, { -2, 2 }, { 2, -2 }, { 2, 2 } };                                              // me.c:405
__constant__ int8_t c_umh_hex4[16][2] = { { -4, 2 }, { -4, 1 }, { -4, 0 }, { -4, -1 }, { -4, -2 }, { 4, -2 }, { 4, -1 }, { 4, 0 }, { 4, 1 }, { 4, 2 },
                                          { 2, 3 }, { 0, 4 }, { -2, 3 }, { -2, -3 }, { 0, -4 }, { 2, -3 } };                       // me.c:412-417
__constant__ int8_t c_umh_range_mul[4][4] = { { 3, 3, 4, 4 }, { 3, 4, 4, 4 }, { 4, 4, 4, 5 }, { 4, 4, 5, 6 } };                    // me.c:360-366
__constant__ int8_t c_pixel_size_shift[7] = { 0, 1, 1, 2, 3, 3, 4 };                                                               // me.c:311

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

    const bool refine_only = method == X264_CUDA_ME_METHOD_REFINE_QPEL; // x264_me_refine_qpel: seed_mv is m->mv (quarter-pel), seed_cost m->cost
    if (method == X264_CUDA_ME_METHOD_SEEDED || refine_only) {
        bmx = job.seed_mv[0]; bmy = job.seed_mv[1]; bcost = job.seed_cost;
    } else {
        bool zero_done = false;
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
            // COST_MV(bmx,bmy) with bcost == COST_MAX, then COST_MV(0,0) (me.c:229): two candidates of ONE round, folded in that order
            const int kk = lane / c.U;
            const int cost = eval_round(c, c.fpel_satd, lane, kk < 2, kk == 0 ? bmx << 2 : 0, kk == 0 ? bmy << 2 : 0);
            const int c0 = cand_cost(c, cost, 0), c1 = cand_cost(c, cost, 1);
            if (c0 < bcost) bcost = c0;
            if (c1 < bcost) { bcost = c1; bmx = 0; bmy = 0; }
            zero_done = true;
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
        if (!zero_done) { // COST_MV(0,0), me.c:229
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
    } else if (method == X264_CUDA_ME_METHOD_HEX || method == X264_CUDA_ME_METHOD_UMH) {
        int range = me_range;
        bool hex2 = true;
        if (method == X264_CUDA_ME_METHOD_UMH) { // me.c:306-447: uneven-cross multi-hexagon-grid search, then me_hex2
            const int shift = c_pixel_size_shift[c.i_pixel];
#define SAD_THRESH(v) (bcost < ((v) >> shift))
            int omx = pmx, omy = pmy, cross_start = 1;
            // CROSS (me.c:128-154): +-i along x for i = start, start+2, .. < xm, then along y.  The reference's unchecked
            // COST_MV_X4 form only runs when every candidate is inside the limits, so one range-checked list is the same thing.
            auto cross = [&](int start, int xm, int ym) {
                const int ox = omx, oy = omy;
                const int nx = xm > start ? (xm - start + 1) / 2 * 2 : 0, ny = ym > start ? (ym - start + 1) / 2 * 2 : 0;
                scan_gen(c, lane, nx, [&](int k, int &mx, int &my) { const int i = start + (k >> 1) * 2; mx = ox + ((k & 1) ? -i : i); my = oy;
                                                                     return (k & 1) ? mx >= x_min : mx <= x_max; }, bcost, bmx, bmy);
                scan_gen(c, lane, ny, [&](int k, int &mx, int &my) { const int i = start + (k >> 1) * 2; mx = ox; my = oy + ((k & 1) ? -i : i);
                                                                     return (k & 1) ? my >= y_min : my <= y_max; }, bcost, bmx, bmy);
            };
            const int ucost1 = bcost;
            scan_fpel(c, lane, c_square1, 0, 4, pmx, pmy, bcost, bmx, bmy);                 // DIA1_ITER( pmx, pmy )
            if (pmx | pmy) scan_fpel(c, lane, c_square1, 0, 4, 0, 0, bcost, bmx, bmy);      // DIA1_ITER( 0, 0 )
            if (c.i_pixel != X264_CUDA_PIXEL_4x4) {
                const int ucost2 = bcost;
                if ((bmx | bmy) && ((bmx - pmx) | (bmy - pmy))) { const int cx = bmx, cy = bmy; scan_fpel(c, lane, c_square1, 0, 4, cx, cy, bcost, bmx, bmy); }
                if (bcost == ucost2) cross_start = 3;
                omx = bmx; omy = bmy;
                bool done = false;
                if (bcost == ucost2 && SAD_THRESH(2000)) { // early termination
                    scan_fpel(c, lane, c_umh_oct, 0, 8, omx, omy, bcost, bmx, bmy);
                    if (bcost == ucost1 && SAD_THRESH(500)) done = true;
                    else if (bcost == ucost2) {
                        const int r = (range >> 1) | 1;
                        cross(3, r, r);
                        scan_fpel(c, lane, c_umh_star, 0, 8, omx, omy, bcost, bmx, bmy);
                        if (bcost == ucost2) done = true;
                        else cross_start = r + 2;
                    }
                }
                if (done) hex2 = false;
                else {
                    const int n_mvc = min((int)job.i_mvc, 12);
                    if (n_mvc) { // adaptive search range from the agreement of the predictors (me.c:354-399)
                        int mvd, denom = 1;
                        if (n_mvc == 1)
                            mvd = c.i_pixel == X264_CUDA_PIXEL_16x16 ? 25 : abs(job.mvp[0] - job.mvc[0][0]) + abs(job.mvp[1] - job.mvc[0][1]);
                        else {
                            denom = n_mvc - 1;
                            mvd = 0;
                            if (c.i_pixel != X264_CUDA_PIXEL_16x16) { mvd = abs(job.mvp[0] - job.mvc[0][0]) + abs(job.mvp[1] - job.mvc[0][1]); denom++; }
                            for (int k = 0; k < n_mvc - 1; k++) mvd += abs(job.mvc[k][0] - job.mvc[k + 1][0]) + abs(job.mvc[k][1] - job.mvc[k + 1][1]);
                        }
                        const int sad_ctx = SAD_THRESH(1000) ? 0 : SAD_THRESH(2000) ? 1 : SAD_THRESH(4000) ? 2 : 3;
                        const int mvd_ctx = mvd < 10 * denom ? 0 : mvd < 20 * denom ? 1 : mvd < 40 * denom ? 2 : 3;
                        range = range * c_umh_range_mul[mvd_ctx][sad_ctx] / 4;
                    }
                    cross(cross_start, range, range / 2);
                    scan_fpel(c, lane, c_umh_diag2, 0, 4, omx, omy, bcost, bmx, bmy);
                    omx = bmx; omy = bmy;
                    int i = 1;
                    do { // hexagon grid, 16 points scaled by i (range-checked: the unchecked form needs all of them inside anyway)
                        const int ox = omx, oy = omy, sc = i;
                        scan_gen(c, lane, 16, [&](int k, int &mx, int &my) { mx = ox + c_umh_hex4[k][0] * sc; my = oy + c_umh_hex4[k][1] * sc;
                                                                             return mx >= x_min && mx <= x_max && my >= y_min && my <= y_max; }, bcost, bmx, bmy);
                    } while (++i <= range / 4);
                    hex2 = bmy <= y_max;
                }
            }
#undef SAD_THRESH
        }
        if (hex2) { // me.c:246-304 (me_hex2)
            int ox = bmx, oy = bmy;
            int dir = scan_fpel(c, lane, c_hex2, 1, 6, ox, oy, bcost, ox, oy); // hex2[1..6] == (-2,0),(-1,2),(1,2),(2,0),(1,-2),(-1,-2)
            if (dir >= 0) {
                bmx += c_hex2[dir + 1][0]; bmy += c_hex2[dir + 1][1];
                for (int i = 1; i < range / 2 && IN_RANGE(bmx, bmy); i++) {
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
    }
    else if (method == X264_CUDA_ME_METHOD_TESA)
        tesa_search(c, job, me_range, lane, bcost, bmx, bmy);
    const int fbmx = bmx, fbmy = bmy;

    // ---- "-> qpel mv", me.c:603-620
    int mvx, mvy, cost;
    if (refine_only) { mvx = bmx; mvy = bmy; cost = bcost; }
    else if (bpred_cost < bcost) { mvx = bpred_mx; mvy = bpred_my; cost = bpred_cost; }
    else { mvx = bmx << 2; mvy = bmy << 2; cost = bcost; }
    int cost_mv = c.cmx[mvx] + c.cmy[mvy];
    if (!refine_only && bmx == pmx && bmy == pmy && subme < 3) cost += cost_mv;

    // ---- refine_subpel(h, m, hpel, qpel, NULL, b_refine_qpel), me.c:680-778: from x264_me_search_ref (columns 2,3 of the iteration
    // table, b_refine_qpel 0, only for subme >= 2) or as x264_me_refine_qpel (me.c:633-643: columns 0,1, b_refine_qpel 1)
    if (subme >= 2 || refine_only) {
        const int hpel_iters = c_subpel_iters[subme][refine_only ? 0 : 2], qpel_iters = c_subpel_iters[subme][refine_only ? 1 : 3];
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
        // !b_refine_qpel (me.c:729-736): the centre is costed again with mbcmp.  Where a round has room for five candidates (U <= 4) that
        // evaluation rides along with the first quarter-pel diamond instead of being a round of its own (the lookahead's searches are a chain
        // of such rounds on the wavefront's critical path)
        bool centre_pending = false;
        if (!refine_only) {
            if (sy > spel_ymax) sy = spel_ymax;
            if (qpel_iters > 0 && 32 / c.U >= 5) centre_pending = true;
            else sc = eval_round_satd(c, lane, true, sx, sy);
        }
        int bdir = -1;
        for (int i = qpel_iters; i > 0; i--) { // quarter-pel diamond with mbcmp, me.c:755-767
            const int odir = bdir, ox = sx, oy = sy;
            const int k = lane / c.U;
            const int dx = k == 2 ? -1 : k == 3 ? 1 : 0, dy = k == 0 ? -1 : k == 1 ? 1 : 0; // k == 4: the centre itself
            const bool valid = (k < 4 && (refine_only || (k ^ 1) != odir)) || (centre_pending && k == 4); // COST_MV_SATD: if( b_refine_qpel || (dir^1) != odir )
            const int v = eval_round_satd(c, lane, valid, ox + dx, oy + dy);
            if (centre_pending) { sc = cand_cost(c, v, 4); centre_pending = false; }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int vj = cand_cost(c, v, j);
                if ((refine_only || (j ^ 1) != odir) && vj < sc) {
                    sc = vj; bdir = j;
                    sx = ox + (j == 2 ? -1 : j == 3 ? 1 : 0); sy = oy + (j == 0 ? -1 : j == 1 ? 1 : 0);
                }
            }
            if (sx == ox && sy == oy) break;
        }
        if (sy > spel_ymax) { sy = spel_ymax; sc = eval_round_satd(c, lane, true, sx, sy); } // me.c:770-775
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

struct Planes {
    const uint8_t *fenc; const uint8_t *ref[4]; int stride; const uint16_t *sums8, *sums4; int2 *scratch; int list_cap;
    const uint8_t *fenc_c[2], *ref_c[2]; int stride_c; // chroma planes (NULL without X264_CUDA_FRAME_CHROMA on both frames)
};

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
    c.chroma_me = (job.flags & X264_CUDA_ME_CHROMA) && ip <= X264_CUDA_PIXEL_8x8; // me.c:686
    if (c.chroma_me && !pl.ref_c[0]) {
        if (lane == 0) { x264_cuda_me_final_t r = { { 0, 0 }, -1, -1, 0, 0 }; results[j] = r; } // frames without chroma planes
        continue;
    }
    c.stride_c = pl.stride_c;
    {
        const size_t offc = (size_t)(job.by >> 1) * pl.stride_c + (job.bx >> 1);
        c.fe_c[0] = pl.fenc_c[0] + offc; c.fe_c[1] = pl.fenc_c[1] + offc; c.ref_c[0] = pl.ref_c[0] + offc; c.ref_c[1] = pl.ref_c[1] + offc;
    }
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
    x264_cuda_enter(ctx);
    if (n_jobs <= 0) return 0;
    if (fenc->g.stride != fref->g.stride || fenc->g.lines != fref->g.lines) {
        snprintf(ctx->err, 256, "x264_cuda_me_search_small: fenc/fref geometry mismatch");
        return -1;
    }
    if (subme < 0 || subme > 9 || me_range < 1 || me_range > 64 ||
        (method != X264_CUDA_ME_METHOD_DIA && method != X264_CUDA_ME_METHOD_HEX && method != X264_CUDA_ME_METHOD_UMH &&
         method != X264_CUDA_ME_METHOD_SEEDED && method != X264_CUDA_ME_METHOD_TESA && method != X264_CUDA_ME_METHOD_REFINE_QPEL)) {
        snprintf(ctx->err, 256, "x264_cuda_me_search_small: bad method %d / subme %d / me_range %d", method, subme, me_range);
        return -1;
    }
    if (subme >= 2 && !(fref->g.flags & X264_CUDA_FRAME_HPEL)) {
        snprintf(ctx->err, 256, "x264_cuda_me_search_small: subme %d needs the half-pel planes (X264_CUDA_FRAME_HPEL)", subme);
        return -1;
    }
    const int16_t *const *d_tabs;
    if (x264_cuda_cost_tables(ctx, &d_tabs)) return -1;
    Planes pl = { fenc->plane[0], { fref->plane[0], fref->plane[1], fref->plane[2], fref->plane[3] }, fenc->g.stride, nullptr, nullptr, nullptr, 0,
                  { nullptr, nullptr }, { nullptr, nullptr }, 0 };
    if (fenc->buf_chroma && fref->buf_chroma && fenc->stride_c == fref->stride_c) {
        pl.fenc_c[0] = fenc->chroma[0]; pl.fenc_c[1] = fenc->chroma[1]; pl.ref_c[0] = fref->chroma[0]; pl.ref_c[1] = fref->chroma[1];
        pl.stride_c = fenc->stride_c;
    }
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
    x264_cuda_enter(ctx);
    if (n_jobs <= 0) return 0;
    const size_t jb = (size_t)n_jobs * sizeof(x264_cuda_me_job_t), rb = (size_t)n_jobs * sizeof(x264_cuda_me_final_t);
    const size_t jb_al = (jb + 255) & ~(size_t)255;
    if (x264_cuda_stage(ctx, jb_al + rb, jb_al + rb)) return -1;
    uint8_t *hs = (uint8_t *)ctx->h_stage, *ds = (uint8_t *)ctx->d_stage;
    if (x264_cuda_jobs_in(ctx, ds, jobs, hs, jb)) return -1;
    if (x264_cuda_me_search_small_dev(ctx, fenc, fref, method, me_range, subme, ds, n_jobs, ds + jb_al)) return -1;
    if (x264_cuda_results_out(ctx, results, ds + jb_al, hs + jb_al, rb)) return -1;
    return 0;
}

extern "C" int x264_cuda_block_cmp(x264_cuda_t *ctx, int metric, int i_pixel, int n, const uint8_t *pix1, const uint8_t *pix2, int *out)
{
    x264_cuda_enter(ctx);
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

// =====================================================================================================================
// Lowres lookahead: x264_slicetype_frame_cost / x264_slicetype_mb_cost (S/encoder/slicetype.c:43-355)
// =====================================================================================================================
namespace {

#define LA_F1(a, b) (((a) + (b) + 1) >> 1)
#define LA_F2(a, b, c) (((a) + 2 * (b) + (c) + 2) >> 2)

// one of the ten predictions of slicetype.c:205-229 at pixel (x,y) of the 8x8 block; tp[-1..15] row above, lf[0..7] left
// column (raw), l[]/t[]/lt the filtered edge of x264_predict_8x8_filter (S/common/predict.c:499-540)
struct Edge { int tp[17]; int lf[8]; int l[8]; int t[16]; int lt; int dc[4]; int pb, pc, pi00; };
__device__ __forceinline__ int EL(const Edge &e, int k) { return k < 8 ? e.l[7 - k] : k == 8 ? e.lt : e.t[k - 9]; } // l7..l0, lt, t0..t7
__device__ int intra_px(const Edge &e, int mode, int x, int y)
{
    switch (mode) {
    case 0: return e.dc[(y >> 2) * 2 + (x >> 2)];                                  // predict_8x8c_dc, predict.c:234-277
    case 1: return e.lf[y];                                                        // _h
    case 2: return e.tp[x + 1];                                                    // _v
    case 3: return clip_u8((e.pi00 + e.pc * y + e.pb * x) >> 5);                   // _p, predict.c:305-336
    case 4: { const int k = x + y; return k == 14 ? LA_F2(e.t[14], e.t[15], e.t[15]) : LA_F2(e.t[k], e.t[k + 1], e.t[k + 2]); } // ddl
    case 5: { const int k = 8 + x - y; return LA_F2(EL(e, k - 1), EL(e, k), EL(e, k + 1)); }                                   // ddr
    case 6: { const int z = 2 * x - y;                                                                                        // vr
              if (z >= 0) { const int k = 8 + x - (y >> 1); return (z & 1) ? LA_F2(EL(e, k - 1), EL(e, k), EL(e, k + 1)) : LA_F1(EL(e, k), EL(e, k + 1)); }
              if (z == -1) return LA_F2(e.l[0], e.lt, e.t[0]);
              const int k = y - 2 * x - 1; return LA_F2(e.l[k], e.l[k - 1], k - 2 >= 0 ? e.l[k - 2] : e.lt); }
    case 7: { const int z = 2 * y - x;                                                                                        // hd
              if (z >= 0) { const int k = 8 - y + (x >> 1); return (z & 1) ? LA_F2(EL(e, k - 1), EL(e, k), EL(e, k + 1)) : LA_F1(EL(e, k - 1), EL(e, k)); }
              if (z == -1) return LA_F2(e.l[0], e.lt, e.t[0]);
              const int k = x - 2 * y - 1; return LA_F2(e.t[k], e.t[k - 1], k - 2 >= 0 ? e.t[k - 2] : e.lt); }
    case 8: { const int k = x + (y >> 1); return (y & 1) ? LA_F2(e.t[k], e.t[k + 1], e.t[k + 2]) : LA_F1(e.t[k], e.t[k + 1]); } // vl
    default: { const int z = x + 2 * y;                                                                                       // hu
               if (z > 13) return e.l[7];
               if (z == 13) return LA_F2(e.l[6], e.l[7], e.l[7]);
               const int k = y + (x >> 1); return (z & 1) ? LA_F2(e.l[k], e.l[k + 1], e.l[k + 2]) : LA_F1(e.l[k], e.l[k + 1]); }
    }
}

// intra cost of every block (fully parallel, no dependencies): one thread per block, all threads of a warp walk the ten
// modes together.  slicetype.c:192-233: min over the predictions of mbcmp(pred, fenc) + 5.
__global__ void __launch_bounds__(64) lowres_intra_kernel(const uint8_t *__restrict__ l0, int stride, int W, int H, int satd, int *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W * H) return;
    const int mx = i % W, my = i / W;
    const uint8_t *src = l0 + (size_t)(8 * my) * stride + 8 * mx;
    Edge e;
    for (int k = 0; k < 17; k++) e.tp[k] = src[-stride - 1 + k];
    for (int k = 0; k < 8; k++) e.lf[k] = src[(ptrdiff_t)k * stride - 1];
    const int *top = e.tp + 1;
    e.lt = LA_F2(top[0], top[-1], e.lf[0]);
    e.l[0] = LA_F2(top[-1], e.lf[0], e.lf[1]);
    for (int y = 1; y < 7; y++) e.l[y] = LA_F2(e.lf[y - 1], e.lf[y], e.lf[y + 1]);
    e.l[7] = (e.lf[6] + 3 * e.lf[7] + 2) >> 2;
    e.t[0] = LA_F2(top[-1], top[0], top[1]);
    for (int x = 1; x < 15; x++) e.t[x] = LA_F2(top[x - 1], top[x], top[x + 1]);
    e.t[15] = (top[14] + 3 * top[15] + 2) >> 2;
    {
        int s0 = 0, s1 = 0, s2 = 0, s3 = 0, Hh = 0, V = 0;
        for (int k = 0; k < 4; k++) {
            s0 += top[k]; s1 += top[k + 4]; s2 += e.lf[k]; s3 += e.lf[k + 4];
            Hh += (k + 1) * (top[4 + k] - top[2 - k]);
            V += (k + 1) * (e.lf[4 + k] - (2 - k >= 0 ? e.lf[2 - k] : top[-1]));
        }
        e.dc[0] = (s0 + s2 + 4) >> 3; e.dc[1] = (s1 + 2) >> 2; e.dc[2] = (s3 + 2) >> 2; e.dc[3] = (s1 + s3 + 4) >> 3;
        const int a = 16 * (e.lf[7] + top[7]);
        e.pb = (17 * Hh + 16) >> 5; e.pc = (17 * V + 16) >> 5; e.pi00 = a - 3 * e.pb - 3 * e.pc + 16;
    }
    uint2 f[8];
    for (int y = 0; y < 8; y++) f[y] = ldg8(src + (size_t)y * stride);
    int best = 1 << 30;
    for (int mode = 0; mode < 10; mode++) {
        uint2 p[8];
        for (int y = 0; y < 8; y++) {
            uint32_t lo = 0, hi = 0;
            for (int x = 0; x < 4; x++) { lo |= (uint32_t)intra_px(e, mode, x, y) << (8 * x); hi |= (uint32_t)intra_px(e, mode, x + 4, y) << (8 * x); }
            p[y] = make_uint2(lo, hi);
        }
        int c;
        if (satd) {
            const uint2 fa[4] = { f[0], f[1], f[2], f[3] }, fb[4] = { f[4], f[5], f[6], f[7] };
            const uint2 pa[4] = { p[0], p[1], p[2], p[3] }, pb[4] = { p[4], p[5], p[6], p[7] };
            c = satd_8x4_rows(pa, fa) + satd_8x4_rows(pb, fb);
        } else {
            uint32_t acc = 0;
            for (int y = 0; y < 8; y++) { acc = sad4_acc(p[y].x, f[y].x, acc); acc = sad4_acc(p[y].y, f[y].y, acc); }
            c = (int)acc;
        }
        best = min(best, c);
    }
    out[i] = best + 5;
}

struct LaArgs {
    const uint8_t *fenc[4], *ref[2][4];
    int stride, W, H;
    int16_t *mvs[2]; int *costs[2]; const int16_t *ref1_mvs; int *intra; unsigned long long *sync[2]; int n_mb; const int *order; int n_order;
    int *ticket, *sums; // sums: score, intra_mbs, intra_cost_sum, score_aq
    const uint16_t *inv_q; int *row_satd; int all; // the VBV form (slicetype.c:300-316): every block evaluated, AQ-weighted per-row sums
    int epoch, b_bidir, b_any_inter, dsf, weight, method, me_range, do_search0, do_search1, mbcmp_satd, fpel_satd;
    const int16_t *tab; // p_cost_mv (qp 12) centre
};

// Hand-over between blocks of one launch: a block publishes, per list, ONE 64-bit word (epoch << 32 | mvy << 16 | mvx) with a
// relaxed store; a dependent block polls that word and takes the vector out of it.  The vector travels inside the flag, so
// neither side needs a fence or a second dependent load (a fenced flag + data hand-over through global memory costs ~1500
// cycles one way on B200, a relaxed one ~900: measured).
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void la_prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// TRY_BIDIR (slicetype.c:96-112): lanes 0/1 own the two 8x4 units of the block
__device__ int bidir_cost(const LaArgs &a, const uint8_t *fe, size_t off, int mv0x, int mv0y, int mv1x, int mv1y, int lane)
{
    int v = 0;
    if (lane < 2) {
        const uint8_t *const p0[4] = { a.ref[0][0] + off, a.ref[0][1] + off, a.ref[0][2] + off, a.ref[0][3] + off };
        const uint8_t *const p1[4] = { a.ref[1][0] + off, a.ref[1][1] + off, a.ref[1][2] + off, a.ref[1][3] + off };
        const QpelSrc s0 = qpel_src(p0, a.stride, mv0x, mv0y), s1 = qpel_src(p1, a.stride, mv1x, mv1y);
        uint2 f[4], r[4];
#pragma unroll
        for (int y = 0; y < 4; y++) {
            const ptrdiff_t o = (ptrdiff_t)(4 * lane + y) * a.stride;
            f[y] = ldg8(fe + o);
            const uint2 x0 = qpel_row8(s0, o), x1 = qpel_row8(s1, o);
            if (a.weight == 32) r[y] = make_uint2(__vavgu4(x0.x, x1.x), __vavgu4(x0.y, x1.y)); // pixel_avg_wxh, mc.c:52-62
            else { // pixel_avg_weight_wxh, mc.c:64-97
                uint32_t w[2] = { 0, 0 };
                const uint32_t xa[2] = { x0.x, x0.y }, xb[2] = { x1.x, x1.y };
                for (int k = 0; k < 8; k++) {
                    const int pa = (xa[k >> 2] >> (8 * (k & 3))) & 255, pb = (xb[k >> 2] >> (8 * (k & 3))) & 255;
                    w[k >> 2] |= (uint32_t)clip_u8((pa * a.weight + pb * (64 - a.weight) + 32) >> 6) << (8 * (k & 3));
                }
                r[y] = make_uint2(w[0], w[1]);
            }
        }
        if (a.mbcmp_satd) v = satd_8x4_rows(f, r);
        else {
            uint32_t acc = 0;
#pragma unroll
            for (int y = 0; y < 4; y++) { acc = sad4_acc(f[y].x, r[y].x, acc); acc = sad4_acc(f[y].y, r[y].y, acc); }
            v = (int)acc;
        }
    }
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return __shfl_sync(0xffffffffu, v, 0);
}

// `evals` = n_evals independent evaluations sharing one launch: ticket t belongs to evaluation t % n_evals (interleaved, so all
// wavefronts advance together) and is that evaluation's block number t / n_evals in wavefront order.
__global__ void __launch_bounds__(128) lowres_cost_kernel(const LaArgs *__restrict__ evals, int n_evals, int *ticket_counter, int n_order)
{
    __shared__ x264_cuda_me_job_t s_job[4];
    __shared__ x264_cuda_me_final_t s_fin[4];
    __shared__ uint32_t s_mvc[4][4];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (;;) {
        int t = 0;
        if (lane == 0) t = atomicAdd(ticket_counter, 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= n_order * n_evals) return;
        const LaArgs &a = evals[t % n_evals];
        t /= n_evals;
        const int xy = a.order[t], mx = xy % a.W, my = xy / a.W;
        const size_t off = (size_t)(8 * my) * a.stride + 8 * mx;
        const uint8_t *fe = a.fenc[0] + off;
        int bcost = COST_MAX;
        if (a.b_any_inter) {
            // while the neighbours finish: pull the full-pel search window of every list to be searched into this SM's L1
            for (int l = 0; l < 1 + a.b_bidir; l++)
                if (l ? a.do_search1 : a.do_search0) {
                    const uint8_t *w0 = a.ref[l][0] + off - (ptrdiff_t)20 * a.stride - 20;
                    for (int r = lane; r < 48; r += 32) { la_prefetch_l1(w0 + (size_t)r * a.stride); la_prefetch_l1(w0 + (size_t)r * a.stride + 47); }
                }
            const int fx_min = -8 * mx - 4, fx_max = 8 * (a.W - mx - 1) + 4, fy_min = -8 * my - 4, fy_max = 8 * (a.H - my - 1) + 4; // slicetype.c:75-85
            const int sx_min = 4 * (fx_min - 8), sx_max = 4 * (fx_max + 8), sy_min = 4 * (fy_min - 8), sy_max = 4 * (fy_max + 8);
            int mvx[2] = { 0, 0 }, mvy[2] = { 0, 0 };
            if (a.b_bidir) { // slicetype.c:121-142
                const int rx = a.ref1_mvs[2 * xy], ry = a.ref1_mvs[2 * xy + 1];
                int d0x = (rx * a.dsf + 128) >> 8, d0y = (ry * a.dsf + 128) >> 8;
                int d1x = d0x - rx, d1y = d0y - ry;
                d0x = clip3i(d0x, sx_min, sx_max); d0y = clip3i(d0y, sy_min, sy_max);
                d1x = clip3i(d1x, sx_min, sx_max); d1y = clip3i(d1y, sy_min, sy_max);
                bcost = min(bcost, bidir_cost(a, fe, off, d0x, d0y, d1x, d1y, lane));
                if (d0x | d0y | d1x | d1y) bcost = min(bcost, bidir_cost(a, fe, off, 0, 0, 0, 0, lane));
            }
            for (int l = 0; l < 1 + a.b_bidir; l++) {
                int cost;
                if (l ? a.do_search1 : a.do_search0) {
                    int n_mvc = 0;
                    __syncwarp();
                    // the vectors this block predicts from: right, below, below-left, below-right (slicetype.c:153-166), one lane
                    // each.  Blocks evaluated in this launch (the interior ones, or all of them for tiny frames) hand theirs over
                    // through the sync word; frame-edge blocks keep whatever the frame's array holds, like in the reference.
                    if (lane < 4) s_mvc[wid][lane] = 0;
                    __syncwarp();
                    {
                        const int dx = lane == 0 ? 1 : lane == 1 ? 0 : lane == 2 ? -1 : 1, dy = lane == 0 ? 0 : 1;
                        const int nx = mx + dx, ny = my + dy;
                        const bool ex = lane < 4 && nx >= 0 && nx < a.W && ny < a.H;
                        uint32_t v = 0;
                        if (ex) {
                            const bool small = a.W <= 2 || a.H <= 2 || a.all;
                            const bool inside = small || (nx >= 1 && nx <= a.W - 2 && ny >= 1 && ny <= a.H - 2);
                            if (inside) {
                                unsigned long long wv;
                                do wv = ld_relaxed_u64(a.sync[l] + nx + ny * a.W); while ((uint32_t)(wv >> 32) != (uint32_t)a.epoch);
                                v = (uint32_t)wv;
                            } else
                                v = *(const uint32_t *)(a.mvs[l] + 2 * (nx + ny * a.W));
                        }
                        const unsigned exm = __ballot_sync(0xffffffffu, ex);
                        if (ex) s_mvc[wid][__popc(exm & ((1u << lane) - 1))] = v;
                        n_mvc = __popc(exm);
                    }
                    __syncwarp();
                    if (lane == 0) {
                        x264_cuda_me_job_t &j = s_job[wid];
                        const int n = n_mvc;
                        int16_t c[4][2];
                        for (int k = 0; k < 4; k++) { c[k][0] = (int16_t)(s_mvc[wid][k] & 0xffff); c[k][1] = (int16_t)(s_mvc[wid][k] >> 16); }
                        j.bx = 8 * mx; j.by = 8 * my; j.i_pixel = X264_CUDA_PIXEL_8x8; j.qp = 12; j.i_mvc = n;
                        j.flags = (a.fpel_satd ? X264_CUDA_ME_FPEL_SATD : 0) | (a.mbcmp_satd ? X264_CUDA_ME_MBCMP_SATD : 0);
                        for (int k = 0; k < 2; k++) { // x264_median_mv(mvc[0], mvc[1], mvc[2])
                            const int p = c[0][k], q = c[1][k], r = c[2][k];
                            j.mvp[k] = (int16_t)max(min(p, q), min(max(p, q), r));
                        }
                        for (int k = 0; k < 4; k++) { j.mvc[k][0] = c[k][0]; j.mvc[k][1] = c[k][1]; }
                        j.mv_min_fpel[0] = fx_min; j.mv_max_fpel[0] = fx_max; j.mv_min_fpel[1] = fy_min; j.mv_max_fpel[1] = fy_max;
                        j.mv_min_spel[0] = sx_min; j.mv_max_spel[0] = sx_max; j.mv_min_spel[1] = sy_min; j.mv_max_spel[1] = sy_max;
                    }
                    __syncwarp();
                    const x264_cuda_me_job_t &job = s_job[wid];
                    SearchCtx c;
                    c.stride = a.stride; c.bw = 8; c.bh = 8; c.units = 2; c.U = 2; c.fe = fe;
#pragma unroll
                    for (int k = 0; k < 4; k++) c.planes[k] = a.ref[l][k] + off;
                    c.cmx = a.tab - job.mvp[0]; c.cmy = a.tab - job.mvp[1];
                    c.fpel_satd = a.fpel_satd; c.mbcmp_satd = a.mbcmp_satd;
                    c.chroma_me = false; // the lookahead has no chroma planes
                    c.i_pixel = X264_CUDA_PIXEL_8x8; c.sums8 = c.sums4 = nullptr; c.list = nullptr; c.list_cap = 0; c.lines_pad = 0;
                    warp_search(c, job, a.method, a.me_range, 4, lane, &s_fin[wid]);
                    __syncwarp();
                    const x264_cuda_me_final_t fin = s_fin[wid];
                    cost = fin.cost - 2;                                  // slicetype.c:169-171
                    if (fin.mv[0] | fin.mv[1]) cost += 5;
                    mvx[l] = fin.mv[0]; mvy[l] = fin.mv[1];
                    if (lane == 0) {
                        const uint32_t mvw = ((uint32_t)(uint16_t)fin.mv[0]) | ((uint32_t)(uint16_t)fin.mv[1] << 16);
                        st_relaxed_u64(a.sync[l] + xy, ((unsigned long long)(uint32_t)a.epoch << 32) | mvw); // hand-over
                        *(uint32_t *)(a.mvs[l] + 2 * xy) = mvw; // the frame's persistent lowres_mvs
                        a.costs[l][xy] = cost;
                    }
                } else {
                    mvx[l] = a.mvs[l][2 * xy]; mvy[l] = a.mvs[l][2 * xy + 1]; cost = a.costs[l][xy];
                }
                bcost = min(bcost, cost);
            }
            if (a.b_bidir && (mvx[0] | mvy[0] | mvx[1] | mvy[1])) bcost = min(bcost, 5 + bidir_cost(a, fe, off, mvx[0], mvy[0], mvx[1], mvy[1], lane));
        }
        int b_intra = 0, icost = 0;
        if (!a.b_bidir) { // slicetype.c:189-245
            icost = a.intra[xy];
            b_intra = icost < bcost;
            if (b_intra) bcost = icost;
        }
        if (lane == 0) {
            const bool interior = mx > 0 && mx < a.W - 1 && my > 0 && my < a.H - 1;
            const bool tiny = a.W <= 2 || a.H <= 2;
            const int aq = a.inv_q ? (bcost * a.inv_q[xy] + 128) >> 8 : bcost; // slicetype.c:307-309, :324-326
            if (tiny) atomicAdd(a.sums + 0, bcost);                            // :293-298: plain sum, no weighted score
            else {
                if (a.all) atomicAdd(a.row_satd + my, aq);                     // :310
                if (interior) { atomicAdd(a.sums + 0, bcost); atomicAdd(a.sums + 3, aq); }
            }
            if (!a.b_bidir && interior) { atomicAdd(a.sums + 1, b_intra); atomicAdd(a.sums + 2, icost); }
        }
    }
}

} // namespace

static std::atomic<int> g_la_epoch{0};

extern "C" int x264_cuda_frame_lookahead_alloc(x264_cuda_t *ctx, x264_cuda_frame_t *f, int n_dist)
{
    x264_cuda_enter(ctx);
    if (!(f->g.flags & X264_CUDA_FRAME_LOWRES) || n_dist < 1 || n_dist > 17) {
        snprintf(ctx->err, 256, "x264_cuda_frame_lookahead_alloc: frame needs X264_CUDA_FRAME_LOWRES and 1 <= n_dist <= 17");
        return -1;
    }
    const size_t n_mb = (size_t)f->g.mb_width * f->g.mb_height;
    cudaFree(f->la_mvs); cudaFree(f->la_costs); cudaFree(f->la_intra); cudaFree(f->la_done);
    f->la_mvs = nullptr; f->la_costs = nullptr; f->la_intra = nullptr; f->la_done = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&f->la_mvs, 2 * n_dist * n_mb * 4));
    CUDA_TRY(ctx, cudaMalloc(&f->la_costs, 2 * n_dist * n_mb * 4));
    CUDA_TRY(ctx, cudaMalloc(&f->la_intra, n_mb * 4));
    CUDA_TRY(ctx, cudaMalloc(&f->la_done, 2 * n_dist * n_mb * 8)); // hand-over words, one set per (list, distance) like the vectors they carry:
    // two evaluations of one batch that search the same frame and list at different distances never share a word (see lowres_cost_kernel)
    CUDA_TRY(ctx, cudaMemsetAsync(f->la_mvs, 0, 2 * n_dist * n_mb * 4, ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(f->la_costs, 0, 2 * n_dist * n_mb * 4, ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(f->la_intra, 0, n_mb * 4, ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(f->la_done, 0, 2 * n_dist * n_mb * 8, ctx->stream));
    f->la_dist = n_dist;
    return 0;
}

static int la_check(x264_cuda_t *ctx, const x264_cuda_frame_t *f, int list, int dist)
{
    if (!f->la_mvs || list < 0 || list > 1 || dist < 0 || dist >= f->la_dist) {
        snprintf(ctx->err, 256, "x264_cuda lookahead: no state for list %d distance %d (x264_cuda_frame_lookahead_alloc)", list, dist);
        return -1;
    }
    return 0;
}

extern "C" int x264_cuda_frame_lookahead_get(x264_cuda_t *ctx, const x264_cuda_frame_t *f, int list, int dist, int16_t *mvs, int *costs,
                                             uint16_t *intra_cost)
{
    x264_cuda_enter(ctx);
    if (la_check(ctx, f, list, dist)) return -1;
    const size_t n_mb = (size_t)f->g.mb_width * f->g.mb_height, o = ((size_t)list * f->la_dist + dist) * n_mb;
    if (mvs) CUDA_TRY(ctx, cudaMemcpyAsync(mvs, f->la_mvs + 2 * o, n_mb * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (costs) CUDA_TRY(ctx, cudaMemcpyAsync(costs, f->la_costs + o, n_mb * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (intra_cost) {
        int *tmp = (int *)malloc(n_mb * 4);
        cudaError_t e = cudaMemcpy(tmp, f->la_intra, n_mb * 4, cudaMemcpyDeviceToHost);
        for (size_t i = 0; i < n_mb; i++) intra_cost[i] = (uint16_t)tmp[i];
        free(tmp);
        if (e != cudaSuccess) return x264_cuda_fail(ctx, "lookahead_get", e);
    }
    return 0;
}

extern "C" int x264_cuda_frame_lookahead_set(x264_cuda_t *ctx, x264_cuda_frame_t *f, int list, int dist, const int16_t *mvs, const int *costs,
                                             const uint16_t *intra_cost)
{
    x264_cuda_enter(ctx);
    if (la_check(ctx, f, list, dist)) return -1;
    const size_t n_mb = (size_t)f->g.mb_width * f->g.mb_height, o = ((size_t)list * f->la_dist + dist) * n_mb;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (mvs) CUDA_TRY(ctx, cudaMemcpy(f->la_mvs + 2 * o, mvs, n_mb * 4, cudaMemcpyHostToDevice));
    if (costs) CUDA_TRY(ctx, cudaMemcpy(f->la_costs + o, costs, n_mb * 4, cudaMemcpyHostToDevice));
    if (intra_cost) {
        int *tmp = (int *)malloc(n_mb * 4);
        for (size_t i = 0; i < n_mb; i++) tmp[i] = intra_cost[i];
        cudaError_t e = cudaMemcpy(f->la_intra, tmp, n_mb * 4, cudaMemcpyHostToDevice);
        free(tmp);
        if (e != cudaSuccess) return x264_cuda_fail(ctx, "lookahead_set", e);
    }
    return 0;
}

#define LA_MAX_BATCH 64

// inv_qscales / row_satds: NULL, or one pointer per evaluation (host arrays: uint16[n_mb] / int[mb_height]; entries may be NULL)
static int lowres_frame_cost_core(x264_cuda_t *ctx, int n_evals, x264_cuda_frame_t *const *fencs, const x264_cuda_frame_t *const *fref0s,
                                  const x264_cuda_frame_t *const *fref1s, const x264_cuda_lowres_params_t *pms,
                                  const uint16_t *const *inv_qscales, x264_cuda_lowres_result_t *results, int *const *row_satds)
{
    x264_cuda_enter(ctx);
    if (n_evals < 1 || n_evals > LA_MAX_BATCH) { snprintf(ctx->err, 256, "x264_cuda_lowres_frame_cost_batch: 1..%d evaluations per call", LA_MAX_BATCH); return -1; }
    const x264_cuda_geom_t &g = fencs[0]->g;
    const int W = g.mb_width, H = g.mb_height;
    const int16_t *const *d_tabs;
    if (!ctx->d_cost_mv[12]) { // the lookahead always works at qp 12 (slicetype.c:35)
        int16_t *t = (int16_t *)malloc((4 * 4 * 2048 + 1) * sizeof(int16_t));
        x264_cuda_host_cost_mv(12, t);
        int rc = x264_cuda_set_cost_mv(ctx, 12, t);
        free(t);
        if (rc) return -1;
    }
    if (x264_cuda_cost_tables(ctx, &d_tabs)) return -1;
    // wavefront order: interior blocks (all blocks for tiny frames) by x + 2y descending
    if (ctx->la_w != W || ctx->la_h != H) {
        const bool small = W <= 2 || H <= 2;
        // two lists in one allocation: the interior blocks (default form), then every block (VBV form, slicetype.c:300-316)
        int *ord = (int *)malloc((size_t)2 * W * H * sizeof(int)), n = 0, n_all = 0;
        for (int pass = 0; pass < 2; pass++)
            for (int v = (W - 1) + 2 * (H - 1); v >= 0; v--)
                for (int y = H - 1; y >= 0; y--) {
                    const int x = v - 2 * y;
                    if (x < 0 || x >= W) continue;
                    if (pass == 0) {
                        if (!small && (x < 1 || x > W - 2 || y < 1 || y > H - 2)) continue;
                        ord[n++] = x + y * W;
                    } else
                        ord[n + n_all++] = x + y * W;
                }
        cudaFree(ctx->d_la_order); ctx->d_la_order = nullptr;
        cudaFree(ctx->d_la_vbv); ctx->d_la_vbv = nullptr;
        // per evaluation 8 ints of sums; then the ticket counter; then the LaArgs array
        if (!ctx->d_la_sums) CUDA_TRY(ctx, cudaMalloc(&ctx->d_la_sums, (8 * LA_MAX_BATCH + 8) * sizeof(int) + LA_MAX_BATCH * sizeof(LaArgs)));
        CUDA_TRY(ctx, cudaMalloc(&ctx->d_la_order, (size_t)(n + n_all) * sizeof(int)));
        CUDA_TRY(ctx, cudaMemcpy(ctx->d_la_order, ord, (size_t)(n + n_all) * sizeof(int), cudaMemcpyHostToDevice));
        free(ord);
        ctx->la_w = W; ctx->la_h = H; ctx->la_n = n;
    }
    const size_t n_mb = (size_t)W * H;
    // VBV / AQ side arrays, per evaluation: row sums int[H_al] | inverse qscale uint16[n_mb_al]
    const size_t vbv_rows = ((size_t)H + 3) & ~(size_t)3, vbv_stride = vbv_rows * sizeof(int) + ((n_mb * sizeof(uint16_t) + 15) & ~(size_t)15);
    bool any_vbv = false, any_all = false;
    for (int e = 0; e < n_evals; e++) {
        const bool all = (pms[e].flags & X264_CUDA_LOWRES_VBV) != 0;
        if (all && !(row_satds && row_satds[e])) { snprintf(ctx->err, 256, "x264_cuda_lowres_frame_cost: X264_CUDA_LOWRES_VBV needs a row_satd array"); return -1; }
        any_all |= all;
        any_vbv |= all || (inv_qscales && inv_qscales[e]);
    }
    if (any_all && n_evals > 1) {
        // one ticket space per launch: evaluations of a batch must all walk the same block list
        for (int e = 0; e < n_evals; e++)
            if (!(pms[e].flags & X264_CUDA_LOWRES_VBV)) { snprintf(ctx->err, 256, "x264_cuda_lowres_frame_cost_batch: VBV and default evaluations cannot share a batch"); return -1; }
    }
    if (any_vbv) {
        if (!ctx->d_la_vbv) CUDA_TRY(ctx, cudaMalloc(&ctx->d_la_vbv, vbv_stride * LA_MAX_BATCH));
        CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_la_vbv, 0, vbv_stride * n_evals, ctx->stream));
        for (int e = 0; e < n_evals; e++)
            if (inv_qscales && inv_qscales[e])
                CUDA_TRY(ctx, cudaMemcpyAsync((uint8_t *)ctx->d_la_vbv + vbv_stride * e + vbv_rows * sizeof(int), inv_qscales[e], n_mb * sizeof(uint16_t), cudaMemcpyHostToDevice, ctx->stream));
    }
    const int n_order = any_all && !(W <= 2 || H <= 2) ? (int)n_mb : ctx->la_n;
    const int *d_order = any_all && !(W <= 2 || H <= 2) ? ctx->d_la_order + ctx->la_n : ctx->d_la_order;
    int *d_ticket = ctx->d_la_sums + 8 * LA_MAX_BATCH;
    LaArgs *d_evals = (LaArgs *)(ctx->d_la_sums + 8 * LA_MAX_BATCH + 8);
    static_assert(sizeof(LaArgs) % 8 == 0, "LaArgs array follows an 8-int header");
    LaArgs *h_evals = (LaArgs *)malloc((size_t)n_evals * sizeof(LaArgs));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_la_sums, 0, (8 * LA_MAX_BATCH + 8) * sizeof(int), ctx->stream));
    for (int e = 0; e < n_evals; e++) {
        x264_cuda_frame_t *fenc = fencs[e];
        const x264_cuda_frame_t *fref0 = fref0s[e], *fref1 = fref1s[e];
        const x264_cuda_lowres_params_t *pm = pms + e;
        const int b_bidir = pm->b < pm->p1, any = !(pm->p0 == pm->p1 && pm->p0 == pm->b);
        if (!fenc->lowres[0] || !fref0->lowres[0] || !fref1->lowres[0] || fref0->g.stride_lowres != g.stride_lowres || fref1->g.stride_lowres != g.stride_lowres ||
            fenc->g.stride_lowres != g.stride_lowres || fenc->g.mb_width != W || fenc->g.mb_height != H) {
            snprintf(ctx->err, 256, "x264_cuda_lowres_frame_cost: frames need X264_CUDA_FRAME_LOWRES and equal geometry");
            free(h_evals);
            return -1;
        }
        const int d0 = pm->b - pm->p0 - 1, d1 = pm->p1 - pm->b - 1;
        if (!fenc->la_mvs || (any && (d0 < 0 || d0 >= fenc->la_dist)) ||
            (b_bidir && (d1 < 0 || d1 >= fenc->la_dist || !fref1->la_mvs || pm->p1 - pm->p0 - 1 >= fref1->la_dist))) {
            snprintf(ctx->err, 256, "x264_cuda_lowres_frame_cost: lookahead state missing or distance out of range");
            free(h_evals);
            return -1;
        }
        if (!b_bidir && !pm->b_intra_calculated) {
            lowres_intra_kernel<<<(int)((n_mb + 63) / 64), 64, 0, ctx->stream>>>(fenc->lowres[0], g.stride_lowres, W, H, !!(pm->flags & X264_CUDA_ME_MBCMP_SATD), fenc->la_intra);
            ctx->launches++;
        }
        LaArgs &a = h_evals[e];
        memset(&a, 0, sizeof(a));
        for (int k = 0; k < 4; k++) { a.fenc[k] = fenc->lowres[k]; a.ref[0][k] = fref0->lowres[k]; a.ref[1][k] = fref1->lowres[k]; }
        a.stride = g.stride_lowres; a.W = W; a.H = H;
        a.mvs[0] = fenc->la_mvs + 2 * ((size_t)0 * fenc->la_dist + (d0 < 0 ? 0 : d0)) * n_mb; a.costs[0] = fenc->la_costs + ((size_t)0 * fenc->la_dist + (d0 < 0 ? 0 : d0)) * n_mb;
        a.mvs[1] = fenc->la_mvs + 2 * ((size_t)1 * fenc->la_dist + (d1 < 0 ? 0 : d1)) * n_mb; a.costs[1] = fenc->la_costs + ((size_t)1 * fenc->la_dist + (d1 < 0 ? 0 : d1)) * n_mb;
        a.ref1_mvs = b_bidir ? fref1->la_mvs + 2 * ((size_t)(pm->p1 - pm->p0 - 1)) * n_mb : nullptr;
        a.intra = fenc->la_intra; a.n_mb = (int)n_mb;
        a.sync[0] = (unsigned long long *)fenc->la_done + ((size_t)0 * fenc->la_dist + (d0 < 0 ? 0 : d0)) * n_mb;
        a.sync[1] = (unsigned long long *)fenc->la_done + ((size_t)1 * fenc->la_dist + (d1 < 0 ? 0 : d1)) * n_mb;
        a.order = d_order; a.n_order = n_order;
        a.ticket = d_ticket; a.sums = ctx->d_la_sums + 8 * e;
        a.all = any_all;
        a.row_satd = any_vbv ? (int *)((uint8_t *)ctx->d_la_vbv + vbv_stride * e) : nullptr;
        a.inv_q = (inv_qscales && inv_qscales[e]) ? (const uint16_t *)((uint8_t *)ctx->d_la_vbv + vbv_stride * e + vbv_rows * sizeof(int)) : nullptr;
        a.epoch = ++g_la_epoch; /* process-wide: the hand-over words live in the FRAME, which several contexts may evaluate in turn */ a.b_bidir = b_bidir; a.b_any_inter = any;
        a.dsf = pm->p1 != pm->p0 ? (((pm->b - pm->p0) << 8) + ((pm->p1 - pm->p0) >> 1)) / (pm->p1 - pm->p0) : 128; // slicetype.c:289-290
        a.weight = (pm->flags & X264_CUDA_LOWRES_WEIGHTED_BIPRED) ? 64 - (a.dsf >> 2) : 32;                          // slicetype.c:57
        a.method = pm->me_method < 1 ? X264_CUDA_ME_METHOD_DIA : X264_CUDA_ME_METHOD_HEX;                            // min(HEX, me), slicetype.c:38
        a.me_range = pm->me_range; a.do_search0 = pm->do_search[0]; a.do_search1 = pm->do_search[1];
        a.mbcmp_satd = !!(pm->flags & X264_CUDA_ME_MBCMP_SATD); a.fpel_satd = !!(pm->flags & X264_CUDA_ME_FPEL_SATD);
        a.tab = ctx->d_cost_mv[12] + 2 * 4 * 2048;
    }
    // the contract of the batch (x264_cuda.h): no two evaluations may search the same (frame, list, distance) state — they would race on
    // the vectors and on the hand-over words, and a waiter could spin on an epoch that never comes.  Refuse instead of hanging.
    for (int e = 0; e < n_evals; e++)
        for (int f = 0; f < e; f++)
            for (int l = 0; l < 2; l++)
                for (int k = 0; k < 2; k++) {
                    const bool se = (l ? h_evals[e].do_search1 && h_evals[e].b_bidir : h_evals[e].do_search0) && h_evals[e].b_any_inter;
                    const bool sf = (k ? h_evals[f].do_search1 && h_evals[f].b_bidir : h_evals[f].do_search0) && h_evals[f].b_any_inter;
                    if (se && sf && h_evals[e].mvs[l] == h_evals[f].mvs[k]) {
                        snprintf(ctx->err, 256, "x264_cuda_lowres_frame_cost_batch: evaluations %d and %d search the same (frame, list, distance)", f, e);
                        free(h_evals);
                        return -1;
                    }
                }
    cudaError_t ce = cudaMemcpyAsync(d_evals, h_evals, (size_t)n_evals * sizeof(LaArgs), cudaMemcpyHostToDevice, ctx->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream); // h_evals is pageable and freed below
    free(h_evals);
    if (ce != cudaSuccess) return x264_cuda_fail(ctx, "lowres_frame_cost: evaluation list upload", ce);
    // persistent warps pulling tickets in wavefront order: every ticket's dependencies hold smaller tickets, so a waiting
    // warp only ever waits for warps that are already running.  A diagonal of one wavefront holds at most ~W/2 blocks; twice
    // that many warps per evaluation keep the next diagonals prefetching while leaving the rest of the GPU to other kernels.
    const int blocks = min((n_order * n_evals + 3) / 4, min(max(8, (W + 3) / 4) * n_evals, ctx->sm_count * 4));
    lowres_cost_kernel<<<blocks, 128, 0, ctx->stream>>>(d_evals, n_evals, d_ticket, n_order);
    LAUNCH_CHECK(ctx, "lowres_cost_kernel");
    int *sums = (int *)malloc((size_t)n_evals * 8 * sizeof(int));
    ce = cudaMemcpyAsync(sums, ctx->d_la_sums, (size_t)n_evals * 8 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    for (int e = 0; e < n_evals && ce == cudaSuccess; e++)
        if (any_all && row_satds && row_satds[e])
            ce = cudaMemcpyAsync(row_satds[e], (uint8_t *)ctx->d_la_vbv + vbv_stride * e, (size_t)H * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream);
    if (ce == cudaSuccess)
        for (int e = 0; e < n_evals; e++) { results[e].score = sums[8 * e]; results[e].intra_mbs = sums[8 * e + 1]; results[e].intra_cost_sum = sums[8 * e + 2]; results[e].score_aq = sums[8 * e + 3]; }
    free(sums);
    if (ce != cudaSuccess) return x264_cuda_fail(ctx, "lowres_frame_cost", ce);
    return 0;
}

extern "C" int x264_cuda_lowres_frame_cost_batch(x264_cuda_t *ctx, int n_evals, x264_cuda_frame_t *const *fencs, const x264_cuda_frame_t *const *fref0s,
                                                 const x264_cuda_frame_t *const *fref1s, const x264_cuda_lowres_params_t *pms,
                                                 x264_cuda_lowres_result_t *results)
{
    return lowres_frame_cost_core(ctx, n_evals, fencs, fref0s, fref1s, pms, nullptr, results, nullptr);
}

extern "C" int x264_cuda_lowres_frame_cost_batch_rc(x264_cuda_t *ctx, int n_evals, x264_cuda_frame_t *const *fencs, const x264_cuda_frame_t *const *fref0s,
                                                    const x264_cuda_frame_t *const *fref1s, const x264_cuda_lowres_params_t *pms,
                                                    const uint16_t *const *inv_qscales, x264_cuda_lowres_result_t *results, int *const *row_satds)
{
    return lowres_frame_cost_core(ctx, n_evals, fencs, fref0s, fref1s, pms, inv_qscales, results, row_satds);
}

extern "C" int x264_cuda_lowres_frame_cost_rc(x264_cuda_t *ctx, x264_cuda_frame_t *fenc, const x264_cuda_frame_t *fref0, const x264_cuda_frame_t *fref1,
                                              const x264_cuda_lowres_params_t *pm, const uint16_t *inv_qscale, x264_cuda_lowres_result_t *result, int *row_satd)
{
    return lowres_frame_cost_core(ctx, 1, &fenc, &fref0, &fref1, pm, &inv_qscale, result, &row_satd);
}

extern "C" int x264_cuda_lowres_frame_cost(x264_cuda_t *ctx, x264_cuda_frame_t *fenc, const x264_cuda_frame_t *fref0,
                                           const x264_cuda_frame_t *fref1, const x264_cuda_lowres_params_t *pm, x264_cuda_lowres_result_t *result)
{
    return x264_cuda_lowres_frame_cost_batch(ctx, 1, &fenc, &fref0, &fref1, pm, result);
}

// =====================================================================================================================
// x264_me_refine_bidir_satd (S/encoder/me.c:843-927): joint quarter-pel refinement of a (list 0, list 1) vector pair against the
// blended prediction.  One warp per partition: the 32 candidate pairs of a pass are spread over the lanes (U lanes per candidate, one
// 8x4 unit each), their costs land in shared memory, and the reference's sequential part — the aliasing `visited` map (indices & 7)
// and the strict-'<' scan in CHECK_BIDIR order — is then replayed over those 32 costs.
// =====================================================================================================================
namespace {

__constant__ int8_t c_bidir_cand[32][4] = {
    { 0, 0, 0, 1 }, { 0, 0, 0, -1 }, { 0, 0, 1, 0 }, { 0, 0, -1, 0 }, { 0, 1, 0, 0 }, { 0, -1, 0, 0 }, { 1, 0, 0, 0 }, { -1, 0, 0, 0 },   // CHECK_BIDIR8( 0, 0, 0, 1 )
    { 0, 0, 1, 1 }, { 0, 0, -1, -1 }, { 0, 1, 1, 0 }, { 0, -1, -1, 0 }, { 1, 1, 0, 0 }, { -1, -1, 0, 0 }, { 1, 0, 0, 1 }, { -1, 0, 0, -1 }, // CHECK_BIDIR8( 0, 0, 1, 1 )
    { 0, 1, 0, 1 }, { 0, -1, 0, -1 }, { 1, 0, 1, 0 }, { -1, 0, -1, 0 },                                                                 // CHECK_BIDIR2 x 2
    { 0, 0, -1, 1 }, { 0, 0, 1, -1 }, { 0, -1, 1, 0 }, { 0, 1, -1, 0 }, { -1, 1, 0, 0 }, { 1, -1, 0, 0 }, { 1, 0, 0, -1 }, { -1, 0, 0, 1 }, // CHECK_BIDIR8( 0, 0,-1, 1 )
    { 0, -1, 0, 1 }, { 0, 1, 0, -1 }, { -1, 0, 1, 0 }, { 1, 0, -1, 0 } };                                                               // CHECK_BIDIR2 x 2

struct BidirPlanes { const uint8_t *fenc; const uint8_t *ref0[4], *ref1[4]; int stride; };

// mbcmp of one 8x4 unit of the blended prediction (h->mc.avg then pixf.mbcmp, me.c:802-803)
__device__ __forceinline__ int bidir_unit(const uint8_t *fe, const QpelSrc &s0, const QpelSrc &s1, int stride, int ux, int uy, int weight, bool satd)
{
    uint2 f[4], r[4];
#pragma unroll
    for (int y = 0; y < 4; y++) {
        const ptrdiff_t o = (ptrdiff_t)(uy + y) * stride + ux;
        f[y] = ldg8(fe + o);
        const uint2 a = qpel_row8(s0, o), b = qpel_row8(s1, o);
        if (weight == 32) r[y] = make_uint2(__vavgu4(a.x, b.x), __vavgu4(a.y, b.y));
        else {
            uint32_t w[2] = { 0, 0 };
            const uint32_t xa[2] = { a.x, a.y }, xb[2] = { b.x, b.y };
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int pa = (xa[k >> 2] >> (8 * (k & 3))) & 255, pb = (xb[k >> 2] >> (8 * (k & 3))) & 255;
                w[k >> 2] |= (uint32_t)clip_u8((pa * weight + pb * (64 - weight) + 32) >> 6) << (8 * (k & 3));
            }
            r[y] = make_uint2(w[0], w[1]);
        }
    }
    if (satd) return satd_8x4_rows(f, r);
    uint32_t acc = 0;
#pragma unroll
    for (int y = 0; y < 4; y++) { acc = sad4_acc(f[y].x, r[y].x, acc); acc = sad4_acc(f[y].y, r[y].y, acc); }
    return (int)acc;
}

__global__ void __launch_bounds__(128) me_bidir_kernel(BidirPlanes pl, const x264_cuda_bidir_job_t *__restrict__ jobs, int n_jobs,
                                                       const int16_t *const *__restrict__ cost_tabs, x264_cuda_bidir_result_t *__restrict__ results)
{
    __shared__ uint8_t s_visited[4][512];
    __shared__ int s_cost[4][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (j >= n_jobs) return;
    const x264_cuda_bidir_job_t job = jobs[j];
    const int ip = min((int)job.i_pixel, 3);
    const int bw = ip <= 1 ? 16 : 8, bh = (ip == 0 || ip == 2) ? 16 : 8;
    const int U = unit_count(bw, bh), per = 32 / U, weight = job.weight;
    const bool satd = job.flags & X264_CUDA_ME_MBCMP_SATD;
    const int16_t *tab = cost_tabs[job.qp > 51 ? 51 : job.qp] + 2 * 4 * 2048;
    const int lo = job.mv_min_spel[0], hi = job.mv_max_spel[0]; // the reference clips BOTH components of mvp with the x limits (me.c:854-857)
    const int16_t *c0x = tab - clip3i(job.mvp0[0], lo, hi), *c0y = tab - clip3i(job.mvp0[1], lo, hi);
    const int16_t *c1x = tab - clip3i(job.mvp1[0], lo, hi), *c1y = tab - clip3i(job.mvp1[1], lo, hi);
    int bm0x = job.mv0[0], bm0y = job.mv0[1], bm1x = job.mv1[0], bm1y = job.mv1[1], bcost = COST_MAX;
    x264_cuda_bidir_result_t res = { { (int16_t)bm0x, (int16_t)bm0y }, { (int16_t)bm1x, (int16_t)bm1y }, COST_MAX };
    // table indices must stay inside p_cost_mv (+-2*4*2048): vectors move by at most 8 quarter-pels from here
    const int lim = 2 * 4 * 2048 - 16;
    const bool bad = abs(bm0x - clip3i(job.mvp0[0], lo, hi)) > lim || abs(bm0y - clip3i(job.mvp0[1], lo, hi)) > lim ||
                     abs(bm1x - clip3i(job.mvp1[0], lo, hi)) > lim || abs(bm1y - clip3i(job.mvp1[1], lo, hi)) > lim;
    if (bad || bm0y > job.mv_max_spel[1] - 8 || bm1y > job.mv_max_spel[1] - 8) { // me.c:874-876 (or unusable input: cost -1)
        if (bad) res.cost = -1;
        if (lane == 0) results[j] = res;
        return;
    }
    const size_t off = (size_t)job.by * pl.stride + job.bx;
    const uint8_t *fe = pl.fenc + off;
    const uint8_t *const p0[4] = { pl.ref0[0] + off, pl.ref0[1] + off, pl.ref0[2] + off, pl.ref0[3] + off };
    const uint8_t *const p1[4] = { pl.ref1[0] + off, pl.ref1[1] + off, pl.ref1[2] + off, pl.ref1[3] + off };
    for (int i = lane; i < 128; i += 32) ((uint32_t *)s_visited[wid])[i] = 0;
    __syncwarp();
    auto pair_cost = [&](bool valid, int m0x, int m0y, int m1x, int m1y) { // all lanes of a candidate's group return its cost
        int v = 0;
        const int u = lane & (U - 1);
        if (valid) {
            int ux, uy;
            unit_pos(bw, u, ux, uy);
            const QpelSrc s0 = qpel_src(p0, pl.stride, m0x, m0y), s1 = qpel_src(p1, pl.stride, m1x, m1y);
            v = bidir_unit(fe, s0, s1, pl.stride, ux, uy, weight, satd);
        }
        for (int o = U >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v + c0x[m0x] + c0y[m0y] + c1x[m1x] + c1y[m1y];
    };
    auto visit = [&](int m0x, int m0y, int m1x, int m1y, int pass) { // COST_BIMV_SATD's gate; returns "evaluate" (uniform)
        uint8_t &cell = s_visited[wid][((m0x & 7) * 8 + (m0y & 7)) * 8 + (m1x & 7)];
        const bool go = pass == 0 || !(cell & (1 << (m1y & 7)));
        __syncwarp();
        if (go && lane == 0) cell |= (uint8_t)(1 << (m1y & 7));
        __syncwarp();
        return go;
    };
    int om0x = bm0x, om0y = bm0y, om1x = bm1x, om1y = bm1y;
    { // CHECK_BIDIR( 0, 0, 0, 0 ) with pass == 0
        visit(om0x, om0y, om1x, om1y, 0);
        bcost = __shfl_sync(0xffffffffu, pair_cost(lane < U, om0x, om0y, om1x, om1y), 0);
    }
    for (int pass = 0; pass < 8; pass++) {
        for (int k0 = 0; k0 < 32; k0 += per) {
            const int k = k0 + lane / U;
            const int c = pair_cost(true, om0x + c_bidir_cand[k][0], om0y + c_bidir_cand[k][1], om1x + c_bidir_cand[k][2], om1y + c_bidir_cand[k][3]);
            if ((lane & (U - 1)) == 0) s_cost[wid][k] = c;
        }
        __syncwarp();
        for (int k = 0; k < 32; k++) { // the sequential part, in CHECK_BIDIR order
            const int m0x = om0x + c_bidir_cand[k][0], m0y = om0y + c_bidir_cand[k][1], m1x = om1x + c_bidir_cand[k][2], m1y = om1y + c_bidir_cand[k][3];
            if (visit(m0x, m0y, m1x, m1y, pass)) {
                const int c = s_cost[wid][k];
                if (c < bcost) { bcost = c; bm0x = m0x; bm0y = m0y; bm1x = m1x; bm1y = m1y; }
            }
        }
        __syncwarp();
        if (om0x == bm0x && om0y == bm0y && om1x == bm1x && om1y == bm1y) break;
        om0x = bm0x; om0y = bm0y; om1x = bm1x; om1y = bm1y;
    }
    if (lane == 0) {
        res.mv0[0] = (int16_t)bm0x; res.mv0[1] = (int16_t)bm0y; res.mv1[0] = (int16_t)bm1x; res.mv1[1] = (int16_t)bm1y; res.cost = bcost;
        results[j] = res;
    }
}

} // namespace

extern "C" int x264_cuda_me_refine_bidir_dev(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, const x264_cuda_frame_t *fref0, const x264_cuda_frame_t *fref1,
                                             const void *d_jobs, int n_jobs, void *d_results)
{
    x264_cuda_enter(ctx);
    if (n_jobs <= 0) return 0;
    if (!(fref0->g.flags & fref1->g.flags & X264_CUDA_FRAME_HPEL) || fref0->g.stride != fenc->g.stride || fref1->g.stride != fenc->g.stride) {
        snprintf(ctx->err, 256, "x264_cuda_me_refine_bidir: both references need the half-pel planes and fenc's geometry");
        return -1;
    }
    const int16_t *const *d_tabs;
    if (x264_cuda_cost_tables(ctx, &d_tabs)) return -1;
    BidirPlanes pl;
    pl.fenc = fenc->plane[0]; pl.stride = fenc->g.stride;
    for (int k = 0; k < 4; k++) { pl.ref0[k] = fref0->plane[k]; pl.ref1[k] = fref1->plane[k]; }
    me_bidir_kernel<<<(n_jobs + 3) / 4, 128, 0, ctx->stream>>>(pl, (const x264_cuda_bidir_job_t *)d_jobs, n_jobs, d_tabs, (x264_cuda_bidir_result_t *)d_results);
    LAUNCH_CHECK(ctx, "me_bidir_kernel");
    return 0;
}

extern "C" int x264_cuda_me_refine_bidir(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, const x264_cuda_frame_t *fref0, const x264_cuda_frame_t *fref1,
                                         const x264_cuda_bidir_job_t *jobs, int n_jobs, x264_cuda_bidir_result_t *results)
{
    x264_cuda_enter(ctx);
    if (n_jobs <= 0) return 0;
    const size_t jb = (size_t)n_jobs * sizeof(x264_cuda_bidir_job_t), rb = (size_t)n_jobs * sizeof(x264_cuda_bidir_result_t);
    const size_t jb_al = (jb + 255) & ~(size_t)255;
    if (x264_cuda_stage(ctx, jb_al + rb, jb_al + rb)) return -1;
    uint8_t *hs = (uint8_t *)ctx->h_stage, *ds = (uint8_t *)ctx->d_stage;
    if (x264_cuda_jobs_in(ctx, ds, jobs, hs, jb)) return -1;
    if (x264_cuda_me_refine_bidir_dev(ctx, fenc, fref0, fref1, ds, n_jobs, ds + jb_al)) return -1;
    return x264_cuda_results_out(ctx, results, ds + jb_al, hs + jb_al, rb);
}
