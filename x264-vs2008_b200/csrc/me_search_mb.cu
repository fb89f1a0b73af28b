// me_search_mb.cu — macroblock-batched exhaustive search: the nine inter-partition ESA searches of one
// macroblock (16x16, 2x16x8, 2x8x16, 4x8x8) in ONE pass over the union of their windows.
//
// ONE WARP PER MACROBLOCK.  Lane l owns candidate column x0+l of the union window and walks the rows top-down with
// a 16-row register ring of its 16-byte reference strip (5 aligned loads + 4 funnel shifts per row).  Per candidate
// position the warp issues the 64 VABSDIFF4.U8.ACC of the whole macroblock into four accumulators — the 8x8
// quadrant SADs TL,TR,BL,BR — and every partition cost is a sum of those (S/common/pixel.c:40-56), plus that
// partition's own lambda*bits(mv - mvp) (S/encoder/me.c:54-63).  A partition only accepts positions inside ITS
// window (seed +- me_range clipped to the MV limits, width rounded up to 4: me.c:451-457); outside it the cost is
// forced above every real cost.  Keys (cost<<12|row) make min() the reference's strict-'<' raster-order update.
//
// Roofline class: integer ALU pipe.  Per candidate position: 64 SAD ops + ~30 ALU ops of bookkeeping serve nine
// (partition, mv) candidates of the reference's search space.
#include "common.cuh"
#include <cstdlib>

namespace {

struct Geo { const uint8_t *fenc; const uint8_t *fref; int stride; };

#define MB_WARPS 4
#define MB_MAX_RANGE 64   // larger ranges go through the per-block kernel (x264_cuda_me_search)
#define MB_MAX_UW 128     // union window limits; beyond them partitions are searched one at a time
#define MB_UNION_SLACK 31 // rows the union may exceed a single window's 2*me_range+1 by (the y-cost table is sized for that)
#define NP X264_CUDA_ME_MB_PARTS
#define NCAND (X264_CUDA_ME_MB_MVC + 2)
#define INVALID_COST 0x3ffff // above any real cost (<= 65280 + 2*~2.6k); three of them still fit the key's 20 cost bits

// key layout: cost (up to 20 bits) << 10 | row (8 bits) << 2 | column chunk (2 bits).  min() over keys is the
// reference's strict-'<' update in raster order at column-chunk granularity; the column inside the chunk is the
// lane's own, resolved by the final warp reduction.
#define KEY_SHIFT 10

struct __align__(16) WarpSmem {
    uint32_t F[16][4];                       // fenc macroblock, 256 B (read as uint4 rows: keep 16-byte aligned)
    x264_cuda_me_mb_job_t job;               // 280 B
    int quad[NP * NCAND][4];                 // predictor stage: 8x8 quadrant SADs of every (partition, candidate)
    int pc_x[NP * NCAND], pc_y[NP * NCAND];
    int ext[8][3];                           // partition 0's predictors 4..10: cost (or -1), x, y
    int seed[NP][3];                         // bmx, bmy, bcost per partition
    int win[NP][4];                          // min_x, min_y, width, rows per partition
};

__device__ __forceinline__ void load_row16(uint32_t (&dst)[4], const uint8_t *p, int sh)
{
    uint32_t w[5];
#pragma unroll
    for (int i = 0; i < 5; i++) w[i] = __ldg((const uint32_t *)p + i);
#pragma unroll
    for (int i = 0; i < 4; i++) dst[i] = __funnelshift_r(w[i], w[i + 1], sh);
}
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// quadrants (bit q: TL,TR,BL,BR) making up partition p
__device__ __forceinline__ unsigned part_quads(int p)
{
    return (unsigned)(0x8421a5c3full >> (4 * p)) & 15; // p0:1111 p1:0011 p2:1100 p3:0101 p4:1010 p5:0001 p6:0010 p7:0100 p8:1000
}

// pair k of the 16 existing (partition, quadrant) pairs: p0 owns q0..3; p1: q0,q1; p2: q2,q3; p3: q0,q2; p4: q1,q3; p5..8: q(p-5)
__device__ __forceinline__ void pair_pq(int k, int &p, int &q)
{
    p = (int)((0x8765443322110000ull >> (4 * k)) & 15);
    q = (int)((0xe4d8e4e4u >> (2 * k)) & 3);
}

// SAD of the 8x8 fenc quadrant q against the reference bytes at `a` (any alignment); fully unrolled, uniform work
__device__ __forceinline__ int sad_quad(const uint32_t (*F)[4], int q, const uint8_t *a, int stride)
{
    const int sh = ((uintptr_t)a & 3) * 8;
    const uint8_t *p = (const uint8_t *)((uintptr_t)a & ~(uintptr_t)3);
    const int w0 = (q & 1) * 2, y0 = (q >> 1) * 8;
    uint32_t acc0 = 0, acc1 = 0;
#pragma unroll
    for (int y = 0; y < 8; y++) {
        const uint32_t *row = (const uint32_t *)(p + (size_t)y * stride);
        const uint32_t a0 = __ldg(row), a1 = __ldg(row + 1), a2 = __ldg(row + 2);
        acc0 = sad4_acc(F[y0 + y][w0], __funnelshift_r(a0, a1, sh), acc0);
        acc1 = sad4_acc(F[y0 + y][w0 + 1], __funnelshift_r(a1, a2, sh), acc1);
    }
    return (int)(acc0 + acc1);
}

// Partition 0 with more than four extra predictors (X264_CUDA_ME_MB_MVC16): predictors 4.. live in slot 3 of partitions 1..7.  Their
// 16x16 SADs are computed here (lane = predictor x quadrant) and lane 0 redoes the reference's sequential strict-'<' selection in its
// order — mvp, mvc[0..], then (0,0) (me.c:207-229) — from the quadrant sums the main stage left in S.quad.  Kept out of line so that the
// common path's code is unaffected.
__device__ __noinline__ void seed_p0_wide(WarpSmem &S, const int16_t *tab, const uint8_t *ref0, int stride, int n_ext, int x_min, int x_max,
                                          int y_min, int y_max, int lane)
{
    const x264_cuda_me_mb_job_t &job = S.job;
    const int k = lane >> 2, q = lane & 3;
    const int mx = (job.mvc[1 + min(k, 6)][X264_CUDA_ME_MB_MVC - 1][0] + 2) >> 2, my = (job.mvc[1 + min(k, 6)][X264_CUDA_ME_MB_MVC - 1][1] + 2) >> 2;
    const bool valid = k < n_ext && (mx | my) != 0;
    const int cx = clip3i(mx, x_min, x_max), cy = clip3i(my, y_min, y_max);
    int v = valid ? sad_quad(S.F, q, ref0 + (ptrdiff_t)((q >> 1) * 8 + cy) * stride + (q & 1) * 8 + cx, stride) : 0;
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    __syncwarp();
    if (q == 0) {
        S.ext[k][0] = valid ? v + tab[(cx << 2) - job.mvp[0][0]] + tab[(cy << 2) - job.mvp[0][1]] : -1;
        S.ext[k][1] = cx; S.ext[k][2] = cy;
    }
    __syncwarp();
    if (lane == 0) {
        int bc = COST_MAX + 1, bx = 0, by = 0;
        for (int c = 0; c < NCAND; c++) {
            if (c == NCAND - 1)
                for (int e = 0; e < n_ext; e++)
                    if (S.ext[e][0] >= 0 && S.ext[e][0] < bc) { bc = S.ext[e][0]; bx = S.ext[e][1]; by = S.ext[e][2]; }
            if (S.pc_y[c] == (1 << 20)) continue;
            int w = S.quad[c][0] + S.quad[c][1] + S.quad[c][2] + S.quad[c][3];
            if (c != 0) w += tab[(S.pc_x[c] << 2) - job.mvp[0][0]] + tab[(S.pc_y[c] << 2) - job.mvp[0][1]];
            if (w < bc) { bc = w; bx = S.pc_x[c]; by = S.pc_y[c]; }
        }
        S.seed[0][0] = bx; S.seed[0][1] = by; S.seed[0][2] = bc;
    }
    __syncwarp();
}

// The exhaustive pass over the union window of the partitions in `mask`.  Returns per-lane best keys in best[].
// cyt: per union row the 9 partition y-costs, pre-shifted, | row<<2 (+3 pad) — dynamic smem of 2*me_range+1+MB_UNION_SLACK+4
// rows (a small table leaves the L1 to the window tiles)
__device__ __forceinline__ void scan_union(WarpSmem &S, uint32_t (*cyt)[12], const int16_t *tab, const uint8_t *ref0, int stride, unsigned mask,
                                           int ux0, int uy0, int uwidth, int urows, int lane, uint32_t (&best)[NP])
{
    const x264_cuda_me_mb_job_t &job = S.job;
    // y-cost table for the whole union: cyt[r][p] = (cost_y << KEY_SHIFT) | r << 2, INVALID outside p's row range.
    // Lane -> (partition p = lane % 9, row phase lane / 9); four rows per trip so that the table loads overlap.
    __syncwarp();
    if (lane < 27) {
        const int p = lane % 9, ph = lane / 9;
        const bool on = mask >> p & 1;
        const int wy0 = S.win[p][1] - uy0, wy1 = wy0 + S.win[p][3];
        const int16_t *ty = tab + ((uy0 << 2) - job.mvp[p][1]);
        for (int r0 = ph; r0 < urows + 3; r0 += 12) {
            uint32_t v[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int r = r0 + 3 * k;
                v[k] = (on && r >= wy0 && r < wy1) ? (uint32_t)ty[r << 2] : INVALID_COST;
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int r = r0 + 3 * k;
                if (r < urows + 3) cyt[r][p] = (v[k] << KEY_SHIFT) | ((uint32_t)min(r, 255) << 2);
            }
        }
    }
    __syncwarp();
    const uint4 *F4 = (const uint4 *)&S.F[0][0];
    for (int c0 = 0, chunk = 0; c0 < uwidth; c0 += 32, chunk++) {
        // lane tiling of this column chunk: a full chunk is 32 columns x 1 row segment; a narrow tail chunk is
        // folded to 16x2 or 8x4 (columns x row segments) so that all 32 lanes stay busy
        const int rem = uwidth - c0;
        const int cw = rem > 16 ? 32 : rem > 8 ? 16 : 8, segs = 32 / cw;
        const int lcol = lane & (cw - 1), seg = lane / cw;
        const int col = c0 + lcol;
        const int mx = ux0 + min(col, uwidth - 1);
        const int seg_rows = (urows + segs - 1) / segs;
        uint32_t cxp[NP]; // (x-cost << KEY_SHIFT) | chunk, INVALID outside the partition's column range
#pragma unroll
        for (int p = 0; p < NP; p++) {
            const int wx0 = S.win[p][0], ww = S.win[p][2];
            const bool in = (mask >> p & 1) && col < uwidth && mx >= wx0 && mx < wx0 + ww;
            cxp[p] = ((in ? (uint32_t)tab[(mx << 2) - job.mvp[p][0]] : INVALID_COST) << KEY_SHIFT) | (uint32_t)chunk;
        }
        // pull the window tile into L1: (urows+15+pad) rows x <=51 bytes, at most two 128-byte lines per row
        const uint8_t *t0 = ref0 + (ptrdiff_t)uy0 * stride + ux0 + c0;
        for (int r = lane; r < seg_rows * segs + 15; r += 32) {
            prefetch_l1(t0 + (size_t)r * stride);
            prefetch_l1(t0 + (size_t)r * stride + cw + 16);
        }
        const int rbeg = seg * seg_rows;
        const uint8_t *a = ref0 + (ptrdiff_t)(uy0 + rbeg) * stride + mx;
        const int sh = ((uintptr_t)a & 3) * 8;
        const uint8_t *pr = (const uint8_t *)((uintptr_t)a & ~(uintptr_t)3);
        uint32_t R[16][4];
#pragma unroll
        for (int y = 0; y < 15; y++) load_row16(R[y], pr + (size_t)y * stride, sh);
        for (int base = 0; base < seg_rows; base += 16) {
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const int r = base + j;
                if (r >= seg_rows) break; // warp-uniform
                load_row16(R[(j + 15) % 16], pr + (size_t)(r + 15) * stride, sh);
                const uint4 *cy4 = (const uint4 *)&cyt[rbeg + r][0];
                const uint4 ca = cy4[0], cb = cy4[1], cc = cy4[2];
                // eight independent accumulator chains (two per 8x8 quadrant) keep the ALU pipe fed
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
                const uint32_t top = tl + tr, bot = bl + br, lft = tl + bl, rgt = tr + br, all = top + bot;
                // key = sad*2^KEY_SHIFT + (cx<<KEY_SHIFT|chunk) + (cy<<KEY_SHIFT|row<<2): one IMAD (FMA pipe) + IADD + MIN
#define UPD(p, sad, cy) best[p] = min(best[p], __umul24((sad), 1u << KEY_SHIFT) + cxp[p] + (cy))
                UPD(0, all, ca.x); UPD(1, top, ca.y); UPD(2, bot, ca.z); UPD(3, lft, ca.w); UPD(4, rgt, cb.x);
                UPD(5, tl, cb.y); UPD(6, tr, cb.z); UPD(7, bl, cb.w); UPD(8, br, cc.x);
#undef UPD
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Second lane mapping (the default): TWO LANES PER CANDIDATE COLUMN.  Lane l = (half h = l >> 4, column l & 15): half 0 owns the left
// eight pixel columns of the macroblock, half 1 the right eight.  A lane keeps an 8-byte reference strip (16-row ring: 32 registers) and
// its half of the source macroblock (32 registers) instead of 64 + 64, which brings the kernel from 168 to ~100 registers: five CTAs
// instead of three per SM, i.e. 20 instead of 12 warps to hide the load and fixed-latency stalls the profile showed.  Per position
// a lane issues 32 VABSDIFF4 into its two quadrant sums (top, bottom of its half), packs them into one word, swaps it with its partner
// lane (one SHFL) and then owns a share of the nine partitions: half 0 updates {16x16, 16x8 top, 8x16 left, TL, BL}, half 1
// {-, 16x8 bottom, 8x16 right, TR, BR} — the same five expressions on both halves with per-lane cost terms.
#define KEY_SHIFT2 11 // key: cost (20 bits) << 11 | row (8 bits) << 3 | 16-column chunk (3 bits)
__device__ __forceinline__ int half_slot_part(int h, int s) // partition updated by slot s of half h (-1: none)
{
    return h ? (s == 0 ? -1 : 2 * s) : (s == 0 ? 0 : 2 * s - 1); // h0: 0,1,3,5,7   h1: -,2,4,6,8
}
__device__ __forceinline__ void load_row8(uint32_t (&dst)[2], const uint8_t *p, int sh)
{
    const uint32_t w0 = __ldg((const uint32_t *)p), w1 = __ldg((const uint32_t *)p + 1), w2 = __ldg((const uint32_t *)p + 2);
    dst[0] = __funnelshift_r(w0, w1, sh);
    dst[1] = __funnelshift_r(w1, w2, sh);
}
// cyt2: per union row 2 x 8 words: [h][slot] = (y cost of the slot's partition << KEY_SHIFT2) | row << 3, INVALID outside its row range
__device__ __forceinline__ void scan_union2(WarpSmem &S, uint32_t (*cyt2)[16], const int16_t *tab, const uint8_t *ref0, int stride, unsigned mask,
                                            int ux0, int uy0, int uwidth, int urows, int lane, uint32_t (&best)[5])
{
    const x264_cuda_me_mb_job_t &job = S.job;
    const int h = lane >> 4, ci = lane & 15;
    __syncwarp();
    if (lane < 30) {
        const int e = lane % 10, ph = lane / 10, eh = e / 5, es = e % 5;
        const int p = half_slot_part(eh, es);
        const bool on = p >= 0 && (mask >> p & 1);
        const int pp = max(p, 0);
        const int wy0 = S.win[pp][1] - uy0, wy1 = wy0 + S.win[pp][3];
        const int16_t *ty = tab + ((uy0 << 2) - job.mvp[pp][1]);
        for (int r0 = ph; r0 < urows + 3; r0 += 12) {
            uint32_t v[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int r = r0 + 3 * k;
                v[k] = (on && r >= wy0 && r < wy1) ? (uint32_t)ty[r << 2] : INVALID_COST;
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int r = r0 + 3 * k;
                if (r < urows + 3) cyt2[r][eh * 8 + es] = (v[k] << KEY_SHIFT2) | ((uint32_t)min(r, 255) << 3);
            }
        }
    }
    __syncwarp();
    // this lane's half of the source macroblock
    uint32_t F[16][2];
#pragma unroll
    for (int y = 0; y < 16; y++) { F[y][0] = S.F[y][2 * h]; F[y][1] = S.F[y][2 * h + 1]; }
    for (int c0 = 0, chunk = 0; c0 < uwidth; c0 += 16, chunk++) {
        // a full chunk is 16 columns x 1 row segment; the narrow tail chunk is folded to 8 columns x 2 row segments
        const int rem = uwidth - c0;
        const int cw = rem > 8 ? 16 : 8, segs = 16 / cw;
        const int lcol = ci & (cw - 1), seg = ci / cw;
        const int col = c0 + lcol;
        const int mx = ux0 + min(col, uwidth - 1);
        const int seg_rows = (urows + segs - 1) / segs;
        uint32_t cxp[5];
#pragma unroll
        for (int sl = 0; sl < 5; sl++) {
            const int p = half_slot_part(h, sl), pp = max(p, 0);
            const int wx0 = S.win[pp][0], ww = S.win[pp][2];
            const bool in = p >= 0 && (mask >> pp & 1) && col < uwidth && mx >= wx0 && mx < wx0 + ww;
            cxp[sl] = ((in ? (uint32_t)tab[(mx << 2) - job.mvp[pp][0]] : INVALID_COST) << KEY_SHIFT2) | (uint32_t)chunk;
        }
        const uint8_t *t0 = ref0 + (ptrdiff_t)uy0 * stride + ux0 + c0;
        for (int r = lane; r < seg_rows * segs + 15; r += 32) {
            prefetch_l1(t0 + (size_t)r * stride);
            prefetch_l1(t0 + (size_t)r * stride + cw + 16);
        }
        const int rbeg = seg * seg_rows;
        const uint8_t *a = ref0 + (ptrdiff_t)(uy0 + rbeg) * stride + mx + 8 * h;
        const int sh = ((uintptr_t)a & 3) * 8;
        const uint8_t *pr = (const uint8_t *)((uintptr_t)a & ~(uintptr_t)3);
        uint32_t R[16][2];
#pragma unroll
        for (int y = 0; y < 15; y++) load_row8(R[y], pr + (size_t)y * stride, sh);
        for (int base = 0; base < seg_rows; base += 16) {
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const int r = base + j;
                if (r >= seg_rows) break; // warp-uniform
                load_row8(R[(j + 15) % 16], pr + (size_t)(r + 15) * stride, sh);
                const uint4 ca = *(const uint4 *)&cyt2[rbeg + r][h * 8];
                const uint32_t cb = cyt2[rbeg + r][h * 8 + 4];
                uint32_t t0a = 0, t1a = 0, b0a = 0, b1a = 0;
#pragma unroll
                for (int y = 0; y < 8; y++) {
                    t0a = sad4_acc(F[y][0], R[(j + y) % 16][0], t0a);
                    b0a = sad4_acc(F[y + 8][0], R[(j + y + 8) % 16][0], b0a);
                    t1a = sad4_acc(F[y][1], R[(j + y) % 16][1], t1a);
                    b1a = sad4_acc(F[y + 8][1], R[(j + y + 8) % 16][1], b1a);
                }
                const uint32_t t = t0a + t1a, b = b0a + b1a;
                const uint32_t o = __shfl_xor_sync(0xffffffffu, t | (b << 16), 16); // the partner half's (top, bottom)
                const uint32_t ot = o & 0xffffu, ob = o >> 16;
                const uint32_t own = t + b, tt = t + ot, bb = b + ob, all = tt + bb;
                const uint32_t s1 = h ? bb : tt;
#define UPD2(sl, sad, cy) best[sl] = min(best[sl], __umul24((sad), 1u << KEY_SHIFT2) + cxp[sl] + (cy))
                UPD2(0, all, ca.x); UPD2(1, s1, ca.y); UPD2(2, own, ca.z); UPD2(3, t, ca.w); UPD2(4, b, cb);
#undef UPD2
            }
        }
    }
}

template <bool HALF>
__global__ void __launch_bounds__(MB_WARPS * 32, HALF ? 5 : 3)
me_search_mb_kernel(Geo geo, const x264_cuda_me_mb_job_t *__restrict__ jobs, int n_jobs,
                    const int16_t *const *__restrict__ cost_tabs, int me_range, int max_ur, int prefetch_dist,
                    x264_cuda_me_mb_result_t *__restrict__ results)
{
    __shared__ __align__(16) WarpSmem s_all[MB_WARPS];
    extern __shared__ __align__(16) uint32_t s_cyt[]; // MB_WARPS x (max_ur + 4) x 12
    const int lane = threadIdx.x & 31;
    WarpSmem &S = s_all[threadIdx.x >> 5];
    uint32_t (*cyt)[12] = (uint32_t (*)[12])(s_cyt + (size_t)(threadIdx.x >> 5) * (max_ur + 4) * 12);    // full-strip mapping
    uint32_t (*cyt2)[16] = (uint32_t (*)[16])(s_cyt + (size_t)(threadIdx.x >> 5) * (max_ur + 4) * 16);   // half-strip mapping
    const int stride = geo.stride;
    const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
    for (int jb = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; jb < n_jobs; jb += warps_per_grid) {
        __syncwarp();
        for (int i = lane; i < (int)(sizeof(x264_cuda_me_mb_job_t) / 4); i += 32)
            ((uint32_t *)&S.job)[i] = __ldg((const uint32_t *)(jobs + jb) + i);
        // the macroblock that will run in this warp's slot roughly one residency later: pull its job and pixels into L2
        // now (first-touch HBM latency is the main cost of the short phases before the scan)
        const int jn = jb + prefetch_dist;
        uint32_t next_pos = 0xffffffffu; // mb_x | mb_y << 16 of that job
        if (jn < n_jobs) {
            next_pos = __ldg((const uint32_t *)(jobs + jn));
            if (lane < 3) prefetch_l2((const uint8_t *)(jobs + jn) + 128 * lane);
        }
        __syncwarp();
        const x264_cuda_me_mb_job_t &job = S.job;
        const int16_t *tab = cost_tabs[job.qp > 51 ? 51 : job.qp] + 2 * 4 * 2048;
        const int x_min = job.mv_min_fpel[0], y_min = job.mv_min_fpel[1];
        const int x_max = job.mv_max_fpel[0], y_max = job.mv_max_fpel[1];
        const unsigned mask = job.part_mask & ((1u << NP) - 1);
        x264_cuda_me_result_t *out = results[jb].part;

        // reject jobs whose table indices could leave p_cost_mv's +-2*4*2048 (analyse.c:196-203) or with bad limits
        int bad = x_min > 0 || x_max < 0 || y_min > 0 || y_max < 0;
        if (lane < NP && (mask >> lane & 1)) {
            const int lim = 2 * 4 * 2048 - 8;
            bad |= max(abs(4 * x_min - job.mvp[lane][0]), abs(4 * x_max - job.mvp[lane][0])) > lim;
            bad |= max(abs(4 * y_min - job.mvp[lane][1]), abs(4 * y_max - job.mvp[lane][1])) > lim;
        }
        // partition 0 may carry up to 4 + 7 predictors; extra k occupies slot 3 of partition 1 + k, which then must not use it itself
        const int n_ext = min(max((int)job.i_mvc[0] - X264_CUDA_ME_MB_MVC, 0), X264_CUDA_ME_MB_MVC16_EXTRA);
        if (lane >= 1 && lane < NP && lane - 1 < n_ext) bad |= job.i_mvc[lane] > X264_CUDA_ME_MB_MVC - 1;
        if (__any_sync(0xffffffffu, bad) || !mask) {
            if (lane < NP) { x264_cuda_me_result_t r = { 0, 0, -1, 0, 0, -1 }; out[lane] = r; }
            continue;
        }
        const uint8_t *fe = geo.fenc + (size_t)job.mb_y * 16 * stride + job.mb_x * 16;
        const uint8_t *ref0 = geo.fref + (size_t)job.mb_y * 16 * stride + job.mb_x * 16;
        for (int i = lane; i < 64; i += 32) S.F[i >> 2][i & 3] = __ldg((const uint32_t *)(fe + (size_t)(i >> 2) * stride) + (i & 3));
        __syncwarp();

        // ---- predictor stage per partition (me.c:207-229).  Every (partition, candidate) cost is a sum of 8x8
        // quadrant SADs at that candidate's position: 54 items x up to 4 quadrants, one uniform 8x8 SAD per lane.
        if (!(job.flags & X264_CUDA_ME_SEEDED)) {
            for (int idx = lane; idx < NP * NCAND; idx += 32) { // candidate positions
                const int p = idx / NCAND, c = idx - p * NCAND;
                int cx = 0, cy = 0, valid = (mask >> p & 1);
                const int n_mvc = min((int)job.i_mvc[p], X264_CUDA_ME_MB_MVC);
                if (c == 0) {
                    cx = (clip3i(job.mvp[p][0], x_min * 4, x_max * 4) + 2) >> 2;
                    cy = (clip3i(job.mvp[p][1], y_min * 4, y_max * 4) + 2) >> 2;
                } else if (c <= X264_CUDA_ME_MB_MVC) {
                    const int mx = (job.mvc[p][c - 1][0] + 2) >> 2, my = (job.mvc[p][c - 1][1] + 2) >> 2;
                    valid = valid && (c - 1 < n_mvc) && (mx | my) != 0; // zero predictors: covered by the (0,0) test
                    cx = clip3i(mx, x_min, x_max); cy = clip3i(my, y_min, y_max);
                }
                S.pc_x[idx] = cx; S.pc_y[idx] = valid ? cy : (1 << 20); // pc_y == 1<<20 marks "skip"
            }
            __syncwarp();
            // (candidate, partition, quadrant) tasks: only the 16 (partition, quadrant) pairs that exist — 6 x 16 = 96 tasks,
            // three per lane, each one uniform 8x8 SAD; quadrants a partition does not own stay 0 (S.quad is cleared first)
            for (int i = lane; i < NP * NCAND * 4; i += 32) (&S.quad[0][0])[i] = 0;
            __syncwarp();
            int v[3];
#pragma unroll
            for (int u = 0; u < 3; u++) {
                const int t = lane + 32 * u, c = t >> 4;
                int p, q;
                pair_pq(t & 15, p, q);
                const int idx = p * NCAND + c;
                v[u] = 0;
                if (S.pc_y[idx] != (1 << 20))
                    v[u] = sad_quad(S.F, q, ref0 + (ptrdiff_t)((q >> 1) * 8 + S.pc_y[idx]) * stride + (q & 1) * 8 + S.pc_x[idx], stride);
            }
#pragma unroll
            for (int u = 0; u < 3; u++) {
                const int t = lane + 32 * u, c = t >> 4;
                int p, q;
                pair_pq(t & 15, p, q);
                S.quad[p * NCAND + c][q] = v[u];
            }
            __syncwarp();
            if (lane < NP) {
                int bc = COST_MAX + 1, bx = 0, by = 0; // sequential strict '<' in candidate order
                for (int c = 0; c < NCAND; c++) {
                    const int idx = lane * NCAND + c;
                    if (S.pc_y[idx] == (1 << 20)) continue;
                    int v = S.quad[idx][0] + S.quad[idx][1] + S.quad[idx][2] + S.quad[idx][3];
                    if (c != 0) v += tab[(S.pc_x[idx] << 2) - job.mvp[lane][0]] + tab[(S.pc_y[idx] << 2) - job.mvp[lane][1]];
                    if (v < bc) { bc = v; bx = S.pc_x[idx]; by = S.pc_y[idx]; }
                }
                S.seed[lane][0] = bx; S.seed[lane][1] = by; S.seed[lane][2] = bc;
            }
            if (n_ext > 0 && (mask & 1)) seed_p0_wide(S, tab, ref0, stride, n_ext, x_min, x_max, y_min, y_max, lane); // warp-uniform, rare
        } else if (lane < NP) {
            S.seed[lane][0] = clip3i(job.seed_mv[lane][0], x_min, x_max);
            S.seed[lane][1] = clip3i(job.seed_mv[lane][1], y_min, y_max);
            S.seed[lane][2] = job.seed_cost[lane];
        }
        __syncwarp();

        // ---- per-partition windows (me.c:451-457)
        if (lane < NP) {
            int min_x = 0, min_y = 0, width = 0, rows = 0;
            if (mask >> lane & 1) {
                const int bmx = S.seed[lane][0], bmy = S.seed[lane][1];
                min_x = max(bmx - me_range, x_min); min_y = max(bmy - me_range, y_min);
                const int max_x = min(bmx + me_range, x_max), max_y = min(bmy + me_range, y_max);
                width = (max_x - min_x + 3) & ~3; rows = max_y - min_y + 1;
            }
            S.win[lane][0] = min_x; S.win[lane][1] = min_y; S.win[lane][2] = width; S.win[lane][3] = rows;
        }
        __syncwarp();

        // ---- that later macroblock: fenc rows and a reference tile displaced like this macroblock's first window
        if (next_pos != 0xffffffffu) {
            const size_t o = (size_t)(next_pos >> 16) * 16 * stride + (next_pos & 0xffff) * 16;
            if (lane < 16) prefetch_l2(geo.fenc + o + (size_t)lane * stride);
            const uint8_t *t = geo.fref + o + (ptrdiff_t)(S.seed[0][1] - me_range) * stride + (S.seed[0][0] - me_range);
            for (int r = lane; r < 2 * me_range + 16; r += 32) { prefetch_l2(t + (size_t)r * stride); prefetch_l2(t + (size_t)r * stride + 2 * me_range + 15); }
        }

        // ---- exhaustive pass(es): all partitions over the union of their windows when it is compact, else one by one
        unsigned todo = mask;
        int my_bmx = 0, my_bmy = 0, my_cost = INVALID_COST; // lane p keeps partition p's window winner
        while (todo) {
            int ux0 = 1 << 20, uy0 = 1 << 20, ux1 = -(1 << 20), uy1 = -(1 << 20);
            if (lane < NP && (todo >> lane & 1)) {
                ux0 = S.win[lane][0]; uy0 = S.win[lane][1]; ux1 = ux0 + S.win[lane][2]; uy1 = uy0 + S.win[lane][3];
            }
            ux0 = __reduce_min_sync(0xffffffffu, ux0); uy0 = __reduce_min_sync(0xffffffffu, uy0);
            ux1 = __reduce_max_sync(0xffffffffu, ux1); uy1 = __reduce_max_sync(0xffffffffu, uy1);
            unsigned group = todo;
            if (ux1 - ux0 > MB_MAX_UW || uy1 - uy0 > max_ur) { // scattered predictors: search the lowest partition alone
                const int p = __ffs(todo) - 1;
                group = 1u << p;
                ux0 = S.win[p][0]; uy0 = S.win[p][1]; ux1 = ux0 + S.win[p][2]; uy1 = uy0 + S.win[p][3];
            }
            todo &= ~group;
            const int uwidth = ux1 - ux0;
            if constexpr (HALF) {
                uint32_t best[5];
#pragma unroll
                for (int sl = 0; sl < 5; sl++) best[sl] = 0xffffffffu;
                scan_union2(S, cyt2, tab, ref0, stride, group, ux0, uy0, uwidth, uy1 - uy0, lane, best);
                // argmin per partition over the 16 lanes of the half that owns it: min key (cost,row,chunk), then the lowest column
#pragma unroll
                for (int p = 0; p < NP; p++) {
                    if (!(group >> p & 1)) continue; // warp-uniform
                    const int hp = p == 0 ? 0 : ((p - 1) & 1), sp = p == 0 ? 0 : (p + 1) / 2;
                    const bool mine = (lane >> 4) == hp;
                    const uint32_t k = __reduce_min_sync(0xffffffffu, mine ? best[sp] : 0xffffffffu);
                    const int chunk = k & 7, rem = uwidth - chunk * 16;
                    const int cw = rem > 8 ? 16 : 8;
                    const uint32_t c = __reduce_min_sync(0xffffffffu, (mine && best[sp] == k) ? (uint32_t)(lane & (cw - 1)) : 0xffffffffu);
                    if (lane == p) {
                        my_cost = (int)(k >> KEY_SHIFT2);
                        my_bmy = uy0 + (int)((k >> 3) & 255);
                        my_bmx = ux0 + chunk * 16 + (int)c;
                    }
                }
            } else {
                uint32_t best[NP];
#pragma unroll
                for (int p = 0; p < NP; p++) best[p] = 0xffffffffu;
                scan_union(S, cyt, tab, ref0, stride, group, ux0, uy0, uwidth, uy1 - uy0, lane, best);
                // warp argmin per partition: min key (cost,row,chunk), then the lowest in-chunk column among its holders
#pragma unroll
                for (int p = 0; p < NP; p++) {
                    if (!(group >> p & 1)) continue; // warp-uniform
                    const uint32_t k = __reduce_min_sync(0xffffffffu, best[p]);
                    const int chunk = k & 3, rem = uwidth - chunk * 32;
                    const int cw = rem > 16 ? 32 : rem > 8 ? 16 : 8;
                    const uint32_t c = __reduce_min_sync(0xffffffffu, best[p] == k ? (uint32_t)(lane & (cw - 1)) : 0xffffffffu);
                    if (lane == p) {
                        my_cost = (int)(k >> KEY_SHIFT);
                        my_bmy = uy0 + (int)((k >> 2) & 255);
                        my_bmx = ux0 + chunk * 32 + (int)c;
                    }
                }
            }
        }
        if (lane < NP) {
            x264_cuda_me_result_t r;
            int bmx = S.seed[lane][0], bmy = S.seed[lane][1], bcost = S.seed[lane][2];
            r.seed_mx = (int16_t)bmx; r.seed_my = (int16_t)bmy; r.seed_cost = bcost;
            if ((mask >> lane & 1) && my_cost < INVALID_COST && my_cost < bcost) { bcost = my_cost; bmy = my_bmy; bmx = my_bmx; }
            if (!(mask >> lane & 1)) { bmx = bmy = 0; bcost = -1; r.seed_mx = r.seed_my = 0; r.seed_cost = -1; }
            r.bmx = (int16_t)bmx; r.bmy = (int16_t)bmy; r.bcost = bcost;
            out[lane] = r;
        }
    }
}


// =====================================================================================================================
// Third form (the default): WARP-SPECIALISED PERSISTENT KERNEL.  The profile of the one-warp-does-everything kernel shows a warp spending
// ~60 % of its residency in the short latency-bound phases before and after the scan (job, predictor SADs, cost-table gathers, ring
// preload) while holding the 168 registers only the scan needs, so the ALU pipe idles half the time at 3 warps per scheduler.  Here one CTA
// per SM runs 4 PRODUCER warps on 56 registers and 8 SCAN warps on 224 (setmaxnreg moves the registers between the warpgroups):
//   producer : takes the next macroblock from a global ticket, runs the predictor stage, seeds, windows, pass grouping and builds the
//              per-pass tables — y-cost word per (row, partition), x-cost word per (chunk, partition, lane) — into a shared-memory slot;
//   scanner  : waits for a full slot (mbarrier), scans the union window exactly like scan_union above with every table coming from the
//              slot, reduces, writes the result and hands the slot back.  It never waits on a global-memory dependency chain other than
//              its own reference rows, which it prefetches into L1 one column chunk / one slot ahead.
// Each producer feeds two scanners, two slots each (ring).  Results are identical by construction: same windows, same keys.
#define V3_PROD 8      // producer warps (setmaxnreg works on whole warpgroups of four)
#define V3_CONS 8      // scan warps: two per scheduler — the scan loop alone saturates the math issue port at that (scratch/scanbench.cu)
#define V3_FEED 1      // scanners fed by one producer
#define V3_SLOTS 2
#define V3_MAX_UW 64   // two 32-column chunks per pass; wider unions are split into per-partition passes
#define V3_UNION_SLACK 15 // rows a pass's union may exceed 2*me_range+1 by
#define V3_PROD_REGS 48 // 8 x 32 x 48 + 8 x 32 x 208 = 65536 = the 128 registers x 512 threads the CTA is launched with
#define V3_CONS_REGS 208
#define V3_FREG_ROWS 16 // source-macroblock rows a scanner keeps in registers; the others are re-read from the slot (uniform LDS.128)

struct __align__(16) SlotHead {
    uint32_t F[16][4];   // source macroblock
    int seed[NP][3];     // bmx, bmy, bcost after the predictor stage
    int ux0, uy0, uwidth, urows;
    unsigned group, mask; // partitions of this pass / of the macroblock
    int jb;              // job index; < 0: no more work for this scanner
    int first, last;     // first / last pass of the macroblock
    int mb_x, mb_y, pad;
};
#define V3_CXP_BYTES (2 * NP * 32 * 2) // x costs as uint16 (0xffff: outside the partition's window)
#define V3_TILE_PITCH 112              // bytes per staged reference row: 15 (alignment) + V3_MAX_UW + 19 (16-byte strip + funnel word) rounded up to 16
#define V3_ROW_PAD 4                   // candidate rows past the union a lane may touch: 3 from folding a narrow chunk into row segments + 1 from row pairs
#define V3_TILE_EXTRA (15 + V3_ROW_PAD) // staged rows below the union: 15 of the macroblock + the pad
__host__ __device__ inline size_t v3_cyt_bytes(int max_ur) { return (size_t)(max_ur + V3_ROW_PAD) * 12 * 4; }
__host__ __device__ inline size_t v3_slot_bytes(int max_ur)
{
    return sizeof(SlotHead) + V3_CXP_BYTES + v3_cyt_bytes(max_ur) + (size_t)(max_ur + V3_TILE_EXTRA) * V3_TILE_PITCH;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint64_t *b)
{
    asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t *b, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) { while (!mbar_test(b, parity)) { } }
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t *b, uint32_t parity) { while (!mbar_test(b, parity)) __nanosleep(256); } // a waiter that is ahead
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }
// a shared-memory read the compiler must repeat where it is written (it would otherwise hoist the 64 source words into registers)
__device__ __forceinline__ uint4 lds128(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

// the lane tiling of a column chunk (shared by the table builder and the scan): a full chunk is 32 columns x 1 row segment,
// a narrow tail chunk is folded to 16x2 or 8x4 (columns x row segments)
__device__ __forceinline__ void chunk_tiling(int uwidth, int c0, int lane, int &cw, int &segs, int &lcol, int &seg)
{
    const int rem = uwidth - c0;
    cw = rem > 16 ? 32 : rem > 8 ? 16 : 8; segs = 32 / cw;
    lcol = lane & (cw - 1); seg = lane / cw;
}

// L1 prefetch of a chunk's window tile: (rows + 15) rows x (cw + 15) bytes, at most two 128-byte lines per row
__device__ __forceinline__ void prefetch_tile(const uint8_t *ref0, int stride, int ux0, int uy0, int uwidth, int urows, int c0, int lane)
{
    int cw, segs, lcol, seg;
    chunk_tiling(uwidth, c0, lane, cw, segs, lcol, seg);
    const int seg_rows = (urows + segs - 1) / segs;
    const uint8_t *t0 = ref0 + (ptrdiff_t)uy0 * stride + ux0 + c0;
    for (int r = lane; r < seg_rows * segs + 15; r += 32) {
        prefetch_l1(t0 + (size_t)r * stride);
        prefetch_l1(t0 + (size_t)r * stride + cw + 16);
    }
}

// ---- producer: everything of one macroblock up to its per-partition windows (the legacy kernel's first half); returns the partition
// mask, 0 when the job was rejected (its result is written here)
__device__ __forceinline__ unsigned v3_setup(WarpSmem &S, const Geo &geo, const x264_cuda_me_mb_job_t *__restrict__ jobs, int jb,
                                             const int16_t *const *__restrict__ cost_tabs, int me_range,
                                             x264_cuda_me_mb_result_t *__restrict__ results, int lane, const int16_t *&tab, const uint8_t *&ref0)
{
    const int stride = geo.stride;
    __syncwarp();
    for (int i = lane; i < (int)(sizeof(x264_cuda_me_mb_job_t) / 4); i += 32) ((uint32_t *)&S.job)[i] = __ldg((const uint32_t *)(jobs + jb) + i);
    __syncwarp();
    const x264_cuda_me_mb_job_t &job = S.job;
    tab = cost_tabs[job.qp > 51 ? 51 : job.qp] + 2 * 4 * 2048;
    const int x_min = job.mv_min_fpel[0], y_min = job.mv_min_fpel[1];
    const int x_max = job.mv_max_fpel[0], y_max = job.mv_max_fpel[1];
    const unsigned mask = job.part_mask & ((1u << NP) - 1);
    int bad = x_min > 0 || x_max < 0 || y_min > 0 || y_max < 0;
    if (lane < NP && (mask >> lane & 1)) {
        const int lim = 2 * 4 * 2048 - 8;
        bad |= max(abs(4 * x_min - job.mvp[lane][0]), abs(4 * x_max - job.mvp[lane][0])) > lim;
        bad |= max(abs(4 * y_min - job.mvp[lane][1]), abs(4 * y_max - job.mvp[lane][1])) > lim;
    }
    const int n_ext = min(max((int)job.i_mvc[0] - X264_CUDA_ME_MB_MVC, 0), X264_CUDA_ME_MB_MVC16_EXTRA);
    if (lane >= 1 && lane < NP && lane - 1 < n_ext) bad |= job.i_mvc[lane] > X264_CUDA_ME_MB_MVC - 1;
    if (__any_sync(0xffffffffu, bad) || !mask) {
        if (lane < NP) { x264_cuda_me_result_t r = { 0, 0, -1, 0, 0, -1 }; results[jb].part[lane] = r; }
        return 0;
    }
    const uint8_t *fe = geo.fenc + (size_t)job.mb_y * 16 * stride + job.mb_x * 16;
    ref0 = geo.fref + (size_t)job.mb_y * 16 * stride + job.mb_x * 16;
    for (int i = lane; i < 64; i += 32) S.F[i >> 2][i & 3] = __ldg((const uint32_t *)(fe + (size_t)(i >> 2) * stride) + (i & 3));
    __syncwarp();
    if (!(job.flags & X264_CUDA_ME_SEEDED)) {
        for (int idx = lane; idx < NP * NCAND; idx += 32) {
            const int p = idx / NCAND, c = idx - p * NCAND;
            int cx = 0, cy = 0, valid = (mask >> p & 1);
            const int n_mvc = min((int)job.i_mvc[p], X264_CUDA_ME_MB_MVC);
            if (c == 0) {
                cx = (clip3i(job.mvp[p][0], x_min * 4, x_max * 4) + 2) >> 2;
                cy = (clip3i(job.mvp[p][1], y_min * 4, y_max * 4) + 2) >> 2;
            } else if (c <= X264_CUDA_ME_MB_MVC) {
                const int mx = (job.mvc[p][c - 1][0] + 2) >> 2, my = (job.mvc[p][c - 1][1] + 2) >> 2;
                valid = valid && (c - 1 < n_mvc) && (mx | my) != 0;
                cx = clip3i(mx, x_min, x_max); cy = clip3i(my, y_min, y_max);
            }
            S.pc_x[idx] = cx; S.pc_y[idx] = valid ? cy : (1 << 20);
        }
        for (int i = lane; i < NP * NCAND * 4; i += 32) (&S.quad[0][0])[i] = 0;
        __syncwarp();
#pragma unroll 1
        for (int u = 0; u < 3; u++) { // 6 candidates x 16 (partition, quadrant) pairs = 96 uniform 8x8 SADs, three per lane
            const int t = lane + 32 * u, c = t >> 4;
            int p, q;
            pair_pq(t & 15, p, q);
            const int idx = p * NCAND + c;
            if (S.pc_y[idx] != (1 << 20))
                S.quad[idx][q] = sad_quad(S.F, q, ref0 + (ptrdiff_t)((q >> 1) * 8 + S.pc_y[idx]) * stride + (q & 1) * 8 + S.pc_x[idx], stride);
        }
        __syncwarp();
        if (lane < NP) {
            int mvc_cost[NCAND]; // all the table gathers first (one round trip), then the sequential strict '<' in candidate order
#pragma unroll
            for (int c = 0; c < NCAND; c++) {
                const int idx = lane * NCAND + c;
                const bool on = c != 0 && S.pc_y[idx] != (1 << 20);
                mvc_cost[c] = on ? tab[(S.pc_x[idx] << 2) - job.mvp[lane][0]] + tab[(S.pc_y[idx] << 2) - job.mvp[lane][1]] : 0;
            }
            int bc = COST_MAX + 1, bx = 0, by = 0;
#pragma unroll
            for (int c = 0; c < NCAND; c++) {
                const int idx = lane * NCAND + c;
                if (S.pc_y[idx] == (1 << 20)) continue;
                const int v = S.quad[idx][0] + S.quad[idx][1] + S.quad[idx][2] + S.quad[idx][3] + mvc_cost[c];
                if (v < bc) { bc = v; bx = S.pc_x[idx]; by = S.pc_y[idx]; }
            }
            S.seed[lane][0] = bx; S.seed[lane][1] = by; S.seed[lane][2] = bc;
        }
        if (n_ext > 0 && (mask & 1)) seed_p0_wide(S, tab, ref0, stride, n_ext, x_min, x_max, y_min, y_max, lane);
    } else if (lane < NP) {
        S.seed[lane][0] = clip3i(job.seed_mv[lane][0], x_min, x_max);
        S.seed[lane][1] = clip3i(job.seed_mv[lane][1], y_min, y_max);
        S.seed[lane][2] = job.seed_cost[lane];
    }
    __syncwarp();
    if (lane < NP) {
        int min_x = 0, min_y = 0, width = 0, rows = 0;
        if (mask >> lane & 1) {
            const int bmx = S.seed[lane][0], bmy = S.seed[lane][1];
            min_x = max(bmx - me_range, x_min); min_y = max(bmy - me_range, y_min);
            const int max_x = min(bmx + me_range, x_max), max_y = min(bmy + me_range, y_max);
            width = (max_x - min_x + 3) & ~3; rows = max_y - min_y + 1;
        }
        S.win[lane][0] = min_x; S.win[lane][1] = min_y; S.win[lane][2] = width; S.win[lane][3] = rows;
    }
    __syncwarp();
    return mask;
}

// ---- producer: the tables of one pass into a slot
__device__ __forceinline__ void v3_fill(SlotHead &H, uint16_t *cxp, uint32_t (*cyt)[12], uint8_t *tile, const WarpSmem &S, const int16_t *tab, const uint8_t *ref0,
                                        int stride, int jb, unsigned mask, unsigned group, int first, int last, int ux0, int uy0, int uwidth, int urows, int lane)
{
    const x264_cuda_me_mb_job_t &job = S.job;
    for (int i = lane; i < 64; i += 32) (&H.F[0][0])[i] = (&S.F[0][0])[i];
    if (lane < NP * 3) (&H.seed[0][0])[lane] = (&S.seed[0][0])[lane];
    if (lane == 0) {
        H.ux0 = ux0; H.uy0 = uy0; H.uwidth = uwidth; H.urows = urows; H.group = group; H.mask = mask; H.jb = jb; H.first = first; H.last = last;
        H.mb_x = job.mb_x; H.mb_y = job.mb_y;
    }
    {   // stage the pass's reference window in the slot: rows uy0 .. uy0+urows+17, 16-byte units from the aligned-down left edge (cp.async:
        // global -> shared without registers; the copies fly while the cost tables below are gathered)
        const int o = ux0 & 15;
        const uint8_t *g0 = ref0 + (ptrdiff_t)uy0 * stride + (ux0 - o);
        const int units = (o + uwidth + 19 + 15) >> 4, trows = urows + V3_TILE_EXTRA; // units <= 7
        const uint32_t t0 = smem_u32(tile);
        const int u = lane & 7;
        if (u < units)
            for (int r = lane >> 3; r < trows; r += 4) cp_async16(t0 + r * V3_TILE_PITCH + u * 16, g0 + (size_t)r * stride + u * 16);
    }
    if (lane < 27) { // y-cost words: lane -> (partition, row phase 0..2); eight gathers in flight per lane
        const int p = lane % 9, ph = lane / 9;
        const bool on = group >> p & 1;
        const int wy0 = S.win[p][1] - uy0, wy1 = wy0 + S.win[p][3];
        const int16_t *ty = tab + ((uy0 << 2) - job.mvp[p][1]);
        for (int r0 = ph; r0 < urows + V3_ROW_PAD; r0 += 24) {
            uint32_t v[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int r = r0 + 3 * k;
                v[k] = (on && r >= wy0 && r < wy1) ? (uint32_t)ty[r << 2] : INVALID_COST;
            }
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int r = r0 + 3 * k;
                if (r < urows + V3_ROW_PAD) cyt[r][p] = (v[k] << KEY_SHIFT) | ((uint32_t)min(r, 255) << 2);
            }
        }
    }
#pragma unroll 1
    for (int c0 = 0, chunk = 0; c0 < uwidth; c0 += 32, chunk++) { // x costs per (chunk, partition, lane)
        int cw, segs, lcol, seg;
        chunk_tiling(uwidth, c0, lane, cw, segs, lcol, seg);
        const int col = c0 + lcol;
        const int mx = ux0 + min(col, uwidth - 1);
        uint32_t v[NP];
#pragma unroll
        for (int p = 0; p < NP; p++) {
            const int wx0 = S.win[p][0], ww = S.win[p][2];
            const bool in = (group >> p & 1) && col < uwidth && mx >= wx0 && mx < wx0 + ww;
            v[p] = in ? (uint32_t)(uint16_t)tab[(mx << 2) - job.mvp[p][0]] : 0xffffu;
        }
#pragma unroll
        for (int p = 0; p < NP; p++) cxp[(chunk * NP + p) * 32 + lane] = (uint16_t)v[p];
    }
    cp_async_wait_all();
}

// ---- scanner: the exhaustive pass over one slot's union window (scan_union with the tables in shared memory)
__device__ __forceinline__ void load_row16_s(uint32_t (&dst)[4], const uint32_t *p, int sh)
{
    uint32_t w[5];
#pragma unroll
    for (int i = 0; i < 5; i++) w[i] = p[i];
#pragma unroll
    for (int i = 0; i < 4; i++) dst[i] = __funnelshift_r(w[i], w[i + 1], sh);
}
// (measured and rejected: the same extraction on the FMA pipe, hi32(w[i] * 2^(32-sh)) + lo32(w[i+1] * 2^(32-sh)) as IMAD.HI + IMAD, to take the
// four SHF per row off the ALU pipe that VABSDIFF4 saturates — 0.111 ms per 1080p launch against 0.101: IMAD.HI issues at a quarter rate)
__device__ __forceinline__ void v3_scan(const SlotHead &H, const uint16_t *cxp_tab, const uint32_t (*cyt)[12], const uint8_t *tile, int lane, uint32_t (&best)[NP])
{
    const int ux0 = H.ux0, uwidth = H.uwidth, urows = H.urows;
    const uint4 *F4 = (const uint4 *)&H.F[0][0];
    const uint32_t F4s = smem_u32(F4);
    uint4 F[16];
#pragma unroll
    for (int y = 0; y < V3_FREG_ROWS; y++) F[y] = F4[y];
    for (int c0 = 0, chunk = 0; c0 < uwidth; c0 += 32, chunk++) {
        int cw, segs, lcol, seg;
        chunk_tiling(uwidth, c0, lane, cw, segs, lcol, seg);
        const int col = c0 + lcol;
        const int seg_rows = (urows + segs - 1) / segs;
        uint32_t cxp[NP];
#pragma unroll
        for (int p = 0; p < NP; p++) {
            const uint32_t v = cxp_tab[(chunk * NP + p) * 32 + lane];
            cxp[p] = ((v == 0xffffu ? INVALID_COST : v) << KEY_SHIFT) | (uint32_t)chunk;
        }
        const int rbeg = seg * seg_rows;
        const int xb = (ux0 & 15) + min(col, uwidth - 1); // byte offset of this lane's strip in a staged row
        const int sh = (xb & 3) * 8;
        const uint32_t *pr = (const uint32_t *)(tile + (size_t)rbeg * V3_TILE_PITCH + (xb & ~3));
        uint32_t R[16][4];
#pragma unroll
        for (int y = 0; y < 15; y++) load_row16_s(R[y], pr + y * (V3_TILE_PITCH / 4), sh);
        for (int base = 0; base < seg_rows; base += 16) {
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const int r = base + j;
                if (r >= seg_rows) break; // warp-uniform
                load_row16_s(R[(j + 15) % 16], pr + (r + 15) * (V3_TILE_PITCH / 4), sh);
                const uint4 *cy4 = (const uint4 *)&cyt[rbeg + r][0];
                const uint4 ca = cy4[0], cb = cy4[1], cc = cy4[2];
                uint32_t tl = 0, tr = 0, bl = 0, br = 0, tl2 = 0, tr2 = 0, bl2 = 0, br2 = 0;
#pragma unroll
                for (int y = 0; y < 8; y++) {
                    uint4 f, g;
                    if (y < V3_FREG_ROWS) f = F[y]; else f = lds128(F4s + 16 * y);
                    if (y + 8 < V3_FREG_ROWS) g = F[y + 8]; else g = lds128(F4s + 16 * (y + 8));
                    tl = sad4_acc(f.x, R[(j + y) % 16][0], tl); tr = sad4_acc(f.z, R[(j + y) % 16][2], tr);
                    bl = sad4_acc(g.x, R[(j + y + 8) % 16][0], bl); br = sad4_acc(g.z, R[(j + y + 8) % 16][2], br);
                    tl2 = sad4_acc(f.y, R[(j + y) % 16][1], tl2); tr2 = sad4_acc(f.w, R[(j + y) % 16][3], tr2);
                    bl2 = sad4_acc(g.y, R[(j + y + 8) % 16][1], bl2); br2 = sad4_acc(g.w, R[(j + y + 8) % 16][3], br2);
                }
                tl += tl2; tr += tr2; bl += bl2; br += br2;
                const uint32_t top = tl + tr, bot = bl + br, lft = tl + bl, rgt = tr + br, all = top + bot;
#define UPD(p, sad, cy) best[p] = min(best[p], __umul24((sad), 1u << KEY_SHIFT) + cxp[p] + (cy))
                UPD(0, all, ca.x); UPD(1, top, ca.y); UPD(2, bot, ca.z); UPD(3, lft, ca.w); UPD(4, rgt, cb.x);
                UPD(5, tl, cb.y); UPD(6, tr, cb.z); UPD(7, bl, cb.w); UPD(8, br, cc.x);
#undef UPD
            }
        }
    }
}

// The same scan, TWO CANDIDATE ROWS PER STEP: rows r and r+1 meet every source word, so consecutive SADs share that operand (operand
// reuse: no third register-bank read) and the loop control, table reads and ring loads are paid once per two rows.  18-row ring.
__device__ __forceinline__ void v3_scan2(const SlotHead &H, const uint16_t *cxp_tab, const uint32_t (*cyt)[12], const uint8_t *tile, int lane, uint32_t (&best)[NP])
{
    const int ux0 = H.ux0, uwidth = H.uwidth, urows = H.urows;
    const uint4 *F4 = (const uint4 *)&H.F[0][0];
    uint4 F[16];
#pragma unroll
    for (int y = 0; y < 16; y++) F[y] = F4[y];
    for (int c0 = 0, chunk = 0; c0 < uwidth; c0 += 32, chunk++) {
        int cw, segs, lcol, seg;
        chunk_tiling(uwidth, c0, lane, cw, segs, lcol, seg);
        const int col = c0 + lcol;
        const int seg_rows = (urows + segs - 1) / segs;
        uint32_t cxp[NP];
#pragma unroll
        for (int p = 0; p < NP; p++) {
            const uint32_t v = cxp_tab[(chunk * NP + p) * 32 + lane];
            cxp[p] = ((v == 0xffffu ? INVALID_COST : v) << KEY_SHIFT) | (uint32_t)chunk;
        }
        const int rbeg = seg * seg_rows;
        const int xb = (ux0 & 15) + min(col, uwidth - 1);
        const int sh = (xb & 3) * 8;
        const uint32_t *pr = (const uint32_t *)(tile + (size_t)rbeg * V3_TILE_PITCH + (xb & ~3));
        uint32_t R[18][4];
#pragma unroll
        for (int y = 0; y < 16; y++) load_row16_s(R[y], pr + y * (V3_TILE_PITCH / 4), sh);
        for (int base = 0; base < seg_rows; base += 18) {
#pragma unroll
            for (int j = 0; j < 18; j += 2) {
                const int r = base + j;
                if (r >= seg_rows) break; // warp-uniform; an odd last row takes its pair from the pad (another segment's row or an invalid one)
                load_row16_s(R[(j + 16) % 18], pr + (r + 16) * (V3_TILE_PITCH / 4), sh);
                load_row16_s(R[(j + 17) % 18], pr + (r + 17) * (V3_TILE_PITCH / 4), sh);
                uint32_t a[2][8];
#pragma unroll
                for (int k = 0; k < 2; k++)
#pragma unroll
                    for (int i = 0; i < 8; i++) a[k][i] = 0;
#pragma unroll
                for (int y = 0; y < 16; y++) {
                    const uint4 f = F[y];
                    const int h = y >> 3; // 0: top quadrants, 1: bottom
#pragma unroll
                    for (int k = 0; k < 2; k++) {
                        a[k][4 * h + 0] = sad4_acc(f.x, R[(j + k + y) % 18][0], a[k][4 * h + 0]);
                        a[k][4 * h + 1] = sad4_acc(f.y, R[(j + k + y) % 18][1], a[k][4 * h + 1]);
                        a[k][4 * h + 2] = sad4_acc(f.z, R[(j + k + y) % 18][2], a[k][4 * h + 2]);
                        a[k][4 * h + 3] = sad4_acc(f.w, R[(j + k + y) % 18][3], a[k][4 * h + 3]);
                    }
                }
#pragma unroll
                for (int k = 0; k < 2; k++) {
                    const uint4 *cy4 = (const uint4 *)&cyt[rbeg + r + k][0];
                    const uint4 ca = cy4[0], cb = cy4[1], cc = cy4[2];
                    const uint32_t tl = a[k][0] + a[k][1], tr = a[k][2] + a[k][3], bl = a[k][4] + a[k][5], br = a[k][6] + a[k][7];
                    const uint32_t top = tl + tr, bot = bl + br, lft = tl + bl, rgt = tr + br, all = top + bot;
#define UPD(p, sad, cy) best[p] = min(best[p], __umul24((sad), 1u << KEY_SHIFT) + cxp[p] + (cy))
                    UPD(0, all, ca.x); UPD(1, top, ca.y); UPD(2, bot, ca.z); UPD(3, lft, ca.w); UPD(4, rgt, cb.x);
                    UPD(5, tl, cb.y); UPD(6, tr, cb.z); UPD(7, bl, cb.w); UPD(8, br, cc.x);
#undef UPD
                }
            }
        }
    }
}

template <int RS>
__global__ void __launch_bounds__((V3_PROD + V3_CONS) * 32, 1)
me_search_mb3_kernel(Geo geo, const x264_cuda_me_mb_job_t *__restrict__ jobs, int n_jobs, const int16_t *const *__restrict__ cost_tabs, int me_range,
                     int max_ur, x264_cuda_me_mb_result_t *__restrict__ results, long long *__restrict__ dbg)
{
    extern __shared__ __align__(16) uint8_t s_dyn[]; // V3_CONS x V3_SLOTS slots | producers' WarpSmem | mbarriers
    const size_t slot_bytes = v3_slot_bytes(max_ur);
    WarpSmem *s_prod = (WarpSmem *)(s_dyn + slot_bytes * V3_CONS * V3_SLOTS);
    uint64_t *s_full = (uint64_t *)(s_prod + V3_CONS / V3_FEED), *s_empty = s_full + V3_CONS * V3_SLOTS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int stride = geo.stride;
    if (threadIdx.x < V3_CONS * V3_SLOTS) { mbar_init(s_full + threadIdx.x, 1); mbar_init(s_empty + threadIdx.x, 1); }
    __syncthreads();
#define SLOT_HEAD(i) (*(SlotHead *)(s_dyn + slot_bytes * (i)))
#define SLOT_CXP(i) ((uint16_t *)(s_dyn + slot_bytes * (i) + sizeof(SlotHead)))
#define SLOT_CYT(i) ((uint32_t (*)[12])(s_dyn + slot_bytes * (i) + sizeof(SlotHead) + V3_CXP_BYTES))
#define SLOT_TILE(i) (s_dyn + slot_bytes * (i) + sizeof(SlotHead) + V3_CXP_BYTES + v3_cyt_bytes(max_ur))
    if (warp < V3_PROD) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(V3_PROD_REGS));
        if (warp < V3_CONS / V3_FEED) {
            WarpSmem &S = s_prod[warp];
            long long t_wait = 0, t_work = 0, t_first = 0; const long long t_begin = dbg ? clock64() : 0; // dbg: cycle accounting (X264_CUDA_MB_DEBUG)
            int kf[V3_FEED]; // slots filled so far for each of this producer's scanners (warp * V3_FEED + e)
#pragma unroll
            for (int e = 0; e < V3_FEED; e++) kf[e] = 0;
            // static work split that keeps neighbours together: CTA b takes the batches b, b + grid, ... of V3_CONS consecutive macroblocks,
            // scanner c the c-th of each — the window tiles in flight on an SM overlap, so the reference rows are shared in L1
            for (int j0 = blockIdx.x * V3_CONS + warp * V3_FEED; j0 < n_jobs; j0 += gridDim.x * V3_CONS) {
#pragma unroll
                for (int e = 0; e < V3_FEED; e++) {
                    const int jb = j0 + e;
                    if (jb >= n_jobs) break;
                    {   // this producer's macroblock after the next: its job record towards L2 now; the next one's source rows (its record
                        // came in a turn ago)
                        const int jn = e + 1 < V3_FEED ? jb + 1 : j0 + gridDim.x * V3_CONS;
                        const int jn2 = e + 2 < V3_FEED ? jb + 2 : j0 + gridDim.x * V3_CONS + (e + 2 - V3_FEED);
                        if (jn2 < n_jobs && lane < 3) prefetch_l2((const uint8_t *)(jobs + jn2) + 128 * lane);
                        if (jn < n_jobs) {
                            const uint32_t pos = __ldg((const uint32_t *)(jobs + jn)); // mb_x | mb_y << 16
                            if (lane < 16) prefetch_l2(geo.fenc + ((size_t)(pos >> 16) * 16 + lane) * stride + (pos & 0xffff) * 16);
                        }
                    }
                    const int16_t *tab; const uint8_t *ref0;
                    const unsigned mask = v3_setup(S, geo, jobs, jb, cost_tabs, me_range, results, lane, tab, ref0);
                    if (!mask) continue;
                    unsigned todo = mask;
                    int first = 1;
                    while (todo) { // pass grouping: all partitions over the union of their windows when it is compact, else one by one
                        int ux0 = 1 << 20, uy0 = 1 << 20, ux1 = -(1 << 20), uy1 = -(1 << 20);
                        if (lane < NP && (todo >> lane & 1)) { ux0 = S.win[lane][0]; uy0 = S.win[lane][1]; ux1 = ux0 + S.win[lane][2]; uy1 = uy0 + S.win[lane][3]; }
                        ux0 = __reduce_min_sync(0xffffffffu, ux0); uy0 = __reduce_min_sync(0xffffffffu, uy0);
                        ux1 = __reduce_max_sync(0xffffffffu, ux1); uy1 = __reduce_max_sync(0xffffffffu, uy1);
                        unsigned group = todo;
                        if (ux1 - ux0 > V3_MAX_UW || uy1 - uy0 > max_ur) {
                            const int p = __ffs(todo) - 1;
                            group = 1u << p;
                            ux0 = S.win[p][0]; uy0 = S.win[p][1]; ux1 = ux0 + S.win[p][2]; uy1 = uy0 + S.win[p][3];
                        }
                        todo &= ~group;
                        const int si = (warp * V3_FEED + e) * V3_SLOTS + kf[e] % V3_SLOTS;
                        const long long tw0 = dbg ? clock64() : 0;
                        mbar_wait_relaxed(s_empty + si, ((kf[e] / V3_SLOTS) & 1) ^ 1);
                        if (dbg) t_wait += clock64() - tw0;
                        v3_fill(SLOT_HEAD(si), SLOT_CXP(si), SLOT_CYT(si), SLOT_TILE(si), S, tab, ref0, stride, jb, mask, group, first, todo == 0, ux0, uy0, ux1 - ux0, uy1 - uy0, lane);
                        __syncwarp();
                        if (lane == 0) mbar_arrive(s_full + si);
                        if (dbg && !t_first) t_first = clock64() - t_begin;
                        kf[e]++;
                        first = 0;
                    }
                }
            }
            if (dbg && lane == 0) { t_work = clock64() - t_begin - t_wait; long long *d = dbg + ((size_t)blockIdx.x * (V3_PROD + V3_CONS) + warp) * 4; d[0] = t_work; d[1] = t_wait; d[2] = t_first; d[3] = kf[0]; }
#pragma unroll
            for (int e = 0; e < V3_FEED; e++) { // no more work: tell the scanners
                const int si = (warp * V3_FEED + e) * V3_SLOTS + kf[e] % V3_SLOTS;
                mbar_wait_relaxed(s_empty + si, ((kf[e] / V3_SLOTS) & 1) ^ 1);
                if (lane == 0) { SLOT_HEAD(si).jb = -1; mbar_arrive(s_full + si); }
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(V3_CONS_REGS));
        const int c = (warp - V3_PROD); // scanner index; fed by producer c / V3_FEED
        int my_bmx = 0, my_bmy = 0, my_cost = INVALID_COST; // lane p keeps partition p's window winner
        long long t_wait = 0, t_first = 0; const long long t_begin = dbg ? clock64() : 0;
        for (int n = 0;; n++) {
            const int si = c * V3_SLOTS + n % V3_SLOTS;
            const long long tw0 = dbg ? clock64() : 0;
            mbar_wait(s_full + si, (n / V3_SLOTS) & 1);
            if (dbg) { const long long w = clock64() - tw0; if (n == 0) t_first = w; else t_wait += w; }
            const SlotHead &H = SLOT_HEAD(si);
            const int jb = H.jb;
            if (jb < 0) {
                if (dbg && lane == 0) { long long *d = dbg + ((size_t)blockIdx.x * (V3_PROD + V3_CONS) + warp) * 4; d[0] = clock64() - t_begin - t_wait - t_first; d[1] = t_wait; d[2] = t_first; d[3] = n; }
                break;
            }
            if (H.first) { my_bmx = my_bmy = 0; my_cost = INVALID_COST; }
            uint32_t best[NP];
#pragma unroll
            for (int p = 0; p < NP; p++) best[p] = 0xffffffffu;
            if (RS == 2) v3_scan2(H, SLOT_CXP(si), SLOT_CYT(si), SLOT_TILE(si), lane, best);
            else v3_scan(H, SLOT_CXP(si), SLOT_CYT(si), SLOT_TILE(si), lane, best);
            const unsigned group = H.group;
            const int ux0 = H.ux0, uy0 = H.uy0, uwidth = H.uwidth;
            {   // warp argmin per partition, branch-free so that the 18 reductions pipeline: min key (cost,row,chunk), then the lowest
                // in-chunk column among its holders
                uint32_t kmin[NP], cmin[NP];
#pragma unroll
                for (int p = 0; p < NP; p++) kmin[p] = __reduce_min_sync(0xffffffffu, best[p]);
#pragma unroll
                for (int p = 0; p < NP; p++) {
                    const int rem = uwidth - (int)(kmin[p] & 3) * 32;
                    const int cw = rem > 16 ? 32 : rem > 8 ? 16 : 8;
                    cmin[p] = __reduce_min_sync(0xffffffffu, best[p] == kmin[p] ? (uint32_t)(lane & (cw - 1)) : 0xffffffffu);
                }
#pragma unroll
                for (int p = 0; p < NP; p++)
                    if (lane == p && (group >> p & 1)) {
                        my_cost = (int)(kmin[p] >> KEY_SHIFT);
                        my_bmy = uy0 + (int)((kmin[p] >> 2) & 255);
                        my_bmx = ux0 + (int)(kmin[p] & 3) * 32 + (int)cmin[p];
                    }
            }
            if (H.last && lane < NP) {
                const unsigned mask = H.mask;
                x264_cuda_me_result_t r;
                int bmx = H.seed[lane][0], bmy = H.seed[lane][1], bcost = H.seed[lane][2];
                r.seed_mx = (int16_t)bmx; r.seed_my = (int16_t)bmy; r.seed_cost = bcost;
                if ((mask >> lane & 1) && my_cost < INVALID_COST && my_cost < bcost) { bcost = my_cost; bmy = my_bmy; bmx = my_bmx; }
                if (!(mask >> lane & 1)) { bmx = bmy = 0; bcost = -1; r.seed_mx = r.seed_my = 0; r.seed_cost = -1; }
                r.bmx = (int16_t)bmx; r.bmy = (int16_t)bmy; r.bcost = bcost;
                results[jb].part[lane] = r;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(s_empty + si);
        }
    }
#undef SLOT_HEAD
#undef SLOT_CXP
#undef SLOT_CYT
#undef SLOT_TILE
}

} // namespace

extern "C" int x264_cuda_me_search_mb_dev(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, const x264_cuda_frame_t *fref,
                                          int me_range, const void *d_jobs, int n_jobs, void *d_results)
{
    x264_cuda_enter(ctx);
    if (n_jobs <= 0) return 0;
    if (fenc->g.stride != fref->g.stride || fenc->g.lines != fref->g.lines) {
        snprintf(ctx->err, 256, "x264_cuda_me_search_mb: fenc/fref geometry mismatch");
        return -1;
    }
    if (me_range < 1 || me_range > MB_MAX_RANGE) {
        snprintf(ctx->err, 256, "x264_cuda_me_search_mb: me_range %d not in 1..%d (use x264_cuda_me_search for larger ranges)", me_range,
                 MB_MAX_RANGE);
        return -1;
    }
    const int16_t *const *d_tabs;
    if (x264_cuda_cost_tables(ctx, &d_tabs)) return -1;
    Geo geo = { fenc->plane[0], fref->plane[0], fenc->g.stride };
    const int max_ur = 2 * me_range + 1 + MB_UNION_SLACK;
    static const int variant = getenv("X264_CUDA_MB_KERNEL") ? atoi(getenv("X264_CUDA_MB_KERNEL")) : 3; // 1: full-strip lanes, 2: half-strip lanes, 3: warp-specialised
    const int blocks = (n_jobs + MB_WARPS - 1) / MB_WARPS;
    const int max_ur3 = 2 * me_range + 1 + V3_UNION_SLACK;
    const size_t dyn3 = v3_slot_bytes(max_ur3) * V3_CONS * V3_SLOTS + sizeof(WarpSmem) * (V3_CONS / V3_FEED) + 2 * V3_CONS * V3_SLOTS * sizeof(uint64_t);
    if (variant == 3 && dyn3 <= 227 * 1024 && 2 * me_range + 4 <= V3_MAX_UW) { // a single partition's window must fit one pass
        static const int rs = getenv("X264_CUDA_MB_RS") ? atoi(getenv("X264_CUDA_MB_RS")) : 1; // candidate rows per scan step
        static bool attr_set[64];
        if (!attr_set[ctx->device & 63]) {
            CUDA_TRY(ctx, cudaFuncSetAttribute(me_search_mb3_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            CUDA_TRY(ctx, cudaFuncSetAttribute(me_search_mb3_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            attr_set[ctx->device & 63] = true;
        }
        const int grid = min(ctx->sm_count, (n_jobs + V3_CONS - 1) / V3_CONS);
        static const bool debug = getenv("X264_CUDA_MB_DEBUG") != nullptr; // cycle accounting of the two roles, printed per launch
        long long *d_dbg = nullptr;
        const size_t n_dbg = (size_t)grid * (V3_PROD + V3_CONS) * 4;
        if (debug) { CUDA_TRY(ctx, cudaMalloc(&d_dbg, n_dbg * 8)); CUDA_TRY(ctx, cudaMemsetAsync(d_dbg, 0, n_dbg * 8, ctx->stream)); }
        (rs == 2 ? me_search_mb3_kernel<2> : me_search_mb3_kernel<1>)<<<grid, (V3_PROD + V3_CONS) * 32, dyn3, ctx->stream>>>(
            geo, (const x264_cuda_me_mb_job_t *)d_jobs, n_jobs, d_tabs, me_range, max_ur3, (x264_cuda_me_mb_result_t *)d_results, d_dbg);
        if (debug) {
            long long *h = (long long *)malloc(n_dbg * 8);
            CUDA_TRY(ctx, cudaMemcpyAsync(h, d_dbg, n_dbg * 8, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
            double s[2][4] = { { 0 } }, mx[2][4] = { { 0 } };
            for (int b = 0; b < grid; b++)
                for (int w = 0; w < V3_PROD + V3_CONS; w++)
                    for (int k = 0; k < 4; k++) {
                        const double v = (double)h[((size_t)b * (V3_PROD + V3_CONS) + w) * 4 + k];
                        s[w >= V3_PROD][k] += v; if (v > mx[w >= V3_PROD][k]) mx[w >= V3_PROD][k] = v;
                    }
            const double np = (double)grid * V3_PROD, nc = (double)grid * V3_CONS;
            fprintf(stderr, "mb3 debug (cycles, mean / max per warp): producer work %.0f / %.0f, wait-empty %.0f / %.0f, first slot ready %.0f / %.0f, slots %.1f | "
                            "scanner busy %.0f / %.0f, wait-full %.0f / %.0f, first wait %.0f / %.0f, slots %.1f\n",
                    s[0][0] / np, mx[0][0], s[0][1] / np, mx[0][1], s[0][2] / np, mx[0][2], s[0][3] / np, s[1][0] / nc, mx[1][0], s[1][1] / nc, mx[1][1],
                    s[1][2] / nc, mx[1][2], s[1][3] / nc);
            free(h); cudaFree(d_dbg);
        }
    } else if (variant == 1) {
        const size_t dyn = (size_t)MB_WARPS * (max_ur + 4) * 12 * sizeof(uint32_t);
        const int prefetch_dist = MB_WARPS * 3 * ctx->sm_count; // warps resident at once (168 registers: three CTAs per SM)
        me_search_mb_kernel<false><<<blocks, MB_WARPS * 32, dyn, ctx->stream>>>(geo, (const x264_cuda_me_mb_job_t *)d_jobs, n_jobs, d_tabs, me_range,
                                                                                max_ur, prefetch_dist, (x264_cuda_me_mb_result_t *)d_results);
    } else {
        const size_t dyn = (size_t)MB_WARPS * (max_ur + 4) * 16 * sizeof(uint32_t);
        const int prefetch_dist = MB_WARPS * 5 * ctx->sm_count; // five CTAs per SM
        me_search_mb_kernel<true><<<blocks, MB_WARPS * 32, dyn, ctx->stream>>>(geo, (const x264_cuda_me_mb_job_t *)d_jobs, n_jobs, d_tabs, me_range,
                                                                               max_ur, prefetch_dist, (x264_cuda_me_mb_result_t *)d_results);
    }
    LAUNCH_CHECK(ctx, "me_search_mb_kernel");
    return 0;
}

extern "C" int x264_cuda_me_search_mb(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, const x264_cuda_frame_t *fref, int me_range,
                                      const x264_cuda_me_mb_job_t *jobs, int n_jobs, x264_cuda_me_mb_result_t *results)
{
    x264_cuda_enter(ctx);
    if (n_jobs <= 0) return 0;
    const size_t jb = (size_t)n_jobs * sizeof(x264_cuda_me_mb_job_t), rb = (size_t)n_jobs * sizeof(x264_cuda_me_mb_result_t);
    const size_t jb_al = (jb + 255) & ~(size_t)255;
    if (x264_cuda_stage(ctx, jb_al + rb, jb_al + rb)) return -1;
    uint8_t *hs = (uint8_t *)ctx->h_stage, *ds = (uint8_t *)ctx->d_stage;
    if (x264_cuda_jobs_in(ctx, ds, jobs, hs, jb)) return -1;
    if (x264_cuda_me_search_mb_dev(ctx, fenc, fref, me_range, ds, n_jobs, ds + jb_al)) return -1;
    if (x264_cuda_results_out(ctx, results, ds + jb_al, hs + jb_al, rb)) return -1;
    return 0;
}
