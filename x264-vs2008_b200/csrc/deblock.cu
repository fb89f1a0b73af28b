// deblock.cu — in-loop deblocking of one progressive frame, bit-exact with x264_frame_deblock_row
// (S/common/frame.c:621-792; edge filters :424-586; tables :377-421).
//
// The filter is order-dependent: a macroblock's left edge reads the left neighbour AFTER that neighbour's horizontal
// edges were filtered, and its top edge reads rows the upper-right neighbour's left edge has already touched.  Exactness
// therefore needs the reference's macroblock order, which leaves the classic 2:1 wavefront: (x,y) may run once (x-1,y)
// and (x+1,y-1) are done.  ONE WARP PER MACROBLOCK ROW sweeps left to right; the left dependency never leaves the warp
// (the previous macroblock's right columns stay in shared memory), the upper dependency is a per-row progress counter in
// global memory.  Lanes 0-15 own the 16 luma lines across the current edge, lanes 16-31 the 8+8 chroma lines.
//
// Roofline class: latency (W + 2H dependent steps per frame); HBM traffic is one read + one write of the planes.
#include "common.cuh"

namespace {

__constant__ uint8_t c_alpha[52] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 4, 4, 5, 6, 7, 8, 9, 10, 12, 13, 15, 17, 20, 22, 25, 28, 32, 36, 40, 45,
                                     50, 56, 63, 71, 80, 90, 101, 113, 127, 144, 162, 182, 203, 226, 255, 255 };
__constant__ uint8_t c_beta[52] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10,
                                    11, 11, 12, 12, 13, 13, 14, 14, 15, 15, 16, 16, 17, 17, 18, 18 };
__constant__ uint8_t c_tc0[52][4] = { // [qp + offset][bS]; bS 0 is never looked up
    { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 },
    { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 1 },
    { 0, 0, 0, 1 }, { 0, 0, 0, 1 }, { 0, 0, 0, 1 }, { 0, 0, 1, 1 }, { 0, 0, 1, 1 }, { 0, 1, 1, 1 }, { 0, 1, 1, 1 }, { 0, 1, 1, 1 }, { 0, 1, 1, 1 },
    { 0, 1, 1, 2 }, { 0, 1, 1, 2 }, { 0, 1, 1, 2 }, { 0, 1, 1, 2 }, { 0, 1, 2, 3 }, { 0, 1, 2, 3 }, { 0, 2, 2, 3 }, { 0, 2, 2, 4 }, { 0, 2, 3, 4 },
    { 0, 2, 3, 4 }, { 0, 3, 3, 5 }, { 0, 3, 4, 6 }, { 0, 3, 4, 6 }, { 0, 4, 5, 7 }, { 0, 4, 5, 8 }, { 0, 4, 6, 9 }, { 0, 5, 7, 10 }, { 0, 6, 8, 11 },
    { 0, 6, 8, 13 }, { 0, 7, 10, 14 }, { 0, 8, 11, 16 }, { 0, 9, 12, 18 }, { 0, 10, 13, 20 }, { 0, 11, 15, 23 }, { 0, 13, 17, 25 } };
__constant__ uint8_t c_chroma_qp[52] = { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29,
                                         29, 30, 31, 32, 32, 33, 34, 34, 35, 35, 36, 36, 37, 37, 37, 38, 38, 38, 39, 39, 39, 39 };

// the reference indexes its tables with qp + offset in -12..63 through 12 entries of padding either side
__device__ __forceinline__ int alpha_of(int q) { return q < 0 ? 0 : c_alpha[min(q, 51)]; }
__device__ __forceinline__ int beta_of(int q) { return q < 0 ? 0 : c_beta[min(q, 51)]; }
__device__ __forceinline__ int tc0_of(int q, int bs) { return q < 0 ? 0 : c_tc0[min(q, 51)][bs]; }
__device__ __forceinline__ int chroma_qp(int q) { return c_chroma_qp[clip3i(q, 0, 51)]; }

struct MbInfo {            // what the filter needs of one macroblock, unpacked from the reference's arrays
    int type, qp, t8;
    unsigned nz;           // bit x+4y: 4x4 block has coefficients, as the deblocker sees it (munge_cavlc_nnz_row, frame.c:336-352)
    int8_t ref[2][4];      // [list][8x8 block]
    int16_t mv[2][16][2];  // [list][4x4 block x+4y]
};

struct DeblockArgs {
    uint8_t *y, *u, *v;
    int stride, stride_c, W, H;
    const int8_t *type, *qp, *t8;
    const uint8_t *nnz;    // [mb][24]
    const int8_t *ref[2];  // frame-wide 8x8 grid, stride 2W
    const int16_t *mv[2];  // frame-wide 4x4 grid, stride 4W, 2 components
    int alpha_off, beta_off, chroma_off, slice_b, psub8x8, cavlc8;
    int *progress;         // [H] macroblocks finished per row; zero before the launch
};

#define LT_STRIDE 32 // luma tile: rows -4..15, columns -4..15 at byte 16 + c (column 0 on a 16-byte boundary)
#define CT_STRIDE 16 // chroma tile: rows -2..7, columns -2..7 at byte 8 + c
struct RowSmem {
    __align__(16) uint8_t L[20][LT_STRIDE];
    __align__(16) uint8_t C[2][10][CT_STRIDE];
    MbInfo cur, left, top;
};

__device__ void load_info(const DeblockArgs &a, int mb_x, int mb_y, MbInfo &m, int lane)
{
    const int mb = mb_y * a.W + mb_x;
    if (lane == 0) {
        m.type = a.type[mb]; m.qp = a.qp[mb]; m.t8 = a.t8[mb];
        unsigned nz = 0;
        const uint8_t *n = a.nnz + (size_t)mb * 24;
        for (int i = 0; i < 16; i++) nz |= (unsigned)(n[i] != 0) << i;
        if (a.cavlc8 && m.t8) { // per-8x8 "any coefficient"
            unsigned o = 0;
            for (int b = 0; b < 4; b++) {
                const int s = (b & 1) * 2 + (b >> 1) * 8;
                if (nz & (0x33u << s)) o |= 0x33u << s;
            }
            nz = o;
        }
        m.nz = nz;
    }
    if (lane < 8) { // ref: 2 lists x 2 rows x 2 entries
        const int l = lane >> 2, k = lane & 3;
        m.ref[l][k] = (l == 0 || a.slice_b) ? a.ref[l][(size_t)(2 * mb_y + (k >> 1)) * 2 * a.W + 2 * mb_x + (k & 1)] : (int8_t)0;
    }
    { // mv: 2 lists x 16 blocks, one 32-bit (x,y) pair per lane
        const int l = lane >> 4, k = lane & 15;
        uint32_t v = 0;
        if (l == 0 || a.slice_b) v = ((const uint32_t *)a.mv[l])[(size_t)(4 * mb_y + (k >> 2)) * 4 * a.W + 4 * mb_x + (k & 3)];
        *(uint32_t *)m.mv[l][k] = v;
    }
}

// bS of the four 4-line groups of one edge (frame.c:697-741); every lane computes all four (they chain)
__device__ unsigned edge_bs(const DeblockArgs &a, const MbInfo &c, const MbInfo &n, int dir, int edge, int no_sub8x8)
{
    unsigned out = 0;
    int prev = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int x = dir == 0 ? edge : i, y = dir == 0 ? i : edge;
        const int xn = dir == 0 ? (x - 1) & 3 : x, yn = dir == 0 ? y : (y - 1) & 3;
        int bs = 0;
        if ((c.nz >> (x + 4 * y) & 1) | (n.nz >> (xn + 4 * yn) & 1)) bs = 2;
        else if (!(edge & no_sub8x8)) {
            if ((i & no_sub8x8) && prev != 2) bs = prev;
            else {
                const int bp = x + 4 * y, bq = xn + 4 * yn, rp = (x >> 1) + (y >> 1) * 2, rq = (xn >> 1) + (yn >> 1) * 2;
                bool diff = c.ref[0][rp] != n.ref[0][rq] || abs(c.mv[0][bp][0] - n.mv[0][bq][0]) >= 4 || abs(c.mv[0][bp][1] - n.mv[0][bq][1]) >= 4;
                if (!diff && a.slice_b)
                    diff = c.ref[1][rp] != n.ref[1][rq] || abs(c.mv[1][bp][0] - n.mv[1][bq][0]) >= 4 || abs(c.mv[1][bp][1] - n.mv[1][bq][1]) >= 4;
                bs = diff;
            }
        }
        prev = bs;
        out |= (unsigned)bs << (2 * i);
    }
    return out;
}

// one line across an edge; p points at q0, xs = byte step across the edge.  bs 4 = the intra macroblock-edge filter.
__device__ void filter_luma(uint8_t *p, int xs, int alpha, int beta, int bs, int tc0)
{
    const int p2 = p[-3 * xs], p1 = p[-2 * xs], p0 = p[-xs], q0 = p[0], q1 = p[xs], q2 = p[2 * xs];
    if (abs(p0 - q0) >= alpha || abs(p1 - p0) >= beta || abs(q1 - q0) >= beta) return;
    if (bs < 4) { // frame.c:424-467
        int tc = tc0;
        const int avg = (p0 + q0 + 1) >> 1;
        if (abs(p2 - p0) < beta) { p[-2 * xs] = (uint8_t)(p1 + clip3i(((p2 + avg) >> 1) - p1, -tc0, tc0)); tc++; }
        if (abs(q2 - q0) < beta) { p[xs] = (uint8_t)(q1 + clip3i(((q2 + avg) >> 1) - q1, -tc0, tc0)); tc++; }
        const int delta = clip3i((((q0 - p0) << 2) + (p1 - q1) + 4) >> 3, -tc, tc);
        p[-xs] = (uint8_t)clip_u8(p0 + delta);
        p[0] = (uint8_t)clip_u8(q0 - delta);
    } else if (abs(p0 - q0) < ((alpha >> 2) + 2)) { // frame.c:507-552
        if (abs(p2 - p0) < beta) {
            const int p3 = p[-4 * xs];
            p[-xs] = (uint8_t)((p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
            p[-2 * xs] = (uint8_t)((p2 + p1 + p0 + q0 + 2) >> 2);
            p[-3 * xs] = (uint8_t)((2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
        } else
            p[-xs] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2);
        if (abs(q2 - q0) < beta) {
            const int q3 = p[3 * xs];
            p[0] = (uint8_t)((p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3);
            p[xs] = (uint8_t)((p0 + q0 + q1 + q2 + 2) >> 2);
            p[2 * xs] = (uint8_t)((2 * q3 + 3 * q2 + q1 + q0 + p0 + 4) >> 3);
        } else
            p[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2);
    } else {
        p[-xs] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2);
        p[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2);
    }
}
__device__ void filter_chroma(uint8_t *p, int xs, int alpha, int beta, int bs, int tc)
{
    const int p1 = p[-2 * xs], p0 = p[-xs], q0 = p[0], q1 = p[xs];
    if (abs(p0 - q0) >= alpha || abs(p1 - p0) >= beta || abs(q1 - q0) >= beta) return;
    if (bs < 4) { // frame.c:470-497
        const int delta = clip3i((((q0 - p0) << 2) + (p1 - q1) + 4) >> 3, -tc, tc);
        p[-xs] = (uint8_t)clip_u8(p0 + delta);
        p[0] = (uint8_t)clip_u8(q0 - delta);
    } else { // frame.c:562-580
        p[-xs] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2);
        p[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2);
    }
}

__global__ void __launch_bounds__(32) deblock_rows_kernel(DeblockArgs a)
{
    __shared__ RowSmem S;
    const int lane = threadIdx.x, mb_y = blockIdx.x;
    const int qp_thresh = 15 - min(a.alpha_off, a.beta_off) - max(0, a.chroma_off);
    uint8_t *rowy = a.y + (size_t)16 * mb_y * a.stride, *rowu = a.u + (size_t)8 * mb_y * a.stride_c, *rowv = a.v + (size_t)8 * mb_y * a.stride_c;
    for (int mb_x = 0; mb_x < a.W; mb_x++) {
        // ---- this macroblock's own pixels and state do not depend on the row above: fetch them before waiting
        uint4 ly = make_uint4(0, 0, 0, 0);
        uint2 lc = make_uint2(0, 0);
        if (lane < 16) {
            ly = __ldcg((const uint4 *)(rowy + (size_t)lane * a.stride + 16 * mb_x));
            lc = __ldcg((const uint2 *)((lane < 8 ? rowu : rowv) + (size_t)(lane & 7) * a.stride_c + 8 * mb_x));
        }
        load_info(a, mb_x, mb_y, S.cur, lane);
        if (mb_y > 0) {
            load_info(a, mb_x, mb_y - 1, S.top, lane);
            if (lane == 0) {
                const int need = min(mb_x + 2, a.W);
                while (*(volatile int *)(a.progress + mb_y - 1) < need) __nanosleep(100);
            }
            __syncwarp();
            __threadfence();
            if (lane < 4) *(uint4 *)&S.L[lane][16] = __ldcg((const uint4 *)(rowy - (size_t)(4 - lane) * a.stride + 16 * mb_x));
            else if (lane < 8) {
                const int pl = (lane - 4) >> 1, r = (lane - 4) & 1; // chroma rows -2, -1 of both planes
                *(uint2 *)&S.C[pl][r][8] = __ldcg((const uint2 *)((pl ? rowv : rowu) - (size_t)(2 - r) * a.stride_c + 8 * mb_x));
            }
        }
        if (lane < 16) {
            *(uint4 *)&S.L[4 + lane][16] = ly;
            *(uint2 *)&S.C[lane >> 3][2 + (lane & 7)][8] = lc;
        }
        __syncwarp();

        const MbInfo &c = S.cur;
        const int intra = c.type >= 0 && c.type <= 3;
        int edge_end = c.type == 6 ? 1 : 4;                       // P_SKIP
        const int no_sub8x8 = c.type != 5 || !a.psub8x8;          // P_8x8
        if (c.qp <= qp_thresh) edge_end = 1;
        for (int dir = 0; dir < 2; dir++) {
            int edge = dir ? mb_y == 0 : mb_x == 0;
            if (edge) edge += c.t8;
            for (; edge < edge_end; edge += c.t8 + 1) {
                const MbInfo &n = edge ? c : dir == 0 ? S.left : S.top;
                const int n_intra = n.type >= 0 && n.type <= 3;
                unsigned bsv;
                if (edge == 0 && (intra | n_intra)) bsv = 0x100; // marks the bS-4 filters
                else if (intra | n_intra) bsv = 0xff;            // bS 3 on all four groups
                else bsv = edge_bs(a, c, n, dir, edge, no_sub8x8);
                if (!bsv) continue;                               // warp-uniform
                const bool chroma_lane = lane >= 16;
                const bool active = !(chroma_lane && (edge & 1)); // chroma has edges 0 and 2 only
                const int qa = chroma_lane ? (chroma_qp(c.qp + a.chroma_off) + chroma_qp(n.qp + a.chroma_off) + 1) >> 1 : (c.qp + n.qp + 1) >> 1;
                const int ia = qa + a.alpha_off, alpha = alpha_of(ia), beta = beta_of(qa + a.beta_off);
                if (active && alpha && beta) {
                    if (!chroma_lane) {
                        const int bs = bsv == 0x100 ? 4 : (bsv >> (2 * (lane >> 2))) & 3;
                        uint8_t *p = dir == 0 ? &S.L[4 + lane][16 + 4 * edge] : &S.L[4 + 4 * edge][16 + lane];
                        if (bs) filter_luma(p, dir == 0 ? 1 : LT_STRIDE, alpha, beta, bs, bs < 4 ? tc0_of(ia, bs) : 0);
                    } else {
                        const int pl = (lane - 16) >> 3, i = lane & 7;
                        const int bs = bsv == 0x100 ? 4 : (bsv >> (2 * (i >> 1))) & 3;
                        uint8_t *p = dir == 0 ? &S.C[pl][2 + i][8 + 2 * edge] : &S.C[pl][2 + 2 * edge][8 + i];
                        if (bs) filter_chroma(p, dir == 0 ? 1 : CT_STRIDE, alpha, beta, bs, bs < 4 ? tc0_of(ia, bs) + 1 : 0);
                    }
                }
                __syncwarp();
            }
            __syncwarp();
        }

        // ---- write back: this macroblock, the 3 (chroma: 1) columns of the left neighbour and rows of the upper one it touched
        if (lane < 16) {
            uint8_t *d = rowy + (size_t)lane * a.stride + 16 * mb_x;
            *(uint4 *)d = *(const uint4 *)&S.L[4 + lane][16];
            if (mb_x > 0) *(uint32_t *)(d - 4) = *(const uint32_t *)&S.L[4 + lane][12];
            uint8_t *dc = (lane < 8 ? rowu : rowv) + (size_t)(lane & 7) * a.stride_c + 8 * mb_x;
            *(uint2 *)dc = *(const uint2 *)&S.C[lane >> 3][2 + (lane & 7)][8];
            if (mb_x > 0) *(uint16_t *)(dc - 2) = *(const uint16_t *)&S.C[lane >> 3][2 + (lane & 7)][6];
        } else if (mb_y > 0) {
            if (lane < 19) *(uint4 *)(rowy - (size_t)(19 - lane) * a.stride + 16 * mb_x) = *(const uint4 *)&S.L[lane - 15][16]; // rows -3..-1
            else if (lane < 21) *(uint2 *)((lane == 19 ? rowu : rowv) - a.stride_c + 8 * mb_x) = *(const uint2 *)&S.C[lane - 19][1][8];
        }
        __syncwarp();
        // ---- the right columns stay in shared memory as the next macroblock's left neighbour
        if (lane < 16) {
            *(uint32_t *)&S.L[4 + lane][12] = *(const uint32_t *)&S.L[4 + lane][28];
            *(uint16_t *)&S.C[lane >> 3][2 + (lane & 7)][6] = *(const uint16_t *)&S.C[lane >> 3][2 + (lane & 7)][14];
        }
        for (int i = lane; i < (int)(sizeof(MbInfo) / 4); i += 32) ((uint32_t *)&S.left)[i] = ((const uint32_t *)&S.cur)[i];
        __threadfence();
        __syncwarp();
        if (lane == 0) *(volatile int *)(a.progress + mb_y) = mb_x + 1;
    }
}

} // namespace

extern "C" int x264_cuda_frame_deblock_dev(x264_cuda_t *ctx, x264_cuda_frame_t *fdec, const x264_cuda_deblock_params_t *pm, const int8_t *d_type,
                                           const int8_t *d_qp, const int8_t *d_transform8x8, const uint8_t *d_nnz, const int8_t *d_ref0,
                                           const int16_t *d_mv0, const int8_t *d_ref1, const int16_t *d_mv1)
{
    if (!fdec->buf_chroma) {
        snprintf(ctx->err, 256, "x264_cuda_frame_deblock: frame needs X264_CUDA_FRAME_CHROMA");
        return -1;
    }
    if (pm->b_slice_b && (!d_ref1 || !d_mv1)) {
        snprintf(ctx->err, 256, "x264_cuda_frame_deblock: B slices need the list-1 ref/mv arrays");
        return -1;
    }
    const int H = fdec->g.mb_height;
    if (H > ctx->deblock_rows) {
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->d_deblock_progress); ctx->d_deblock_progress = nullptr; ctx->deblock_rows = 0;
        CUDA_TRY(ctx, cudaMalloc(&ctx->d_deblock_progress, (size_t)H * sizeof(int)));
        ctx->deblock_rows = H;
    }
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_deblock_progress, 0, (size_t)H * sizeof(int), ctx->stream));
    DeblockArgs a;
    a.y = fdec->plane[0]; a.u = fdec->chroma[0]; a.v = fdec->chroma[1];
    a.stride = fdec->g.stride; a.stride_c = fdec->stride_c; a.W = fdec->g.mb_width; a.H = H;
    a.type = d_type; a.qp = d_qp; a.t8 = d_transform8x8; a.nnz = d_nnz;
    a.ref[0] = d_ref0; a.ref[1] = d_ref1; a.mv[0] = d_mv0; a.mv[1] = d_mv1;
    a.alpha_off = pm->alpha_c0_offset; a.beta_off = pm->beta_offset; a.chroma_off = pm->chroma_qp_offset;
    a.slice_b = !!pm->b_slice_b; a.psub8x8 = !!pm->b_psub8x8; a.cavlc8 = !!pm->b_cavlc_8x8dct;
    a.progress = ctx->d_deblock_progress;
    // one single-warp CTA per macroblock row; all of them are resident at once (<= a few hundred), so a row can always
    // wait for the row above
    deblock_rows_kernel<<<H, 32, 0, ctx->stream>>>(a);
    LAUNCH_CHECK(ctx, "deblock_rows_kernel");
    return 0;
}

extern "C" int x264_cuda_frame_deblock(x264_cuda_t *ctx, x264_cuda_frame_t *fdec, const x264_cuda_deblock_params_t *pm, const int8_t *type,
                                       const int8_t *qp, const int8_t *transform8x8, const uint8_t (*nnz)[24], const int8_t *ref0,
                                       const int16_t (*mv0)[2], const int8_t *ref1, const int16_t (*mv1)[2])
{
    const size_t n = (size_t)fdec->g.mb_width * fdec->g.mb_height;
    const bool b = pm->b_slice_b && ref1 && mv1;
    // staging layout (256-byte aligned pieces): type, qp, t8, nnz, ref0, mv0, ref1, mv1
    const size_t sz[8] = { n, n, n, n * 24, n * 4, n * 64, b ? n * 4 : 0, b ? n * 64 : 0 };
    const void *src[8] = { type, qp, transform8x8, nnz, ref0, mv0, ref1, mv1 };
    size_t off[8], total = 0;
    for (int i = 0; i < 8; i++) { off[i] = total; total += (sz[i] + 255) & ~(size_t)255; }
    if (x264_cuda_stage(ctx, total, total)) return -1;
    uint8_t *hs = (uint8_t *)ctx->h_stage, *ds = (uint8_t *)ctx->d_stage;
    for (int i = 0; i < 8; i++)
        if (sz[i] && x264_cuda_jobs_in(ctx, ds + off[i], src[i], hs + off[i], sz[i])) return -1;
    if (x264_cuda_frame_deblock_dev(ctx, fdec, pm, (const int8_t *)(ds + off[0]), (const int8_t *)(ds + off[1]), (const int8_t *)(ds + off[2]), ds + off[3],
                                    (const int8_t *)(ds + off[4]), (const int16_t *)(ds + off[5]), b ? (const int8_t *)(ds + off[6]) : nullptr,
                                    b ? (const int16_t *)(ds + off[7]) : nullptr))
        return -1;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)); // the staging buffer is reused by the next call
    return 0;
}
