// deblock.cu — in-loop deblocking of one progressive frame, bit-exact with x264_frame_deblock_row
// (S/common/frame.c:621-792; edge filters :424-586; tables :377-421).
//
// The filter is order-dependent: a macroblock's left edge reads the left neighbour AFTER that neighbour's horizontal
// edges were filtered, and its top edge reads rows the upper-right neighbour's left edge has already touched.  Exactness
// therefore needs the reference's macroblock order, which leaves the classic 2:1 wavefront: (x,y) may run once (x-1,y)
// and (x+1,y-1) are done.  ONE WARP PER MACROBLOCK ROW sweeps left to right; the left dependency never leaves the warp
// (the previous macroblock's right columns stay in shared memory), the upper dependency travels through distributed shared
// memory inside a thread-block cluster of eight rows and through a global progress counter between clusters.  Lanes 0-15 own the 16 luma lines across the current edge, lanes 16-31 the 8+8 chroma lines.
// Everything that does not depend on pixels — bS of all 8 edges of every macroblock, alpha/beta/tc0 lookups — is done
// beforehand by a fully parallel kernel (one thread per macroblock edge) so that the dependent chain only filters.
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

// what the wavefront needs to know about one edge (dir*4 + edge) of one macroblock: 16 bytes
struct __align__(16) EdgeRec {
    uint8_t mode;          // 0: edge not filtered, 1: bS < 4 filters, 2: bS 4 (intra macroblock edge) filters
    uint8_t alpha, beta;   // luma thresholds (0 disables, like the reference's early return)
    uint8_t alpha_c, beta_c;
    uint8_t tc[4];         // luma tc0 per 4-line group, 0xff = bS 0
    uint8_t tc_c[4];       // chroma tc (tc0 + 1) per 2-line group, 0 = bS 0
    uint8_t pad[3];
};

__device__ void read_info(const DeblockArgs &a, int mb_x, int mb_y, MbInfo &m)
{
    const int mb = mb_y * a.W + mb_x;
    m.type = a.type[mb]; m.qp = a.qp[mb]; m.t8 = a.t8[mb];
    const uint2 *n2 = (const uint2 *)(a.nnz + (size_t)mb * 24); // 24-byte rows: 8-byte aligned
    const uint2 n0 = n2[0], n1 = n2[1];
    const uint32_t w[4] = { n0.x, n0.y, n1.x, n1.y };
    unsigned nz = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) nz |= (unsigned)(((w[i >> 2] >> (8 * (i & 3))) & 255) != 0) << i;
    if (a.cavlc8 && m.t8) { // per-8x8 "any coefficient"
        unsigned o = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int s = (b & 1) * 2 + (b >> 1) * 8;
            if (nz & (0x33u << s)) o |= 0x33u << s;
        }
        nz = o;
    }
    m.nz = nz;
}

// one thread per (macroblock, direction, edge): frame.c:644-657 (which edges), :697-741 (bS), :588-604 (thresholds)
__global__ void __launch_bounds__(256) deblock_prep_kernel(DeblockArgs a, EdgeRec *__restrict__ recs)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.W * a.H * 8) return;
    const int mb = t >> 3, dir = (t >> 2) & 1, edge = t & 3, mb_x = mb % a.W, mb_y = mb / a.W;
    EdgeRec r;
    *(uint4 *)&r = make_uint4(0, 0, 0, 0);
    MbInfo c;
    read_info(a, mb_x, mb_y, c);
    const int intra = c.type >= 0 && c.type <= 3;
    int edge_end = c.type == 6 ? 1 : 4;                       // P_SKIP
    const int no_sub8x8 = c.type != 5 || !a.psub8x8;          // P_8x8
    if (c.qp <= 15 - min(a.alpha_off, a.beta_off) - max(0, a.chroma_off)) edge_end = 1;
    const bool has_nb = dir ? mb_y > 0 : mb_x > 0;
    const bool on = edge == 0 ? has_nb : (edge < edge_end && !(c.t8 && (edge & 1)));
    if (on) {
        const int nx = mb_x - (edge == 0 && dir == 0), ny = mb_y - (edge == 0 && dir == 1);
        MbInfo n;
        read_info(a, nx, ny, n);
        const int n_intra = n.type >= 0 && n.type <= 3;
        unsigned bsv = 0;
        if (edge == 0 && (intra | n_intra)) r.mode = 2;
        else if (intra | n_intra) bsv = 0xff;
        else {
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
                        bool diff = false;
                        for (int l = 0; l < 1 + a.slice_b && !diff; l++) {
                            const int rp = a.ref[l][(size_t)(2 * mb_y + (y >> 1)) * 2 * a.W + 2 * mb_x + (x >> 1)];
                            const int rq = a.ref[l][(size_t)(2 * ny + (yn >> 1)) * 2 * a.W + 2 * nx + (xn >> 1)];
                            const uint32_t vp = ((const uint32_t *)a.mv[l])[(size_t)(4 * mb_y + y) * 4 * a.W + 4 * mb_x + x];
                            const uint32_t vq = ((const uint32_t *)a.mv[l])[(size_t)(4 * ny + yn) * 4 * a.W + 4 * nx + xn];
                            diff = rp != rq || abs((int16_t)(vp & 0xffff) - (int16_t)(vq & 0xffff)) >= 4 || abs((int16_t)(vp >> 16) - (int16_t)(vq >> 16)) >= 4;
                        }
                        bs = diff;
                    }
                }
                prev = bs;
                bsv |= (unsigned)bs << (2 * i);
            }
        }
        if (r.mode == 2 || bsv) {
            if (!r.mode) r.mode = 1;
            const int ql = (c.qp + n.qp + 1) >> 1, ia = ql + a.alpha_off;
            r.alpha = (uint8_t)alpha_of(ia); r.beta = (uint8_t)beta_of(ql + a.beta_off);
            const int qc = (chroma_qp(c.qp + a.chroma_off) + chroma_qp(n.qp + a.chroma_off) + 1) >> 1, iac = qc + a.alpha_off;
            r.alpha_c = (uint8_t)alpha_of(iac); r.beta_c = (uint8_t)beta_of(qc + a.beta_off);
            if (edge & 1) r.alpha_c = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int bs = (bsv >> (2 * i)) & 3;
                r.tc[i] = bs ? (uint8_t)tc0_of(ia, bs) : (uint8_t)0xff;
                r.tc_c[i] = bs ? (uint8_t)(tc0_of(iac, bs) + 1) : (uint8_t)0;
            }
        }
    }
    *(uint4 *)&recs[t] = *(const uint4 *)&r;
}


__device__ __forceinline__ int ld_acquire(const int *p) { int v; asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_release(int *p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// One line across an edge held in two packed words: P = p3,p2,p1,p0 (byte 0..3), Q = q0,q1,q2,q3.  Branch-free, and the SAME
// instruction stream serves luma and chroma lanes (a warp that branches on lane type runs both paths one after the other):
//   bS < 4 (frame.c:424-467, :470-497): chroma is the luma filter without the p1/q1 taps and with tc given directly;
//   bS 4   (frame.c:507-552, :562-580): chroma is the luma filter's "large step" branch.
// `strong` is warp-uniform (a property of the edge).  `on` = this lane filters; `luma` = lane holds a luma line.
__device__ __forceinline__ void filter_w(uint32_t &P, uint32_t &Q, int alpha, int beta, bool strong, int tc0, bool luma, bool on)
{
    const int p3 = P & 255, p2 = (P >> 8) & 255, p1 = (P >> 16) & 255, p0 = P >> 24;
    const int q0 = Q & 255, q1 = (Q >> 8) & 255, q2 = (Q >> 16) & 255, q3 = Q >> 24;
    const bool ok = on && abs(p0 - q0) < alpha && abs(p1 - p0) < beta && abs(q1 - q0) < beta;
    const bool ap = luma && abs(p2 - p0) < beta, aq = luma && abs(q2 - q0) < beta;
    int n2 = p2, n1 = p1, n0, m0, m1 = q1, m2 = q2;
    if (!strong) {
        const int tc = tc0 + (luma ? (int)ap + (int)aq : 0);
        const int avg = (p0 + q0 + 1) >> 1;
        const int d1 = clip3i(((p2 + avg) >> 1) - p1, -tc0, tc0), e1 = clip3i(((q2 + avg) >> 1) - q1, -tc0, tc0);
        if (ap) n1 = p1 + d1;
        if (aq) m1 = q1 + e1;
        const int delta = clip3i((((q0 - p0) << 2) + (p1 - q1) + 4) >> 3, -tc, tc);
        n0 = clip_u8(p0 + delta);
        m0 = clip_u8(q0 - delta);
    } else {
        const bool small = luma && abs(p0 - q0) < ((alpha >> 2) + 2);
        const bool sp = small && ap, sq = small && aq;
        n0 = sp ? (p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3 : (2 * p1 + p0 + q1 + 2) >> 2;
        m0 = sq ? (p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3 : (2 * q1 + q0 + p1 + 2) >> 2;
        if (sp) { n1 = (p2 + p1 + p0 + q0 + 2) >> 2; n2 = (2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3; }
        if (sq) { m1 = (p0 + q0 + q1 + q2 + 2) >> 2; m2 = (2 * q3 + 3 * q2 + q1 + q0 + p0 + 4) >> 3; }
    }
    if (ok) {
        P = (uint32_t)p3 | (uint32_t)n2 << 8 | (uint32_t)n1 << 16 | (uint32_t)n0 << 24;
        Q = (uint32_t)m0 | (uint32_t)m1 << 8 | (uint32_t)m2 << 16 | (uint32_t)q3 << 24;
    }
}

// the four (luma) / two (chroma) edges of one direction on the line this lane holds in registers.  Luma lines are w[0..4]
// (edge e between w[e] and w[e+1]); chroma lines are w[0..2] (edge 0 between w[0], w[1]; edge 2 between w[1], w[2]).
__device__ __forceinline__ void filter_line(uint32_t (&w)[5], const EdgeRec *rec, int dir, int lane)
{
    const bool luma = lane < 16;
    const int grp = luma ? lane >> 2 : (lane & 7) >> 1;
#pragma unroll
    for (int e = 0; e < 4; e++) {
        const EdgeRec &r = rec[dir * 4 + e];
        if (!r.mode) continue; // warp-uniform
        const bool strong = r.mode == 2;
        const int alpha = luma ? r.alpha : r.alpha_c, beta = luma ? r.beta : r.beta_c;
        const int tc0 = luma ? r.tc[grp] : r.tc_c[grp];
        // luma: tc0 0xff marks bS 0; chroma: tc 0 marks bS 0 and only even edges exist
        const bool on = strong ? (luma || !(e & 1)) : luma ? tc0 != 0xff : (tc0 != 0 && !(e & 1));
        if (e == 2) { // chroma's second edge sits one word earlier than luma's third
            uint32_t P = luma ? w[2] : w[1], Q = luma ? w[3] : w[2];
            filter_w(P, Q, alpha, beta, strong, tc0, luma, on);
            if (luma) { w[2] = P; w[3] = Q; } else { w[1] = P; w[2] = Q; }
        } else
            filter_w(w[e], w[e + 1], alpha, beta, strong, tc0, luma, on);
    }
}

// ---- the wavefront: eight consecutive macroblock rows form a thread-block cluster.  Inside a cluster a row PUSHES the
// final bottom rows of each finished macroblock into the next row's shared memory (DSMEM stores, ~200 cycles) and bumps a
// counter there; only the first row of a cluster reads the previous cluster's rows and progress through global memory
// (a global-memory flag handoff costs ~900 cycles relaxed / ~1500 with release-acquire one way on B200, measured), and only
// the last row of a cluster publishes one.  This row's own pixels and edge records never depend on other rows: they are
// always one macroblock ahead, in registers.
#define DB_CLUSTER 8
#define DB_RING 8          // macroblocks of top rows buffered per row
struct __align__(16) TopSlot { uint8_t y[4][16]; uint8_t c[2][2][8]; }; // luma rows 12..15, chroma rows 6,7 of U and V: 96 bytes

struct ClusterSmem {
    __align__(16) uint8_t L[20][LT_STRIDE];
    __align__(16) uint8_t C[2][10][CT_STRIDE];
    __align__(16) EdgeRec rec[8];
    __align__(16) TopSlot ring[DB_RING];   // written by the row above (remote), read here
    __align__(16) TopSlot prev;            // bottom rows of this row's previous macroblock, pushed once its right columns are final
    int top_avail;                         // macroblocks delivered into `ring` (written remotely, release)
    int credit;                            // macroblocks the row below has taken out of ITS ring (written remotely, release)
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t map_rank(uint32_t addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint4 v)
{
    asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_cluster_v2(uint32_t addr, uint2 v)
{
    asm volatile("st.shared::cluster.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void st_cluster_release(uint32_t addr, int v)
{
    asm volatile("st.release.cluster.shared::cluster.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_cluster_acquire(const int *p)
{
    int v;
    asm volatile("ld.acquire.cluster.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}

__global__ void __cluster_dims__(DB_CLUSTER, 1, 1) __launch_bounds__(32) deblock_rows_cluster_kernel(DeblockArgs a, const EdgeRec *__restrict__ recs)
{
    __shared__ ClusterSmem S;
    const int lane = threadIdx.x, mb_y = blockIdx.x;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if (lane == 0) { S.top_avail = 0; S.credit = 0; }
    // every CTA of the cluster has initialised its counters before anyone pushes into them
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (mb_y < a.H) {
        const bool top_local = mb_y > 0 && rank > 0;                       // top rows arrive through DSMEM
        const bool top_global = mb_y > 0 && rank == 0;                     // ... or through global memory (previous cluster)
        const bool push_down = mb_y + 1 < a.H && rank + 1 < DB_CLUSTER;    // the row below is in this cluster
        const bool publish = mb_y + 1 < a.H && rank + 1 == DB_CLUSTER;     // the row below is the next cluster's first
        const uint32_t down_ring = map_rank(smem_u32(&S.ring[0]), rank + (push_down ? 1 : 0));
        const uint32_t down_avail = map_rank(smem_u32(&S.top_avail), rank + (push_down ? 1 : 0));
        const uint32_t up_credit = map_rank(smem_u32(&S.credit), rank - (top_local ? 1 : 0));
        uint8_t *rowy = a.y + (size_t)16 * mb_y * a.stride, *rowu = a.u + (size_t)8 * mb_y * a.stride_c, *rowv = a.v + (size_t)8 * mb_y * a.stride_c;
        const uint4 *rrow = (const uint4 *)(recs + (size_t)mb_y * a.W * 8);
        const int *above = a.progress + mb_y - 1;
        const int cpl = (lane >> 3) & 1, cln = lane & 7;
        uint4 ly = make_uint4(0, 0, 0, 0), rc = make_uint4(0, 0, 0, 0), tl = make_uint4(0, 0, 0, 0);
        uint2 lc = make_uint2(0, 0), tc2 = make_uint2(0, 0);
        if (lane < 16) {
            ly = __ldcg((const uint4 *)(rowy + (size_t)lane * a.stride));
            lc = __ldcg((const uint2 *)((lane < 8 ? rowu : rowv) + (size_t)(lane & 7) * a.stride_c));
        }
        if (lane < 8) rc = __ldg(rrow + lane);
        int seen = 0, avail = 0, credit = 0;
        bool have_top = false;
        for (int mb_x = 0; mb_x < a.W; mb_x++) {
            if (lane < 16) {
                *(uint4 *)&S.L[4 + lane][16] = ly;
                *(uint2 *)&S.C[lane >> 3][2 + (lane & 7)][8] = lc;
            }
            if (lane < 8) *(uint4 *)&S.rec[lane] = rc;
            if (mb_x + 1 < a.W) {
                if (lane < 16) {
                    ly = __ldcg((const uint4 *)(rowy + (size_t)lane * a.stride + 16 * (mb_x + 1)));
                    lc = __ldcg((const uint2 *)((lane < 8 ? rowu : rowv) + (size_t)(lane & 7) * a.stride_c + 8 * (mb_x + 1)));
                }
                if (lane < 8) rc = __ldg(rrow + (size_t)(mb_x + 1) * 8 + lane);
            }
            int flag_probe = -1;
            if (top_local) { // the row above delivered (or will deliver) this macroblock's top rows into the ring
                while (avail <= mb_x) avail = ld_cluster_acquire(&S.top_avail);
                const TopSlot &t = S.ring[mb_x % DB_RING];
                if (lane < 4) *(uint4 *)&S.L[lane][16] = *(const uint4 *)t.y[lane];
                else if (lane < 8) *(uint2 *)&S.C[(lane - 4) >> 1][(lane - 4) & 1][8] = *(const uint2 *)t.c[(lane - 4) >> 1][(lane - 4) & 1];
            } else if (top_global) {
                if (!have_top) {
                    const int need = min(mb_x + 2, a.W);
                    while (seen < need) seen = ld_acquire(above);
                    if (lane < 4) tl = __ldcg((const uint4 *)(rowy - (size_t)(4 - lane) * a.stride + 16 * mb_x));
                    else if (lane < 8) tc2 = __ldcg((const uint2 *)(((lane - 4) >> 1 ? rowv : rowu) - (size_t)(2 - ((lane - 4) & 1)) * a.stride_c + 8 * mb_x));
                }
                if (lane < 4) *(uint4 *)&S.L[lane][16] = tl;
                else if (lane < 8) *(uint2 *)&S.C[(lane - 4) >> 1][(lane - 4) & 1][8] = tc2;
                have_top = false;
                if (mb_x + 1 < a.W) {
                    if (seen >= min(mb_x + 3, a.W)) {
                        if (lane < 4) tl = __ldcg((const uint4 *)(rowy - (size_t)(4 - lane) * a.stride + 16 * (mb_x + 1)));
                        else if (lane < 8) tc2 = __ldcg((const uint2 *)(((lane - 4) >> 1 ? rowv : rowu) - (size_t)(2 - ((lane - 4) & 1)) * a.stride_c + 8 * (mb_x + 1)));
                        have_top = true;
                    } else
                        flag_probe = ld_acquire(above);
                }
            }
            __syncwarp();

            // ---- vertical edges: the lane's line (luma row / chroma row) lives in registers across all of them; then horizontal
            // edges on the lane's column, gathered from the tile
            uint32_t w[5];
            if (lane < 16) {
                w[0] = *(const uint32_t *)&S.L[4 + lane][12];
                const uint4 v = *(const uint4 *)&S.L[4 + lane][16];
                w[1] = v.x; w[2] = v.y; w[3] = v.z; w[4] = v.w;
            } else {
                w[0] = *(const uint32_t *)&S.C[cpl][2 + cln][4];
                const uint2 v = *(const uint2 *)&S.C[cpl][2 + cln][8];
                w[1] = v.x; w[2] = v.y; w[3] = w[4] = 0;
            }
            filter_line(w, S.rec, 0, lane);
            if (lane < 16) {
                *(uint32_t *)&S.L[4 + lane][12] = w[0];
                *(uint4 *)&S.L[4 + lane][16] = make_uint4(w[1], w[2], w[3], w[4]);
            } else {
                *(uint32_t *)&S.C[cpl][2 + cln][4] = w[0];
                *(uint2 *)&S.C[cpl][2 + cln][8] = make_uint2(w[1], w[2]);
            }
            __syncwarp();
            if (lane < 16) {
#pragma unroll
                for (int k = 0; k < 5; k++)
                    w[k] = (uint32_t)S.L[4 * k][16 + lane] | (uint32_t)S.L[4 * k + 1][16 + lane] << 8 | (uint32_t)S.L[4 * k + 2][16 + lane] << 16 |
                           (uint32_t)S.L[4 * k + 3][16 + lane] << 24;
            } else {
                w[0] = (uint32_t)S.C[cpl][0][8 + cln] << 16 | (uint32_t)S.C[cpl][1][8 + cln] << 24;
#pragma unroll
                for (int k = 1; k < 3; k++)
                    w[k] = (uint32_t)S.C[cpl][4 * k - 2][8 + cln] | (uint32_t)S.C[cpl][4 * k - 1][8 + cln] << 8 | (uint32_t)S.C[cpl][4 * k][8 + cln] << 16 |
                           (uint32_t)S.C[cpl][4 * k + 1][8 + cln] << 24;
            }
            filter_line(w, S.rec, 1, lane);
            if (lane < 16) {
#pragma unroll
                for (int k = 0; k < 5; k++)
#pragma unroll
                    for (int b = (k == 0 ? 1 : 0); b < (k == 4 ? 3 : 4); b++) S.L[4 * k + b][16 + lane] = (uint8_t)(w[k] >> (8 * b));
            } else {
                S.C[cpl][1][8 + cln] = (uint8_t)(w[0] >> 24);
#pragma unroll
                for (int k = 1; k < 3; k++)
#pragma unroll
                    for (int b = 0; b < 4; b++) S.C[cpl][4 * k - 2 + b][8 + cln] = (uint8_t)(w[k] >> (8 * b));
            }
            __syncwarp();

            // Release stores wait for every earlier store of the thread to be acknowledged; this step's global write-back is
            // therefore issued AFTER them (measured: a release right behind the write-back costs ~1 us per step), and the
            // previous step's write-back has had the whole filter time to drain.
            if (top_local && lane == 0) st_cluster_release(up_credit, mb_x + 1); // the ring slot may be reused
            // ---- hand the now-final bottom rows of the PREVIOUS macroblock (its right columns were just filtered) to the row
            // below, and at the end of the row this macroblock's too
            if (push_down) {
                const bool last = mb_x + 1 == a.W;
                if (mb_x > 0 || last) {
                    const int n_push = (mb_x > 0) + last, first = mb_x > 0 ? mb_x - 1 : mb_x;
                    while (credit < first + n_push - DB_RING) credit = ld_cluster_acquire(&S.credit); // ring full: wait for the row below
                    if (mb_x > 0) { // S.prev holds MB x-1's rows 12..15 / chroma rows 6,7 with stale right columns: patch them from the tile
                        if (lane < 4) *(uint32_t *)&S.prev.y[lane][12] = *(const uint32_t *)&S.L[16 + lane][12];
                        else if (lane < 8) *(uint16_t *)&S.prev.c[(lane - 4) >> 1][(lane - 4) & 1][6] = *(const uint16_t *)&S.C[(lane - 4) >> 1][8 + ((lane - 4) & 1)][6];
                        __syncwarp();
                        const uint32_t slot = down_ring + (uint32_t)((mb_x - 1) % DB_RING) * (uint32_t)sizeof(TopSlot);
                        if (lane < 4) st_cluster_v4(slot + 16 * lane, *(const uint4 *)S.prev.y[lane]);
                        else if (lane < 8) st_cluster_v2(slot + 64 + 8 * (lane - 4), *(const uint2 *)S.prev.c[(lane - 4) >> 1][(lane - 4) & 1]);
                    }
                    if (last) { // no right neighbour: this macroblock is final already
                        const uint32_t slot = down_ring + (uint32_t)(mb_x % DB_RING) * (uint32_t)sizeof(TopSlot);
                        if (lane < 4) st_cluster_v4(slot + 16 * lane, *(const uint4 *)&S.L[16 + lane][16]);
                        else if (lane < 8) st_cluster_v2(slot + 64 + 8 * (lane - 4), *(const uint2 *)&S.C[(lane - 4) >> 1][8 + ((lane - 4) & 1)][8]);
                    }
                    __syncwarp();
                    if (lane == 0) st_cluster_release(down_avail, mb_x + last);
                }
                __syncwarp();
                if (lane < 4) *(uint4 *)S.prev.y[lane] = *(const uint4 *)&S.L[16 + lane][16];
                else if (lane < 8) *(uint2 *)S.prev.c[(lane - 4) >> 1][(lane - 4) & 1] = *(const uint2 *)&S.C[(lane - 4) >> 1][8 + ((lane - 4) & 1)][8];
            }
            __syncwarp();
            // ---- write back to global memory (every row: the frame must end up there)
            // (luma rows 13..15 and chroma row 7 are left to the in-cluster row below: it writes their final values after its
            // top-edge filter, and an earlier value written from here after the hand-over could land on top of them)
            if (lane < 16) {
                if (!(push_down && lane >= 13)) {
                    uint8_t *d = rowy + (size_t)lane * a.stride + 16 * mb_x;
                    *(uint4 *)d = *(const uint4 *)&S.L[4 + lane][16];
                    if (mb_x > 0) *(uint32_t *)(d - 4) = *(const uint32_t *)&S.L[4 + lane][12];
                }
                if (!(push_down && (lane & 7) == 7)) {
                    uint8_t *dc = (lane < 8 ? rowu : rowv) + (size_t)(lane & 7) * a.stride_c + 8 * mb_x;
                    *(uint2 *)dc = *(const uint2 *)&S.C[lane >> 3][2 + (lane & 7)][8];
                    if (mb_x > 0) *(uint16_t *)(dc - 2) = *(const uint16_t *)&S.C[lane >> 3][2 + (lane & 7)][6];
                }
            } else if (mb_y > 0) {
                if (lane < 19) *(uint4 *)(rowy - (size_t)(19 - lane) * a.stride + 16 * mb_x) = *(const uint4 *)&S.L[lane - 15][16]; // rows -3..-1
                else if (lane < 21) *(uint2 *)((lane == 19 ? rowu : rowv) - a.stride_c + 8 * mb_x) = *(const uint2 *)&S.C[lane - 19][1][8];
            }
            __syncwarp();
            if (lane < 16) { // the right columns stay in shared memory as the next macroblock's left neighbour
                *(uint32_t *)&S.L[4 + lane][12] = *(const uint32_t *)&S.L[4 + lane][28];
                *(uint32_t *)&S.C[lane >> 3][2 + (lane & 7)][4] = *(const uint32_t *)&S.C[lane >> 3][2 + (lane & 7)][12];
            }
            __syncwarp();
            if (publish && lane == 0) st_release(a.progress + mb_y, mb_x + 1);
            if (flag_probe >= 0) {
                seen = max(seen, flag_probe);
                if (seen >= min(mb_x + 3, a.W)) {
                    if (lane < 4) tl = __ldcg((const uint4 *)(rowy - (size_t)(4 - lane) * a.stride + 16 * (mb_x + 1)));
                    else if (lane < 8) tc2 = __ldcg((const uint2 *)(((lane - 4) >> 1 ? rowv : rowu) - (size_t)(2 - ((lane - 4) & 1)) * a.stride_c + 8 * (mb_x + 1)));
                    have_top = true;
                }
            }
        }
    }
    // nobody leaves while a neighbour may still write credits / rows into its shared memory
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

} // namespace

extern "C" int x264_cuda_frame_deblock_dev(x264_cuda_t *ctx, x264_cuda_frame_t *fdec, const x264_cuda_deblock_params_t *pm, const int8_t *d_type,
                                           const int8_t *d_qp, const int8_t *d_transform8x8, const uint8_t *d_nnz, const int8_t *d_ref0,
                                           const int16_t *d_mv0, const int8_t *d_ref1, const int16_t *d_mv1)
{
    x264_cuda_enter(ctx);
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
    const size_t n_edges = (size_t)fdec->g.mb_width * H * 8;
    if (n_edges * sizeof(EdgeRec) > ctx->d_deblock_recs_size) {
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->d_deblock_recs); ctx->d_deblock_recs = nullptr; ctx->d_deblock_recs_size = 0;
        CUDA_TRY(ctx, cudaMalloc(&ctx->d_deblock_recs, n_edges * sizeof(EdgeRec)));
        ctx->d_deblock_recs_size = n_edges * sizeof(EdgeRec);
    }
    DeblockArgs a;
    a.y = fdec->plane[0]; a.u = fdec->chroma[0]; a.v = fdec->chroma[1];
    a.stride = fdec->g.stride; a.stride_c = fdec->stride_c; a.W = fdec->g.mb_width; a.H = H;
    a.type = d_type; a.qp = d_qp; a.t8 = d_transform8x8; a.nnz = d_nnz;
    a.ref[0] = d_ref0; a.ref[1] = d_ref1; a.mv[0] = d_mv0; a.mv[1] = d_mv1;
    a.alpha_off = pm->alpha_c0_offset; a.beta_off = pm->beta_offset; a.chroma_off = pm->chroma_qp_offset;
    a.slice_b = !!pm->b_slice_b; a.psub8x8 = !!pm->b_psub8x8; a.cavlc8 = !!pm->b_cavlc_8x8dct;
    a.progress = ctx->d_deblock_progress;
    // one single-warp CTA per macroblock row in clusters of eight; all of them are resident at once (<= a few hundred), so
    // a row can always wait for the row above
    deblock_prep_kernel<<<(int)((n_edges + 255) / 256), 256, 0, ctx->stream>>>(a, (EdgeRec *)ctx->d_deblock_recs);
    LAUNCH_CHECK(ctx, "deblock_prep_kernel");
    deblock_rows_cluster_kernel<<<(H + DB_CLUSTER - 1) / DB_CLUSTER * DB_CLUSTER, 32, 0, ctx->stream>>>(a, (const EdgeRec *)ctx->d_deblock_recs);
    LAUNCH_CHECK(ctx, "deblock_rows_cluster_kernel");
    return 0;
}

extern "C" int x264_cuda_frame_deblock(x264_cuda_t *ctx, x264_cuda_frame_t *fdec, const x264_cuda_deblock_params_t *pm, const int8_t *type,
                                       const int8_t *qp, const int8_t *transform8x8, const uint8_t (*nnz)[24], const int8_t *ref0,
                                       const int16_t (*mv0)[2], const int8_t *ref1, const int16_t (*mv1)[2])
{
    x264_cuda_enter(ctx);
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
