// pixel_dev.cuh — device-side block metrics and sub-pel sample fetch shared by the search kernels.
//
//  * SAD  : VABSDIFF4.U8.ACC per 4 pixels (S/common/pixel.c:40-65).
//  * SATD : the reference's own packed arithmetic (two 16-bit lanes per 32-bit word, HADAMARD4 + abs2,
//           S/common/pixel.c:164-231) executed verbatim on the integer pipe, so lane overflow behaviour is identical.
//  * SA8D : same idea for the 8x8 Hadamard (pixel.c:256-303).
//  * qpel : get_ref / mc_luma semantics (S/common/mc.c:157-202): a sample is either one of the four half-pel planes
//           or the rounded byte average of two of them (== __vavgu4).
#pragma once
#include "common.cuh"

// unaligned 8 / 4 byte fetch through the read-only path
__device__ __forceinline__ uint2 ldg8(const uint8_t *a)
{
    const int sh = ((uintptr_t)a & 3) * 8;
    const uint32_t *p = (const uint32_t *)((uintptr_t)a & ~(uintptr_t)3);
    const uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
    return make_uint2(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh));
}
__device__ __forceinline__ uint32_t ldg4(const uint8_t *a)
{
    const int sh = ((uintptr_t)a & 3) * 8;
    const uint32_t *p = (const uint32_t *)((uintptr_t)a & ~(uintptr_t)3);
    return __funnelshift_r(__ldg(p), __ldg(p + 1), sh);
}

// A block of reference samples at a quarter-pel position: p1 (+ optional p2 to average with), mc.c:181-202
struct QpelSrc { const uint8_t *p1, *p2; };
__device__ __forceinline__ QpelSrc qpel_src(const uint8_t *const (&planes)[4], int stride, int mvx, int mvy)
{
    // hpel_ref0 / hpel_ref1 (mc.c:157-158) packed as sixteen 2-bit fields
    const int qidx = ((mvy & 3) << 2) + (mvx & 3);
    const int i0 = (0x54fe5454u >> (2 * qidx)) & 3, i1 = (0xbababa00u >> (2 * qidx)) & 3;
    const ptrdiff_t off = (ptrdiff_t)(mvy >> 2) * stride + (mvx >> 2);
    QpelSrc s;
    s.p1 = planes[i0] + off + ((mvy & 3) == 3) * stride;
    s.p2 = (qidx & 5) ? planes[i1] + off + ((mvx & 3) == 3) : nullptr;
    return s;
}
__device__ __forceinline__ uint2 qpel_row8(const QpelSrc &s, ptrdiff_t o)
{
    uint2 a = ldg8(s.p1 + o);
    if (s.p2) { const uint2 b = ldg8(s.p2 + o); a.x = __vavgu4(a.x, b.x); a.y = __vavgu4(a.y, b.y); }
    return a;
}
__device__ __forceinline__ uint32_t qpel_row4(const QpelSrc &s, ptrdiff_t o)
{
    uint32_t a = ldg4(s.p1 + o);
    if (s.p2) a = __vavgu4(a, ldg4(s.p2 + o));
    return a;
}

// ---- packed 2x16-bit helpers, pixel.c:164-181
__device__ __forceinline__ uint32_t abs2(uint32_t a)
{
    const uint32_t s = ((a >> 15) & 0x10001u) * 0xffffu;
    return (a + s) ^ s;
}
#define HADAMARD4_PK(d0, d1, d2, d3, s0, s1, s2, s3) { \
    const uint32_t t0_ = (s0) + (s1), t1_ = (s0) - (s1), t2_ = (s2) + (s3), t3_ = (s2) - (s3); \
    d0 = t0_ + t2_; d2 = t0_ - t2_; d1 = t1_ + t3_; d3 = t1_ - t3_; }

// (byte k of lo) | (byte k of hi) << 16, for k = 0..3
__device__ __forceinline__ void unpack_pairs(uint32_t lo, uint32_t hi, uint32_t (&o)[4])
{
    const uint32_t x = __byte_perm(lo, hi, 0x5140), y = __byte_perm(lo, hi, 0x7362); // (l0,h0,l1,h1), (l2,h2,l3,h3)
    o[0] = __byte_perm(x, 0, 0x4140); o[1] = __byte_perm(x, 0, 0x4342);
    o[2] = __byte_perm(y, 0, 0x4140); o[3] = __byte_perm(y, 0, 0x4342);
}

// x264_pixel_satd_8x4 (pixel.c:212-231) on 4 rows of (fenc 8 px, ref 8 px) given as word pairs; returns the
// UNSHIFTED packed sum folded to an int: caller applies the single >>1
__device__ __forceinline__ int satd_8x4_rows(const uint2 (&f)[4], const uint2 (&r)[4])
{
    uint32_t tmp[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint32_t a[4], b[4];
        unpack_pairs(f[i].x, f[i].y, a);
        unpack_pairs(r[i].x, r[i].y, b);
        HADAMARD4_PK(tmp[i][0], tmp[i][1], tmp[i][2], tmp[i][3], a[0] - b[0], a[1] - b[1], a[2] - b[2], a[3] - b[3]);
    }
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint32_t a0, a1, a2, a3;
        HADAMARD4_PK(a0, a1, a2, a3, tmp[0][i], tmp[1][i], tmp[2][i], tmp[3][i]);
        sum += abs2(a0) + abs2(a1) + abs2(a2) + abs2(a3);
    }
    return (int)(((sum & 0xffff) + (sum >> 16)) >> 1);
}
// x264_pixel_satd_4x4 (pixel.c:187-210): one 4x4; rows as single words
__device__ __forceinline__ int satd_4x4_rows(const uint32_t (&f)[4], const uint32_t (&r)[4])
{
    // same value as the reference: sum |H4 D H4| >> 1.  Pack rows (0,1 | 2,3)?  No: keep the reference's structure,
    // lanes = (a0+a1, a0-a1): build per row b0 = (d0+d1) + ((d0-d1)<<16), b1 likewise for d2,d3.
    uint32_t tmp[4][2];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int d0 = (int)(f[i] & 255) - (int)(r[i] & 255), d1 = (int)((f[i] >> 8) & 255) - (int)((r[i] >> 8) & 255);
        const int d2 = (int)((f[i] >> 16) & 255) - (int)((r[i] >> 16) & 255), d3 = (int)(f[i] >> 24) - (int)(r[i] >> 24);
        const uint32_t b0 = (uint32_t)(d0 + d1) + ((uint32_t)(d0 - d1) << 16);
        const uint32_t b1 = (uint32_t)(d2 + d3) + ((uint32_t)(d2 - d3) << 16);
        tmp[i][0] = b0 + b1; tmp[i][1] = b0 - b1;
    }
    int sum = 0;
#pragma unroll
    for (int i = 0; i < 2; i++) {
        uint32_t a0, a1, a2, a3;
        HADAMARD4_PK(a0, a1, a2, a3, tmp[0][i], tmp[1][i], tmp[2][i], tmp[3][i]);
        a0 = abs2(a0) + abs2(a1) + abs2(a2) + abs2(a3);
        sum += (int)(a0 & 0xffff) + (int)(a0 >> 16);
    }
    return sum >> 1;
}

// pixel_hadamard_ac (pixel.c:306-341) of one 8x8 block, the reference's packed 2x16-bit arithmetic executed verbatim
__device__ inline uint64_t hadamard_ac_8x8(const uint8_t *pix, int stride)
{
    uint32_t tmp[32];
    uint32_t sum4 = 0, sum8 = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint2 r = __ldg((const uint2 *)(pix + (size_t)i * stride));
        const uint32_t p0 = r.x & 255, p1 = (r.x >> 8) & 255, p2 = (r.x >> 16) & 255, p3 = r.x >> 24, p4 = r.y & 255, p5 = (r.y >> 8) & 255,
                       p6 = (r.y >> 16) & 255, p7 = r.y >> 24;
        const int t = (i & 3) + (i & 4) * 4;
        const uint32_t a0 = (p0 + p1) + ((p0 - p1) << 16), a1 = (p2 + p3) + ((p2 - p3) << 16);
        tmp[t] = a0 + a1; tmp[t + 4] = a0 - a1;
        const uint32_t a2 = (p4 + p5) + ((p4 - p5) << 16), a3 = (p6 + p7) + ((p6 - p7) << 16);
        tmp[t + 8] = a2 + a3; tmp[t + 12] = a2 - a3;
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t a0, a1, a2, a3;
        HADAMARD4_PK(a0, a1, a2, a3, tmp[i * 4 + 0], tmp[i * 4 + 1], tmp[i * 4 + 2], tmp[i * 4 + 3]);
        tmp[i * 4 + 0] = a0; tmp[i * 4 + 1] = a1; tmp[i * 4 + 2] = a2; tmp[i * 4 + 3] = a3;
        sum4 += abs2(a0) + abs2(a1) + abs2(a2) + abs2(a3);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t a0, a1, a2, a3;
        HADAMARD4_PK(a0, a1, a2, a3, tmp[i], tmp[8 + i], tmp[16 + i], tmp[24 + i]);
        sum8 += abs2(a0) + abs2(a1) + abs2(a2) + abs2(a3);
    }
    const uint32_t dc = (uint16_t)(tmp[0] + tmp[8] + tmp[16] + tmp[24]);
    const int s4 = (int)((uint16_t)sum4 + (sum4 >> 16) - dc), s8 = (int)((uint16_t)sum8 + (sum8 >> 16) - dc);
    return ((uint64_t)s8 << 32) + s4; // int -> uint64 conversions as in the reference (sum4 is added sign-extended)
}

// sa8d_8x8 (pixel.c:256-288): raw sum, rows as word pairs
__device__ __forceinline__ int sa8d_8x8_rows(const uint2 (&f)[8], const uint2 (&r)[8])
{
    uint32_t tmp[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        int d[8];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            d[k] = (int)((f[i].x >> (8 * k)) & 255) - (int)((r[i].x >> (8 * k)) & 255);
            d[k + 4] = (int)((f[i].y >> (8 * k)) & 255) - (int)((r[i].y >> (8 * k)) & 255);
        }
        const uint32_t b0 = (uint32_t)(d[0] + d[1]) + ((uint32_t)(d[0] - d[1]) << 16), b1 = (uint32_t)(d[2] + d[3]) + ((uint32_t)(d[2] - d[3]) << 16);
        const uint32_t b2 = (uint32_t)(d[4] + d[5]) + ((uint32_t)(d[4] - d[5]) << 16), b3 = (uint32_t)(d[6] + d[7]) + ((uint32_t)(d[6] - d[7]) << 16);
        HADAMARD4_PK(tmp[i][0], tmp[i][1], tmp[i][2], tmp[i][3], b0, b1, b2, b3);
    }
    int sum = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint32_t a0, a1, a2, a3, a4, a5, a6, a7;
        HADAMARD4_PK(a0, a1, a2, a3, tmp[0][i], tmp[1][i], tmp[2][i], tmp[3][i]);
        HADAMARD4_PK(a4, a5, a6, a7, tmp[4][i], tmp[5][i], tmp[6][i], tmp[7][i]);
        uint32_t b0 = abs2(a0 + a4) + abs2(a0 - a4);
        b0 += abs2(a1 + a5) + abs2(a1 - a5);
        b0 += abs2(a2 + a6) + abs2(a2 - a6);
        b0 += abs2(a3 + a7) + abs2(a3 - a7);
        sum += (int)(b0 & 0xffff) + (int)(b0 >> 16);
    }
    return sum;
}

// ---- one "unit" of a block metric: the piece one lane computes.  Blocks with w >= 8 are tiled by 8x4 units, 4-wide
// blocks by 4x4 units (exactly the granularity at which the reference halves its SATD sums, PIXEL_SATD_C :233-253).
// fe: fenc pixel pointer of the unit (aligned to 4, given stride); src: qpel source positioned at the BLOCK origin.
__device__ __forceinline__ int unit_cost(bool satd, int bw, const uint8_t *fe, int fstride, const QpelSrc &src, int rstride, int ux, int uy)
{
    if (bw >= 8) {
        uint2 f[4], r[4];
#pragma unroll
        for (int y = 0; y < 4; y++) {
            const uint32_t *fp = (const uint32_t *)(fe + (size_t)(uy + y) * fstride + ux);
            f[y] = make_uint2(__ldg(fp), __ldg(fp + 1));
            r[y] = qpel_row8(src, (ptrdiff_t)(uy + y) * rstride + ux);
        }
        if (satd) return satd_8x4_rows(f, r);
        uint32_t acc = 0;
#pragma unroll
        for (int y = 0; y < 4; y++) { acc = sad4_acc(f[y].x, r[y].x, acc); acc = sad4_acc(f[y].y, r[y].y, acc); }
        return (int)acc;
    } else {
        uint32_t f[4], r[4];
#pragma unroll
        for (int y = 0; y < 4; y++) {
            f[y] = __ldg((const uint32_t *)(fe + (size_t)(uy + y) * fstride + ux));
            r[y] = qpel_row4(src, (ptrdiff_t)(uy + y) * rstride + ux);
        }
        if (satd) return satd_4x4_rows(f, r);
        uint32_t acc = 0;
#pragma unroll
        for (int y = 0; y < 4; y++) acc = sad4_acc(f[y], r[y], acc);
        return (int)acc;
    }
}
// number of units of a bw x bh block and the position of unit u
__device__ __forceinline__ int unit_count(int bw, int bh) { return (bw >= 8 ? bw / 8 : 1) * (bh / 4); }
__device__ __forceinline__ void unit_pos(int bw, int u, int &ux, int &uy)
{
    const int per_row = bw >= 8 ? bw / 8 : 1;
    ux = (u % per_row) * 8; uy = (u / per_row) * 4;
}
