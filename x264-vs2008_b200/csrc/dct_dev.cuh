// dct_dev.cuh — device-side H.264 integer transforms and (de)quantisation primitives (S/common/dct.c, quant.c),
// shared by the batched residual kernels and the per-call plugin-table entries.
#pragma once
#include "common.cuh"

__device__ __forceinline__ int s16(int v) { return (int)(int16_t)v; }

// ---- 4x4 core transform (dct.c:122-155).  d: 16 residuals row-major; out: coefficient block, reference layout
__device__ __forceinline__ void fwd4x4(const int (&d)[16], int (&o)[16])
{
    int t[16];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int s03 = d[i * 4] + d[i * 4 + 3], s12 = d[i * 4 + 1] + d[i * 4 + 2];
        const int d03 = d[i * 4] - d[i * 4 + 3], d12 = d[i * 4 + 1] - d[i * 4 + 2];
        t[0 * 4 + i] = s03 + s12; t[1 * 4 + i] = 2 * d03 + d12; t[2 * 4 + i] = s03 - s12; t[3 * 4 + i] = d03 - 2 * d12;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int s03 = t[i * 4] + t[i * 4 + 3], s12 = t[i * 4 + 1] + t[i * 4 + 2];
        const int d03 = t[i * 4] - t[i * 4 + 3], d12 = t[i * 4 + 1] - t[i * 4 + 2];
        o[i * 4 + 0] = s03 + s12; o[i * 4 + 1] = 2 * d03 + d12; o[i * 4 + 2] = s03 - s12; o[i * 4 + 3] = d03 - 2 * d12;
    }
}
// inverse (dct.c:174-216): returns the 16 residuals to add, row-major (d[y][x])
__device__ __forceinline__ void inv4x4(const int (&c)[16], int (&r)[16])
{
    int t[16];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int s02 = c[0 * 4 + i] + c[2 * 4 + i], d02 = c[0 * 4 + i] - c[2 * 4 + i];
        const int s13 = c[1 * 4 + i] + (c[3 * 4 + i] >> 1), d13 = (c[1 * 4 + i] >> 1) - c[3 * 4 + i];
        t[i * 4 + 0] = s16(s02 + s13); t[i * 4 + 1] = s16(d02 + d13); t[i * 4 + 2] = s16(d02 - d13); t[i * 4 + 3] = s16(s02 - s13);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int s02 = t[0 * 4 + i] + t[2 * 4 + i], d02 = t[0 * 4 + i] - t[2 * 4 + i];
        const int s13 = t[1 * 4 + i] + (t[3 * 4 + i] >> 1), d13 = (t[1 * 4 + i] >> 1) - t[3 * 4 + i];
        r[0 * 4 + i] = s16((s02 + s13 + 32) >> 6); r[1 * 4 + i] = s16((d02 + d13 + 32) >> 6);
        r[2 * 4 + i] = s16((d02 - d13 + 32) >> 6); r[3 * 4 + i] = s16((s02 - s13 + 32) >> 6);
    }
}
// quant.c:33-40 on one coefficient
__device__ __forceinline__ int quant1(int c, int mf, int f) { return s16(c > 0 ? ((f + c) * mf >> 16) : -((f - c) * mf >> 16)); }
// quant.c:76-109 / :111-146 on one coefficient
__device__ __forceinline__ int dequant1(int c, int dmf, int qbits)
{
    return qbits >= 0 ? s16((c * dmf) << qbits) : s16((c * dmf + (1 << (-qbits - 1))) >> (-qbits));
}

// ---- 8-point transforms (dct.c:238-263, :295-320)
__device__ __forceinline__ void fwd8(const int (&s)[8], int (&o)[8])
{
    const int s07 = s[0] + s[7], s16_ = s[1] + s[6], s25 = s[2] + s[5], s34 = s[3] + s[4];
    const int a0 = s07 + s34, a1 = s16_ + s25, a2 = s07 - s34, a3 = s16_ - s25;
    const int d07 = s[0] - s[7], d16 = s[1] - s[6], d25 = s[2] - s[5], d34 = s[3] - s[4];
    const int a4 = d16 + d25 + (d07 + (d07 >> 1)), a5 = d07 - d34 - (d25 + (d25 >> 1));
    const int a6 = d07 + d34 - (d16 + (d16 >> 1)), a7 = d16 - d25 + (d34 + (d34 >> 1));
    o[0] = a0 + a1; o[1] = a4 + (a7 >> 2); o[2] = a2 + (a3 >> 1); o[3] = a5 + (a6 >> 2);
    o[4] = a0 - a1; o[5] = a6 - (a5 >> 2); o[6] = (a2 >> 1) - a3; o[7] = (a4 >> 2) - a7;
}
__device__ __forceinline__ void inv8(const int (&s)[8], int (&o)[8])
{
    const int a0 = s[0] + s[4], a2 = s[0] - s[4], a4 = (s[2] >> 1) - s[6], a6 = (s[6] >> 1) + s[2];
    const int b0 = a0 + a6, b2 = a2 + a4, b4 = a2 - a4, b6 = a0 - a6;
    const int a1 = -s[3] + s[5] - s[7] - (s[7] >> 1), a3 = s[1] + s[7] - s[3] - (s[3] >> 1);
    const int a5 = -s[1] + s[7] + s[5] + (s[5] >> 1), a7 = s[3] + s[5] + s[1] + (s[1] >> 1);
    const int b1 = (a7 >> 2) + a1, b3 = a3 + (a5 >> 2), b5 = (a3 >> 2) - a5, b7 = a7 - (a1 >> 2);
    o[0] = b0 + b7; o[1] = b2 + b5; o[2] = b4 + b3; o[3] = b6 + b1;
    o[4] = b6 - b1; o[5] = b4 - b3; o[6] = b2 - b5; o[7] = b0 - b7;
}
// sub8x8_dct8 (dct.c:265-285): d row-major residuals -> c in reference layout (c[x*8+i])
static __device__ void fwd8x8(int *d /*64, clobbered*/, int *c)
{
    for (int i = 0; i < 8; i++) { // columns, in place
        int s[8], o[8];
#pragma unroll
        for (int k = 0; k < 8; k++) s[k] = d[k * 8 + i];
        fwd8(s, o);
#pragma unroll
        for (int k = 0; k < 8; k++) d[k * 8 + i] = s16(o[k]);
    }
    for (int i = 0; i < 8; i++) { // rows, written transposed
        int s[8], o[8];
#pragma unroll
        for (int k = 0; k < 8; k++) s[k] = d[i * 8 + k];
        fwd8(s, o);
#pragma unroll
        for (int k = 0; k < 8; k++) c[k * 8 + i] = s16(o[k]);
    }
}
// add8x8_idct8 (dct.c:322-341): c (clobbered) -> r residuals with r[k*8+i] added to pixel (row k, col i)
static __device__ void inv8x8(int *c, int *r)
{
    c[0] = s16(c[0] + 32);
    for (int i = 0; i < 8; i++) {
        int s[8], o[8];
#pragma unroll
        for (int k = 0; k < 8; k++) s[k] = c[k * 8 + i];
        inv8(s, o);
#pragma unroll
        for (int k = 0; k < 8; k++) c[k * 8 + i] = s16(o[k]);
    }
    for (int i = 0; i < 8; i++) {
        int s[8], o[8];
#pragma unroll
        for (int k = 0; k < 8; k++) s[k] = c[i * 8 + k];
        inv8(s, o);
#pragma unroll
        for (int k = 0; k < 8; k++) r[k * 8 + i] = o[k] >> 6; // dst[i + k*FDEC_STRIDE]
    }
}

// dct4x4dc / idct4x4dc (dct.c:39-105)
__device__ __forceinline__ void hadamard_dc(int (&d)[16], bool fwd)
{
    int t[16];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int s01 = d[i * 4] + d[i * 4 + 1], d01 = d[i * 4] - d[i * 4 + 1];
        const int s23 = d[i * 4 + 2] + d[i * 4 + 3], d23 = d[i * 4 + 2] - d[i * 4 + 3];
        t[0 * 4 + i] = s16(s01 + s23); t[1 * 4 + i] = s16(s01 - s23); t[2 * 4 + i] = s16(d01 - d23); t[3 * 4 + i] = s16(d01 + d23);
    }
    const int r = fwd ? 1 : 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int s01 = t[i * 4] + t[i * 4 + 1], d01 = t[i * 4] - t[i * 4 + 1];
        const int s23 = t[i * 4 + 2] + t[i * 4 + 3], d23 = t[i * 4 + 2] - t[i * 4 + 3];
        d[i * 4 + 0] = s16((s01 + s23 + r) >> r); d[i * 4 + 1] = s16((s01 - s23 + r) >> r);
        d[i * 4 + 2] = s16((d01 - d23 + r) >> r); d[i * 4 + 3] = s16((d01 + d23 + r) >> r);
    }
}

