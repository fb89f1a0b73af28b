// intra_dev.cuh — device-side intra predictors for whole 16x16 luma / 8x8 chroma blocks (S/common/predict.c:40-170, :172-336), produced
// 8x4 unit by 8x4 unit from an edge vector so that a warp can spread (candidate, unit) pairs over its lanes.  Shared by the frame-batched
// kernel (intra.cu) and the per-call table entries (tables.cu).
#pragma once
#include "pixel_dev.cuh"

// per-warp edge store: [3] corner, [4..4+n) row above, [4+n..4+2n) left column
struct IntraEdges { uint8_t y[36], u[20], v[20]; };

template <int N> __device__ __forceinline__ int e_top(const uint8_t *e, int k) { return k < 0 ? e[3] : e[4 + k]; }
template <int N> __device__ __forceinline__ int e_left(const uint8_t *e, int k) { return k < 0 ? e[3] : e[4 + N + k]; }

// one 8x4 unit (rows y0..y0+3, columns x0..x0+7) of predictor `kind` for an NxN block: 0 V, 1 H, 2 DC (the value in dcq[]: four
// quadrant values for chroma, one for luma), 3 plane
template <int N>
__device__ __forceinline__ void predict_unit(const uint8_t *e, int kind, const int (&dcq)[4], int x0, int y0, uint2 (&r)[4])
{
    if (kind == 0) {
        const uint2 t = make_uint2(*(const uint32_t *)(e + 4 + x0), *(const uint32_t *)(e + 8 + x0));
#pragma unroll
        for (int k = 0; k < 4; k++) r[k] = t;
    } else if (kind == 1) {
#pragma unroll
        for (int k = 0; k < 4; k++) { const uint32_t v = e[4 + N + y0 + k] * 0x01010101u; r[k] = make_uint2(v, v); }
    } else if (kind == 2) {
        const int q = N == 8 ? (y0 >> 2) * 2 : 0;
        const uint2 v = make_uint2((uint32_t)dcq[q] * 0x01010101u, (uint32_t)dcq[N == 8 ? q + 1 : 0] * 0x01010101u);
#pragma unroll
        for (int k = 0; k < 4; k++) r[k] = v;
    } else { // predict_16x16_p (predict.c:134-167) / predict_8x8c_p (:305-336)
        constexpr int HALF = N / 2;
        int H = 0, V = 0;
#pragma unroll
        for (int i = 0; i < HALF; i++) {
            H += (i + 1) * (e_top<N>(e, HALF + i) - e_top<N>(e, HALF - 2 - i));
            V += (i + 1) * (e_left<N>(e, HALF + i) - e_left<N>(e, HALF - 2 - i));
        }
        const int a = 16 * (e[4 + N + N - 1] + e[4 + N - 1]);
        const int b = N == 16 ? (5 * H + 32) >> 6 : (17 * H + 16) >> 5, c = N == 16 ? (5 * V + 32) >> 6 : (17 * V + 16) >> 5;
        const int i00 = a - (HALF - 1) * (b + c) + 16;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint32_t w[2] = { 0, 0 };
#pragma unroll
            for (int x = 0; x < 8; x++) w[x >> 2] |= (uint32_t)clip_u8((i00 + b * (x0 + x) + c * (y0 + k)) >> 5) << (8 * (x & 3));
            r[k] = make_uint2(w[0], w[1]);
        }
    }
}

__device__ __forceinline__ int unit_metric(bool satd, const uint2 (&f)[4], const uint2 (&r)[4])
{
    if (satd) return satd_8x4_rows(f, r);
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { s = sad4_acc(f[k].x, r[k].x, s); s = sad4_acc(f[k].y, r[k].y, s); }
    return (int)s;
}

