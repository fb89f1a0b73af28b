// frame_filter.cu — whole-frame streaming kernels (HBM-bound class): border replication, the three 6-tap
// half-pel planes, the integral image(s) and the four half-resolution lookahead planes.
//
// Reference semantics (all on the padded plane layout of S/common/frame.c:29-152):
//   border   : x264_frame_expand_border_mod16 + x264_frame_expand_border   (frame.c:304-331, :218-267)
//   hpel     : hpel_filter driven by x264_frame_filter(h,f,0,1)             (mc.c:133-155, :404-426)
//              + x264_frame_expand_border_filtered                           (frame.c:269-295)
//   integral : integral_init4h/8h/4v/8v as sequenced by x264_frame_filter   (mc.c:270-304, :428-461)
//   lowres   : frame_init_lowres_core + x264_frame_expand_border_lowres     (mc.c:333-357, frame.c:297-302)
//
// Every border operation of the reference is a replication, so "compute core, then replicate" collapses to
// out[y][x] = core[clamp(y)][clamp(x)], which is how the kernels below are written (one generic clamp-copy).
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------------------
// Generic replicate-border kernel: for every 4-byte word of the extent [X0,X1) x [Y0,Y1) that is not entirely
// inside the core [cx0,cx1] x [cy0,cy1], write core[clamp(y)][clamp(x)].  x coordinates are relative to the
// plane's pixel (0,0); X0 is a multiple of 4.  `zero_from`: columns >= zero_from (and <= cx1) of core rows are
// treated as never-written zeros (lowres planes of odd-mb-width frames, see x264_cuda_frame_init_lowres).
struct BorderArgs { uint8_t *p[4]; int stride, cx0, cx1, cy0, cy1, X0, X1, Y0, Y1, words, nl, wr0, side, band_items, items; };

// Clamp-copy of everything outside [cx0,cx1] x [cy0,cy1] inside [X0,X1) x [Y0,Y1): one thread per BORDER word only.  Rows above
// and below the valid area are whole rows of `words` words; rows inside it only have `side` words (nl on the left, the rest from
// word wr0 on the right), so a 1080p plane is 49k threads instead of 571k.
__global__ void __launch_bounds__(256) replicate_border_kernel(BorderArgs a)
{
    uint8_t *plane = a.p[blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.items) return;
    int y, wi;
    if (i < a.band_items) {
        const int r = i / a.words, top = a.cy0 - a.Y0;
        wi = i - r * a.words;
        y = r < top ? a.Y0 + r : a.cy1 + 1 + (r - top);
    } else {
        const int j = i - a.band_items, r = j / a.side, k = j - r * a.side;
        y = a.cy0 + r;
        wi = k < a.nl ? k : a.wr0 + (k - a.nl);
    }
    const int x = a.X0 + wi * 4;
    const uint8_t *row = plane + (ptrdiff_t)clip3i(y, a.cy0, a.cy1) * a.stride;
    uint32_t v = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) v |= (uint32_t)row[clip3i(x + k, a.cx0, a.cx1)] << (8 * k);
    *(uint32_t *)(plane + (ptrdiff_t)y * a.stride + x) = v;
}

// ---------------------------------------------------------------------------------------------------------
// Half-pel planes (hpel_filter, S/common/mc.c:133-155).  One thread = 8 adjacent output columns (one 64-bit store per plane) walking
// down R rows.  The arithmetic is packed so that the kernel is bound by HBM rather than by the integer pipe:
//   * the source rows are unpacked ONCE into pairs of 16-bit lanes (7 registers = columns x-2 .. x+11) and kept in a 6-row register ring;
//   * the vertical 6-tap runs on both lanes of a register at a time: Vb = (a+f) + 20(c+d) - 5(b+e) + 2560 per lane — the bias keeps every
//     lane non-negative, so the 32-bit multiply-adds never borrow across lanes (|v| <= 10710 is the reference's int16 `buf`);
//   * V plane: ((Vb+16)>>5) - 80 == (v+16)>>5, clamped with the packed 16-bit min/max instructions;
//   * centre plane: horizontal 6-tap over the 16-bit vertical sums as three DP2A per pixel (pairs (1,-5) (20,20) (-5,1));
//   * H plane: horizontal 6-tap over the source bytes as two DP4A per pixel ((1,-5,20,20) and (-5,1,0,0));
//   * clamp-and-pack of four results = two I2IP (cvt.pack.sat.u8.s32).
__device__ __forceinline__ int dp2a_lo_ss(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// bytes (sat(p0), sat(p1), sat(p2), sat(p3)), low byte first
__device__ __forceinline__ uint32_t pack4_sat(int p0, int p1, int p2, int p3)
{
    uint32_t t, d;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(t) : "r"(p3), "r"(p2), "r"(0));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(p1), "r"(p0), "r"(t));
    return d;
}

template <int R>
__global__ void __launch_bounds__(128) hpel_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dh,
                                                   uint8_t *__restrict__ dv, uint8_t *__restrict__ dc, int stride,
                                                   int x_begin, int n_cols, int y_begin, int y_end)
{
    const int ti = blockIdx.x * blockDim.x + threadIdx.x;
    if (ti >= n_cols) return;
    const int x = x_begin + ti * 8; // multiple of 8 relative to pixel 0 => 8-byte aligned (PADH = 32, stride % 128 == 0)
    const int y0 = y_begin + blockIdx.y * R;
    uint32_t U[6][7]; // rows y-2 .. y+3; U[.][k] = (p[x-2+2k], p[x-1+2k]) as two 16-bit lanes
    auto load_unpack = [&](uint32_t (&u)[7], int y) {
        const uint8_t *p = src + (ptrdiff_t)y * stride + x;
        const uint2 m = __ldg((const uint2 *)p);
        const uint32_t w0 = __ldg((const uint32_t *)(p - 4)), w3 = __ldg((const uint32_t *)(p + 8));
        u[0] = __byte_perm(w0, 0, 0x4342);
        u[1] = __byte_perm(m.x, 0, 0x4140); u[2] = __byte_perm(m.x, 0, 0x4342);
        u[3] = __byte_perm(m.y, 0, 0x4140); u[4] = __byte_perm(m.y, 0, 0x4342);
        u[5] = __byte_perm(w3, 0, 0x4140); u[6] = __byte_perm(w3, 0, 0x4342);
    };
#pragma unroll
    for (int k = 0; k < 5; k++) load_unpack(U[k], y0 - 2 + k);
    // The row loop has no early exit (rows past y_end are computed from clamped addresses and not stored) and the raw words of the
    // rows two and three iterations ahead are already in flight: with a `break` per row the compiler cannot move a row's loads above
    // the previous row's stores, and every row then pays a full memory round trip.
    uint32_t raw[2][4];
    auto fetch = [&](uint32_t (&w)[4], int y) {
        const uint8_t *p = src + (ptrdiff_t)min(y, y_end + 2) * stride + x;
        const uint2 m = __ldg((const uint2 *)p);
        w[0] = __ldg((const uint32_t *)(p - 4)); w[1] = m.x; w[2] = m.y; w[3] = __ldg((const uint32_t *)(p + 8));
    };
    auto unpack = [&](uint32_t (&u)[7], const uint32_t (&w)[4]) {
        u[0] = __byte_perm(w[0], 0, 0x4342);
        u[1] = __byte_perm(w[1], 0, 0x4140); u[2] = __byte_perm(w[1], 0, 0x4342);
        u[3] = __byte_perm(w[2], 0, 0x4140); u[4] = __byte_perm(w[2], 0, 0x4342);
        u[5] = __byte_perm(w[3], 0, 0x4140); u[6] = __byte_perm(w[3], 0, 0x4342);
    };
    fetch(raw[0], y0 + 3); fetch(raw[1], y0 + 4);
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int y = y0 + r;
        const bool live = y < y_end;
        unpack(U[(r + 5) % 6], raw[r % 2]);
        if (r + 2 < R) fetch(raw[r % 2], y + 5);
        const ptrdiff_t o = (ptrdiff_t)min(y, y_end - 1) * stride + x;
        // ---- vertical 6-tap, two columns per register, biased by 2560
        uint32_t Vb[7];
#pragma unroll
        for (int k = 0; k < 7; k++) {
            const uint32_t af = U[r % 6][k] + U[(r + 5) % 6][k] + 0x0a000a00u;
            const uint32_t be = U[(r + 1) % 6][k] + U[(r + 4) % 6][k], cd = U[(r + 2) % 6][k] + U[(r + 3) % 6][k];
            Vb[k] = af + cd * 20u + be * 0xfffffffbu;
        }
        {   // ---- V plane: columns x .. x+7 are pairs 1..4
            uint32_t m[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t s = ((Vb[k + 1] + 0x00100010u) >> 5) & 0x07ff07ffu;
                m[k] = __viaddmin_s16x2_relu(s, 0xffb0ffb0u, 0x00ff00ffu); // max(min(s - 80, 255), 0) per lane
            }
            if (live) *(uint2 *)(dv + o) = make_uint2(__byte_perm(m[0], m[1], 0x6420), __byte_perm(m[2], m[3], 0x6420));
        }
        {   // ---- centre plane: 6-tap over the vertical sums; S[k] = (Vb[k].hi, Vb[k+1].lo) serves the odd columns
            uint32_t S[6];
#pragma unroll
            for (int k = 0; k < 6; k++) S[k] = __byte_perm(Vb[k], Vb[k + 1], 0x5432);
            int c[8];
            const int K = 512 - 32 * 2560; // rounding, minus the bias seen through coefficients that sum to 32
#pragma unroll
            for (int mm = 0; mm < 4; mm++) {
                c[2 * mm] = dp2a_lo_ss(Vb[mm + 2], 0x01fbu, dp2a_lo_ss(Vb[mm + 1], 0x1414u, dp2a_lo_ss(Vb[mm], 0xfb01u, K))) >> 10;
                c[2 * mm + 1] = dp2a_lo_ss(S[mm + 2], 0x01fbu, dp2a_lo_ss(S[mm + 1], 0x1414u, dp2a_lo_ss(S[mm], 0xfb01u, K))) >> 10;
            }
            if (live) *(uint2 *)(dc + o) = make_uint2(pack4_sat(c[0], c[1], c[2], c[3]), pack4_sat(c[4], c[5], c[6], c[7]));
        }
        {   // ---- H plane: 6-tap over the bytes of row y (re-read: an L1 hit, and cheaper than carrying the raw words through the ring)
            const uint8_t *p = src + o;
            const uint2 m = __ldg((const uint2 *)p);
            const uint32_t w[5] = { __ldg((const uint32_t *)(p - 4)), m.x, m.y, __ldg((const uint32_t *)(p + 8)), 0 };
            uint32_t A[12]; // A[j] = bytes x+j-2 .. x+j+1
#pragma unroll
            for (int j = 0; j < 12; j++) {
                const int q = (j + 2) >> 2, sh = ((j + 2) & 3) * 8;
                A[j] = sh ? __funnelshift_r(w[q], w[q + 1], sh) : w[q];
            }
            int h[8];
#pragma unroll
            for (int i = 0; i < 8; i++) h[i] = dp4a_us(A[i + 4], 0x000001fbu, dp4a_us(A[i], 0x1414fb01u, 16)) >> 5;
            if (live) *(uint2 *)(dh + o) = make_uint2(pack4_sat(h[0], h[1], h[2], h[3]), pack4_sat(h[4], h[5], h[6], h[7]));
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Integral image: sum8[y][x] = sum of the 8x8 pixel box with top-left (x,y), sum4 likewise for 4x4, both
// mod 2^16 like the reference's uint16 arithmetic.  One thread = 4 adjacent columns (64-bit stores) walking down R rows with a
// running vertical sum (add the entering row's horizontal sum, subtract the leaving one).
template <int R, bool SUB4>
__global__ void __launch_bounds__(128) integral_kernel(const uint8_t *__restrict__ src, uint16_t *__restrict__ sum8,
                                                       uint16_t *__restrict__ sum4, int stride, int x_begin, int n_quads,
                                                       int y_begin, int y_end)
{
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= n_quads) return;
    const int x = x_begin + qi * 4; // multiple of 4: the twelve bytes x .. x+11 a row contributes are three aligned words
    const int y0 = y_begin + blockIdx.y * R;
    // horizontal sums of one row for columns x .. x+3 from its three words, packed two columns per register: h8 = 8 pixels, h4 = 4 pixels
    auto fetch = [&](uint32_t (&w)[3], int y) {
        const uint32_t *p = (const uint32_t *)(src + (ptrdiff_t)min(y, y_end + 7) * stride + x); // y_end + 7: the last row any stored sum needs
        w[0] = __ldg(p); w[1] = __ldg(p + 1); w[2] = __ldg(p + 2);
    };
    auto hsum = [&](const uint32_t (&w)[3], uint32_t (&h8)[2], uint32_t (&h4)[2]) {
        uint32_t l[4], h[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t lo = k ? __funnelshift_r(w[0], w[1], 8 * k) : w[0], hi = k ? __funnelshift_r(w[1], w[2], 8 * k) : w[1];
            l[k] = sad4_acc(lo, 0, 0);
            h[k] = sad4_acc(hi, 0, l[k]);
        }
        h4[0] = l[0] | (l[1] << 16); h4[1] = l[2] | (l[3] << 16);
        h8[0] = h[0] | (h[1] << 16); h8[1] = h[2] | (h[3] << 16);
    };
    // running vertical sums, one register per column (the uint16 wrap of the reference is applied when storing)
    uint32_t s8[4] = { 0, 0, 0, 0 }, s4[4] = { 0, 0, 0, 0 };
    uint32_t ring8[8][2], ring4[8][2]; // horizontal sums of rows y..y+7 (slot (r+k)%8 <-> row y+k)
    uint32_t pre[8][3], pf[4][3];      // the first eight rows and the next four are requested before anything is computed
#pragma unroll
    for (int k = 0; k < 8; k++) fetch(pre[k], y0 + k);
#pragma unroll
    for (int k = 0; k < 4; k++) fetch(pf[k], y0 + 8 + k);
#pragma unroll
    for (int k = 0; k < 8; k++) {
        hsum(pre[k], ring8[k], ring4[k]);
#pragma unroll
        for (int c = 0; c < 4; c++) {
            s8[c] += (ring8[k][c >> 1] >> (16 * (c & 1))) & 0xffff;
            if (SUB4 && k < 4) s4[c] += (ring4[k][c >> 1] >> (16 * (c & 1))) & 0xffff;
        }
    }
    // No early exit in the row loop (rows past y_end are computed from clamped addresses and not stored) and a row's words are requested
    // four iterations before they are summed: with a `break` per row every row paid a full memory round trip.
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int y = y0 + r;
        if (y < y_end) {
            const ptrdiff_t o = (ptrdiff_t)y * stride + x;
            *(uint2 *)(sum8 + o) = make_uint2((s8[0] & 0xffff) | (s8[1] << 16), (s8[2] & 0xffff) | (s8[3] << 16));
            if (SUB4) *(uint2 *)(sum4 + o) = make_uint2((s4[0] & 0xffff) | (s4[1] << 16), (s4[2] & 0xffff) | (s4[3] << 16));
        }
        if (r + 1 < R) {
            uint32_t h8[2], h4[2];
            hsum(pf[r % 4], h8, h4); // row y + 8
            if (r + 5 < R) fetch(pf[r % 4], y + 12);
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int sh = 16 * (c & 1);
                s8[c] += ((h8[c >> 1] >> sh) & 0xffff) - ((ring8[r % 8][c >> 1] >> sh) & 0xffff);
                if (SUB4) s4[c] += ((ring4[(r + 4) % 8][c >> 1] >> sh) & 0xffff) - ((ring4[r % 8][c >> 1] >> sh) & 0xffff);
            }
            ring8[r % 8][0] = h8[0]; ring8[r % 8][1] = h8[1]; ring4[r % 8][0] = h4[0]; ring4[r % 8][1] = h4[1];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Half-resolution planes: per output pixel the nested rounding average of a 2x2 source neighbourhood
// (mc.c:343-349).  (a+b+1)>>1 per byte is exactly __vavgu4, so one thread produces 4 output pixels of all four
// planes from 3 rows x 9 source bytes.
template <int RL>
__global__ void __launch_bounds__(128) lowres_kernel(const uint8_t *__restrict__ src, int src_stride, uint8_t *__restrict__ l0,
                                                     uint8_t *__restrict__ lh, uint8_t *__restrict__ lv, uint8_t *__restrict__ lc,
                                                     int dst_stride, int n_cols, int lines_lowres)
{
    // one thread = 8 output pixels (16 source bytes + one word) x RL output rows; a source row serves two output rows
    const int ci = blockIdx.x * blockDim.x + threadIdx.x;
    if (ci >= n_cols) return;
    const int y0 = blockIdx.y * RL;
    const uint8_t *r0 = src + (ptrdiff_t)(2 * y0) * src_stride + 16 * ci;
    auto load = [&](uint32_t (&w)[5], int k) {
        const uint8_t *p = r0 + (ptrdiff_t)k * src_stride;
        const uint4 q = __ldg((const uint4 *)p);
        w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w; w[4] = __ldg((const uint32_t *)(p + 16));
    };
    uint32_t w[2 * RL + 1][5]; // every source row of the tile first: the loads are independent, so one memory round trip serves all
#pragma unroll
    for (int k = 0; k < 2 * RL + 1; k++) load(w[k], k);
#pragma unroll
    for (int r = 0; r < RL; r++) {
        const int y = y0 + r;
        if (y >= lines_lowres) break;
        uint32_t a01[5], a12[5];
#pragma unroll
        for (int i = 0; i < 5; i++) { a01[i] = __vavgu4(w[2 * r][i], w[2 * r + 1][i]); a12[i] = __vavgu4(w[2 * r + 1][i], w[2 * r + 2][i]); }
        const ptrdiff_t o = (ptrdiff_t)y * dst_stride + 8 * ci;
        auto emit = [&](const uint32_t (&a)[5], uint8_t *d0, uint8_t *d1) {
            uint32_t e[2], f[2];
#pragma unroll
            for (int g = 0; g < 2; g++) {
                const uint32_t even = __byte_perm(a[2 * g], a[2 * g + 1], 0x6420), odd = __byte_perm(a[2 * g], a[2 * g + 1], 0x7531);
                const uint32_t even1 = __byte_perm(even, a[2 * g + 2], 0x4321);
                e[g] = __vavgu4(even, odd); f[g] = __vavgu4(odd, even1);
            }
            *(uint2 *)(d0 + o) = make_uint2(e[0], e[1]);
            *(uint2 *)(d1 + o) = make_uint2(f[0], f[1]);
        };
        emit(a01, l0, lh);
        emit(a12, lv, lc);
    }
}

int launch_border(x264_cuda_t *ctx, uint8_t *const planes[], int n_planes, int stride, int cx0, int cx1, int cy0, int cy1,
                  int X0, int X1, int Y0, int Y1)
{
    BorderArgs a;
    for (int i = 0; i < 4; i++) a.p[i] = planes[i < n_planes ? i : 0];
    a.stride = stride; a.cx0 = cx0; a.cx1 = cx1; a.cy0 = cy0; a.cy1 = cy1; a.X0 = X0; a.X1 = X1; a.Y0 = Y0; a.Y1 = Y1;
    a.words = (X1 - X0) >> 2;
    a.nl = (cx0 - X0 + 3) >> 2;                  // words holding a pixel left of the valid area
    a.wr0 = max(a.nl, (cx1 + 1 - X0) >> 2);      // first word holding a pixel right of it
    a.side = a.nl + (a.words - a.wr0);
    a.band_items = ((cy0 - Y0) + (Y1 - 1 - cy1)) * a.words;
    a.items = a.band_items + (cy1 - cy0 + 1) * a.side;
    if (a.items <= 0) return 0;
    dim3 grid((a.items + 255) / 256, n_planes);
    replicate_border_kernel<<<grid, 256, 0, ctx->stream>>>(a);
    LAUNCH_CHECK(ctx, "replicate_border_kernel");
    return 0;
}

} // namespace

extern "C" int x264_cuda_frame_expand_border(x264_cuda_t *ctx, x264_cuda_frame_t *f)
{
    x264_cuda_enter(ctx);
    const x264_cuda_geom_t &g = f->g;
    uint8_t *planes[1] = { f->plane[0] };
    // mod16 padding and the 32-px borders are both replications of the picture edge (frame.c:304-331, :218-267)
    if (launch_border(ctx, planes, 1, g.stride, 0, g.width - 1, 0, g.height - 1, -PADH, g.mb_width * 16 + PADH, -PADV, g.lines + PADV)) return -1;
    if (f->buf_chroma) { // planes 1 and 2: half size, half padding (frame.c:229-236)
        uint8_t *cp[2] = { f->chroma[0], f->chroma[1] };
        return launch_border(ctx, cp, 2, f->stride_c, 0, g.width / 2 - 1, 0, g.height / 2 - 1, -PADH / 2, g.mb_width * 8 + PADH / 2, -PADV / 2,
                             g.lines / 2 + PADV / 2);
    }
    return 0;
}

// x264_frame_expand_border_mod16 (frame.c:304-331): what the reference does to fenc — only the padding up to a multiple of 16
extern "C" int x264_cuda_frame_expand_border_mod16(x264_cuda_t *ctx, x264_cuda_frame_t *f)
{
    x264_cuda_enter(ctx);
    const x264_cuda_geom_t &g = f->g;
    uint8_t *planes[1] = { f->plane[0] };
    if (launch_border(ctx, planes, 1, g.stride, 0, g.width - 1, 0, g.height - 1, 0, g.mb_width * 16, 0, g.lines)) return -1;
    if (f->buf_chroma) {
        uint8_t *cp[2] = { f->chroma[0], f->chroma[1] };
        return launch_border(ctx, cp, 2, f->stride_c, 0, g.width / 2 - 1, 0, g.height / 2 - 1, 0, g.mb_width * 8, 0, g.lines / 2);
    }
    return 0;
}

extern "C" int x264_cuda_frame_filter(x264_cuda_t *ctx, x264_cuda_frame_t *f)
{
    x264_cuda_enter(ctx);
    const x264_cuda_geom_t &g = f->g;
    const int w16 = g.mb_width * 16;
    if (!(g.flags & X264_CUDA_FRAME_HPEL)) {
        snprintf(ctx->err, 256, "x264_cuda_frame_filter: frame was created without X264_CUDA_FRAME_HPEL");
        return -1;
    }
    {
        // hpel_filter covers columns [-8, w16+8) and rows [-8, lines+8) (mc.c:409-426); expand_border_filtered
        // then keeps only columns [-4, w16+4) of it (frame.c:278-281), which is all we compute.
        // (the kernel computes [-8, w16+8); the border pass below overwrites the outer four columns on each side with the replication)
        const int x_begin = -8, n_cols = (w16 + 16) / 8, y_begin = -8, y_end = g.lines + 8;
        if (g.lines >= 1440) { // rows per thread: 12 where the frame has enough threads to fill the machine, else 6
            constexpr int R = 12;
            dim3 grid((n_cols + 127) / 128, (y_end - y_begin + R - 1) / R);
            hpel_kernel<R><<<grid, 128, 0, ctx->stream>>>(f->plane[0], f->plane[1], f->plane[2], f->plane[3], g.stride, x_begin, n_cols, y_begin, y_end);
        } else {
            constexpr int R = 6;
            dim3 grid((n_cols + 127) / 128, (y_end - y_begin + R - 1) / R);
            hpel_kernel<R><<<grid, 128, 0, ctx->stream>>>(f->plane[0], f->plane[1], f->plane[2], f->plane[3], g.stride, x_begin, n_cols, y_begin, y_end);
        }
        LAUNCH_CHECK(ctx, "hpel_kernel");
        uint8_t *planes[3] = { f->plane[1], f->plane[2], f->plane[3] };
        if (launch_border(ctx, planes, 3, g.stride, -4, w16 + 3, -8, g.lines + 7, -PADH, w16 + PADH, -PADV, g.lines + PADV))
            return -1;
    }
    if (f->integral) {
        // valid region of the reference's integral planes: rows [-31, lines+23], columns [-32, w16+24)
        // (mc.c:428-461 with stride_ref = w16+64; the h pass stops 8 short of the row end)
        constexpr int R = 16;
        const int x_begin = -PADH, n_quads = (w16 + 24 + PADH) / 4, y_begin = -PADV + 1, y_end = g.lines + 24;
        dim3 grid((n_quads + 127) / 128, (y_end - y_begin + R - 1) / R);
        if (g.flags & X264_CUDA_FRAME_INTEGRAL4)
            integral_kernel<R, true><<<grid, 128, 0, ctx->stream>>>(f->plane[0], f->integral, f->integral + f->plane_size, g.stride,
                                                                     x_begin, n_quads, y_begin, y_end);
        else
            integral_kernel<R, false><<<grid, 128, 0, ctx->stream>>>(f->plane[0], f->integral, nullptr, g.stride, x_begin, n_quads,
                                                                      y_begin, y_end);
        LAUNCH_CHECK(ctx, "integral_kernel");
    }
    return 0;
}

extern "C" int x264_cuda_frame_init_lowres(x264_cuda_t *ctx, x264_cuda_frame_t *f)
{
    x264_cuda_enter(ctx);
    const x264_cuda_geom_t &g = f->g;
    if (!(g.flags & X264_CUDA_FRAME_LOWRES)) {
        snprintf(ctx->err, 256, "x264_cuda_frame_init_lowres: frame was created without X264_CUDA_FRAME_LOWRES");
        return -1;
    }
    // The reference first duplicates the last column/row of the source (mc.c:315-317); on a border-expanded
    // device plane those bytes already hold exactly that replication, so the core reads them as they are.
    constexpr int RL = 2;
    const int n_cols = g.width_lowres / 8; // width_lowres = 8 * mb_width
    dim3 grid((n_cols + 127) / 128, (g.lines_lowres + RL - 1) / RL);
    lowres_kernel<RL><<<grid, 128, 0, ctx->stream>>>(f->plane[0], g.stride, f->lowres[0], f->lowres[1], f->lowres[2], f->lowres[3],
                                                      g.stride_lowres, n_cols, g.lines_lowres);
    LAUNCH_CHECK(ctx, "lowres_kernel");
    // x264_frame_expand_border_lowres passes width = i_stride_lowres - 2*PADH of the REFERENCE layout
    // (frame.c:301), which exceeds width_lowres when mb_width is odd; those extra columns are never written by
    // the reference (zero-filled allocation) and the right border replicates the last of them.
    const int ref_w = ((g.width_lowres + 2 * PADH + 15) & ~15) - 2 * PADH;
    if (ref_w > g.width_lowres) {
        for (int i = 0; i < 4; i++)
            CUDA_TRY(ctx, cudaMemset2DAsync(f->lowres[i] + g.width_lowres, g.stride_lowres, 0, ref_w - g.width_lowres,
                                            g.lines_lowres, ctx->stream));
    }
    return launch_border(ctx, f->lowres, 4, g.stride_lowres, 0, ref_w - 1, 0, g.lines_lowres - 1, -PADH, ref_w + PADH, -PADV,
                         g.lines_lowres + PADV);
}
