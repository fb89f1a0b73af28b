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
// Half-pel planes.  One thread = 4 adjacent output columns (one 32-bit store per plane) walking down R rows.
// It keeps the last 6 source rows of its 9-byte neighbourhood (columns x-2..x+6) in registers, so each new
// row costs 3 aligned word loads; the vertical 6-tap (int, |v| <= 10710 fits the reference's int16 buf) is
// evaluated for those 9 columns and feeds both the V plane and the horizontal pass of the centre plane.
__device__ __forceinline__ int tap6(int a, int b, int c, int d, int e, int f) { return a + f - 5 * (b + e) + 20 * (c + d); }
__device__ __forceinline__ uint32_t pack4(int a, int b, int c, int d)
{
    return (uint32_t)a | ((uint32_t)b << 8) | ((uint32_t)c << 16) | ((uint32_t)d << 24);
}

template <int R>
__global__ void __launch_bounds__(128) hpel_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dh,
                                                   uint8_t *__restrict__ dv, uint8_t *__restrict__ dc, int stride,
                                                   int x_begin, int n_words, int y_begin, int y_end)
{
    const int wi = blockIdx.x * blockDim.x + threadIdx.x;
    if (wi >= n_words) return;
    const int x = x_begin + wi * 4; // multiple of 4 relative to pixel 0 => aligned words (PADH, stride % 4 == 0)
    const int y0 = y_begin + blockIdx.y * R;
    // ring of 6 rows x 12 bytes (x-4 .. x+7), kept as the 9 needed byte values x-2..x+6
    int px[6][9];
    auto load = [&](int (&dst)[9], int y) {
        const uint32_t *p = (const uint32_t *)(src + (ptrdiff_t)y * stride + x - 4);
        const uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
        dst[0] = (w0 >> 16) & 255; dst[1] = w0 >> 24;
        dst[2] = w1 & 255; dst[3] = (w1 >> 8) & 255; dst[4] = (w1 >> 16) & 255; dst[5] = w1 >> 24;
        dst[6] = w2 & 255; dst[7] = (w2 >> 8) & 255; dst[8] = (w2 >> 16) & 255;
    };
#pragma unroll
    for (int k = 0; k < 5; k++) load(px[k], y0 - 2 + k);
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int y = y0 + r;
        if (y >= y_end) break;
        // rows y-2..y+3 live in px[(r+k)%6], k=0..5 (static after unrolling)
        load(px[(r + 5) % 6], y + 3);
        int v[9];
#pragma unroll
        for (int i = 0; i < 9; i++)
            v[i] = tap6(px[r % 6][i], px[(r + 1) % 6][i], px[(r + 2) % 6][i], px[(r + 3) % 6][i], px[(r + 4) % 6][i],
                        px[(r + 5) % 6][i]);
        const int(&s)[9] = px[(r + 2) % 6]; // source row y
        int oh[4], ov[4], oc[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            // output column x+i sits at neighbourhood index i+2
            oh[i] = clip_u8((tap6(s[i], s[i + 1], s[i + 2], s[i + 3], s[i + 4], s[i + 5]) + 16) >> 5);
            ov[i] = clip_u8((v[i + 2] + 16) >> 5);
            oc[i] = clip_u8((tap6(v[i], v[i + 1], v[i + 2], v[i + 3], v[i + 4], v[i + 5]) + 512) >> 10);
        }
        const ptrdiff_t o = (ptrdiff_t)y * stride + x;
        *(uint32_t *)(dh + o) = pack4(oh[0], oh[1], oh[2], oh[3]);
        *(uint32_t *)(dv + o) = pack4(ov[0], ov[1], ov[2], ov[3]);
        *(uint32_t *)(dc + o) = pack4(oc[0], oc[1], oc[2], oc[3]);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Integral image: sum8[y][x] = sum of the 8x8 pixel box with top-left (x,y), sum4 likewise for 4x4, both
// mod 2^16 like the reference's uint16 arithmetic.  One thread = 2 adjacent columns walking down R rows with a
// running vertical sum (add the entering row's horizontal sum, subtract the leaving one).
template <int R, bool SUB4>
__global__ void __launch_bounds__(128) integral_kernel(const uint8_t *__restrict__ src, uint16_t *__restrict__ sum8,
                                                       uint16_t *__restrict__ sum4, int stride, int x_begin, int n_pairs,
                                                       int y_begin, int y_end)
{
    const int pi = blockIdx.x * blockDim.x + threadIdx.x;
    if (pi >= n_pairs) return;
    const int x = x_begin + pi * 2; // even; loads below use the enclosing aligned words
    const int y0 = y_begin + blockIdx.y * R;
    // horizontal sums of row y for columns x and x+1: h8 = 8 pixels, h4 = 4 pixels
    auto hsum = [&](int y, uint32_t &h8, uint32_t &h4) {
        const uint8_t *a = src + (ptrdiff_t)y * stride + x;
        const int sh = ((uintptr_t)a & 3) * 8; // 0 or 16
        const uint32_t *p = (const uint32_t *)((uintptr_t)a & ~(uintptr_t)3);
        const uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
        const uint32_t a0 = __funnelshift_r(w0, w1, sh), a1 = __funnelshift_r(w1, w2, sh);         // bytes x..x+7
        const uint32_t b0 = __funnelshift_r(w0, w1, sh + 8), b1 = __funnelshift_r(w1, w2, sh + 8); // bytes x+1..x+8
        const uint32_t l0 = sad4_acc(a0, 0, 0), l1 = sad4_acc(b0, 0, 0);
        h4 = l0 | (l1 << 16);
        h8 = sad4_acc(a1, 0, l0) | (sad4_acc(b1, 0, l1) << 16);
    };
    // running sums hold two 16-bit lanes (columns x, x+1); lanes never carry into each other because every
    // true value is < 2^16 (8x8 box <= 16320) and we only add/subtract whole lane-pairs of equal structure:
    // keep the lanes in separate registers to stay safe.
    uint32_t s8a = 0, s8b = 0, s4a = 0, s4b = 0;
    uint32_t ring8[8], ring4[8]; // horizontal 8- and 4-pixel sums of rows y..y+7 (slot (r+k)%8 <-> row y+k)
#pragma unroll
    for (int k = 0; k < 8; k++) {
        hsum(y0 + k, ring8[k], ring4[k]);
        s8a += ring8[k] & 0xffff; s8b += ring8[k] >> 16;
        if (SUB4 && k < 4) { s4a += ring4[k] & 0xffff; s4b += ring4[k] >> 16; }
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int y = y0 + r;
        if (y >= y_end) break;
        const ptrdiff_t o = (ptrdiff_t)y * stride + x;
        *(uint32_t *)(sum8 + o) = (s8a & 0xffff) | (s8b << 16);
        if (SUB4) *(uint32_t *)(sum4 + o) = (s4a & 0xffff) | (s4b << 16);
        if (r + 1 < R) {
            uint32_t h8, h4;
            hsum(y + 8, h8, h4);
            const uint32_t old8 = ring8[r % 8], old4 = ring4[r % 8], in4 = ring4[(r + 4) % 8];
            s8a += (h8 & 0xffff) - (old8 & 0xffff); s8b += (h8 >> 16) - (old8 >> 16);
            if (SUB4) { s4a += (in4 & 0xffff) - (old4 & 0xffff); s4b += (in4 >> 16) - (old4 >> 16); }
            ring8[r % 8] = h8; ring4[r % 8] = h4;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Half-resolution planes: per output pixel the nested rounding average of a 2x2 source neighbourhood
// (mc.c:343-349).  (a+b+1)>>1 per byte is exactly __vavgu4, so one thread produces 4 output pixels of all four
// planes from 3 rows x 9 source bytes.
__global__ void __launch_bounds__(128) lowres_kernel(const uint8_t *__restrict__ src, int src_stride, uint8_t *__restrict__ l0,
                                                     uint8_t *__restrict__ lh, uint8_t *__restrict__ lv, uint8_t *__restrict__ lc,
                                                     int dst_stride, int n_words, int lines_lowres)
{
    const int wi = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (wi >= n_words || y >= lines_lowres) return;
    const uint8_t *r0 = src + (ptrdiff_t)(2 * y) * src_stride + 8 * wi;
    uint32_t w[3][3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const uint32_t *p = (const uint32_t *)(r0 + (ptrdiff_t)k * src_stride);
        w[k][0] = __ldg(p); w[k][1] = __ldg(p + 1); w[k][2] = __ldg(p + 2);
    }
    uint32_t a01[3], a12[3];
#pragma unroll
    for (int i = 0; i < 3; i++) { a01[i] = __vavgu4(w[0][i], w[1][i]); a12[i] = __vavgu4(w[1][i], w[2][i]); }
    auto emit = [&](const uint32_t (&a)[3], uint8_t *d0, uint8_t *d1) {
        const uint32_t even = __byte_perm(a[0], a[1], 0x6420), odd = __byte_perm(a[0], a[1], 0x7531);
        const uint32_t even1 = __byte_perm(even, a[2], 0x4321);
        const ptrdiff_t o = (ptrdiff_t)y * dst_stride + 4 * wi;
        *(uint32_t *)(d0 + o) = __vavgu4(even, odd);
        *(uint32_t *)(d1 + o) = __vavgu4(odd, even1);
    };
    emit(a01, l0, lh);
    emit(a12, lv, lc);
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
        constexpr int R = 8;
        const int x_begin = -4, n_words = (w16 + 8) / 4, y_begin = -8, y_end = g.lines + 8;
        dim3 grid((n_words + 127) / 128, (y_end - y_begin + R - 1) / R);
        hpel_kernel<R><<<grid, 128, 0, ctx->stream>>>(f->plane[0], f->plane[1], f->plane[2], f->plane[3], g.stride, x_begin,
                                                       n_words, y_begin, y_end);
        LAUNCH_CHECK(ctx, "hpel_kernel");
        uint8_t *planes[3] = { f->plane[1], f->plane[2], f->plane[3] };
        if (launch_border(ctx, planes, 3, g.stride, -4, w16 + 3, -8, g.lines + 7, -PADH, w16 + PADH, -PADV, g.lines + PADV))
            return -1;
    }
    if (f->integral) {
        // valid region of the reference's integral planes: rows [-31, lines+23], columns [-32, w16+24)
        // (mc.c:428-461 with stride_ref = w16+64; the h pass stops 8 short of the row end)
        constexpr int R = 16;
        const int x_begin = -PADH, n_pairs = (w16 + 24 + PADH) / 2, y_begin = -PADV + 1, y_end = g.lines + 24;
        dim3 grid((n_pairs + 127) / 128, (y_end - y_begin + R - 1) / R);
        if (g.flags & X264_CUDA_FRAME_INTEGRAL4)
            integral_kernel<R, true><<<grid, 128, 0, ctx->stream>>>(f->plane[0], f->integral, f->integral + f->plane_size, g.stride,
                                                                     x_begin, n_pairs, y_begin, y_end);
        else
            integral_kernel<R, false><<<grid, 128, 0, ctx->stream>>>(f->plane[0], f->integral, nullptr, g.stride, x_begin, n_pairs,
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
    const int n_words = g.width_lowres / 4;
    dim3 grid((n_words + 127) / 128, g.lines_lowres);
    lowres_kernel<<<grid, 128, 0, ctx->stream>>>(f->plane[0], g.stride, f->lowres[0], f->lowres[1], f->lowres[2], f->lowres[3],
                                                  g.stride_lowres, n_words, g.lines_lowres);
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
