// tables.cu — per-call entries for the reference's function tables (include/x264_cuda_tables.h).
//
// Each entry stages its operands into pinned memory, copies them to the device, runs ONE kernel that applies the same
// device functions the frame-batched kernels use (dct_dev.cuh / pixel_dev.cuh), copies the result back and returns it.
// That is a PCIe round trip per call — fine for tools/checkasm-style verification and for correctness of any caller,
// not a performance path (the frame-batched x264_cuda_* entry points are).  No CPU arithmetic happens here.
#include <atomic>
#include <mutex>
#include <vector>
#include <cstdlib>
#include "dct_dev.cuh"
#include "pixel_dev.cuh"
#include "intra_dev.cuh"
#include "../../include/x264_cuda_tables.h"

namespace {

// One context (stream, staging buffers) PER HOST THREAD: the reference copies its tables into every frame thread's x264_t
// (S/encoder/encoder.c:780) and calls the entries concurrently from up to X264_THREAD_MAX = 128 pthreads (S/common/common.h:50), so
// the entries must not queue on a process-wide lock.  Contexts are created on a thread's first call and closed when the process exits.
std::mutex g_reg_mu;                 // guards the registry only (taken once per thread, and by the counters)
std::vector<x264_cuda_t *> g_all_ctx;
std::atomic<long long> g_launches{0};
thread_local x264_cuda_t *t_ctx = nullptr;
thread_local int t_gen = 0;          // registry generation t_ctx belongs to (x264_cuda_tables_shutdown starts a new one)
std::atomic<int> g_gen{0};
#define g_ctx t_ctx

// The table signatures cannot report failure (S/common/pixel.h:26-28).  A device error is recorded ONCE here (sticky), every later entry
// returns without touching the device (results zero), and the encoder-side hook turns x264_cuda_tables_error() != NULL into
// x264_encoder_encode() returning -1 — the reference's own error convention (S/x264.c:759-762) instead of abort().
std::atomic<bool> g_failed{false};
char g_fail_msg[512] = "";

void fail(const char *what)
{
    bool expected = false;
    if (g_failed.compare_exchange_strong(expected, true)) {
        snprintf(g_fail_msg, sizeof(g_fail_msg), "x264_cuda table entry %s: %s", what, g_ctx ? x264_cuda_error(g_ctx) : x264_cuda_error(nullptr));
        fprintf(stderr, "%s\n", g_fail_msg);
        if (getenv("X264_CUDA_TABLES_ABORT")) abort(); // opt-in: die at the point of failure (debugging)
    }
}

int ensure_ctx()
{
    if (g_failed.load(std::memory_order_relaxed)) return -1;
    if (g_ctx && t_gen != g_gen.load(std::memory_order_acquire)) g_ctx = nullptr; // closed by x264_cuda_tables_shutdown on another thread
    if (g_ctx) { x264_cuda_enter(g_ctx); return 0; } // this thread may never have selected the device (X264_CUDA_DEVICE != 0)
    const char *e = getenv("X264_CUDA_DEVICE");
    if (x264_cuda_open(&g_ctx, e ? atoi(e) : 0) != 0) { g_ctx = nullptr; fail("x264_cuda_open"); return -1; }
    std::lock_guard<std::mutex> lk(g_reg_mu);
    // no atexit hook: by the time exit handlers run the CUDA runtime may already be torn down (a close then faults); the process's
    // end releases everything, and x264_cuda_tables_shutdown() is there for callers that want an orderly release earlier
    g_all_ctx.push_back(g_ctx);
    t_gen = g_gen.load(std::memory_order_acquire);
    return 0;
}

// ---- generic staging: host bytes in -> device -> kernel -> host bytes out
struct Stage {
    uint8_t *h, *d;
    size_t in_bytes, total;
    bool ok;
    std::vector<uint8_t> dead; // after a failure: host scratch, so that the entries' packing / unpacking code still has memory to touch
    Stage(size_t in_b, size_t out_b) : in_bytes((in_b + 255) & ~(size_t)255), total(((in_b + 255) & ~(size_t)255) + out_b), ok(true)
    {
        if (ensure_ctx()) ok = false;
        else if (x264_cuda_stage(g_ctx, total, total)) { fail("staging"); ok = false; }
        if (ok) { h = (uint8_t *)g_ctx->h_stage; d = (uint8_t *)g_ctx->d_stage; }
        else { dead.assign(total, 0); h = d = dead.data(); }
    }
    cudaStream_t stream() const { return ok ? g_ctx->stream : nullptr; }
    void up() { if (ok && cudaMemcpyAsync(d, h, in_bytes, cudaMemcpyHostToDevice, g_ctx->stream) != cudaSuccess) { fail("H2D"); ok = false; } }
    void down()
    {
        if (!ok) { memset(h + in_bytes, 0, total - in_bytes); return; }
        g_launches++;
        g_ctx->launches++;
        if (cudaGetLastError() != cudaSuccess || cudaMemcpyAsync(h + in_bytes, d + in_bytes, total - in_bytes, cudaMemcpyDeviceToHost, g_ctx->stream) != cudaSuccess ||
            cudaStreamSynchronize(g_ctx->stream) != cudaSuccess) {
            fail("kernel/D2H");
            ok = false;
            memset(h + in_bytes, 0, total - in_bytes);
        }
    }
    uint8_t *hout() { return h + in_bytes; }
    uint8_t *dout() { return d + in_bytes; }
};

// =================================================================================================== pixel
static const int kW[7] = { 16, 16, 8, 8, 8, 4, 4 }, kH[7] = { 16, 8, 16, 8, 4, 8, 4 };

// n comparisons of one w x h block pair each; tiles are 16x16 bytes (stride 16)
__global__ void cmp_tiles_kernel(int metric, int bw, int bh, int n, const uint8_t *a, const uint8_t *b, int a_step, int *out)
{
    const int i = threadIdx.x;
    if (i >= n) return;
    const uint8_t *pa = a + (size_t)i * a_step, *pb = b + (size_t)i * 256;
    int sum = 0;
    if (metric == 0 || metric == 1) {
        for (int y = 0; y < bh; y++)
            for (int x = 0; x < bw; x += 4) {
                const uint32_t wa = *(const uint32_t *)(pa + y * 16 + x), wb = *(const uint32_t *)(pb + y * 16 + x);
                if (metric == 0) sum = (int)sad4_acc(wa, wb, (uint32_t)sum);
                else
                    for (int k = 0; k < 4; k++) { const int d = (int)((wa >> (8 * k)) & 255) - (int)((wb >> (8 * k)) & 255); sum += d * d; }
            }
    } else if (metric == 2) {
        const uint8_t *const pl[4] = { pb, pb, pb, pb };
        const QpelSrc src = qpel_src(pl, 16, 0, 0);
        for (int u = 0; u < unit_count(bw, bh); u++) {
            int ux, uy;
            unit_pos(bw, u, ux, uy);
            sum += unit_cost(true, bw, pa, 16, src, 16, ux, uy);
        }
    } else {
        for (int y0 = 0; y0 < bh; y0 += 8)
            for (int x0 = 0; x0 < bw; x0 += 8) {
                uint2 f[8], r[8];
                for (int y = 0; y < 8; y++) { f[y] = *(const uint2 *)(pa + (y0 + y) * 16 + x0); r[y] = *(const uint2 *)(pb + (y0 + y) * 16 + x0); }
                sum += sa8d_8x8_rows(f, r);
            }
        sum = (sum + 2) >> 2;
    }
    out[i] = sum;
}

void pack_tile(uint8_t *dst, const uint8_t *src, int stride, int w, int h)
{
    for (int y = 0; y < h; y++) memcpy(dst + 16 * y, src + (size_t)y * stride, w);
}

// one fenc block against n reference blocks
void cmp_n(int metric, int ip, const uint8_t *p1, int s1, const uint8_t *const *p2, int s2, int n, int *scores)
{
    Stage st(256 * (1 + n), n * sizeof(int));
    pack_tile(st.h, p1, s1, kW[ip], kH[ip]);
    for (int i = 0; i < n; i++) pack_tile(st.h + 256 * (1 + i), p2[i], s2, kW[ip], kH[ip]);
    st.up();
    if (st.ok) cmp_tiles_kernel<<<1, 32, 0, st.stream()>>>(metric, kW[ip], kH[ip], n, st.d, st.d + 256, 0, (int *)st.dout());
    st.down();
    memcpy(scores, st.hout(), n * sizeof(int));
}

template <int METRIC, int IP> int cmp1(uint8_t *p1, int s1, uint8_t *p2, int s2)
{
    int v;
    const uint8_t *r[1] = { p2 };
    cmp_n(METRIC, IP, p1, s1, r, s2, 1, &v);
    return v;
}
template <int METRIC, int IP> void cmp_x3(uint8_t *fenc, uint8_t *a, uint8_t *b, uint8_t *c, int stride, int scores[3])
{
    const uint8_t *r[3] = { a, b, c };
    cmp_n(METRIC, IP, fenc, 16, r, stride, 3, scores); // fenc at FENC_STRIDE (pixel.c:364-385)
}
template <int METRIC, int IP> void cmp_x4(uint8_t *fenc, uint8_t *a, uint8_t *b, uint8_t *c, uint8_t *d, int stride, int scores[4])
{
    const uint8_t *r[4] = { a, b, c, d };
    cmp_n(METRIC, IP, fenc, 16, r, stride, 4, scores);
}

// ADS (pixel.c:515-559): one warp, ballot compaction keeps the ascending-i order of the C loop
__global__ void ads_kernel(int terms, int dc0, int dc1, int dc2, int dc3, const uint16_t *s_lo, const uint16_t *s_hi, const uint16_t *cost,
                           int width, int thresh, int16_t *mvs, int *count)
{
    const int lane = threadIdx.x;
    int n = 0;
    for (int c0 = 0; c0 < width; c0 += 32) {
        const int i = c0 + lane;
        bool keep = false;
        if (i < width) {
            int ads = abs(dc0 - (int)s_lo[i]) + (int)cost[i];
            if (terms == 2) ads += abs(dc1 - (int)s_hi[i]);
            if (terms == 4) ads += abs(dc1 - (int)s_lo[i + 8]) + abs(dc2 - (int)s_hi[i]) + abs(dc3 - (int)s_hi[i + 8]);
            keep = ads < thresh;
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) mvs[n + __popc(m & ((1u << lane) - 1))] = (int16_t)i;
        n += __popc(m);
    }
    if (lane == 0) *count = n;
}
template <int TERMS> int ads_entry(int enc_dc[4], uint16_t *sums, int delta, uint16_t *cost_mvx, int16_t *mvs, int width, int thresh)
{
    const int w8 = width + 8;
    const size_t row = (size_t)w8 * 2, in = 3 * ((row + 15) & ~(size_t)15);
    Stage st(in, 16 + (size_t)width * 2);
    const size_t rs = (row + 15) & ~(size_t)15;
    memcpy(st.h, sums, row);                                   // sums[i], sums[i+8]
    if (TERMS >= 2) memcpy(st.h + rs, sums + delta, row);      // sums[i+delta], sums[i+delta+8]
    memcpy(st.h + 2 * rs, cost_mvx, (size_t)width * 2);
    st.up();
    if (st.ok) ads_kernel<<<1, 32, 0, st.stream()>>>(TERMS, enc_dc[0], enc_dc[1], TERMS == 4 ? enc_dc[2] : 0, TERMS == 4 ? enc_dc[3] : 0, (const uint16_t *)st.d,
                                            (const uint16_t *)(st.d + rs), (const uint16_t *)(st.d + 2 * rs), width, thresh,
                                            (int16_t *)(st.dout() + 16), (int *)st.dout());
    st.down();
    const int n = *(int *)st.hout();
    memcpy(mvs, st.hout() + 16, (size_t)n * 2);
    return n;
}

// =================================================================================================== dct / quant
enum { OP_SUB4 = 0, OP_ADD4, OP_SUB8, OP_ADD8, OP_DC_FWD, OP_DC_INV, OP_ADD_DC, OP_QUANT, OP_QUANT_DC, OP_DEQUANT, OP_DEQUANT_DC };

// one thread per block; `a`/`b` are the staged operands, `o` the staged result
__global__ void blockop_kernel(int op, int n, int p0, int p1, const uint8_t *a, const uint8_t *b, uint8_t *o)
{
    const int i = threadIdx.x;
    if (i >= n) return;
    switch (op) {
    case OP_SUB4: { // a: fenc 4x4 bytes x n, b: pred 4x4 bytes x n -> int16[16] x n
        int d[16], c[16];
        for (int k = 0; k < 16; k++) d[k] = (int)a[i * 16 + k] - (int)b[i * 16 + k];
        fwd4x4(d, c);
        for (int k = 0; k < 16; k++) ((int16_t *)o)[i * 16 + k] = (int16_t)c[k];
    } break;
    case OP_ADD4: { // a: int16[16] x n, b: dst 4x4 bytes x n -> bytes
        int c[16], r[16];
        for (int k = 0; k < 16; k++) c[k] = ((const int16_t *)a)[i * 16 + k];
        inv4x4(c, r);
        for (int k = 0; k < 16; k++) o[i * 16 + k] = (uint8_t)clip_u8((int)b[i * 16 + k] + r[k]);
    } break;
    case OP_SUB8: {
        int d[64], c[64];
        for (int k = 0; k < 64; k++) d[k] = (int)a[i * 64 + k] - (int)b[i * 64 + k];
        fwd8x8(d, c);
        for (int k = 0; k < 64; k++) ((int16_t *)o)[i * 64 + k] = (int16_t)c[k];
    } break;
    case OP_ADD8: { // the reference leaves its in-place intermediate in dct (dct.c:326-333); callers never read it back
        int c[64], r[64];
        for (int k = 0; k < 64; k++) c[k] = ((const int16_t *)a)[i * 64 + k];
        inv8x8(c, r);
        for (int k = 0; k < 64; k++) o[i * 64 + k] = (uint8_t)clip_u8((int)b[i * 64 + k] + r[k]);
    } break;
    case OP_DC_FWD: case OP_DC_INV: {
        int d[16];
        for (int k = 0; k < 16; k++) d[k] = ((const int16_t *)a)[i * 16 + k];
        hadamard_dc(d, op == OP_DC_FWD);
        for (int k = 0; k < 16; k++) ((int16_t *)o)[i * 16 + k] = (int16_t)d[k];
    } break;
    case OP_ADD_DC: { // a: int16 dc x n, b: 4x4 dst bytes x n (dct.c:351-382)
        const int v = s16((((const int16_t *)a)[i] + 32) >> 6);
        for (int k = 0; k < 16; k++) o[i * 16 + k] = (uint8_t)clip_u8((int)b[i * 16 + k] + v);
    } break;
    case OP_QUANT: { // p0 = coefficients per block; a: int16 coef, b: mf[p0] then bias[p0] (uint16) -> coef, then nz flag word
        int nz = 0;
        const uint16_t *mf = (const uint16_t *)b, *bias = mf + p0;
        for (int k = 0; k < p0; k++) { const int q = quant1(((const int16_t *)a)[k], mf[k], bias[k]); ((int16_t *)o)[k] = (int16_t)q; nz |= q; }
        ((int16_t *)o)[p0] = nz != 0;
    } break;
    case OP_QUANT_DC: { // p0 = count, mf = p1 & 0xffff... passed split: b holds {int mf, int bias}
        int nz = 0;
        const int mf = ((const int *)b)[0], bias = ((const int *)b)[1];
        for (int k = 0; k < p0; k++) { const int q = quant1(((const int16_t *)a)[k], mf, bias); ((int16_t *)o)[k] = (int16_t)q; nz |= q; }
        ((int16_t *)o)[p0] = nz != 0;
    } break;
    case OP_DEQUANT: { // p0 = count, p1 = qbits; b: int dmf[p0] of the right qp%6 row
        for (int k = 0; k < p0; k++) ((int16_t *)o)[k] = (int16_t)dequant1(((const int16_t *)a)[k], ((const int *)b)[k], p1);
    } break;
    case OP_DEQUANT_DC: { // p1 = qbits, b: dmf[0]
        const int dmf0 = ((const int *)b)[0];
        for (int k = 0; k < 16; k++) {
            const int c = ((const int16_t *)a)[k];
            ((int16_t *)o)[k] = (int16_t)(p1 >= 0 ? s16(c * (dmf0 << p1)) : s16((c * dmf0 + (1 << (-p1 - 1))) >> (-p1)));
        }
    } break;
    }
}

void run_op(int op, int n, int p0, int p1, const void *a, size_t a_bytes, const void *b, size_t b_bytes, void *out, size_t out_bytes)
{
    const size_t a_al = (a_bytes + 255) & ~(size_t)255;
    Stage st(a_al + b_bytes, out_bytes);
    memcpy(st.h, a, a_bytes);
    if (b_bytes) memcpy(st.h + a_al, b, b_bytes);
    st.up();
    if (st.ok) blockop_kernel<<<1, 32, 0, st.stream()>>>(op, n, p0, p1, st.d, st.d + a_al, st.dout());
    st.down();
    memcpy(out, st.hout(), out_bytes);
}

// block positions inside 8x8 / 16x16 areas in the order of dct.c:157-171 / :218-232
void gather4(uint8_t *dst, const uint8_t *src, int stride, int nblk)
{
    for (int b = 0; b < nblk; b++) {
        const int i8 = b >> 2, i4 = b & 3;
        const int x = (nblk == 16 ? (i8 & 1) * 8 : 0) + (i4 & 1) * 4, y = (nblk == 16 ? (i8 >> 1) * 8 : 0) + (i4 >> 1) * 4;
        for (int r = 0; r < 4; r++) memcpy(dst + b * 16 + r * 4, src + (size_t)(y + r) * stride + x, 4);
    }
}
void scatter4(uint8_t *dst, int stride, const uint8_t *src, int nblk)
{
    for (int b = 0; b < nblk; b++) {
        const int i8 = b >> 2, i4 = b & 3;
        const int x = (nblk == 16 ? (i8 & 1) * 8 : 0) + (i4 & 1) * 4, y = (nblk == 16 ? (i8 >> 1) * 8 : 0) + (i4 >> 1) * 4;
        for (int r = 0; r < 4; r++) memcpy(dst + (size_t)(y + r) * stride + x, src + b * 16 + r * 4, 4);
    }
}
void gather8(uint8_t *dst, const uint8_t *src, int stride, int nblk)
{
    for (int b = 0; b < nblk; b++)
        for (int r = 0; r < 8; r++) memcpy(dst + b * 64 + r * 8, src + (size_t)((b >> 1) * 8 + r) * stride + (b & 1) * 8, 8);
}
void scatter8(uint8_t *dst, int stride, const uint8_t *src, int nblk)
{
    for (int b = 0; b < nblk; b++)
        for (int r = 0; r < 8; r++) memcpy(dst + (size_t)((b >> 1) * 8 + r) * stride + (b & 1) * 8, src + b * 64 + r * 8, 8);
}

template <int NBLK> void sub_dct4(int16_t *dct, uint8_t *pix1, uint8_t *pix2)
{
    uint8_t a[16 * 16], b[16 * 16];
    gather4(a, pix1, 16, NBLK); gather4(b, pix2, 32, NBLK);
    run_op(OP_SUB4, NBLK, 0, 0, a, NBLK * 16, b, NBLK * 16, dct, NBLK * 32);
}
template <int NBLK> void add_idct4(uint8_t *dst, int16_t *dct)
{
    uint8_t b[16 * 16], o[16 * 16];
    gather4(b, dst, 32, NBLK);
    run_op(OP_ADD4, NBLK, 0, 0, dct, NBLK * 32, b, NBLK * 16, o, NBLK * 16);
    scatter4(dst, 32, o, NBLK);
}
template <int NBLK> void sub_dct8(int16_t *dct, uint8_t *pix1, uint8_t *pix2)
{
    uint8_t a[256], b[256];
    gather8(a, pix1, 16, NBLK); gather8(b, pix2, 32, NBLK);
    run_op(OP_SUB8, NBLK, 0, 0, a, NBLK * 64, b, NBLK * 64, dct, NBLK * 128);
}
template <int NBLK> void add_idct8(uint8_t *dst, int16_t *dct)
{
    uint8_t b[256], o[256];
    gather8(b, dst, 32, NBLK);
    run_op(OP_ADD8, NBLK, 0, 0, dct, NBLK * 128, b, NBLK * 64, o, NBLK * 64);
    scatter8(dst, 32, o, NBLK);
}
// add8x8_idct_dc: dct[2][2] raster over the four 4x4s; add16x16_idct_dc: dct[4][4] raster (dct.c:366-382)
template <int NBLK> void add_idct_dc(uint8_t *dst, int16_t *dc)
{
    uint8_t b[256], o[256];
    const int per_row = NBLK == 4 ? 2 : 4;
    for (int k = 0; k < NBLK; k++)
        for (int r = 0; r < 4; r++) memcpy(b + k * 16 + r * 4, dst + (size_t)((k / per_row) * 4 + r) * 32 + (k % per_row) * 4, 4);
    run_op(OP_ADD_DC, NBLK, 0, 0, dc, NBLK * 2, b, NBLK * 16, o, NBLK * 16);
    for (int k = 0; k < NBLK; k++)
        for (int r = 0; r < 4; r++) memcpy(dst + (size_t)((k / per_row) * 4 + r) * 32 + (k % per_row) * 4, o + k * 16 + r * 4, 4);
}
void t_sub4x4_dct(int16_t dct[4][4], uint8_t *p1, uint8_t *p2) { sub_dct4<1>(&dct[0][0], p1, p2); }
void t_add4x4_idct(uint8_t *d, int16_t dct[4][4]) { add_idct4<1>(d, &dct[0][0]); }
void t_sub8x8_dct(int16_t dct[4][4][4], uint8_t *p1, uint8_t *p2) { sub_dct4<4>(&dct[0][0][0], p1, p2); }
void t_add8x8_idct(uint8_t *d, int16_t dct[4][4][4]) { add_idct4<4>(d, &dct[0][0][0]); }
void t_add8x8_idct_dc(uint8_t *d, int16_t dct[2][2]) { add_idct_dc<4>(d, &dct[0][0]); }
void t_sub16x16_dct(int16_t dct[16][4][4], uint8_t *p1, uint8_t *p2) { sub_dct4<16>(&dct[0][0][0], p1, p2); }
void t_add16x16_idct(uint8_t *d, int16_t dct[16][4][4]) { add_idct4<16>(d, &dct[0][0][0]); }
void t_add16x16_idct_dc(uint8_t *d, int16_t dct[4][4]) { add_idct_dc<16>(d, &dct[0][0]); }
void t_sub8x8_dct8(int16_t dct[8][8], uint8_t *p1, uint8_t *p2) { sub_dct8<1>(&dct[0][0], p1, p2); }
void t_add8x8_idct8(uint8_t *d, int16_t dct[8][8]) { add_idct8<1>(d, &dct[0][0]); }
void t_sub16x16_dct8(int16_t dct[4][8][8], uint8_t *p1, uint8_t *p2) { sub_dct8<4>(&dct[0][0][0], p1, p2); }
void t_add16x16_idct8(uint8_t *d, int16_t dct[4][8][8]) { add_idct8<4>(d, &dct[0][0][0]); }
void t_dct4x4dc(int16_t d[4][4]) { run_op(OP_DC_FWD, 1, 0, 0, d, 32, nullptr, 0, d, 32); }
void t_idct4x4dc(int16_t d[4][4]) { run_op(OP_DC_INV, 1, 0, 0, d, 32, nullptr, 0, d, 32); }

int quant_n(int16_t *dct, const uint16_t *mf, const uint16_t *bias, int n)
{
    uint16_t tb[128];
    int16_t out[65];
    memcpy(tb, mf, n * 2); memcpy(tb + n, bias, n * 2);
    run_op(OP_QUANT, 1, n, 0, dct, n * 2, tb, n * 4, out, (n + 1) * 2);
    memcpy(dct, out, n * 2);
    return out[n];
}
int quant_dc_n(int16_t *dct, int mf, int bias, int n)
{
    int mb[2] = { mf, bias };
    int16_t out[17];
    run_op(OP_QUANT_DC, 1, n, 0, dct, n * 2, mb, 8, out, (n + 1) * 2);
    memcpy(dct, out, n * 2);
    return out[n];
}
int t_quant_8x8(int16_t dct[8][8], uint16_t mf[64], uint16_t bias[64]) { return quant_n(&dct[0][0], mf, bias, 64); }
int t_quant_4x4(int16_t dct[4][4], uint16_t mf[16], uint16_t bias[16]) { return quant_n(&dct[0][0], mf, bias, 16); }
int t_quant_4x4_dc(int16_t dct[4][4], int mf, int bias) { return quant_dc_n(&dct[0][0], mf, bias, 16); }
int t_quant_2x2_dc(int16_t dct[2][2], int mf, int bias) { return quant_dc_n(&dct[0][0], mf, bias, 4); }
void t_dequant_8x8(int16_t dct[8][8], int dq[6][8][8], int qp) { run_op(OP_DEQUANT, 1, 64, qp / 6 - 6, dct, 128, dq[qp % 6], 256, dct, 128); }
void t_dequant_4x4(int16_t dct[4][4], int dq[6][4][4], int qp) { run_op(OP_DEQUANT, 1, 16, qp / 6 - 4, dct, 32, dq[qp % 6], 64, dct, 32); }
void t_dequant_4x4_dc(int16_t dct[4][4], int dq[6][4][4], int qp) { run_op(OP_DEQUANT_DC, 1, 16, qp / 6 - 6, dct, 32, dq[qp % 6], 4, dct, 32); }

// =================================================================================================== mc
// mc_luma / get_ref (mc.c:157-202): w x h samples at a qpel position from the four half-pel planes.  The planes are host
// memory: stage the (w+1) x (h+1) neighbourhood of the two planes the position needs, average on the device.
__global__ void qpel_kernel(const uint8_t *s1, const uint8_t *s2, int stride, int w, int h, uint8_t *dst)
{
    for (int i = threadIdx.x; i < w * h; i += blockDim.x) {
        const int x = i % w, y = i / w;
        const int a = s1[y * stride + x];
        dst[i] = s2 ? (uint8_t)((a + s2[y * stride + x] + 1) >> 1) : (uint8_t)a;
    }
}
void qpel_fetch(uint8_t *dst, int dst_stride, uint8_t **src, int i_src, int mvx, int mvy, int w, int h)
{
    static const uint8_t ref0[16] = { 0, 1, 1, 1, 0, 1, 1, 1, 2, 3, 3, 3, 0, 1, 1, 1 }, ref1[16] = { 0, 0, 0, 0, 2, 2, 3, 2, 2, 2, 3, 2, 2, 2, 3, 2 };
    const int qidx = ((mvy & 3) << 2) + (mvx & 3);
    const ptrdiff_t off = (ptrdiff_t)(mvy >> 2) * i_src + (mvx >> 2);
    const uint8_t *p1 = src[ref0[qidx]] + off + ((mvy & 3) == 3) * i_src;
    const uint8_t *p2 = (qidx & 5) ? src[ref1[qidx]] + off + ((mvx & 3) == 3) : nullptr;
    const int ts = (w + 15) & ~15;
    Stage st(2 * (size_t)ts * h, (size_t)w * h);
    for (int y = 0; y < h; y++) {
        memcpy(st.h + (size_t)y * ts, p1 + (ptrdiff_t)y * i_src, w);
        if (p2) memcpy(st.h + (size_t)(h + y) * ts, p2 + (ptrdiff_t)y * i_src, w);
    }
    st.up();
    if (st.ok) qpel_kernel<<<1, 128, 0, st.stream()>>>(st.d, p2 ? st.d + (size_t)h * ts : nullptr, ts, w, h, st.dout());
    st.down();
    for (int y = 0; y < h; y++) memcpy(dst + (ptrdiff_t)y * dst_stride, st.hout() + (size_t)y * w, w);
}
void t_mc_luma(uint8_t *dst, int i_dst, uint8_t **src, int i_src, int mvx, int mvy, int w, int h) { qpel_fetch(dst, i_dst, src, i_src, mvx, mvy, w, h); }
uint8_t *t_get_ref(uint8_t *dst, int *i_dst, uint8_t **src, int i_src, int mvx, int mvy, int w, int h)
{
    // the C version may return a pointer into the source plane when no averaging is needed (mc.c:197-200); returning the
    // samples in dst with the caller's stride is equally valid for every caller (they only read w x h through the result)
    qpel_fetch(dst, *i_dst, src, i_src, mvx, mvy, w, h);
    return dst;
}

// hpel_filter (mc.c:133-155) on host rows: dsth/dstc get columns [0,width), dstv columns [-2,width+3) like the C loop
__global__ void hpel_rows_kernel(const uint8_t *src, int sstride, int width, int height, uint8_t *dh, uint8_t *dv, uint8_t *dc, int ostride)
{
    // src points at (row 0, col 0) of a staged tile with 2 rows above, 3 below, 5 columns left, 6 right
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < height * (width + 5); i += gridDim.x * blockDim.x) {
        const int y = i / (width + 5), x = i % (width + 5) - 2; // x in [-2, width+3)
        const uint8_t *s = src + (ptrdiff_t)y * sstride;
        auto vtap = [&](int xx) {
            return (int)s[xx - 2 * sstride] + s[xx + 3 * sstride] - 5 * ((int)s[xx - sstride] + s[xx + 2 * sstride]) + 20 * ((int)s[xx] + s[xx + sstride]);
        };
        dv[(ptrdiff_t)y * ostride + x + 2] = (uint8_t)clip_u8((vtap(x) + 16) >> 5);
        if (x >= 0 && x < width) {
            const int c = vtap(x - 2) + vtap(x + 3) - 5 * (vtap(x - 1) + vtap(x + 2)) + 20 * (vtap(x) + vtap(x + 1));
            dc[(ptrdiff_t)y * ostride + x + 2] = (uint8_t)clip_u8((c + 512) >> 10);
            const int hh = (int)s[x - 2] + s[x + 3] - 5 * ((int)s[x - 1] + s[x + 2]) + 20 * ((int)s[x] + s[x + 1]);
            dh[(ptrdiff_t)y * ostride + x + 2] = (uint8_t)clip_u8((hh + 16) >> 5);
        }
    }
}
void t_hpel_filter(uint8_t *dsth, uint8_t *dstv, uint8_t *dstc, uint8_t *src, int stride, int width, int height, int16_t *buf)
{
    (void)buf; // the C body's int16 scratch row is not needed on the device
    const int ts = (width + 11 + 15) & ~15, os = (width + 5 + 15) & ~15;
    const size_t in = (size_t)ts * (height + 5), plane = (size_t)os * height;
    Stage st(in, 3 * plane);
    for (int y = -2; y < height + 3; y++) memcpy(st.h + (size_t)(y + 2) * ts, src + (ptrdiff_t)y * stride - 5, width + 11);
    st.up();
    uint8_t *o = st.dout();
    const int threads = 256, blocks = (height * (width + 5) + threads - 1) / threads;
    if (st.ok) hpel_rows_kernel<<<blocks < 2048 ? blocks : 2048, threads, 0, st.stream()>>>(st.d + 2 * ts + 5, ts, width, height, o, o + plane, o + 2 * plane, os);
    st.down();
    const uint8_t *h = st.hout();
    for (int y = 0; y < height; y++) {
        memcpy(dsth + (ptrdiff_t)y * stride, h + (size_t)y * os + 2, width);
        memcpy(dstv + (ptrdiff_t)y * stride - 2, h + plane + (size_t)y * os, width + 5);
        memcpy(dstc + (ptrdiff_t)y * stride, h + 2 * plane + (size_t)y * os + 2, width);
    }
}

// frame_init_lowres_core (mc.c:333-357)
__global__ void lowres_rows_kernel(const uint8_t *src, int sstride, int width, int height, uint8_t *o, size_t plane, int ostride)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < width * height; i += gridDim.x * blockDim.x) {
        const int x = i % width, y = i / width;
        const uint8_t *r0 = src + (size_t)(2 * y) * sstride, *r1 = r0 + sstride, *r2 = r1 + sstride;
#define AVG2(a, b) (((a) + (b) + 1) >> 1)
#define F4(a, b, c, d) AVG2(AVG2(a, b), AVG2(c, d))
        o[(size_t)y * ostride + x] = (uint8_t)F4(r0[2 * x], r1[2 * x], r0[2 * x + 1], r1[2 * x + 1]);
        o[plane + (size_t)y * ostride + x] = (uint8_t)F4(r0[2 * x + 1], r1[2 * x + 1], r0[2 * x + 2], r1[2 * x + 2]);
        o[2 * plane + (size_t)y * ostride + x] = (uint8_t)F4(r1[2 * x], r2[2 * x], r1[2 * x + 1], r2[2 * x + 1]);
        o[3 * plane + (size_t)y * ostride + x] = (uint8_t)F4(r1[2 * x + 1], r2[2 * x + 1], r1[2 * x + 2], r2[2 * x + 2]);
#undef F4
#undef AVG2
    }
}
void t_frame_init_lowres_core(uint8_t *src0, uint8_t *dst0, uint8_t *dsth, uint8_t *dstv, uint8_t *dstc, int src_stride, int dst_stride,
                              int width, int height)
{
    const int ts = (2 * width + 2 + 15) & ~15, os = (width + 15) & ~15;
    const size_t in = (size_t)ts * (2 * height + 1), plane = (size_t)os * height;
    Stage st(in, 4 * plane);
    for (int y = 0; y < 2 * height + 1; y++) memcpy(st.h + (size_t)y * ts, src0 + (ptrdiff_t)y * src_stride, 2 * width + 1);
    st.up();
    const int threads = 256, blocks = (width * height + threads - 1) / threads;
    if (st.ok) lowres_rows_kernel<<<blocks < 4096 ? blocks : 4096, threads, 0, st.stream()>>>(st.d, ts, width, height, st.dout(), plane, os);
    st.down();
    uint8_t *dsts[4] = { dst0, dsth, dstv, dstc };
    for (int p = 0; p < 4; p++)
        for (int y = 0; y < height; y++) memcpy(dsts[p] + (ptrdiff_t)y * dst_stride, st.hout() + p * plane + (size_t)y * os, width);
}

template <int METRIC> void fill_cmp(x264_cuda_pixel_cmp_t (&t)[7])
{
    t[0] = cmp1<METRIC, 0>; t[1] = cmp1<METRIC, 1>; t[2] = cmp1<METRIC, 2>; t[3] = cmp1<METRIC, 3>;
    t[4] = cmp1<METRIC, 4>; t[5] = cmp1<METRIC, 5>; t[6] = cmp1<METRIC, 6>;
}
template <int METRIC> void fill_x3(x264_cuda_pixel_cmp_x3_t (&t)[7])
{
    t[0] = cmp_x3<METRIC, 0>; t[1] = cmp_x3<METRIC, 1>; t[2] = cmp_x3<METRIC, 2>; t[3] = cmp_x3<METRIC, 3>;
    t[4] = cmp_x3<METRIC, 4>; t[5] = cmp_x3<METRIC, 5>; t[6] = cmp_x3<METRIC, 6>;
}
template <int METRIC> void fill_x4(x264_cuda_pixel_cmp_x4_t (&t)[7])
{
    t[0] = cmp_x4<METRIC, 0>; t[1] = cmp_x4<METRIC, 1>; t[2] = cmp_x4<METRIC, 2>; t[3] = cmp_x4<METRIC, 3>;
    t[4] = cmp_x4<METRIC, 4>; t[5] = cmp_x4<METRIC, 5>; t[6] = cmp_x4<METRIC, 6>;
}

// intra_{satd,sad}_x3_16x16 / intra_satd_x3_8x8c (S/common/pixel.h:97-100): the costs of the first three modes of the respective enum
// (16x16: V H DC; chroma: DC H V) of one block against its neighbours.  One warp: lane = mode (0..2) x 8x4 unit.
template <int N>
__global__ void intra_x3_kernel(bool satd, const uint8_t *edges, const uint8_t *fenc, int *out)
{
    constexpr int UNITS = N == 16 ? 8 : 2;
    const int lane = threadIdx.x, mi = lane / UNITS, u = lane % UNITS;
    const int x0 = N == 16 ? (u & 1) * 8 : 0, y0 = N == 16 ? (u >> 1) * 4 : u * 4;
    int cost = 0;
    if (mi < 3) {
        const int kind = N == 16 ? mi : (mi == 0 ? 2 : mi == 1 ? 1 : 0);
        int dcq[4] = { 0, 0, 0, 0 };
        if (kind == 2) {
            if (N == 16) {
                int s = 16;
                for (int i = 0; i < 32; i++) s += edges[4 + i];
                dcq[0] = s >> 5;
            } else {
                int s0 = 0, s1 = 0, s2 = 0, s3 = 0;
                for (int i = 0; i < 4; i++) { s0 += edges[4 + i]; s1 += edges[8 + i]; s2 += edges[12 + i]; s3 += edges[16 + i]; }
                dcq[0] = (s0 + s2 + 4) >> 3; dcq[1] = (s1 + 2) >> 2; dcq[2] = (s3 + 2) >> 2; dcq[3] = (s1 + s3 + 4) >> 3;
            }
        }
        uint2 f[4], r[4];
        for (int k = 0; k < 4; k++) f[k] = *(const uint2 *)(fenc + (y0 + k) * 16 + x0);
        predict_unit<N>(edges, kind, dcq, x0, y0, r);
        cost = unit_metric(satd, f, r);
    }
    for (int o = 1; o < UNITS; o <<= 1) cost += __shfl_xor_sync(0xffffffffu, cost, o);
    if (mi < 3 && u == 0) out[mi] = cost;
}

// ---- whole-block statistics on a staged 16x16 tile: var (pixel.c:141-160), hadamard_ac (:306-355), ssim_4x4x2_core (:431-458)
__global__ void block_stat_kernel(int op, int w, int h, const uint8_t *a, const uint8_t *b, unsigned long long *out)
{
    if (op == 0) { // PIXEL_VAR_C: w x w, shift = log2(w*w)
        uint32_t sum = 0, sqr = 0;
        for (int y = 0; y < w; y++)
            for (int x = 0; x < w; x += 4) { const uint32_t v = *(const uint32_t *)(a + y * 16 + x); sum = __dp4a(v, 0x01010101u, sum); sqr = __dp4a(v, v, sqr); }
        out[0] = sqr - (sum * sum >> (w == 16 ? 8 : 6));
    } else if (op == 1) { // HADAMARD_AC(w, h)
        unsigned long long sum = 0;
        for (int y = 0; y < h; y += 8)
            for (int x = 0; x < w; x += 8) sum += hadamard_ac_8x8(a + y * 16 + x, 16);
        out[0] = ((sum >> 34) << 32) + ((uint32_t)sum >> 1);
    } else { // two horizontally adjacent 4x4 blocks: s1, s2, ss, s12 each
        int *o = (int *)out;
        for (int z = 0; z < 2; z++) {
            uint32_t s1 = 0, s2 = 0, ss = 0, s12 = 0;
            for (int y = 0; y < 4; y++) {
                const uint32_t va = *(const uint32_t *)(a + y * 16 + 4 * z), vb = *(const uint32_t *)(b + y * 16 + 4 * z);
                s1 = __dp4a(va, 0x01010101u, s1); s2 = __dp4a(vb, 0x01010101u, s2);
                ss = __dp4a(va, va, ss); ss = __dp4a(vb, vb, ss); s12 = __dp4a(va, vb, s12);
            }
            o[4 * z] = (int)s1; o[4 * z + 1] = (int)s2; o[4 * z + 2] = (int)ss; o[4 * z + 3] = (int)s12;
        }
    }
}
unsigned long long block_stat(int op, int w, int h, const uint8_t *p1, int s1, const uint8_t *p2, int s2, int *sums)
{
    Stage st(512, 32);
    pack_tile(st.h, p1, s1, w, h);
    if (p2) pack_tile(st.h + 256, p2, s2, w, h);
    st.up();
    if (st.ok) block_stat_kernel<<<1, 1, 0, st.stream()>>>(op, w, h, st.d, st.d + 256, (unsigned long long *)st.dout());
    st.down();
    if (sums) memcpy(sums, st.hout(), 32);
    unsigned long long v;
    memcpy(&v, st.hout(), 8);
    return v;
}
template <int W> int t_var(uint8_t *pix, int stride) { return (int)(uint32_t)block_stat(0, W, W, pix, stride, nullptr, 0, nullptr); }
template <int W, int H> uint64_t t_hadamard_ac(uint8_t *pix, int stride) { return block_stat(1, W, H, pix, stride, nullptr, 0, nullptr); }
void t_ssim_4x4x2_core(const uint8_t *pix1, int stride1, const uint8_t *pix2, int stride2, int sums[2][4])
{
    block_stat(2, 8, 4, pix1, stride1, pix2, stride2, &sums[0][0]);
}

// ---- mc_chroma (mc.c:205-236) and the ten avg entries (mc.c:52-125) on staged tiles (stride 32 in, 16 out)
__global__ void chroma_avg_kernel(int op, int w, int h, int p0, int p1, const uint8_t *a, const uint8_t *b, uint8_t *out)
{
    const int i = threadIdx.x;
    if (i >= w * h) return;
    const int y = i / w, x = i - y * w;
    if (op == 0) { // p0, p1 = the 1/8-pel fractions
        const int cA = (8 - p0) * (8 - p1), cB = p0 * (8 - p1), cC = (8 - p0) * p1, cD = p0 * p1;
        const uint8_t *s = a + y * 32 + x;
        out[y * 16 + x] = (uint8_t)((cA * s[0] + cB * s[1] + cC * s[32] + cD * s[33] + 32) >> 6);
    } else { // p0 = weight of a
        const int va = a[y * 16 + x], vb = b[y * 16 + x];
        out[y * 16 + x] = p0 == 32 ? (uint8_t)((va + vb + 1) >> 1) : (uint8_t)clip_u8((va * p0 + vb * (64 - p0) + 32) >> 6);
    }
}
void t_mc_chroma(uint8_t *dst, int i_dst, uint8_t *src, int i_src, int mvx, int mvy, int w, int h)
{
    Stage st(32 * 17, 256);
    const uint8_t *s = src + (ptrdiff_t)(mvy >> 3) * i_src + (mvx >> 3);
    for (int y = 0; y <= h; y++) memcpy(st.h + 32 * y, s + (ptrdiff_t)y * i_src, w + 1);
    st.up();
    if (st.ok) chroma_avg_kernel<<<1, 256, 0, st.stream()>>>(0, w, h, mvx & 7, mvy & 7, st.d, nullptr, st.dout());
    st.down();
    for (int y = 0; y < h; y++) memcpy(dst + (ptrdiff_t)y * i_dst, st.hout() + 16 * y, w);
}
template <int W, int H> void t_avg(uint8_t *dst, int i_dst, uint8_t *src1, int i_src1, uint8_t *src2, int i_src2, int weight)
{
    Stage st(512, 256);
    pack_tile(st.h, src1, i_src1, W, H);
    pack_tile(st.h + 256, src2, i_src2, W, H);
    st.up();
    if (st.ok) chroma_avg_kernel<<<1, 256, 0, st.stream()>>>(1, W, H, weight, 0, st.d, st.d + 256, st.dout());
    st.down();
    for (int y = 0; y < H; y++) memcpy(dst + (ptrdiff_t)y * i_dst, st.hout() + 16 * y, W);
}

// fenc at FENC_STRIDE, fdec at FDEC_STRIDE with its neighbours in place (row -1, column -1, corner)
template <int N, bool SATD> void intra_x3(uint8_t *fenc, uint8_t *fdec, int res[3])
{
    Stage st(256 + 64, 3 * sizeof(int));
    pack_tile(st.h, fenc, 16, N, N);
    uint8_t *e = st.h + 256; // IntraEdges layout: [3] corner, [4..4+N) top, [4+N..4+2N) left
    e[3] = fdec[-32 - 1];
    for (int i = 0; i < N; i++) { e[4 + i] = fdec[-32 + i]; e[4 + N + i] = fdec[i * 32 - 1]; }
    st.up();
    if (st.ok) intra_x3_kernel<N><<<1, 32, 0, st.stream()>>>(SATD, st.d + 256, st.d, (int *)st.dout());
    st.down();
    memcpy(res, st.hout(), 3 * sizeof(int));
}

} // namespace

extern "C" int x264_pixel_init_cuda(x264_cuda_pixel_function_t *pixf)
{
    if (ensure_ctx()) return -1;
    fill_cmp<0>(pixf->sad); fill_cmp<0>(pixf->sad_aligned); fill_cmp<1>(pixf->ssd); fill_cmp<2>(pixf->satd);
    pixf->sa8d[X264_CUDA_PIXEL_16x16] = cmp1<3, 0>; pixf->sa8d[X264_CUDA_PIXEL_8x8] = cmp1<3, 3>; // the only two the C table has (pixel.c:607-608)
    fill_x3<0>(pixf->sad_x3); fill_x4<0>(pixf->sad_x4); fill_x3<2>(pixf->satd_x3); fill_x4<2>(pixf->satd_x4);
    // ads4 / ads2 / ads1 by partition, as x264_pixel_init wires them (pixel.c:591-594, :793-796)
    pixf->ads[X264_CUDA_PIXEL_16x16] = ads_entry<4>;
    pixf->ads[X264_CUDA_PIXEL_16x8] = pixf->ads[X264_CUDA_PIXEL_8x16] = pixf->ads[X264_CUDA_PIXEL_8x4] = pixf->ads[X264_CUDA_PIXEL_4x8] = ads_entry<2>;
    pixf->ads[X264_CUDA_PIXEL_8x8] = pixf->ads[X264_CUDA_PIXEL_4x4] = ads_entry<1>;
    // the merged intra cost entries the C table leaves NULL (pixel.c:664-667 sets them for mmxext): x264_mb_analyse_intra(_chroma) uses
    // them when present (analyse.c:560, :627); mbcmp_init (encoder.c:608-618) re-aliases intra_mbcmp_x3_16x16 to the sad or satd one
    pixf->intra_satd_x3_16x16 = pixf->intra_mbcmp_x3_16x16 = intra_x3<16, true>;
    pixf->intra_sad_x3_16x16 = intra_x3<16, false>;
    pixf->intra_satd_x3_8x8c = intra_x3<8, true>;
    // AQ / psy / SSIM primitives (pixel.c:597-613 wiring)
    pixf->var[X264_CUDA_PIXEL_16x16] = t_var<16>; pixf->var[X264_CUDA_PIXEL_8x8] = t_var<8>;
    pixf->hadamard_ac[X264_CUDA_PIXEL_16x16] = t_hadamard_ac<16, 16>; pixf->hadamard_ac[X264_CUDA_PIXEL_16x8] = t_hadamard_ac<16, 8>;
    pixf->hadamard_ac[X264_CUDA_PIXEL_8x16] = t_hadamard_ac<8, 16>; pixf->hadamard_ac[X264_CUDA_PIXEL_8x8] = t_hadamard_ac<8, 8>;
    pixf->ssim_4x4x2_core = t_ssim_4x4x2_core;
    return 0;
}

extern "C" int x264_dct_init_cuda(x264_cuda_dct_function_t *d)
{
    if (ensure_ctx()) return -1;
    d->sub4x4_dct = t_sub4x4_dct; d->add4x4_idct = t_add4x4_idct; d->sub8x8_dct = t_sub8x8_dct; d->add8x8_idct = t_add8x8_idct;
    d->add8x8_idct_dc = t_add8x8_idct_dc; d->sub16x16_dct = t_sub16x16_dct; d->add16x16_idct = t_add16x16_idct;
    d->add16x16_idct_dc = t_add16x16_idct_dc; d->sub8x8_dct8 = t_sub8x8_dct8; d->add8x8_idct8 = t_add8x8_idct8;
    d->sub16x16_dct8 = t_sub16x16_dct8; d->add16x16_idct8 = t_add16x16_idct8; d->dct4x4dc = t_dct4x4dc; d->idct4x4dc = t_idct4x4dc;
    return 0;
}

extern "C" int x264_quant_init_cuda(x264_cuda_quant_function_t *q)
{
    if (ensure_ctx()) return -1;
    q->quant_8x8 = t_quant_8x8; q->quant_4x4 = t_quant_4x4; q->quant_4x4_dc = t_quant_4x4_dc; q->quant_2x2_dc = t_quant_2x2_dc;
    q->dequant_8x8 = t_dequant_8x8; q->dequant_4x4 = t_dequant_4x4; q->dequant_4x4_dc = t_dequant_4x4_dc;
    return 0;
}

extern "C" int x264_mc_init_cuda(x264_cuda_mc_functions_t *m)
{
    if (ensure_ctx()) return -1;
    m->mc_luma = t_mc_luma; m->get_ref = t_get_ref; m->hpel_filter = t_hpel_filter; m->frame_init_lowres_core = t_frame_init_lowres_core;
    // chroma MC and the bi-prediction averages, in the table's PIXEL_* order (mc.c:379-389)
    m->mc_chroma = t_mc_chroma;
    m->avg[0] = t_avg<16, 16>; m->avg[1] = t_avg<16, 8>; m->avg[2] = t_avg<8, 16>; m->avg[3] = t_avg<8, 8>; m->avg[4] = t_avg<8, 4>;
    m->avg[5] = t_avg<4, 8>; m->avg[6] = t_avg<4, 4>; m->avg[7] = t_avg<4, 2>; m->avg[8] = t_avg<2, 4>; m->avg[9] = t_avg<2, 2>;
    return 0;
}

extern "C" long long x264_cuda_tables_launches(void) { return g_launches.load(); }
extern "C" const char *x264_cuda_tables_error(void) { return g_failed.load() ? g_fail_msg : nullptr; }

extern "C" void x264_cuda_tables_shutdown(void)
{
    // closes every thread's context; call it when no thread is inside a table entry (threads that call an entry later open a new one)
    std::lock_guard<std::mutex> lk(g_reg_mu);
    for (x264_cuda_t *c : g_all_ctx) x264_cuda_close(c);
    g_all_ctx.clear();
    g_gen.fetch_add(1, std::memory_order_release);
    g_ctx = nullptr;
}
