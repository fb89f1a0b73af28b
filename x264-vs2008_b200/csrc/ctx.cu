// ctx.cu — context, device frames, host<->device plumbing of the x264_cuda C ABI (include/x264_cuda.h).
#include "common.cuh"
#include <cmath>
#include <cstdlib>

static char g_open_error[256] = "";

int x264_cuda_fail(x264_cuda_t *ctx, const char *what, cudaError_t e)
{
    char *dst = ctx ? ctx->err : g_open_error;
    snprintf(dst, 256, "x264_cuda: %s failed: %s", what, cudaGetErrorString(e));
    return -1;
}

extern "C" const char *x264_cuda_error(const x264_cuda_t *ctx) { return ctx ? ctx->err : g_open_error; }

extern "C" int x264_cuda_open(x264_cuda_t **pctx, int device)
{
    *pctx = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        snprintf(g_open_error, 256, "x264_cuda: no CUDA device (%s); this back-end has no CPU fallback",
                 e != cudaSuccess ? cudaGetErrorString(e) : "count 0");
        return -1;
    }
    if (device < 0 || device >= n) {
        snprintf(g_open_error, 256, "x264_cuda: device %d out of range (0..%d)", device, n - 1);
        return -1;
    }
    // how host threads wait inside synchronising calls: the CUDA default spins; with many encoder threads per device
    // (integration/x264_b200_gops.c) yielding or sleeping waiters leave the cores and the driver's locks to the threads that have work.
    // Must be chosen before the device's context exists, hence an environment switch read by the first open of the process.
    if (const char *sch = getenv("X264_CUDA_SCHED")) {
        const unsigned fl = !strcmp(sch, "blocking") ? cudaDeviceScheduleBlockingSync : !strcmp(sch, "yield") ? cudaDeviceScheduleYield
                          : !strcmp(sch, "spin") ? cudaDeviceScheduleSpin : cudaDeviceScheduleAuto;
        if (cudaSetDeviceFlags(fl) != cudaSuccess) cudaGetLastError(); // a context that already exists keeps its mode
    }
    cudaDeviceProp prop;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
        return x264_cuda_fail(nullptr, "cudaSetDevice", e);
    if (prop.major != 10) {
        snprintf(g_open_error, 256, "x264_cuda: device %d is sm_%d%d; this library is built for sm_100a only", device,
                 prop.major, prop.minor);
        return -1;
    }
    x264_cuda_t *ctx = (x264_cuda_t *)calloc(1, sizeof(x264_cuda_t));
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess) {
        free(ctx);
        return x264_cuda_fail(nullptr, "cudaStreamCreate", e);
    }
    ctx->stream = ctx->own_stream;
    *pctx = ctx;
    return 0;
}

extern "C" void x264_cuda_close(x264_cuda_t *ctx)
{
    x264_cuda_enter(ctx);
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->wait_event) cudaEventDestroy(ctx->wait_event);
    for (int i = 0; i < 52; i++) cudaFree(ctx->d_cost_mv[i]);
    cudaFree(ctx->d_cost_ptrs);
    cudaFree(ctx->d_qt);
    cudaFree(ctx->d_stage);
    cudaFree(ctx->d_scratch);
    cudaFree(ctx->d_mb_ticket);
    cudaFree(ctx->d_deblock_progress);
    cudaFree(ctx->d_deblock_recs);
    cudaFree(ctx->d_la_order); cudaFree(ctx->d_la_sums); cudaFree(ctx->d_la_vbv); cudaFree(ctx->d_i16_state);
    cudaFreeHost(ctx->h_stage);
    cudaFree(ctx->d_ring);
    if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
    for (int i = 0; i < ctx->n_fence_pool; i++) cudaEventDestroy(ctx->fence_pool[i]);
    cudaStreamDestroy(ctx->own_stream);
    free(ctx);
}

extern "C" int x264_cuda_set_stream(x264_cuda_t *ctx, void *s)
{
    x264_cuda_enter(ctx);
    ctx->stream = s ? (cudaStream_t)s : ctx->own_stream;
    return 0;
}
extern "C" void *x264_cuda_get_stream(x264_cuda_t *ctx) { return (void *)ctx->stream; }
extern "C" int x264_cuda_synchronize(x264_cuda_t *ctx)
{
    x264_cuda_enter(ctx);
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
extern "C" long long x264_cuda_launch_count(const x264_cuda_t *ctx) { return ctx->launches; }
extern "C" int x264_cuda_sm_count(const x264_cuda_t *ctx) { return ctx->sm_count; }

int x264_cuda_stage(x264_cuda_t *ctx, size_t dev_bytes, size_t host_bytes)
{
    if (dev_bytes > ctx->d_stage_size) {
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->d_stage);
        ctx->d_stage = nullptr; ctx->d_stage_size = 0;
        size_t sz = dev_bytes + dev_bytes / 2;
        CUDA_TRY(ctx, cudaMalloc(&ctx->d_stage, sz));
        ctx->d_stage_size = sz;
    }
    if (host_bytes > ctx->h_stage_size) {
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFreeHost(ctx->h_stage);
        ctx->h_stage = nullptr; ctx->h_stage_size = 0;
        size_t sz = host_bytes + host_bytes / 2;
        CUDA_TRY(ctx, cudaMallocHost(&ctx->h_stage, sz));
        ctx->h_stage_size = sz;
    }
    return 0;
}

// Device memory for calls that return before their work is done: slices are handed out round-robin from one ring; when the ring wraps,
// the stream is drained first, so a slice is never reused while an earlier call on this context may still be using it.
void *x264_cuda_grid_ring(x264_cuda_t *ctx, size_t bytes)
{
    bytes = (bytes + 255) & ~(size_t)255;
    if (bytes * 4 > ctx->ring_size) {
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return nullptr;
        cudaFree(ctx->d_ring);
        ctx->d_ring = nullptr; ctx->ring_size = ctx->ring_pos = 0;
        const size_t sz = bytes * 4 > ((size_t)64 << 20) ? bytes * 4 : ((size_t)64 << 20);
        cudaError_t e = cudaMalloc(&ctx->d_ring, sz);
        if (e != cudaSuccess) { x264_cuda_fail(ctx, "grid ring allocation", e); return nullptr; }
        ctx->ring_size = sz;
    }
    if (ctx->ring_pos + bytes > ctx->ring_size) {
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return nullptr;
        ctx->ring_pos = 0;
    }
    void *p = ctx->d_ring + ctx->ring_pos;
    ctx->ring_pos += bytes;
    return p;
}

// Callers that know their working set (the live encoder: two frames' worth of grids) size the ring once, so that it wraps about once
// per frame pair — when everything older has long been consumed — instead of draining the stream in the middle of a frame.
extern "C" int x264_cuda_grid_ring_reserve(x264_cuda_t *ctx, size_t bytes)
{
    x264_cuda_enter(ctx);
    if (bytes <= ctx->ring_size) return 0;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->d_ring);
    ctx->d_ring = nullptr; ctx->ring_size = ctx->ring_pos = 0;
    CUDA_TRY(ctx, cudaMalloc(&ctx->d_ring, bytes));
    ctx->ring_size = bytes;
    return 0;
}

// Fences: completion markers on the context's stream for the asynchronous entry points.
extern "C" void *x264_cuda_fence_record(x264_cuda_t *ctx)
{
    x264_cuda_enter(ctx);
    cudaEvent_t ev;
    if (ctx->n_fence_pool > 0) ev = ctx->fence_pool[--ctx->n_fence_pool];
    else if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) { x264_cuda_fail(ctx, "cudaEventCreate", cudaGetLastError()); return nullptr; }
    if (cudaEventRecord(ev, ctx->stream) != cudaSuccess) { x264_cuda_fail(ctx, "cudaEventRecord", cudaGetLastError()); return nullptr; }
    return (void *)ev;
}
extern "C" int x264_cuda_fence_wait(x264_cuda_t *ctx, void *fence)
{
    x264_cuda_enter(ctx);
    if (!fence) return 0;
    cudaEvent_t ev = (cudaEvent_t)fence;
    CUDA_TRY(ctx, cudaEventSynchronize(ev));
    if (ctx->n_fence_pool < 256) ctx->fence_pool[ctx->n_fence_pool++] = ev; else cudaEventDestroy(ev);
    return 0;
}

static bool is_pinned(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}
int x264_cuda_jobs_in(x264_cuda_t *ctx, void *d, const void *h, void *hs, size_t n)
{
    if (!is_pinned(h)) { memcpy(hs, h, n); h = hs; }
    CUDA_TRY(ctx, cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}
// waits for everything queued on the context's stream: spinning (lowest latency, the default) or, with x264_cuda_set_blocking_wait, on a
// blocking-sync event so that the host thread sleeps — the right choice when more frame threads than host cores wait at once
int x264_cuda_wait(x264_cuda_t *ctx)
{
    if (!ctx->blocking_wait) {
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        return 0;
    }
    if (!ctx->wait_event) CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->wait_event, cudaEventBlockingSync | cudaEventDisableTiming));
    CUDA_TRY(ctx, cudaEventRecord(ctx->wait_event, ctx->stream));
    CUDA_TRY(ctx, cudaEventSynchronize(ctx->wait_event));
    return 0;
}
extern "C" int x264_cuda_set_blocking_wait(x264_cuda_t *ctx, int on)
{
    ctx->blocking_wait = on != 0;
    return 0;
}

int x264_cuda_results_out(x264_cuda_t *ctx, void *h, const void *d, void *hs, size_t n)
{
    const bool direct = is_pinned(h);
    CUDA_TRY(ctx, cudaMemcpyAsync(direct ? h : hs, d, n, cudaMemcpyDeviceToHost, ctx->stream));
    if (x264_cuda_wait(ctx)) return -1;
    if (!direct) memcpy(h, hs, n);
    return 0;
}
extern "C" void *x264_cuda_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
extern "C" void x264_cuda_host_free(void *p) { if (p) cudaFreeHost(p); }
extern "C" int x264_cuda_host_register(void *p, size_t bytes)
{
    if (cudaHostRegister(p, bytes, cudaHostRegisterDefault) != cudaSuccess) { cudaGetLastError(); return -1; }
    return 0;
}
extern "C" int x264_cuda_host_unregister(void *p)
{
    if (cudaHostUnregister(p) != cudaSuccess) { cudaGetLastError(); return -1; }
    return 0;
}

// ------------------------------------------------------------------ frames
static inline int align_up(int x, int a) { return (x + a - 1) & ~(a - 1); }

extern "C" x264_cuda_frame_t *x264_cuda_frame_new(x264_cuda_t *ctx, int width, int height, int flags)
{
    x264_cuda_enter(ctx);
    if (width <= 0 || height <= 0) {
        snprintf(ctx->err, 256, "x264_cuda: invalid frame size %dx%d", width, height);
        return nullptr;
    }
    cudaSetDevice(ctx->device);
    x264_cuda_frame_t *f = (x264_cuda_frame_t *)calloc(1, sizeof(*f));
    f->ctx = ctx;
    x264_cuda_geom_t &g = f->g;
    g.width = width; g.height = height;
    g.mb_width = (width + 15) / 16; g.mb_height = (height + 15) / 16;
    g.lines = g.mb_height * 16;
    g.stride = align_up(g.mb_width * 16 + 2 * PADH, 128);
    g.width_lowres = g.mb_width * 8;
    g.lines_lowres = g.lines / 2;
    g.stride_lowres = align_up(g.width_lowres + 2 * PADH, 128);
    g.flags = flags;
    f->plane_size = (size_t)g.stride * (g.lines + 2 * PADV);
    f->plane_size_lowres = (size_t)g.stride_lowres * (g.lines_lowres + 2 * PADV);
    int nplanes = (flags & X264_CUDA_FRAME_HPEL) ? 4 : 1;
    // +256 slack: kernels read whole aligned words/vectors that may straddle the last row's end
    cudaError_t e = cudaMalloc(&f->buf, nplanes * f->plane_size + 256);
    if (e == cudaSuccess) e = cudaMemsetAsync(f->buf, 0, nplanes * f->plane_size + 256, ctx->stream);
    for (int i = 0; i < nplanes && e == cudaSuccess; i++)
        f->plane[i] = f->buf + i * f->plane_size + (size_t)g.stride * PADV + PADH;
    if (e == cudaSuccess && (flags & X264_CUDA_FRAME_LOWRES)) {
        e = cudaMalloc(&f->buf_lowres, 4 * f->plane_size_lowres + 256);
        if (e == cudaSuccess) e = cudaMemsetAsync(f->buf_lowres, 0, 4 * f->plane_size_lowres + 256, ctx->stream);
        for (int i = 0; i < 4; i++)
            f->lowres[i] = f->buf_lowres + i * f->plane_size_lowres + (size_t)g.stride_lowres * PADV + PADH;
    }
    if (e == cudaSuccess && (flags & X264_CUDA_FRAME_CHROMA)) {
        f->stride_c = g.stride / 2;
        const size_t csz = (size_t)f->stride_c * (g.lines / 2 + 2 * 16);
        e = cudaMalloc(&f->buf_chroma, 2 * csz + 256);
        if (e == cudaSuccess) e = cudaMemsetAsync(f->buf_chroma, 0, 2 * csz + 256, ctx->stream);
        for (int i = 0; i < 2; i++) f->chroma[i] = f->buf_chroma + i * csz + (size_t)f->stride_c * 16 + 16;
    }
    if (e == cudaSuccess && (flags & (X264_CUDA_FRAME_INTEGRAL | X264_CUDA_FRAME_INTEGRAL4))) {
        size_t n = f->plane_size * ((flags & X264_CUDA_FRAME_INTEGRAL4) ? 2 : 1);
        e = cudaMalloc(&f->buf_integral, n * sizeof(uint16_t) + 256);
        if (e == cudaSuccess) e = cudaMemsetAsync(f->buf_integral, 0, n * sizeof(uint16_t) + 256, ctx->stream);
        f->integral = f->buf_integral + (size_t)g.stride * PADV + PADH;
    }
    if (e != cudaSuccess) {
        x264_cuda_fail(ctx, "frame allocation", e);
        x264_cuda_frame_delete(f);
        return nullptr;
    }
    return f;
}

extern "C" void x264_cuda_frame_delete(x264_cuda_frame_t *f)
{
    if (!f) return;
    x264_cuda_enter(f->ctx);
    cudaStreamSynchronize(f->ctx->stream);
    cudaFree(f->buf);
    cudaFree(f->buf_lowres);
    cudaFree(f->buf_chroma);
    cudaFree(f->la_mvs); cudaFree(f->la_costs); cudaFree(f->la_intra); cudaFree(f->la_done);
    cudaFree(f->buf_integral);
    free(f);
}

extern "C" void x264_cuda_frame_geometry(const x264_cuda_frame_t *f, x264_cuda_geom_t *g) { *g = f->g; }

extern "C" void *x264_cuda_frame_plane(const x264_cuda_frame_t *f, int plane)
{
    if (plane >= 0 && plane < 4) return f->plane[plane];
    if (plane >= X264_CUDA_PLANE_LOWRES && plane < X264_CUDA_PLANE_LOWRES + 4) return f->lowres[plane - X264_CUDA_PLANE_LOWRES];
    if (plane == X264_CUDA_PLANE_CB || plane == X264_CUDA_PLANE_CR) return f->chroma[plane - X264_CUDA_PLANE_CB];
    if (plane == X264_CUDA_PLANE_INTEGRAL) return f->integral;
    if (plane == X264_CUDA_PLANE_INTEGRAL4)
        return (f->g.flags & X264_CUDA_FRAME_INTEGRAL4) ? f->integral + f->plane_size : nullptr;
    return nullptr;
}

// a picture larger than the frame's macroblock-padded area would overwrite the neighbouring planes of the allocation
static int upload_bounds(x264_cuda_t *ctx, const char *who, int cols, int rows, int max_cols, int max_rows)
{
    if (cols < 0 || rows < 0 || cols > max_cols || rows > max_rows) {
        snprintf(ctx->err, 256, "%s: %d x %d does not fit the frame's %d x %d plane", who, cols, rows, max_cols, max_rows);
        return -1;
    }
    return 0;
}

extern "C" int x264_cuda_frame_upload(x264_cuda_t *ctx, x264_cuda_frame_t *f, const uint8_t *src, int src_stride,
                                      int cols, int rows)
{
    x264_cuda_enter(ctx);
    if (upload_bounds(ctx, "x264_cuda_frame_upload", cols, rows, f->g.mb_width * 16, f->g.lines)) return -1;
    CUDA_TRY(ctx, cudaMemcpy2DAsync(f->plane[0], f->g.stride, src, src_stride, cols, rows, cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}
extern "C" int x264_cuda_frame_upload_chroma(x264_cuda_t *ctx, x264_cuda_frame_t *f, int plane, const uint8_t *src, int src_stride,
                                             int cols, int rows)
{
    x264_cuda_enter(ctx);
    if (!f->buf_chroma || (plane != X264_CUDA_PLANE_CB && plane != X264_CUDA_PLANE_CR)) {
        snprintf(ctx->err, 256, "x264_cuda_frame_upload_chroma: frame has no chroma plane %d", plane);
        return -1;
    }
    if (upload_bounds(ctx, "x264_cuda_frame_upload_chroma", cols, rows, f->g.mb_width * 8, f->g.lines / 2)) return -1;
    CUDA_TRY(ctx, cudaMemcpy2DAsync(f->chroma[plane - X264_CUDA_PLANE_CB], f->stride_c, src, src_stride, cols, rows,
                                    cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}
extern "C" int x264_cuda_frame_upload_dev(x264_cuda_t *ctx, x264_cuda_frame_t *f, const void *dsrc, int src_stride,
                                          int cols, int rows)
{
    x264_cuda_enter(ctx);
    if (upload_bounds(ctx, "x264_cuda_frame_upload_dev", cols, rows, f->g.mb_width * 16, f->g.lines)) return -1;
    CUDA_TRY(ctx, cudaMemcpy2DAsync(f->plane[0], f->g.stride, dsrc, src_stride, cols, rows, cudaMemcpyDeviceToDevice, ctx->stream));
    return 0;
}

extern "C" int x264_cuda_frame_download(x264_cuda_t *ctx, const x264_cuda_frame_t *f, int plane, void *dst, int dst_stride)
{
    x264_cuda_enter(ctx);
    const x264_cuda_geom_t &g = f->g;
    const void *p00 = x264_cuda_frame_plane(f, plane);
    if (!p00) {
        snprintf(ctx->err, 256, "x264_cuda: frame has no plane %d", plane);
        return -1;
    }
    bool lowres = plane >= X264_CUDA_PLANE_LOWRES && plane < X264_CUDA_PLANE_LOWRES + 4;
    bool chroma = plane == X264_CUDA_PLANE_CB || plane == X264_CUDA_PLANE_CR;
    bool integ = plane == X264_CUDA_PLANE_INTEGRAL || plane == X264_CUDA_PLANE_INTEGRAL4;
    int es = integ ? 2 : 1;
    int stride = lowres ? g.stride_lowres : chroma ? f->stride_c : g.stride;
    int lines = lowres ? g.lines_lowres : chroma ? g.lines / 2 : g.lines;
    int w = lowres ? g.width_lowres : chroma ? g.mb_width * 8 : g.mb_width * 16;
    int padh = chroma ? 16 : PADH, padv = chroma ? 16 : PADV;
    const uint8_t *src = (const uint8_t *)p00 - ((size_t)stride * padv + padh) * es;
    int cols = w + 2 * padh;
    if (cols > dst_stride) cols = dst_stride;
    CUDA_TRY(ctx, cudaMemcpy2DAsync(dst, (size_t)dst_stride * es, src, (size_t)stride * es, (size_t)cols * es, lines + 2 * padv,
                                    cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ------------------------------------------------------------------ MV cost tables
extern "C" int x264_cuda_set_cost_mv(x264_cuda_t *ctx, int qp, const int16_t *table)
{
    x264_cuda_enter(ctx);
    if (qp < 0 || qp > 51) { snprintf(ctx->err, 256, "x264_cuda: qp %d out of range", qp); return -1; }
    const size_t n = (4 * 4 * 2048 + 1) * sizeof(int16_t);
    if (!ctx->d_cost_mv[qp]) CUDA_TRY(ctx, cudaMalloc(&ctx->d_cost_mv[qp], n + 16));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_cost_mv[qp], table, n, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)); // `table` may be pageable and short-lived
    ctx->cost_ptrs_dirty = 1;
    return 0;
}

int x264_cuda_cost_tables(x264_cuda_t *ctx, const int16_t *const **d_ptrs)
{
    if (!ctx->d_cost_ptrs) {
        CUDA_TRY(ctx, cudaMalloc(&ctx->d_cost_ptrs, 52 * sizeof(void *)));
        ctx->cost_ptrs_dirty = 1;
    }
    // A job may name any qp: tables the caller has not uploaded (x264_cuda_set_cost_mv) are filled with the library's own
    // x264_cuda_host_cost_mv — the same expression as x264_mb_analyse_load_costs, verified equal to the reference's for all 52 qps —
    // so that no kernel ever dereferences a missing table.
    for (int q = 0; q < 52; q++)
        if (!ctx->d_cost_mv[q]) {
            int16_t *t = (int16_t *)malloc((4 * 4 * 2048 + 1) * sizeof(int16_t));
            x264_cuda_host_cost_mv(q, t);
            const int rc = x264_cuda_set_cost_mv(ctx, q, t);
            free(t);
            if (rc) return -1;
        }
    if (ctx->cost_ptrs_dirty) {
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_cost_ptrs, ctx->d_cost_mv, 52 * sizeof(void *), cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->cost_ptrs_dirty = 0;
    }
    *d_ptrs = (const int16_t *const *)ctx->d_cost_ptrs;
    return 0;
}
