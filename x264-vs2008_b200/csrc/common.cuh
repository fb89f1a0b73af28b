// common.cuh — internal declarations shared by the sm_100a translation units.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include "../../include/x264_cuda.h"

#define PADH 32
#define PADV 32
#define COST_MAX (1 << 28)

struct x264_cuda_t {
    int device;
    int sm_count;
    cudaStream_t own_stream;
    cudaStream_t stream;          // the stream every launch goes to
    int blocking_wait;            // wait for results on a blocking-sync event (the host thread sleeps) instead of spinning in cudaStreamSynchronize
    cudaEvent_t wait_event;
    long long launches;
    char err[256];
    int16_t *d_cost_mv[52];       // device copies of p_cost_mv (base pointers, 4*4*2048+1 entries)
    struct QuantTables *d_qt;     // device quantiser tables (x264_cuda_set_quant_tables)
    int have_qt8;
    void *d_cost_ptrs;            // device array of the 52 pointers above
    int cost_ptrs_dirty;
    // staging for the host-pointer entry points
    int *d_la_order; int la_w, la_h, la_n; int la_epoch; int *d_la_sums; // lookahead wavefront order (interior list, then all blocks) + result cells
    void *d_i16_state; int i16_state_n; unsigned i16_epoch;              // intra 16x16 encode wavefront: per-macroblock state words + ticket
    void *d_la_vbv;                                                      // per-evaluation row sums + inverse qscale factors (VBV / AQ form)
    int *d_deblock_progress; int deblock_rows; void *d_deblock_recs; size_t d_deblock_recs_size;   // per-row progress counters of x264_cuda_frame_deblock
    int resid_no_dct8;   // one-shot hint from x264_cuda_residual_inter to its _dev call
    void *d_scratch; size_t d_scratch_size; // kernel-private scratch (TESA candidate lists)
    int *d_mb_ticket;    // work counter of the persistent macroblock-batched search kernel
    void *d_stage; size_t d_stage_size;
    void *h_stage; size_t h_stage_size; // pinned
    // asynchronous grid calls: device ring the per-call job copies and grids are carved from, and a pool of fence events
    uint8_t *d_ring; size_t ring_size, ring_pos;
    cudaStream_t aux_stream;      // high-priority stream of the direct (zero-copy) grid call
    cudaEvent_t fence_pool[256]; int n_fence_pool;
};

struct QuantTables {
    uint16_t q4mf[4][52][16], q4bias[4][52][16];
    int dq4[4][6][16];
    uint16_t q8mf[2][52][64], q8bias[2][52][64];
    int dq8[2][6][64];
};

struct x264_cuda_frame_t {
    x264_cuda_t *ctx;
    x264_cuda_geom_t g;
    size_t plane_size;            // stride*(lines+2*PADV)
    size_t plane_size_lowres;
    uint8_t *buf;                 // 1 or 4 luma planes, contiguous (frame.c:66-77)
    uint8_t *plane[4];            // pixel (0,0) of filtered[0..3]
    uint8_t *buf_lowres;
    uint8_t *lowres[4];
    uint8_t *buf_chroma;          // U then V, each stride_c*(lines/2+2*16)
    uint8_t *chroma[2];           // pixel (0,0) of U, V
    int stride_c;
    // lookahead state (x264_cuda_frame_lookahead_alloc): lowres_mvs[2][n_dist][mb], lowres_mv_costs[2][n_dist][mb], i_intra_cost[mb]
    int la_dist; int16_t *la_mvs; int *la_costs; int *la_intra; int *la_done;
    uint16_t *buf_integral;
    uint16_t *integral;           // element (0,0) of the 8x8-sum plane; the 4x4 plane follows
};

int x264_cuda_fail(x264_cuda_t *ctx, const char *what, cudaError_t e);
// Every public entry point starts with this: the calling host thread may never have touched CUDA (the reference runs one
// pthread per frame in flight, S/encoder/encoder.c:1569-1608) and would otherwise launch on device 0 with another device's stream.
static inline void x264_cuda_enter(const x264_cuda_t *ctx) { if (ctx) cudaSetDevice(ctx->device); }
int x264_cuda_stage(x264_cuda_t *ctx, size_t dev_bytes, size_t host_bytes);
// caller memory <-> device.  Page-locked caller memory (x264_cuda_host_alloc / x264_cuda_host_register) is DMA'd directly;
// pageable memory goes through the pinned stage at `hs`.  results_out also waits for the stream (the call's completion point).
int x264_cuda_jobs_in(x264_cuda_t *ctx, void *d, const void *h, void *hs, size_t n);
int x264_cuda_results_out(x264_cuda_t *ctx, void *h, const void *d, void *hs, size_t n);
int x264_cuda_wait(x264_cuda_t *ctx);
void *x264_cuda_grid_ring(x264_cuda_t *ctx, size_t bytes); // slice of the device ring for one asynchronous call (nullptr on failure)
int x264_cuda_cost_tables(x264_cuda_t *ctx, const int16_t *const **d_ptrs); // device array of 52 table pointers

#define CUDA_TRY(ctx, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return x264_cuda_fail((ctx), #call, e_); } while (0)
#define LAUNCH_CHECK(ctx, name) do { (ctx)->launches++; cudaError_t e_ = cudaGetLastError(); \
    if (e_ != cudaSuccess) return x264_cuda_fail((ctx), name, e_); } while (0)

// ---- device helpers ----
__device__ __forceinline__ uint32_t sad4_acc(uint32_t a, uint32_t b, uint32_t acc)
{
    uint32_t r; // VABSDIFF4.U8.ACC : sum of |a.b[i]-b.b[i]| + acc, one ALU-pipe instruction on sm_100a
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(acc));
    return r;
}
__device__ __forceinline__ int clip3i(int v, int lo, int hi) { return min(max(v, lo), hi); }
__device__ __forceinline__ int clip_u8(int v) { return min(max(v, 0), 255); }
