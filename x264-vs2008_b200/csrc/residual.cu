// residual.cu — batched H.264 integer transforms and (de)quantisation: S/common/dct.c, S/common/quant.c and the
// inter-macroblock sequencing of S/encoder/macroblock.c:596-742 (+ chroma :272-363).
//
// Roofline class: HBM (80 B of traffic per 4x4 block, 320 B per 8x8 block against a few hundred integer ops).
// Layout conventions follow the reference: coefficient blocks are stored TRANSPOSED (dct[i][k], dct.c:131-154),
// every intermediate that the reference keeps in an int16_t array is narrowed to int16 at the same point.
#include "dct_dev.cuh"
#include "pixel_dev.cuh"

namespace {

// ---- zig-zag (frame) orders as flat indices into the reference's transposed blocks (dct.c:488-560)
__constant__ uint8_t c_zz8[64] = { 0,  8,  1,  2,  9,  16, 24, 17, 10, 3,  4,  11, 18, 25, 32, 40, 33, 26, 19, 12, 5,  6,
                                   13, 20, 27, 34, 41, 48, 56, 49, 42, 35, 28, 21, 14, 7,  15, 22, 29, 36, 43, 50, 57, 58,
                                   51, 44, 37, 30, 23, 31, 38, 45, 52, 59, 60, 53, 46, 39, 47, 54, 61, 62, 55, 63 };
__constant__ uint8_t c_dec4[16] = { 3, 2, 2, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
__constant__ uint8_t c_dec8[64] = { 3, 3, 3, 3, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1 };

// quant.c:219-252 on a zig-zag ordered block in local memory/registers; first = 1 skips the DC (score15)
__device__ int decimate_score(const int16_t *lv, int first, int n)
{
    const uint8_t *tab = n == 64 ? c_dec8 : c_dec4;
    int idx = n - 1, score = 0;
    while (idx >= first && lv[idx] == 0) idx--;
    while (idx >= first) {
        if ((unsigned)(lv[idx--] + 1) > 2) return 9;
        int run = 0;
        while (idx >= first && lv[idx] == 0) { idx--; run++; }
        score += tab[run];
    }
    return score;
}

// zig-zag position k of a 4x4 block as a compile-time constant (the loops that use it are fully unrolled, so level k is a plain
// register move instead of an indexed read of the coefficient array in local memory)
__host__ __device__ constexpr int zz4(int k)
{
    return k == 0 ? 0 : k == 1 ? 4 : k == 2 ? 1 : k == 3 ? 2 : k == 4 ? 5 : k == 5 ? 8 : k == 6 ? 12 : k == 7 ? 9 : k == 8 ? 6 : k == 9 ? 3 : k == 10 ? 7 :
           k == 11 ? 10 : k == 12 ? 13 : k == 13 ? 14 : k == 14 ? 11 : 15;
}

// x264_decimate_score15 / 16 (quant.c:219-252) from two bit masks over the zig-zag levels: nzm = level != 0, big = |level| > 1.
// Walks the non-zero levels from the last one down; each scores by the run of zeros below it (down to `first`).
__device__ __forceinline__ int decimate_score4(unsigned nzm, unsigned big, int first)
{
    if (first) { nzm &= ~1u; big &= ~1u; }
    if (big) return 9;
    int score = 0;
    while (nzm) {
        const int pos = 31 - __clz(nzm);
        nzm ^= 1u << pos;
        const int next = nzm ? 31 - __clz(nzm) : first - 1;
        const int run = pos - next - 1;
        score += run < 1 ? 3 : run < 3 ? 2 : run < 6 ? 1 : 0; // x264_decimate_table4
    }
    return score;
}

// quant_4x4 of the transposed coefficient block c[] (quant.c:33-58) with the list's tables fetched as four 16-byte loads, then — when
// something survives — the zig-zag levels packed two per word (lvw), the decimation score (want_score; first = 1 for the AC-only
// score15) and dequant_4x4 in place (quant.c:82-109).  lvw is all zero when nothing survives.  Returns nz (0/1).
__device__ __forceinline__ int quant_block4(const QuantTables *__restrict__ qt, int list, int qp, int (&c)[16], bool want_score, int first,
                                            uint32_t (&lvw)[8], int &score)
{
    const uint4 *mf4 = (const uint4 *)qt->q4mf[list][qp], *bs4 = (const uint4 *)qt->q4bias[list][qp];
    const uint4 m0 = __ldg(mf4), m1 = __ldg(mf4 + 1), b0 = __ldg(bs4), b1 = __ldg(bs4 + 1);
    const uint32_t mw[8] = { m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w }, bw[8] = { b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w };
    int nz = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const int sh = 16 * (k & 1);
        c[k] = quant1(c[k], (int)((mw[k >> 1] >> sh) & 0xffff), (int)((bw[k >> 1] >> sh) & 0xffff));
        nz |= c[k];
    }
    score = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) lvw[k] = 0;
    if (!nz) return 0;
    unsigned nzm = 0, big = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const int v = c[zz4(k)];
        lvw[k >> 1] |= (uint32_t)(uint16_t)v << (16 * (k & 1));
        nzm |= (unsigned)(v != 0) << k;
        big |= (unsigned)((unsigned)(v + 1) > 2u) << k;
    }
    if (want_score) score = decimate_score4(nzm, big, first);
    const int4 *dq = (const int4 *)qt->dq4[list][qp % 6];
    const int qbits = qp / 6 - 4;
#pragma unroll
    for (int g = 0; g < 4; g++) {
        const int4 d = __ldg(dq + g);
        c[4 * g] = dequant1(c[4 * g], d.x, qbits); c[4 * g + 1] = dequant1(c[4 * g + 1], d.y, qbits);
        c[4 * g + 2] = dequant1(c[4 * g + 2], d.z, qbits); c[4 * g + 3] = dequant1(c[4 * g + 3], d.w, qbits);
    }
    return 1;
}

// =========================================================================================================
// function-level batches over packed blocks
__global__ void __launch_bounds__(128) block_residual4_kernel(const QuantTables *__restrict__ qt, int n, const uint8_t *__restrict__ fenc,
                                                              const uint8_t *__restrict__ pred, const uint8_t *__restrict__ qp,
                                                              const uint8_t *__restrict__ cat, int16_t *dct_out, int16_t *level_out,
                                                              uint8_t *nz_out, uint8_t *recon_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 fe = __ldg((const uint4 *)fenc + i), pr = __ldg((const uint4 *)pred + i);
    const uint32_t fw[4] = { fe.x, fe.y, fe.z, fe.w }, pw[4] = { pr.x, pr.y, pr.z, pr.w };
    int d[16], c[16], p[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        p[k] = (pw[k >> 2] >> (8 * (k & 3))) & 255;
        d[k] = (int)((fw[k >> 2] >> (8 * (k & 3))) & 255) - p[k];
    }
    fwd4x4(d, c);
    if (dct_out)
#pragma unroll
        for (int k = 0; k < 16; k++) dct_out[i * 16 + k] = (int16_t)c[k];
    const int q = min((int)qp[i], 51), l = cat[i] & 3;
    const uint16_t *mf = qt->q4mf[l][q], *bias = qt->q4bias[l][q];
    int nz = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) { c[k] = quant1(c[k], mf[k], bias[k]); nz |= c[k]; }
    if (level_out)
#pragma unroll
        for (int k = 0; k < 16; k++) level_out[i * 16 + k] = (int16_t)c[k];
    if (nz_out) nz_out[i] = nz != 0;
    if (recon_out) {
        if (nz) {
            const int *dmf = qt->dq4[l][q % 6];
            const int qbits = q / 6 - 4;
            int r[16];
#pragma unroll
            for (int k = 0; k < 16; k++) c[k] = dequant1(c[k], dmf[k], qbits);
            inv4x4(c, r);
#pragma unroll
            for (int k = 0; k < 16; k++) p[k] = clip_u8(p[k] + r[k]);
        }
        uint32_t w[4] = { 0, 0, 0, 0 };
#pragma unroll
        for (int k = 0; k < 16; k++) w[k >> 2] |= (uint32_t)p[k] << (8 * (k & 3));
        ((uint4 *)recon_out)[i] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

__global__ void __launch_bounds__(64) block_residual8_kernel(const QuantTables *__restrict__ qt, int n, const uint8_t *__restrict__ fenc,
                                                             const uint8_t *__restrict__ pred, const uint8_t *__restrict__ qp,
                                                             const uint8_t *__restrict__ cat, int16_t *dct_out, int16_t *level_out,
                                                             uint8_t *nz_out, uint8_t *recon_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int d[64], c[64];
    for (int k = 0; k < 64; k++) d[k] = (int)fenc[(size_t)i * 64 + k] - (int)pred[(size_t)i * 64 + k];
    fwd8x8(d, c);
    if (dct_out) for (int k = 0; k < 64; k++) dct_out[(size_t)i * 64 + k] = (int16_t)c[k];
    const int q = min((int)qp[i], 51), l = cat[i] & 1;
    const uint16_t *mf = qt->q8mf[l][q], *bias = qt->q8bias[l][q];
    int nz = 0;
    for (int k = 0; k < 64; k++) { c[k] = quant1(c[k], mf[k], bias[k]); nz |= c[k]; }
    if (level_out) for (int k = 0; k < 64; k++) level_out[(size_t)i * 64 + k] = (int16_t)c[k];
    if (nz_out) nz_out[i] = nz != 0;
    if (recon_out) {
        if (nz) {
            const int *dmf = qt->dq8[l][q % 6];
            const int qbits = q / 6 - 6;
            for (int k = 0; k < 64; k++) c[k] = dequant1(c[k], dmf[k], qbits);
            inv8x8(c, d);
            for (int k = 0; k < 64; k++) recon_out[(size_t)i * 64 + k] = (uint8_t)clip_u8((int)pred[(size_t)i * 64 + k] + d[k]);
        } else
            for (int k = 0; k < 64; k++) recon_out[(size_t)i * 64 + k] = pred[(size_t)i * 64 + k];
    }
}

__global__ void __launch_bounds__(128) block_dc_kernel(const QuantTables *__restrict__ qt, int n, const int16_t *__restrict__ dc_in,
                                                       const uint8_t *__restrict__ qp, const uint8_t *__restrict__ cat, int16_t *fwd_out,
                                                       int16_t *level_out, uint8_t *nz_out, int16_t *deq_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int d[16];
#pragma unroll
    for (int k = 0; k < 16; k++) d[k] = dc_in[i * 16 + k];
    hadamard_dc(d, true);
    if (fwd_out)
#pragma unroll
        for (int k = 0; k < 16; k++) fwd_out[i * 16 + k] = (int16_t)d[k];
    const int q = min((int)qp[i], 51), l = cat[i] & 3;
    const int mf = qt->q4mf[l][q][0] >> 1, bias = qt->q4bias[l][q][0] << 1; // macroblock.c:250
    int nz = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) { d[k] = quant1(d[k], mf, bias); nz |= d[k]; }
    if (level_out)
#pragma unroll
        for (int k = 0; k < 16; k++) level_out[i * 16 + k] = (int16_t)d[k];
    if (nz_out) nz_out[i] = nz != 0;
    if (deq_out) {
        hadamard_dc(d, false);
        const int qbits = q / 6 - 6, dmf0 = qt->dq4[l][q % 6][0]; // quant.c:148-178
#pragma unroll
        for (int k = 0; k < 16; k++)
            d[k] = qbits >= 0 ? s16(d[k] * (dmf0 << qbits)) : s16((d[k] * dmf0 + (1 << (-qbits - 1))) >> (-qbits));
#pragma unroll
        for (int k = 0; k < 16; k++) deq_out[i * 16 + k] = (int16_t)d[k];
    }
}

// =========================================================================================================
// Inter macroblock residual: ONE WARP PER MACROBLOCK.  Lanes 0..15 own the luma 4x4 blocks in the reference's
// block order (block_idx_x/y, S/common/macroblock.h:195-202) — or lanes 0..3 the four 8x8 blocks with 8x8dct —
// and lanes 16..23 the chroma AC blocks (U0..3, V0..3).  Decimation sums and the chroma 2x2 DC travel by shuffles.
struct FrameRefs { const uint8_t *fe_y, *fe_u, *fe_v; uint8_t *fd_y, *fd_u, *fd_v; int stride, stride_c; };

__device__ __forceinline__ void load4x4(const uint8_t *p, int stride, int (&v)[16])
{
#pragma unroll
    for (int y = 0; y < 4; y++) {
        const uint32_t w = *(const uint32_t *)(p + (size_t)y * stride);
#pragma unroll
        for (int x = 0; x < 4; x++) v[y * 4 + x] = (w >> (8 * x)) & 255;
    }
}
__device__ __forceinline__ void store4x4(uint8_t *p, int stride, const int (&v)[16])
{
#pragma unroll
    for (int y = 0; y < 4; y++)
        *(uint32_t *)(p + (size_t)y * stride) = (uint32_t)v[y * 4] | ((uint32_t)v[y * 4 + 1] << 8) | ((uint32_t)v[y * 4 + 2] << 16) |
                                                ((uint32_t)v[y * 4 + 3] << 24);
}

// x264_mb_encode_8x8_chroma (macroblock.c:272-363) on the chroma lanes 16..23 of a macroblock's warp (U0..3, V0..3): f = source block,
// p = prediction (both valid on those lanes), dst = where the lane's reconstructed 4x4 goes (always stored when store_pred: the intra
// caller has the prediction in registers only).  list = CQM_4IC + b_inter; decim as computed at :275.  zero_uncoded: AC levels of blocks
// whose nnz ends up 0 (DC-only planes) are left zero in the record.  Returns i_cbp_chroma on every lane.
__device__ __forceinline__ int chroma_blocks(const QuantTables *__restrict__ qt, int list, int cqp, bool decim, bool intra, int lane, const int (&f)[16],
                                             int (&p)[16], uint8_t *dst, int stride_c, x264_cuda_mb_coeffs_t *out)
{
    const unsigned FULL = 0xffffffffu;
    const int cl = lane - 16;
    const bool mine = lane >= 16 && lane < 24;
    const int ch = (cl >> 2) & 1, bi = cl & 3;
    int c[16], nz = 0, score = 0, dc0 = 0;
    uint32_t lvw[8];
    if (mine) {
        int d[16];
#pragma unroll
        for (int k = 0; k < 16; k++) d[k] = f[k] - p[k];
        fwd4x4(d, c);
        dc0 = c[0]; c[0] = 0; // dct2x2dc takes the DCs out (macroblock.c:72-85)
        nz = quant_block4(qt, list, cqp, c, decim, 1, lvw, score);
    }
    // gather the four DCs / scores / nz of this lane's channel
    const int cb = 16 + ch * 4;
    const int b0 = __shfl_sync(FULL, dc0, cb), b1 = __shfl_sync(FULL, dc0, cb + 1), b2 = __shfl_sync(FULL, dc0, cb + 2),
              b3 = __shfl_sync(FULL, dc0, cb + 3);
    int tot = 0, nz_ac = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { tot += __shfl_sync(FULL, score, cb + k); nz_ac |= __shfl_sync(FULL, nz, cb + k); }
    // dct2x2dc: d[0][0], d[1][0], d[0][1], d[1][1] (macroblock.c:72-80), flat order d[0][0],d[0][1],d[1][0],d[1][1]
    const int e0 = b0 + b1, e1 = b2 + b3, e2 = b0 - b1, e3 = b2 - b3;
    int dc[4] = { s16(e0 + e1), s16(e0 - e1), s16(e2 + e3), s16(e2 - e3) };
    const int mf0 = qt->q4mf[list][cqp][0] >> 1, bias0 = qt->q4bias[list][cqp][0] << 1;
    int nz_dc = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { dc[k] = quant1(dc[k], mf0, bias0); nz_dc |= dc[k]; }
    nz_dc = nz_dc != 0;
    // IDCT_DEQUANT_START (macroblock.c:42-53)
    const int g0 = dc[0] + dc[1], g1 = dc[2] + dc[3], g2 = dc[0] - dc[1], g3 = dc[2] - dc[3];
    int dmf = qt->dq4[list][cqp % 6][0], qbits = cqp / 6 - 5;
    if (qbits > 0) { dmf <<= qbits; qbits = 0; }
    const int o4[4] = { s16((g0 + g1) * dmf >> -qbits), s16((g0 - g1) * dmf >> -qbits), s16((g2 + g3) * dmf >> -qbits),
                        s16((g2 - g3) * dmf >> -qbits) };
    const bool dc_only = (decim && tot < 7) || !nz_ac;
    if (mine) {
        {   // every chroma block's 32 bytes are written exactly once: the levels, or zeros (nothing survived / intra plane that stays DC-only)
            const bool keep = nz && !(intra && dc_only);
            uint4 *o4 = (uint4 *)&out->chroma_ac[cl][0];
            o4[0] = keep ? make_uint4(lvw[0], lvw[1], lvw[2], lvw[3]) : make_uint4(0, 0, 0, 0);
            o4[1] = keep ? make_uint4(lvw[4], lvw[5], lvw[6], lvw[7]) : make_uint4(0, 0, 0, 0);
        }
        out->nnz[16 + cl] = (uint8_t)(dc_only ? 0 : nz);
        if (bi == 0) {
            out->nnz[25 + ch] = (uint8_t)nz_dc;
            if (nz_dc) { // zigzag_scan_2x2_dc: level[i] = dct[x][y]
                out->chroma_dc[ch][0] = (int16_t)dc[0]; out->chroma_dc[ch][1] = (int16_t)dc[2];
                out->chroma_dc[ch][2] = (int16_t)dc[1]; out->chroma_dc[ch][3] = (int16_t)dc[3];
            }
        }
        if (dc_only) {
            if (nz_dc) { // add8x8_idct_dc: block bi gets dct[bi>>1][bi&1] == o4[bi]
                const int v = s16((o4[bi] + 32) >> 6);
#pragma unroll
                for (int k = 0; k < 16; k++) p[k] = clip_u8(p[k] + v);
                store4x4(dst, stride_c, p);
            } else if (intra)
                store4x4(dst, stride_c, p);
        } else {
            if (nz_dc) c[0] = o4[bi]; // idct_dequant_2x2_dc -> dct4x4[bi][0][0]
            int r[16];
            inv4x4(c, r);
#pragma unroll
            for (int k = 0; k < 16; k++) p[k] = clip_u8(p[k] + r[k]);
            store4x4(dst, stride_c, p);
        }
    }
    const int ac_u = __shfl_sync(FULL, (int)!dc_only, 16), ac_v = __shfl_sync(FULL, (int)!dc_only, 20);
    const int dcn_u = __shfl_sync(FULL, nz_dc, 16), dcn_v = __shfl_sync(FULL, nz_dc, 20);
    return (ac_u | ac_v) ? 2 : (dcn_u | dcn_v) ? 1 : 0;
}

// ALLOW8 = false is the variant for batches without 8x8-transform macroblocks: without the 64-coefficient path it needs half the
// registers, i.e. twice the resident warps for a kernel that is bound by instruction latency
template <bool ALLOW8>
__global__ void __launch_bounds__(128, ALLOW8 ? 3 : 8) residual_inter_kernel(const QuantTables *__restrict__ qt, FrameRefs fr,
                                                             const x264_cuda_resid_job_t *__restrict__ jobs, int n_jobs,
                                                             x264_cuda_mb_coeffs_t *__restrict__ outs)
{
    const int lane = threadIdx.x & 31;
    const int jb = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (jb >= n_jobs) return;
    const x264_cuda_resid_job_t job = jobs[jb];
    x264_cuda_mb_coeffs_t *out = outs + jb;
    const int qp = min((int)job.qp, 51), cqp = min((int)job.chroma_qp, 51);
    const bool dct8 = ALLOW8 && (job.flags & X264_CUDA_RESID_8x8DCT), decim = job.flags & X264_CUDA_RESID_DECIMATE;
    const unsigned FULL = 0xffffffffu;

    // Uncoded blocks read as zero, like a cleared h->dct.  The 4x4 paths write every block's 32 bytes exactly once (levels or zeros), so only
    // the tail of the record (chroma_dc, nnz, cbp: 48 bytes of scattered byte stores) is cleared up front; the 8x8 path clears the luma too.
    {
        constexpr int tail0 = offsetof(x264_cuda_mb_coeffs_t, chroma_dc) / 4, words = sizeof(x264_cuda_mb_coeffs_t) / 4;
        for (int i = (dct8 ? 0 : tail0) + lane; i < words; i += 32)
            if (i < (int)(offsetof(x264_cuda_mb_coeffs_t, chroma_ac) / 4) || i >= tail0) ((uint32_t *)out)[i] = 0;
    }
    __syncwarp();

    int cbp_luma = 0;
    if (!dct8) {
        // ------------------------------------------------------------------ 4x4 transform: luma AND chroma in one instruction stream
        // (macroblock.c:678-742 and :272-363).  Lanes 0..15 hold the luma blocks, lanes 16..23 the chroma blocks; transform, quantisation,
        // zig-zag, decimation score and dequantisation are the same code with a per-lane table list / qp, so the warp runs them once instead
        // of once per plane type (the kernel is bound by instruction issue).
        static_assert(offsetof(x264_cuda_mb_coeffs_t, chroma_ac) == 16 * 16 * sizeof(int16_t), "chroma AC levels follow the luma levels");
        const bool isl = lane < 16, act = lane < 24;
        const int cl = lane - 16, ch = (cl >> 2) & 1, bi = cl & 3;
        const int bx = (lane & 1) + ((lane >> 2) & 1) * 2, by = ((lane >> 1) & 1) + ((lane >> 3) & 1) * 2; // block_idx_x/y
        const int st = isl ? fr.stride : fr.stride_c;
        const size_t off = isl ? ((size_t)job.mb_y * 16 + by * 4) * fr.stride + job.mb_x * 16 + bx * 4
                               : ((size_t)job.mb_y * 8 + (bi >> 1) * 4) * fr.stride_c + job.mb_x * 8 + (bi & 1) * 4;
        const uint8_t *src = (isl ? fr.fe_y : ch ? fr.fe_v : fr.fe_u) + off;
        uint8_t *dst = (isl ? fr.fd_y : ch ? fr.fd_v : fr.fd_u) + off;
        int c[16], p[16], nz = 0, score = 0, dc0 = 0;
        if (act) {
            int f[16], d[16];
            uint32_t lvw[8];
            load4x4(src, st, f);
            load4x4(dst, st, p);
#pragma unroll
            for (int k = 0; k < 16; k++) d[k] = f[k] - p[k];
            fwd4x4(d, c);
            if (!isl) { dc0 = c[0]; c[0] = 0; } // dct2x2dc takes the chroma DCs out (macroblock.c:72-85)
            nz = quant_block4(qt, isl ? 1 /* CQM_4PY */ : 3 /* CQM_4PC */, isl ? qp : cqp, c, decim, isl ? 0 : 1, lvw, score);
            uint4 *o4 = (uint4 *)&out->luma[lane * 16]; // lanes 16..23 land in chroma_ac[0..7]; all zero when nothing survived
            o4[0] = make_uint4(lvw[0], lvw[1], lvw[2], lvw[3]); o4[1] = make_uint4(lvw[4], lvw[5], lvw[6], lvw[7]);
        }
        // ---- luma decisions.  Per 8x8: i_decimate_8x8 accumulates the scores of its blocks in order while it is still < 6 (:704-705)
        const int base = lane & ~3;
        int dec8 = 0, any = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int s = __shfl_sync(FULL, score, base + k), z = __shfl_sync(FULL, nz, base + k);
            if (z && dec8 < 6) dec8 += s;
            any |= z;
        }
        const int keep8 = decim ? (dec8 >= 4) : any; // this 8x8's cbp bit before the MB-level test
        int dec_mb = 0, cbp = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            dec_mb += __shfl_sync(FULL, dec8, 4 * k);
            cbp |= __shfl_sync(FULL, keep8, 4 * k) << k;
        }
        if (decim && dec_mb < 6) cbp = 0;
        cbp_luma = cbp;
        // ---- chroma decisions (x264_mb_encode_8x8_chroma with b_inter = 1): the four DCs / scores / nz of this lane's plane
        const int cb = 16 + ch * 4;
        const int b0 = __shfl_sync(FULL, dc0, cb), b1 = __shfl_sync(FULL, dc0, cb + 1), b2 = __shfl_sync(FULL, dc0, cb + 2), b3 = __shfl_sync(FULL, dc0, cb + 3);
        int tot = 0, nz_ac = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) { tot += __shfl_sync(FULL, score, cb + k); nz_ac |= __shfl_sync(FULL, nz, cb + k); }
        // dct2x2dc: d[0][0], d[1][0], d[0][1], d[1][1] (macroblock.c:72-80), flat order d[0][0],d[0][1],d[1][0],d[1][1]
        const int e0 = b0 + b1, e1 = b2 + b3, e2 = b0 - b1, e3 = b2 - b3;
        int dc[4] = { s16(e0 + e1), s16(e0 - e1), s16(e2 + e3), s16(e2 - e3) };
        const int mf0 = qt->q4mf[3][cqp][0] >> 1, bias0 = qt->q4bias[3][cqp][0] << 1;
        int nz_dc = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) { dc[k] = quant1(dc[k], mf0, bias0); nz_dc |= dc[k]; }
        nz_dc = nz_dc != 0;
        // IDCT_DEQUANT_START (macroblock.c:42-53)
        const int g0 = dc[0] + dc[1], g1 = dc[2] + dc[3], g2 = dc[0] - dc[1], g3 = dc[2] - dc[3];
        int dmf = qt->dq4[3][cqp % 6][0], qbits = cqp / 6 - 5;
        if (qbits > 0) { dmf <<= qbits; qbits = 0; }
        const int o4v[4] = { s16((g0 + g1) * dmf >> -qbits), s16((g0 - g1) * dmf >> -qbits), s16((g2 + g3) * dmf >> -qbits), s16((g2 - g3) * dmf >> -qbits) };
        const bool dc_only = (decim && tot < 7) || !nz_ac;
        // ---- bookkeeping bytes
        if (isl) {
            const int coded = (cbp >> (lane >> 2)) & 1;
            // nnz: the quant result, cleared for decimated 8x8s / decimated MBs (STORE_8x8_NNZ, :717; :731-735)
            out->nnz[lane] = (uint8_t)((decim ? coded : 1) && nz);
        } else if (act) {
            out->nnz[16 + cl] = (uint8_t)(dc_only ? 0 : nz);
            if (bi == 0) {
                out->nnz[25 + ch] = (uint8_t)nz_dc;
                if (nz_dc) { // zigzag_scan_2x2_dc: level[i] = dct[x][y]
                    out->chroma_dc[ch][0] = (int16_t)dc[0]; out->chroma_dc[ch][1] = (int16_t)dc[2];
                    out->chroma_dc[ch][2] = (int16_t)dc[1]; out->chroma_dc[ch][3] = (int16_t)dc[3];
                }
            }
        }
        // ---- reconstruction, again one stream: luma adds the inverse transform of every surviving block of a coded 8x8 (add8x8_idct;
        // all-zero blocks add nothing); chroma either the DC-only shortcut (add8x8_idct_dc) or the full block with the dequantised DC put back
        bool inv = false;
        int dcv = 0;
        if (isl) inv = ((cbp >> (lane >> 2)) & 1) && nz;
        else if (act) {
            if (dc_only) dcv = nz_dc ? s16((o4v[bi] + 32) >> 6) : 0; // block bi gets dct[bi>>1][bi&1] == o4v[bi]
            else { inv = true; if (nz_dc) c[0] = o4v[bi]; }        // idct_dequant_2x2_dc -> dct4x4[bi][0][0]
        }
        if (inv) {
            int r[16];
            inv4x4(c, r);
#pragma unroll
            for (int k = 0; k < 16; k++) p[k] = clip_u8(p[k] + r[k]);
            store4x4(dst, st, p);
        } else if (dcv) {
#pragma unroll
            for (int k = 0; k < 16; k++) p[k] = clip_u8(p[k] + dcv);
            store4x4(dst, st, p);
        }
        const int ac_u = __shfl_sync(FULL, (int)!dc_only, 16), ac_v = __shfl_sync(FULL, (int)!dc_only, 20);
        const int dcn_u = __shfl_sync(FULL, nz_dc, 16), dcn_v = __shfl_sync(FULL, nz_dc, 20);
        if (lane == 0) {
            out->cbp_luma = (uint8_t)cbp_luma;
            out->cbp_chroma = (uint8_t)((ac_u | ac_v) ? 2 : (dcn_u | dcn_v) ? 1 : 0);
        }
        return;
    } else { // macroblock.c:627-677
        int nz = 0, score = 0;
        int c[64];
        uint8_t *dst = fr.fd_y + ((size_t)job.mb_y * 16 + (lane >> 1) * 8) * fr.stride + job.mb_x * 16 + (lane & 1) * 8;
        if (lane < 4) {
            int d[64];
            const uint8_t *src = fr.fe_y + ((size_t)job.mb_y * 16 + (lane >> 1) * 8) * fr.stride + job.mb_x * 16 + (lane & 1) * 8;
            for (int y = 0; y < 8; y++)
                for (int x = 0; x < 8; x++) d[y * 8 + x] = (int)src[(size_t)y * fr.stride + x] - (int)dst[(size_t)y * fr.stride + x];
            fwd8x8(d, c);
            const uint16_t *mf = qt->q8mf[1][qp], *bias = qt->q8bias[1][qp]; // CQM_8PY
            for (int k = 0; k < 64; k++) { c[k] = quant1(c[k], mf[k], bias[k]); nz |= c[k]; }
            nz = nz != 0;
            if (nz) {
                int16_t lv[64];
                for (int k = 0; k < 64; k++) lv[k] = (int16_t)c[c_zz8[k]];
                for (int k = 0; k < 64; k++) out->luma[lane * 64 + k] = lv[k];
                if (decim) score = decimate_score(lv, 0, 64);
            }
        }
        int dec_mb = 0, cbp = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int s = __shfl_sync(FULL, score, k), z = __shfl_sync(FULL, nz, k);
            if (z) { dec_mb += decim ? s : 0; if (!decim || s >= 4) cbp |= 1 << k; }
        }
        if (decim && dec_mb < 6) cbp = 0;
        cbp_luma = cbp;
        if (lane < 4 && ((cbp >> lane) & 1)) {
            int r[64];
            const int *dmf = qt->dq8[1][qp % 6];
            const int qbits = qp / 6 - 6;
            for (int k = 0; k < 64; k++) c[k] = dequant1(c[k], dmf[k], qbits);
            inv8x8(c, r);
            for (int y = 0; y < 8; y++)
                for (int x = 0; x < 8; x++) dst[(size_t)y * fr.stride + x] = (uint8_t)clip_u8((int)dst[(size_t)y * fr.stride + x] + r[y * 8 + x]);
        }
        if (lane < 16) out->nnz[lane] = (uint8_t)((cbp >> (lane >> 2)) & 1); // STORE_8x8_NNZ
    }

    // ------------------------------------------------------------------ chroma (b_inter = 1), macroblock.c:272-363
    {
        const int cl = lane - 16;                 // 0..7 on the chroma lanes
        const bool mine = lane >= 16 && lane < 24;
        const int ch = (cl >> 2) & 1, bi = cl & 3;
        int f[16], p[16];
        const uint8_t *fe = (ch ? fr.fe_v : fr.fe_u) + ((size_t)job.mb_y * 8 + (bi >> 1) * 4) * fr.stride_c + job.mb_x * 8 + (bi & 1) * 4;
        uint8_t *dst = (ch ? fr.fd_v : fr.fd_u) + ((size_t)job.mb_y * 8 + (bi >> 1) * 4) * fr.stride_c + job.mb_x * 8 + (bi & 1) * 4;
        if (mine) { load4x4(fe, fr.stride_c, f); load4x4(dst, fr.stride_c, p); }
        const int cbp_chroma = chroma_blocks(qt, 3 /* CQM_4PC */, cqp, decim, false, lane, f, p, dst, fr.stride_c, out);
        if (lane == 0) {
            out->cbp_luma = (uint8_t)cbp_luma;
            out->cbp_chroma = (uint8_t)cbp_chroma;
        }
    }
}

} // namespace

// ------------------------------------------------------------------------------------------------------------
extern "C" int x264_cuda_set_quant_tables(x264_cuda_t *ctx, const uint16_t *const q4mf[4], const uint16_t *const q4bias[4],
                                          const int *const dq4[4], const uint16_t *const q8mf[2], const uint16_t *const q8bias[2],
                                          const int *const dq8[2])
{
    x264_cuda_enter(ctx);
    QuantTables *h = (QuantTables *)calloc(1, sizeof(QuantTables));
    for (int l = 0; l < 4; l++) {
        memcpy(h->q4mf[l], q4mf[l], sizeof(h->q4mf[l]));
        memcpy(h->q4bias[l], q4bias[l], sizeof(h->q4bias[l]));
        memcpy(h->dq4[l], dq4[l], sizeof(h->dq4[l]));
    }
    ctx->have_qt8 = q8mf && q8mf[0] && q8bias && dq8;
    if (ctx->have_qt8)
        for (int l = 0; l < 2; l++) {
            memcpy(h->q8mf[l], q8mf[l], sizeof(h->q8mf[l]));
            memcpy(h->q8bias[l], q8bias[l], sizeof(h->q8bias[l]));
            memcpy(h->dq8[l], dq8[l], sizeof(h->dq8[l]));
        }
    cudaError_t e = cudaSuccess;
    if (!ctx->d_qt) e = cudaMalloc(&ctx->d_qt, sizeof(QuantTables));
    if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->d_qt, h, sizeof(QuantTables), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    free(h);
    if (e != cudaSuccess) return x264_cuda_fail(ctx, "x264_cuda_set_quant_tables", e);
    return 0;
}

extern "C" int x264_cuda_set_quant_preset(x264_cuda_t *ctx, int cqm_preset)
{
    x264_cuda_enter(ctx);
    QuantTables *h = (QuantTables *)calloc(1, sizeof(QuantTables));
    x264_cuda_host_cqm_tables(cqm_preset, h->q4mf, h->q4bias, h->dq4, h->q8mf, h->q8bias, h->dq8);
    const uint16_t *a[4], *b[4], *e[2], *f[2];
    const int *c[4], *g[2];
    for (int l = 0; l < 4; l++) { a[l] = &h->q4mf[l][0][0]; b[l] = &h->q4bias[l][0][0]; c[l] = &h->dq4[l][0][0]; }
    for (int l = 0; l < 2; l++) { e[l] = &h->q8mf[l][0][0]; f[l] = &h->q8bias[l][0][0]; g[l] = &h->dq8[l][0][0]; }
    int rc = x264_cuda_set_quant_tables(ctx, a, b, c, e, f, g);
    free(h);
    return rc;
}

static int need_tables(x264_cuda_t *ctx, bool want8)
{
    if (!ctx->d_qt || (want8 && !ctx->have_qt8)) {
        snprintf(ctx->err, 256, "x264_cuda: quantiser tables not set (x264_cuda_set_quant_tables)%s", want8 ? " for the 8x8 transform" : "");
        return -1;
    }
    return 0;
}

extern "C" int x264_cuda_block_residual(x264_cuda_t *ctx, int kind, int n, const uint8_t *fenc, const uint8_t *pred, const uint8_t *qp,
                                        const uint8_t *cat, int16_t *dct_out, int16_t *level_out, uint8_t *nz_out, uint8_t *recon_out)
{
    x264_cuda_enter(ctx);
    if (n <= 0) return 0;
    if (need_tables(ctx, kind == 1)) return -1;
    const size_t bs = kind ? 64 : 16;
    // device layout: fenc | pred | qp | cat | dct | level | nz | recon
    size_t off[9], sz[8] = { n * bs, n * bs, (size_t)n, (size_t)n, n * bs * 2, n * bs * 2, (size_t)n, n * bs };
    off[0] = 0;
    for (int i = 0; i < 8; i++) off[i + 1] = (off[i] + sz[i] + 255) & ~(size_t)255;
    if (x264_cuda_stage(ctx, off[8], off[8])) return -1;
    uint8_t *hs = (uint8_t *)ctx->h_stage, *ds = (uint8_t *)ctx->d_stage;
    memcpy(hs + off[0], fenc, sz[0]); memcpy(hs + off[1], pred, sz[1]); memcpy(hs + off[2], qp, sz[2]); memcpy(hs + off[3], cat, sz[3]);
    CUDA_TRY(ctx, cudaMemcpyAsync(ds, hs, off[4], cudaMemcpyHostToDevice, ctx->stream));
    if (kind == 0)
        block_residual4_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_qt, n, ds + off[0], ds + off[1], ds + off[2], ds + off[3],
                                                                         (int16_t *)(ds + off[4]), (int16_t *)(ds + off[5]), ds + off[6], ds + off[7]);
    else
        block_residual8_kernel<<<(n + 63) / 64, 64, 0, ctx->stream>>>(ctx->d_qt, n, ds + off[0], ds + off[1], ds + off[2], ds + off[3],
                                                                      (int16_t *)(ds + off[4]), (int16_t *)(ds + off[5]), ds + off[6], ds + off[7]);
    LAUNCH_CHECK(ctx, "block_residual_kernel");
    CUDA_TRY(ctx, cudaMemcpyAsync(hs + off[4], ds + off[4], off[8] - off[4], cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (dct_out) memcpy(dct_out, hs + off[4], sz[4]);
    if (level_out) memcpy(level_out, hs + off[5], sz[5]);
    if (nz_out) memcpy(nz_out, hs + off[6], sz[6]);
    if (recon_out) memcpy(recon_out, hs + off[7], sz[7]);
    return 0;
}

extern "C" int x264_cuda_block_dc(x264_cuda_t *ctx, int n, const int16_t *dc_in, const uint8_t *qp, const uint8_t *cat, int16_t *fwd_out,
                                  int16_t *level_out, uint8_t *nz_out, int16_t *deq_out)
{
    x264_cuda_enter(ctx);
    if (n <= 0) return 0;
    if (need_tables(ctx, false)) return -1;
    size_t off[8], sz[7] = { (size_t)n * 32, (size_t)n, (size_t)n, (size_t)n * 32, (size_t)n * 32, (size_t)n, (size_t)n * 32 };
    off[0] = 0;
    for (int i = 0; i < 7; i++) off[i + 1] = (off[i] + sz[i] + 255) & ~(size_t)255;
    if (x264_cuda_stage(ctx, off[7], off[7])) return -1;
    uint8_t *hs = (uint8_t *)ctx->h_stage, *ds = (uint8_t *)ctx->d_stage;
    memcpy(hs + off[0], dc_in, sz[0]); memcpy(hs + off[1], qp, sz[1]); memcpy(hs + off[2], cat, sz[2]);
    CUDA_TRY(ctx, cudaMemcpyAsync(ds, hs, off[3], cudaMemcpyHostToDevice, ctx->stream));
    block_dc_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_qt, n, (const int16_t *)(ds + off[0]), ds + off[1], ds + off[2],
                                                              (int16_t *)(ds + off[3]), (int16_t *)(ds + off[4]), ds + off[5], (int16_t *)(ds + off[6]));
    LAUNCH_CHECK(ctx, "block_dc_kernel");
    CUDA_TRY(ctx, cudaMemcpyAsync(hs + off[3], ds + off[3], off[7] - off[3], cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (fwd_out) memcpy(fwd_out, hs + off[3], sz[3]);
    if (level_out) memcpy(level_out, hs + off[4], sz[4]);
    if (nz_out) memcpy(nz_out, hs + off[5], sz[5]);
    if (deq_out) memcpy(deq_out, hs + off[6], sz[6]);
    return 0;
}

extern "C" int x264_cuda_residual_inter_dev(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, x264_cuda_frame_t *fdec, const void *d_jobs,
                                            int n_jobs, void *d_coeffs)
{
    x264_cuda_enter(ctx);
    if (n_jobs <= 0) return 0;
    if (need_tables(ctx, true)) return -1;
    if (!fenc->buf_chroma || !fdec->buf_chroma || fenc->g.stride != fdec->g.stride) {
        snprintf(ctx->err, 256, "x264_cuda_residual_inter: frames need X264_CUDA_FRAME_CHROMA and equal geometry");
        return -1;
    }
    FrameRefs fr = { fenc->plane[0], fenc->chroma[0], fenc->chroma[1], fdec->plane[0], fdec->chroma[0], fdec->chroma[1], fenc->g.stride,
                     fenc->stride_c };
    if (ctx->resid_no_dct8)
        residual_inter_kernel<false><<<(n_jobs + 3) / 4, 128, 0, ctx->stream>>>(ctx->d_qt, fr, (const x264_cuda_resid_job_t *)d_jobs, n_jobs,
                                                                                (x264_cuda_mb_coeffs_t *)d_coeffs);
    else
        residual_inter_kernel<true><<<(n_jobs + 3) / 4, 128, 0, ctx->stream>>>(ctx->d_qt, fr, (const x264_cuda_resid_job_t *)d_jobs, n_jobs,
                                                                               (x264_cuda_mb_coeffs_t *)d_coeffs);
    ctx->resid_no_dct8 = 0;
    LAUNCH_CHECK(ctx, "residual_inter_kernel");
    return 0;
}

extern "C" int x264_cuda_residual_inter(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, x264_cuda_frame_t *fdec,
                                        const x264_cuda_resid_job_t *jobs, int n_jobs, x264_cuda_mb_coeffs_t *coeffs)
{
    x264_cuda_enter(ctx);
    if (n_jobs <= 0) return 0;
    const size_t jb = (size_t)n_jobs * sizeof(x264_cuda_resid_job_t), rb = (size_t)n_jobs * sizeof(x264_cuda_mb_coeffs_t);
    const size_t jb_al = (jb + 255) & ~(size_t)255;
    if (x264_cuda_stage(ctx, jb_al + rb, jb_al + rb)) return -1;
    uint8_t *hs = (uint8_t *)ctx->h_stage, *ds = (uint8_t *)ctx->d_stage;
    if (x264_cuda_jobs_in(ctx, ds, jobs, hs, jb)) return -1;
    int any8 = 0; // the job list is host memory here: pick the leaner kernel when no macroblock uses the 8x8 transform
    for (int i = 0; i < n_jobs && !any8; i++) any8 = jobs[i].flags & X264_CUDA_RESID_8x8DCT;
    ctx->resid_no_dct8 = !any8;
    if (x264_cuda_residual_inter_dev(ctx, fenc, fdec, ds, n_jobs, ds + jb_al)) return -1;
    if (x264_cuda_results_out(ctx, coeffs, ds + jb_al, hs + jb_al, rb)) return -1;
    return 0;
}

// =========================================================================================================
// I_16x16 macroblocks through x264_macroblock_encode: predict_16x16[mode] + x264_mb_encode_i16x16 (S/encoder/macroblock.c:512-530,
// :184-270), predict_8x8c[mode] + x264_mb_encode_8x8_chroma(b_inter = 0) (:744-760, :272-363).  The prediction of a macroblock reads the
// RECONSTRUCTION of its left / top / top-left neighbours, so the macroblocks of a call form a wavefront: ONE WARP PER MACROBLOCK, persistent
// warps pull jobs by ticket in list order (the list must name a macroblock after the neighbours it depends on: raster order does), and a
// per-macroblock state word (epoch-stamped: pending / done) tells a warp which neighbours belong to this call and when they are final.
// A waiting warp only ever waits for smaller tickets, which running warps hold, so the scheme cannot deadlock whatever the grid size.
// Lane roles as in residual_inter_kernel: lanes 0..15 the luma 4x4 blocks (block_idx order), lanes 16..23 the chroma blocks; the 4x4 luma
// DC transform (dct4x4dc / quant_4x4_dc / idct4x4dc / dequant_4x4_dc) runs on lane 0 over values gathered by shuffles.
namespace {

// polling load: relaxed (an acquire load invalidates the SM's whole L1 on every poll — CCTL.IVALL — and the warps that are working on the
// same SM then miss on every quantiser-table read: measured 12.5 us per wavefront step).  The acquire is ONE fence after the flag is seen.
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned *p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

__global__ void intra16_mark_kernel(const x264_cuda_intra16_job_t *__restrict__ jobs, int n_jobs, int mb_width, unsigned *state, unsigned epoch)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_jobs) state[jobs[i].mb_y * mb_width + jobs[i].mb_x] = epoch << 1;
}

// predictor kinds of a mode number: luma enum intra16x16_pred_e (V H DC P DC_LEFT DC_TOP DC_128), chroma enum intra_chroma_pred_e
// (DC H V P DC_LEFT DC_TOP DC_128) -> 0 V, 1 H, 2 DC-like, 3 plane
__device__ __forceinline__ int pred_kind(int mode, bool chroma) { return mode == 1 ? 1 : mode == 3 ? 3 : mode == (chroma ? 2 : 0) ? 0 : 2; }

struct Intra16Edges { uint8_t y[40], u[24], v[24]; }; // [3] corner, [4..4+n) row above, [4+n..4+2n) left column

__global__ void __launch_bounds__(128) residual_intra16_kernel(const QuantTables *__restrict__ qt, FrameRefs fr, int mb_width,
                                                               const x264_cuda_intra16_job_t *__restrict__ jobs, int n_jobs,
                                                               x264_cuda_mb_coeffs_i16_t *__restrict__ outs, unsigned *state, unsigned epoch,
                                                               int *ticket, const int *__restrict__ order)
{
    __shared__ __align__(4) Intra16Edges s_edges[4];
    __shared__ __align__(16) int s_dc[4][20];   // dequantised luma DCs + [16] the DC block's nz flag
    __shared__ __align__(16) int s_dcin[4][32]; // every block's DC coefficient, by lane
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned FULL = 0xffffffffu;
    Intra16Edges &E = s_edges[wid];
    for (;;) {
        int jb = 0;
        if (lane == 0) jb = atomicAdd(ticket, 1);
        jb = __shfl_sync(FULL, jb, 0);
        if (jb >= n_jobs) return;
        if (order) jb = __ldg(order + jb); // ticket t works on list entry order[t]: anti-diagonals first, see the host entry
        const x264_cuda_intra16_job_t job = jobs[jb];
        x264_cuda_mb_coeffs_i16_t *out = outs + jb;
        const int qp = min((int)job.qp, 51), cqp = min((int)job.chroma_qp, 51);
        const bool decim = job.flags & X264_CUDA_RESID_DECIMATE;
        const int kind_y = pred_kind(job.mode16, false), kind_c = pred_kind(job.mode_chroma, true);
        // every 4x4 block's levels are written exactly once below (levels or zeros); the tail (chroma_dc, nnz, cbp, luma_dc) is cleared here
        for (int i = (int)(offsetof(x264_cuda_mb_coeffs_t, chroma_dc) / 4) + lane; i < (int)(sizeof(x264_cuda_mb_coeffs_i16_t) / 4); i += 32) ((uint32_t *)out)[i] = 0;
        // source blocks first: they do not depend on anybody
        const int bx = (lane & 1) + ((lane >> 2) & 1) * 2, by = ((lane >> 1) & 1) + ((lane >> 3) & 1) * 2; // block_idx_x/y
        const int cl = lane - 16, ch = (cl >> 2) & 1, bi = cl & 3;
        int f[16];
        if (lane < 16) load4x4(fr.fe_y + ((size_t)job.mb_y * 16 + by * 4) * fr.stride + job.mb_x * 16 + bx * 4, fr.stride, f);
        else if (lane < 24) load4x4((ch ? fr.fe_v : fr.fe_u) + ((size_t)job.mb_y * 8 + (bi >> 1) * 4) * fr.stride_c + job.mb_x * 8 + (bi & 1) * 4, fr.stride_c, f);
        // ---- wait for the neighbours that belong to this call: left, top, top-left (lanes 0..2)
        if (lane < 3) {
            const int nx = job.mb_x - (lane != 1), ny = job.mb_y - (lane != 0);
            if (nx >= 0 && ny >= 0) {
                const unsigned *w = state + ny * mb_width + nx;
                unsigned v = ld_relaxed_u32(w);
                while ((v >> 1) == epoch && !(v & 1)) { __nanosleep(64); v = ld_relaxed_u32(w); }
            }
            asm volatile("fence.acq_rel.gpu;" ::: "memory"); // acquire side of the neighbours' release stores (not __threadfence(): that is fence.sc)
        }
        __syncwarp();
        // ---- neighbour pixels straight from L2 (another SM wrote them a moment ago)
        {
            const uint8_t *py = fr.fd_y + (size_t)job.mb_y * 16 * fr.stride + job.mb_x * 16;
            if (lane < 16) { E.y[4 + lane] = __ldcg(py - (ptrdiff_t)fr.stride + lane); E.y[20 + lane] = __ldcg(py + (ptrdiff_t)lane * fr.stride - 1); }
            if (lane == 16) E.y[3] = __ldcg(py - (ptrdiff_t)fr.stride - 1);
            const size_t co = (size_t)job.mb_y * 8 * fr.stride_c + job.mb_x * 8;
            if (lane >= 16) {
                const int k = lane & 7;
                const uint8_t *pc = ((lane & 8) ? fr.fd_v : fr.fd_u) + co;
                uint8_t *e = (lane & 8) ? E.v : E.u;
                e[4 + k] = __ldcg(pc - (ptrdiff_t)fr.stride_c + k); e[12 + k] = __ldcg(pc + (ptrdiff_t)k * fr.stride_c - 1);
                if (k == 0) e[3] = __ldcg(pc - (ptrdiff_t)fr.stride_c - 1);
            }
        }
        __syncwarp();
        // ---- prediction of this lane's 4x4 block
        int p[16];
        if (lane < 16) { // predict.c:40-170
            const uint8_t *e = E.y;
            const int x0 = bx * 4, y0 = by * 4;
            if (kind_y == 0) {
#pragma unroll
                for (int k = 0; k < 16; k++) p[k] = e[4 + x0 + (k & 3)];
            } else if (kind_y == 1) {
#pragma unroll
                for (int k = 0; k < 16; k++) p[k] = e[20 + y0 + (k >> 2)];
            } else if (kind_y == 2) {
                int st = 0, sl = 0;
#pragma unroll
                for (int i = 0; i < 16; i++) { st += e[4 + i]; sl += e[20 + i]; }
                const int m = job.mode16;
                const int dc = m == 2 ? (st + sl + 16) >> 5 : m == 4 ? (sl + 8) >> 4 : m == 5 ? (st + 8) >> 4 : 128;
#pragma unroll
                for (int k = 0; k < 16; k++) p[k] = dc;
            } else {
                int H = 0, V = 0;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    H += (i + 1) * ((int)e[4 + 8 + i] - (int)(6 - i >= 0 ? e[4 + 6 - i] : e[3]));
                    V += (i + 1) * ((int)e[20 + 8 + i] - (int)(6 - i >= 0 ? e[20 + 6 - i] : e[3]));
                }
                const int a = 16 * (e[20 + 15] + e[4 + 15]), b = (5 * H + 32) >> 6, c = (5 * V + 32) >> 6, i00 = a - 7 * (b + c) + 16;
#pragma unroll
                for (int k = 0; k < 16; k++) p[k] = clip_u8((i00 + b * (x0 + (k & 3)) + c * (y0 + (k >> 2))) >> 5);
            }
        } else if (lane < 24) { // predict.c:172-336
            const uint8_t *e = ch ? E.v : E.u;
            const int x0 = (bi & 1) * 4, y0 = (bi >> 1) * 4;
            if (kind_c == 0) {
#pragma unroll
                for (int k = 0; k < 16; k++) p[k] = e[4 + x0 + (k & 3)];
            } else if (kind_c == 1) {
#pragma unroll
                for (int k = 0; k < 16; k++) p[k] = e[12 + y0 + (k >> 2)];
            } else if (kind_c == 2) {
                int s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
                for (int i = 0; i < 4; i++) { s0 += e[4 + i]; s1 += e[8 + i]; s2 += e[12 + i]; s3 += e[16 + i]; }
                int dcq[4] = { 128, 128, 128, 128 };
                const int m = job.mode_chroma;
                if (m == 0) { dcq[0] = (s0 + s2 + 4) >> 3; dcq[1] = (s1 + 2) >> 2; dcq[2] = (s3 + 2) >> 2; dcq[3] = (s1 + s3 + 4) >> 3; } // predict.c:234-277
                else if (m == 4) { dcq[0] = dcq[1] = (s2 + 2) >> 2; dcq[2] = dcq[3] = (s3 + 2) >> 2; }                                // :184-212
                else if (m == 5) { dcq[0] = dcq[2] = (s0 + 2) >> 2; dcq[1] = dcq[3] = (s1 + 2) >> 2; }                                // :213-233
                const int dc = dcq[bi];
#pragma unroll
                for (int k = 0; k < 16; k++) p[k] = dc;
            } else {
                int H = 0, V = 0;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    H += (i + 1) * ((int)e[4 + 4 + i] - (int)(2 - i >= 0 ? e[4 + 2 - i] : e[3]));
                    V += (i + 1) * ((int)e[12 + 4 + i] - (int)(2 - i >= 0 ? e[12 + 2 - i] : e[3]));
                }
                const int a = 16 * (e[12 + 7] + e[4 + 7]), b = (17 * H + 16) >> 5, c = (17 * V + 16) >> 5, i00 = a - 3 * (b + c) + 16;
#pragma unroll
                for (int k = 0; k < 16; k++) p[k] = clip_u8((i00 + b * (x0 + (k & 3)) + c * (y0 + (k >> 2))) >> 5);
            }
        }
        // ---- transform + quantisation of all 24 blocks in ONE instruction stream (luma x264_mb_encode_i16x16 :215-233 with CQM_4IY and the
        // DC taken out, chroma x264_mb_encode_8x8_chroma :305-323 with CQM_4IC, b_inter = 0: no decimation): a macroblock is one step of the
        // wavefront's critical path, so its instruction count is the frame's latency
        const bool isl = lane < 16, act = lane < 24;
        int c[16], nz = 0, score = 0, dc0 = 0;
        uint32_t lvw[8];
        if (act) {
            int d[16];
#pragma unroll
            for (int k = 0; k < 16; k++) d[k] = f[k] - p[k];
            fwd4x4(d, c);
            dc0 = c[0]; c[0] = 0;                                                   // :218-220 / dct2x2dc :72-85
            nz = quant_block4(qt, isl ? 0 /* CQM_4IY */ : 2 /* CQM_4IC */, isl ? qp : cqp, c, isl && decim, 1, lvw, score); // decimate_score15, :230
        }
        // Cross-lane traffic goes through one vote, one REDUX and shared memory: inside this persistent loop every __shfl_sync compiles to a
        // convergence call + SHFL, and the ~40 of them were half of a step's working time in the ncu source view.
        const unsigned nzb = __ballot_sync(FULL, act && nz);                        // bit l: lane l's block kept coefficients
        if (act) s_dcin[wid][lane] = dc0;
        // luma: the running "if (decimate_score < 6) decimate_score += ..." of :230 ends below 6 exactly when the total does (scores are >= 0)
        const int tot = __reduce_add_sync(FULL, isl ? score : 0);
        const int cbp_luma = ((nzb & 0xffffu) && !(decim && tot < 6)) ? 0xf : 0;     // :232, :238-245
        __syncwarp();
        // luma 4x4 DC block on lane 0: dct_dc4x4[0][block_idx_xy_1d[i]] = dct4x4[i][0][0]
        if (lane == 0) {
            int dcs[16], nzd = 0;
#pragma unroll
            for (int i = 0; i < 16; i++) dcs[((i & 1) + ((i >> 2) & 1) * 2) + 4 * (((i >> 1) & 1) + ((i >> 3) & 1) * 2)] = s_dcin[wid][i];
            hadamard_dc(dcs, true);                                                 // dct4x4dc, :247
            const int mf0 = qt->q4mf[0][qp][0] >> 1, bias0 = qt->q4bias[0][qp][0] << 1; // :251
#pragma unroll
            for (int k = 0; k < 16; k++) { dcs[k] = quant1(dcs[k], mf0, bias0); nzd |= dcs[k]; }
            nzd = nzd != 0;
            if (nzd) {
                uint32_t w[8];
#pragma unroll
                for (int k = 0; k < 8; k++) w[k] = (uint32_t)(uint16_t)dcs[zz4(2 * k)] | (uint32_t)(uint16_t)dcs[zz4(2 * k + 1)] << 16; // zigzag scan_4x4, :256
                uint4 *o4 = (uint4 *)out->luma_dc;
                o4[0] = make_uint4(w[0], w[1], w[2], w[3]); o4[1] = make_uint4(w[4], w[5], w[6], w[7]);
                hadamard_dc(dcs, false);                                            // idct4x4dc, :259
                const int qbits = qp / 6 - 6, dmf0 = qt->dq4[0][qp % 6][0];         // dequant_4x4_dc, quant.c:148-178
#pragma unroll
                for (int k = 0; k < 16; k++)
                    dcs[k] = qbits >= 0 ? s16(dcs[k] * (dmf0 << qbits)) : s16((dcs[k] * dmf0 + (1 << (-qbits - 1))) >> (-qbits));
            }
#pragma unroll
            for (int k = 0; k < 16; k++) s_dc[wid][k] = dcs[k];
            s_dc[wid][16] = nzd;
        }
        // chroma 2x2 DC of this lane's plane (every lane computes it; only lanes 16..23 use it): dct2x2dc, quant_2x2_dc, IDCT_DEQUANT_START
        const int cb = 16 + ch * 4;
        const int4 bq = *(const int4 *)&s_dcin[wid][cb];
        const int b0 = bq.x, b1 = bq.y, b2 = bq.z, b3 = bq.w;
        const int nz_ac = (nzb >> cb) & 0xf;
        const int e0 = b0 + b1, e1 = b2 + b3, e2 = b0 - b1, e3 = b2 - b3;
        int cdc[4] = { s16(e0 + e1), s16(e0 - e1), s16(e2 + e3), s16(e2 - e3) };
        const int cmf0 = qt->q4mf[2][cqp][0] >> 1, cbias0 = qt->q4bias[2][cqp][0] << 1;
        int nz_cdc = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) { cdc[k] = quant1(cdc[k], cmf0, cbias0); nz_cdc |= cdc[k]; }
        nz_cdc = nz_cdc != 0;
        const int g0 = cdc[0] + cdc[1], g1 = cdc[2] + cdc[3], g2 = cdc[0] - cdc[1], g3 = cdc[2] - cdc[3];
        int cdmf = qt->dq4[2][cqp % 6][0], cqbits = cqp / 6 - 5;
        if (cqbits > 0) { cdmf <<= cqbits; cqbits = 0; }
        const int o4v[4] = { s16((g0 + g1) * cdmf >> -cqbits), s16((g0 - g1) * cdmf >> -cqbits), s16((g2 + g3) * cdmf >> -cqbits), s16((g2 - g3) * cdmf >> -cqbits) };
        const bool dc_only = !nz_ac;                                                // :334 with b_decimate = 0
        __syncwarp();
        const int nz_dc = s_dc[wid][16];
        // ---- levels (each block's 32 bytes once: lanes 16..23 land in chroma_ac), bookkeeping bytes, reconstruction — one stream again
        bool inv = false;
        int dcv = 0;
        if (act) {
            const bool keep = isl ? (cbp_luma && nz) : (nz && !dc_only);
            uint4 *o4 = (uint4 *)&out->c.luma[lane * 16];
            o4[0] = keep ? make_uint4(lvw[0], lvw[1], lvw[2], lvw[3]) : make_uint4(0, 0, 0, 0);
            o4[1] = keep ? make_uint4(lvw[4], lvw[5], lvw[6], lvw[7]) : make_uint4(0, 0, 0, 0);
            if (isl) {
                const int my_dc = s_dc[wid][bx + 4 * by];
                if (cbp_luma) { out->c.nnz[lane] = (uint8_t)nz; inv = true; if (nz_dc) c[0] = my_dc; } // :261-267 add16x16_idct
                else if (nz_dc) dcv = s16((my_dc + 32) >> 6);                                          // :269 add16x16_idct_dc
            } else {
                out->c.nnz[16 + cl] = (uint8_t)(dc_only ? 0 : nz);
                if (bi == 0) {
                    out->c.nnz[25 + ch] = (uint8_t)nz_cdc;
                    if (nz_cdc) { // zigzag_scan_2x2_dc: level[i] = dct[x][y]
                        out->c.chroma_dc[ch][0] = (int16_t)cdc[0]; out->c.chroma_dc[ch][1] = (int16_t)cdc[2];
                        out->c.chroma_dc[ch][2] = (int16_t)cdc[1]; out->c.chroma_dc[ch][3] = (int16_t)cdc[3];
                    }
                }
                if (dc_only) dcv = nz_cdc ? s16((o4v[bi] + 32) >> 6) : 0;            // add8x8_idct_dc
                else { inv = true; if (nz_cdc) c[0] = o4v[bi]; }                    // idct_dequant_2x2_dc + add8x8_idct
            }
            if (inv) {
                int r[16];
                inv4x4(c, r);
#pragma unroll
                for (int k = 0; k < 16; k++) p[k] = clip_u8(p[k] + r[k]);
            } else if (dcv) {
#pragma unroll
                for (int k = 0; k < 16; k++) p[k] = clip_u8(p[k] + dcv);
            }
            // the prediction exists in registers only: every block is stored, changed or not
            uint8_t *dst = isl ? fr.fd_y + ((size_t)job.mb_y * 16 + by * 4) * fr.stride + job.mb_x * 16 + bx * 4
                               : (ch ? fr.fd_v : fr.fd_u) + ((size_t)job.mb_y * 8 + (bi >> 1) * 4) * fr.stride_c + job.mb_x * 8 + (bi & 1) * 4;
            store4x4(dst, isl ? fr.stride : fr.stride_c, p);
        }
        {
            const unsigned ac = __ballot_sync(FULL, act && !isl && !dc_only), dcn = __ballot_sync(FULL, act && !isl && nz_cdc);
            if (lane == 0) {
                out->c.nnz[24] = (uint8_t)nz_dc;
                out->c.cbp_luma = (uint8_t)cbp_luma;
                out->c.cbp_chroma = (uint8_t)(ac ? 2 : dcn ? 1 : 0);
            }
        }
        // ---- publish: every lane's pixel stores happen before lane 0's release store through the warp barrier (causality order is cumulative
        // across bar.warp.sync), so ONE st.release.gpu publishes the macroblock.  A __threadfence() on all lanes here (fence.sc.gpu) was two
        // thirds of the working time of a step in the ncu source view.
        __syncwarp();
        if (lane == 0) st_release_u32(state + job.mb_y * mb_width + job.mb_x, (epoch << 1) | 1);
    }
}

} // namespace

// d_order: NULL (tickets follow the list) or a permutation of 0..n_jobs-1 giving the order in which list entries are taken
static int intra16_launch(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, x264_cuda_frame_t *fdec, const void *d_jobs, int n_jobs, void *d_coeffs,
                          const int *d_order)
{
    x264_cuda_enter(ctx);
    if (n_jobs <= 0) return 0;
    if (need_tables(ctx, false)) return -1;
    if (!fenc->buf_chroma || !fdec->buf_chroma || fenc->g.stride != fdec->g.stride) {
        snprintf(ctx->err, 256, "x264_cuda_residual_intra16: frames need X264_CUDA_FRAME_CHROMA and equal geometry");
        return -1;
    }
    const int n_mb = fenc->g.mb_width * fenc->g.mb_height;
    if (ctx->i16_state_n < n_mb + 1) { // state words of every macroblock + the ticket counter
        cudaFree(ctx->d_i16_state); ctx->d_i16_state = nullptr;
        CUDA_TRY(ctx, cudaMalloc(&ctx->d_i16_state, (size_t)(n_mb + 1) * sizeof(unsigned)));
        CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_i16_state, 0, (size_t)(n_mb + 1) * sizeof(unsigned), ctx->stream));
        ctx->i16_state_n = n_mb + 1; ctx->i16_epoch = 0;
    }
    const unsigned epoch = ++ctx->i16_epoch & 0x7fffffffu; // 31 bits: bit 0 of the word is the done flag
    if (epoch == 0) { CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_i16_state, 0, (size_t)n_mb * sizeof(unsigned), ctx->stream)); return intra16_launch(ctx, fenc, fdec, d_jobs, n_jobs, d_coeffs, d_order); }
    unsigned *state = (unsigned *)ctx->d_i16_state;
    int *ticket = (int *)(state + ctx->i16_state_n - 1);
    CUDA_TRY(ctx, cudaMemsetAsync(ticket, 0, sizeof(int), ctx->stream));
    FrameRefs fr = { fenc->plane[0], fenc->chroma[0], fenc->chroma[1], fdec->plane[0], fdec->chroma[0], fdec->chroma[1], fenc->g.stride,
                     fenc->stride_c };
    intra16_mark_kernel<<<(n_jobs + 255) / 256, 256, 0, ctx->stream>>>((const x264_cuda_intra16_job_t *)d_jobs, n_jobs, fenc->g.mb_width, state, epoch);
    ctx->launches++;
    // an anti-diagonal of a W x H frame holds at most min(W, H) macroblocks: a few warps per SM are plenty
    const int blocks = min((n_jobs + 3) / 4, ctx->sm_count * 2);
    residual_intra16_kernel<<<blocks, 128, 0, ctx->stream>>>(ctx->d_qt, fr, fenc->g.mb_width, (const x264_cuda_intra16_job_t *)d_jobs, n_jobs,
                                                             (x264_cuda_mb_coeffs_i16_t *)d_coeffs, state, epoch, ticket, d_order);
    LAUNCH_CHECK(ctx, "residual_intra16_kernel");
    return 0;
}

extern "C" int x264_cuda_residual_intra16_dev(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, x264_cuda_frame_t *fdec, const void *d_jobs,
                                              int n_jobs, void *d_coeffs)
{
    return intra16_launch(ctx, fenc, fdec, d_jobs, n_jobs, d_coeffs, nullptr);
}

extern "C" int x264_cuda_residual_intra16(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, x264_cuda_frame_t *fdec,
                                          const x264_cuda_intra16_job_t *jobs, int n_jobs, x264_cuda_mb_coeffs_i16_t *coeffs)
{
    x264_cuda_enter(ctx);
    if (n_jobs <= 0) return 0;
    for (int i = 0; i < n_jobs; i++)
        if (jobs[i].mb_x < 0 || jobs[i].mb_x >= fenc->g.mb_width || jobs[i].mb_y < 0 || jobs[i].mb_y >= fenc->g.mb_height || jobs[i].mode16 > 6 ||
            jobs[i].mode_chroma > 6) {
            snprintf(ctx->err, 256, "x264_cuda_residual_intra16: job %d: macroblock (%d,%d) / modes (%d,%d) out of range", i, jobs[i].mb_x, jobs[i].mb_y,
                     jobs[i].mode16, jobs[i].mode_chroma);
            return -1;
        }
    // Tickets are taken in ANTI-DIAGONAL order (x + y ascending, stable): every macroblock of a diagonal has its three neighbours on earlier
    // diagonals, so a whole diagonal is in flight at once.  Taken in raster order only (resident warps) / mb_width rows would ever be
    // active — measured 6.5 ms for a 1080p frame against the diagonal order's W + H steps.  Results stay in the caller's list order.
    const size_t jb = (size_t)n_jobs * sizeof(x264_cuda_intra16_job_t), rb = (size_t)n_jobs * sizeof(x264_cuda_mb_coeffs_i16_t);
    const size_t jb_al = (jb + 255) & ~(size_t)255, ob_al = ((size_t)n_jobs * sizeof(int) + 255) & ~(size_t)255;
    if (x264_cuda_stage(ctx, jb_al + ob_al + rb, jb_al + ob_al + rb)) return -1;
    uint8_t *hs = (uint8_t *)ctx->h_stage, *ds = (uint8_t *)ctx->d_stage;
    {
        int *ord = (int *)(hs + jb_al);
        const int n_diag = fenc->g.mb_width + fenc->g.mb_height;
        int *start = (int *)calloc((size_t)n_diag + 1, sizeof(int));
        if (!start) { snprintf(ctx->err, 256, "x264_cuda_residual_intra16: out of memory"); return -1; }
        for (int i = 0; i < n_jobs; i++) start[jobs[i].mb_x + jobs[i].mb_y + 1]++;
        for (int d = 0; d < n_diag; d++) start[d + 1] += start[d];
        for (int i = 0; i < n_jobs; i++) ord[start[jobs[i].mb_x + jobs[i].mb_y]++] = i; // counting sort: stable within a diagonal
        free(start);
    }
    memcpy(hs, jobs, jb);
    CUDA_TRY(ctx, cudaMemcpyAsync(ds, hs, jb_al + (size_t)n_jobs * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    if (intra16_launch(ctx, fenc, fdec, ds, n_jobs, ds + jb_al + ob_al, (const int *)(ds + jb_al))) return -1;
    if (x264_cuda_results_out(ctx, coeffs, ds + jb_al + ob_al, hs + jb_al + ob_al, rb)) return -1;
    return 0;
}

// =========================================================================================================
// x264_macroblock_probe_skip (S/encoder/macroblock.c:797-883): would this macroblock quantise to nothing against the
// P-skip (or B-direct) prediction?  ONE WARP PER MACROBLOCK with the lane roles of residual_inter_kernel: lanes 0..15
// the luma 4x4 blocks, lanes 16..23 the chroma blocks.  The prediction is formed on the fly from the half-pel planes and
// the 1/8-pel chroma taps (mc_luma / mc_chroma, macroblock.c:816-818, :851-856) or read from fdec (b_bidir = 1).  The
// reference's early exits are sums of non-negative scores, so the warp evaluates every block and compares the totals.
namespace {

struct SkipPlanes {
    const uint8_t *fe_y, *fe_u, *fe_v;
    const uint8_t *ref[4], *ref_u, *ref_v;
    uint8_t *fd_y, *fd_u, *fd_v;
    int stride, stride_c;
    uint16_t ssd_thresh[52]; // (x264_lambda2_tab[chroma qp] + 32) >> 6, macroblock.c:843
};

__global__ void __launch_bounds__(128) probe_skip_kernel(const QuantTables *__restrict__ qt, SkipPlanes pl,
                                                         const x264_cuda_skip_job_t *__restrict__ jobs, int n_jobs, uint8_t *__restrict__ skip)
{
    const int lane = threadIdx.x & 31;
    const int jb = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (jb >= n_jobs) return;
    const x264_cuda_skip_job_t job = jobs[jb];
    const int qp = min((int)job.qp, 51), cqp = min((int)job.chroma_qp, 51);
    const bool in_fdec = job.flags & X264_CUDA_SKIP_PRED_IN_FDEC;
    const bool store = !in_fdec && (job.flags & X264_CUDA_SKIP_STORE_PRED) && pl.fd_y;
    const unsigned FULL = 0xffffffffu;

    // Lanes 0..15 form the luma predictions (qpel fetch), lanes 16..23 the chroma ones (1/8-pel bilinear); the residual, transform,
    // quantisation and decimation score of all 24 blocks then run as ONE instruction stream with a per-lane table list and qp.
    int score = 0, ssd = 0, dc0 = 0;
    int f[16], p[16];
    const bool isl = lane < 16, act = lane < 24;
    if (isl) { // luma, macroblock.c:822-840
        const int bx = (lane & 1) + ((lane >> 2) & 1) * 2, by = ((lane >> 1) & 1) + ((lane >> 3) & 1) * 2;
        const size_t off = ((size_t)job.mb_y * 16 + by * 4) * pl.stride + job.mb_x * 16 + bx * 4;
        load4x4(pl.fe_y + off, pl.stride, f);
        if (in_fdec)
            load4x4(pl.fd_y + off, pl.stride, p);
        else {
            const uint8_t *const planes[4] = { pl.ref[0] + off, pl.ref[1] + off, pl.ref[2] + off, pl.ref[3] + off };
            const QpelSrc src = qpel_src(planes, pl.stride, job.mvx, job.mvy);
#pragma unroll
            for (int y = 0; y < 4; y++) {
                const uint32_t w = qpel_row4(src, (ptrdiff_t)y * pl.stride);
#pragma unroll
                for (int x = 0; x < 4; x++) p[y * 4 + x] = (w >> (8 * x)) & 255;
                if (store) *(uint32_t *)(pl.fd_y + off + (size_t)y * pl.stride) = w;
            }
        }
    } else if (act) { // chroma, macroblock.c:845-879
        const int cl = lane - 16, ch = cl >> 2, bi = cl & 3;
        const size_t off = ((size_t)job.mb_y * 8 + (bi >> 1) * 4) * pl.stride_c + job.mb_x * 8 + (bi & 1) * 4;
        load4x4((ch ? pl.fe_v : pl.fe_u) + off, pl.stride_c, f);
        uint8_t *fd = (ch ? pl.fd_v : pl.fd_u);
        if (in_fdec)
            load4x4(fd + off, pl.stride_c, p);
        else { // mc_chroma, mc.c:205-236
            const int d8x = job.mvx & 7, d8y = job.mvy & 7;
            const int cA = (8 - d8x) * (8 - d8y), cB = d8x * (8 - d8y), cC = (8 - d8x) * d8y, cD = d8x * d8y;
            const uint8_t *s = (ch ? pl.ref_v : pl.ref_u) + off + (ptrdiff_t)(job.mvy >> 3) * pl.stride_c + (job.mvx >> 3);
            int top[5];
#pragma unroll
            for (int x = 0; x < 5; x++) top[x] = s[x];
#pragma unroll
            for (int y = 0; y < 4; y++) {
                int bot[5];
#pragma unroll
                for (int x = 0; x < 5; x++) bot[x] = s[(size_t)(y + 1) * pl.stride_c + x];
#pragma unroll
                for (int x = 0; x < 4; x++) p[y * 4 + x] = (cA * top[x] + cB * top[x + 1] + cC * bot[x] + cD * bot[x + 1] + 32) >> 6;
#pragma unroll
                for (int x = 0; x < 5; x++) top[x] = bot[x];
            }
            if (store) store4x4(fd + off, pl.stride_c, p);
        }
    }
    if (act) {
        int d[16], c[16];
        uint32_t lvw[8];
#pragma unroll
        for (int k = 0; k < 16; k++) { d[k] = f[k] - p[k]; ssd += d[k] * d[k]; } // the SSD only matters on the chroma lanes (:843)
        fwd4x4(d, c);
        if (!isl) { dc0 = c[0]; c[0] = 0; } // dct2x2dc takes the chroma DCs out (macroblock.c:72-85)
        quant_block4(qt, isl ? 1 /* CQM_4PY */ : 3 /* CQM_4PC */, isl ? qp : cqp, c, true, isl ? 0 : 1, lvw, score);
    }
    if (isl) ssd = 0;

    // luma total over lanes 0..15, per-plane chroma totals over lanes 16..19 / 20..23
    int luma = lane < 16 ? score : 0;
#pragma unroll
    for (int o = 8; o; o >>= 1) luma += __shfl_xor_sync(FULL, luma, o);
    luma = __shfl_sync(FULL, luma, 0);
    int ac = score, sq = ssd;
#pragma unroll
    for (int o = 2; o; o >>= 1) { ac += __shfl_xor_sync(FULL, ac, o); sq += __shfl_xor_sync(FULL, sq, o); }
    const int cb = lane & ~3;
    const int b0 = __shfl_sync(FULL, dc0, cb), b1 = __shfl_sync(FULL, dc0, cb + 1), b2 = __shfl_sync(FULL, dc0, cb + 2),
              b3 = __shfl_sync(FULL, dc0, cb + 3);
    // dct2x2dc (macroblock.c:72-80) then quant_2x2_dc with the halved mf / doubled bias (:866)
    const int d0 = b0 + b1, d1 = b2 + b3, d2 = b0 - b1, d3 = b2 - b3;
    const int mfdc = qt->q4mf[3][cqp][0] >> 1, biasdc = qt->q4bias[3][cqp][0] << 1;
    const int nz_dc = quant1(s16(d0 + d1), mfdc, biasdc) | quant1(s16(d0 - d1), mfdc, biasdc) | quant1(s16(d2 + d3), mfdc, biasdc) |
                      quant1(s16(d2 - d3), mfdc, biasdc);
    const bool plane_fails = lane >= 16 && lane < 24 && sq >= (int)pl.ssd_thresh[cqp] && (nz_dc != 0 || ac >= 7);
    const unsigned fails = __ballot_sync(FULL, plane_fails);
    if (lane == 0) skip[jb] = (uint8_t)(luma < 6 && fails == 0);
}
} // namespace

extern "C" int x264_cuda_probe_skip_dev(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, const x264_cuda_frame_t *fref, x264_cuda_frame_t *fdec,
                                        const void *d_jobs, int n_jobs, int any_mc, int any_fdec, void *d_skip)
{
    x264_cuda_enter(ctx);
    if (n_jobs <= 0) return 0;
    if (need_tables(ctx, false)) return -1;
    if (!fenc->buf_chroma) {
        snprintf(ctx->err, 256, "x264_cuda_probe_skip: fenc needs X264_CUDA_FRAME_CHROMA");
        return -1;
    }
    if (any_mc && (!fref || !(fref->g.flags & X264_CUDA_FRAME_HPEL) || !fref->buf_chroma || fref->g.stride != fenc->g.stride)) {
        snprintf(ctx->err, 256, "x264_cuda_probe_skip: the reference needs X264_CUDA_FRAME_HPEL | X264_CUDA_FRAME_CHROMA and fenc's geometry");
        return -1;
    }
    if (any_fdec && (!fdec || !fdec->buf_chroma || fdec->g.stride != fenc->g.stride)) {
        snprintf(ctx->err, 256, "x264_cuda_probe_skip: fdec needs X264_CUDA_FRAME_CHROMA and fenc's geometry");
        return -1;
    }
    SkipPlanes pl;
    memset(&pl, 0, sizeof(pl));
    pl.fe_y = fenc->plane[0]; pl.fe_u = fenc->chroma[0]; pl.fe_v = fenc->chroma[1];
    if (fref) {
        for (int k = 0; k < 4; k++) pl.ref[k] = fref->plane[k];
        pl.ref_u = fref->chroma[0]; pl.ref_v = fref->chroma[1];
    }
    if (fdec && fdec->buf_chroma && fdec->g.stride == fenc->g.stride) { pl.fd_y = fdec->plane[0]; pl.fd_u = fdec->chroma[0]; pl.fd_v = fdec->chroma[1]; }
    pl.stride = fenc->g.stride; pl.stride_c = fenc->stride_c;
    for (int q = 0; q < 52; q++) pl.ssd_thresh[q] = (uint16_t)((x264_cuda_host_lambda2(q) + 32) >> 6);
    probe_skip_kernel<<<(n_jobs + 3) / 4, 128, 0, ctx->stream>>>(ctx->d_qt, pl, (const x264_cuda_skip_job_t *)d_jobs, n_jobs, (uint8_t *)d_skip);
    LAUNCH_CHECK(ctx, "probe_skip_kernel");
    return 0;
}

extern "C" int x264_cuda_probe_skip(x264_cuda_t *ctx, const x264_cuda_frame_t *fenc, const x264_cuda_frame_t *fref, x264_cuda_frame_t *fdec,
                                    const x264_cuda_skip_job_t *jobs, int n_jobs, uint8_t *skip)
{
    x264_cuda_enter(ctx);
    if (n_jobs <= 0) return 0;
    const size_t jb = (size_t)n_jobs * sizeof(x264_cuda_skip_job_t), jb_al = (jb + 255) & ~(size_t)255;
    if (x264_cuda_stage(ctx, jb_al + n_jobs, jb_al + n_jobs)) return -1;
    uint8_t *hs = (uint8_t *)ctx->h_stage, *ds = (uint8_t *)ctx->d_stage;
    int any_mc = 0, any_fdec = 0;
    for (int i = 0; i < n_jobs; i++) {
        if (jobs[i].flags & X264_CUDA_SKIP_PRED_IN_FDEC) any_fdec = 1;
        else { any_mc = 1; any_fdec |= jobs[i].flags & X264_CUDA_SKIP_STORE_PRED; }
    }
    if (x264_cuda_jobs_in(ctx, ds, jobs, hs, jb)) return -1;
    if (x264_cuda_probe_skip_dev(ctx, fenc, fref, fdec, ds, n_jobs, any_mc, any_fdec, ds + jb_al)) return -1;
    if (x264_cuda_results_out(ctx, skip, ds + jb_al, hs + jb_al, (size_t)n_jobs)) return -1;
    return 0;
}
