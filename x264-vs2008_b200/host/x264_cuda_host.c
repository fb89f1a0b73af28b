/* x264_cuda_host.c — host-side C companions of the CUDA back-end: things the reference computes on the host
 * and hands to the device as data.  Compiled with the reference's own flags (-O4 -ffast-math). */
#include <math.h>
#include <stdint.h>
#include "../../include/x264_cuda.h"

/* lambda = 2^(qp/6-2), the reference's hand-rounded table (S/encoder/analyse.c:140-148) */
int x264_cuda_host_lambda(int qp)
{
    static const uint8_t lambda_of_qp[52] = {
        1,  1,  1,  1,  1,  1,  1,  1,  1,  1,  1,  1,  1,  1,  1,  1,  2,  2,  2,  2,  3,  3,  3,  4,  4,  4,
        5,  6,  6,  7,  8,  9,  10, 11, 13, 14, 16, 18, 20, 23, 25, 29, 32, 36, 40, 45, 51, 57, 64, 72, 81, 91 };
    return qp < 0 ? 1 : qp > 51 ? 91 : lambda_of_qp[qp];
}

/* p_cost_mv[qp] exactly as x264_mb_analyse_load_costs builds it (S/encoder/analyse.c:40,192-203): the
 * reference's log2f is a macro around double log(), narrowed to float before the division by log(2).
 * In a drop-in integration the reference's own table is uploaded instead (x264_cuda_set_cost_mv). */
void x264_cuda_host_cost_mv(int qp, int16_t *table)
{
    const int lambda = x264_cuda_host_lambda(qp);
    int16_t *centre = table + 2 * 4 * 2048;
    for (int i = 0; i <= 2 * 4 * 2048; i++) {
        float l2 = (float)log((double)(i + 1));
        centre[i] = centre[-i] = lambda * (l2 / (log((double)2)) * 2 + 0.718f + !!i) + .5f;
    }
}
