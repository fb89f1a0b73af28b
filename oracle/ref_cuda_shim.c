/* ref_cuda_shim.c — TEST INFRASTRUCTURE ONLY.  Lets the UNMODIFIED reference encoder run with its four function tables overridden by
 * the CUDA back-end, the way INTEGRATION.md describes, without touching a reference source file: oracle/Makefile compiles
 * S/common/{pixel,mc,dct,quant}.c a second time with -Dx264_<t>_init=x264_<t>_init_c (a rename of the one symbol each file exports for
 * its table), and this file supplies x264_<t>_init: C table first, then the x264_<t>_init_cuda overrides on top
 * (include/x264_cuda_tables.h).  The resulting CLI (oracle/_ref/x264_cuda) must write a byte-identical stream
 * (tests/test_gpu_stream.py): SURVEY 8c "stream level" parity. */
#include <stdio.h>
#include <stdlib.h>
#include "common/common.h"
#include "x264_cuda_tables.h"

void x264_pixel_init_c(int cpu, x264_pixel_function_t *pixf);
void x264_mc_init_c(int cpu, x264_mc_functions_t *pf);
void x264_dct_init_c(int cpu, x264_dct_function_t *dctf);
void x264_quant_init_c(x264_t *h, int cpu, x264_quant_function_t *pf);

/* the mirrors must be layout-identical to the reference's structs */
typedef char chk_pixel[sizeof(x264_cuda_pixel_function_t) == sizeof(x264_pixel_function_t) ? 1 : -1];
typedef char chk_mc[sizeof(x264_cuda_mc_functions_t) == sizeof(x264_mc_functions_t) ? 1 : -1];
typedef char chk_dct[sizeof(x264_cuda_dct_function_t) == sizeof(x264_dct_function_t) ? 1 : -1];
typedef char chk_quant[sizeof(x264_cuda_quant_function_t) == sizeof(x264_quant_function_t) ? 1 : -1];

static void need(int rc, const char *what)
{
    if (rc) { fprintf(stderr, "ref_cuda_shim: %s failed (no CUDA device?)\n", what); exit(3); }
}
static void report(void) { fprintf(stderr, "ref_cuda_shim: %lld device launches\n", x264_cuda_tables_launches()); }
static int enabled(const char *table) /* X264_CUDA_TABLES=pixel,mc,dct,quant (default: all) selects which tables are overridden */
{
    const char *e = getenv("X264_CUDA_TABLES");
    return !e || strstr(e, table);
}

void x264_pixel_init(int cpu, x264_pixel_function_t *pixf)
{
    static int once;
    if (!once++) atexit(report);
    x264_pixel_init_c(cpu, pixf);
    if (enabled("pixel")) need(x264_pixel_init_cuda((x264_cuda_pixel_function_t *)pixf), "x264_pixel_init_cuda");
}
void x264_mc_init(int cpu, x264_mc_functions_t *pf)
{
    x264_mc_init_c(cpu, pf);
    if (enabled("mc")) need(x264_mc_init_cuda((x264_cuda_mc_functions_t *)pf), "x264_mc_init_cuda");
}
void x264_dct_init(int cpu, x264_dct_function_t *dctf)
{
    x264_dct_init_c(cpu, dctf);
    if (enabled("dct")) need(x264_dct_init_cuda((x264_cuda_dct_function_t *)dctf), "x264_dct_init_cuda");
}
void x264_quant_init(x264_t *h, int cpu, x264_quant_function_t *pf)
{
    x264_quant_init_c(h, cpu, pf);
    if (enabled("quant")) need(x264_quant_init_cuda((x264_cuda_quant_function_t *)pf), "x264_quant_init_cuda");
}
