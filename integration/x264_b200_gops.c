/* x264_b200_gops.c — GOP-parallel front end of the performance-mode encoder: ONE process, T encoder threads, one B200.
 *
 * The reference's sequential macroblock loop is host-bound; one encoder instance keeps the device busy for about a
 * millisecond per frame.  Closed GOPs are independent units of work, so this front end
 * runs T instances of the UNMODIFIED reference encoder (public API of S/x264.h, used exactly as S/x264.c:752-901 uses it) side by side
 * in one process: each thread encodes a run of consecutive GOPs with its own x264_t, its own device context and stream
 * (integration/x264_b200_hooks.c keeps its state per thread), all sharing one CUDA context — the grids, deblocking and half-pel passes
 * of different encoders overlap on the device instead of time-slicing between processes.  The output is the concatenation of the
 * runs in order, minus the version SEI every run but the first starts with: byte for byte the stream ONE `x264 --threads 1` process
 * writes for the same options (tests/test_gpu_encode.py, bench.py).
 *
 *   x264_b200_gops [x264 options] --keyint K [--workers T] [--frames N] -o out.264 in.yuv WIDTHxHEIGHT
 *
 * Options are handed to x264_param_parse (S/common/common.c:206-587); --min-keyint K and --scenecut -1 are implied (a fixed IDR cadence
 * is what makes GOPs independent).  Rate control must be constant-QP (--qp), as for any bit-exact sharding.
 *
 * Two properties of this fork of the reference shape the code: the motion-vector cost tables are process-wide, built lazily without a
 * lock (S/encoder/analyse.c:184-217) — they are built here for all 52 qps before the threads start; and x264_encoder_close frees those
 * process-wide tables (S/encoder/encoder.c:2136-2145), so no instance is closed while another may still run (the process exits instead).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include <stdarg.h>
#include "x264.h"

void x264_b200_set_gop_seed(int idr_pic_id, int coded_frames); /* x264_b200_hooks.c: per-thread seeds of the next x264_encoder_open */
void x264_b200_disable_for_this_thread(void);
void x264_b200_report(void);
void x264_b200_warm_device(void);

typedef struct {
    int index, first_gop, first_frame, n_frames;
    x264_param_t param;
    const char *in_path;
    uint8_t *out; size_t out_size, out_cap;
    int rc;
    pthread_t th;
} worker_t;

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + ts.tv_nsec * 1e-9;
}

static void put_nals(worker_t *w, x264_nal_t *nal, int i_nal)
{
    for (int i = 0; i < i_nal; i++) {
        int size = nal[i].i_payload * 3 / 2 + 64;
        if (w->out_size + size > w->out_cap) {
            w->out_cap = (w->out_size + size) * 2;
            w->out = realloc(w->out, w->out_cap);
        }
        x264_nal_encode(w->out + w->out_size, &size, 1, &nal[i]); /* Annex B, as S/x264.c:775 */
        w->out_size += size;
    }
}

static void *worker_main(void *arg)
{
    worker_t *w = arg;
    x264_picture_t pic, pic_out;
    x264_nal_t *nal;
    int i_nal;
    FILE *f = fopen(w->in_path, "rb");
    const int wd = w->param.i_width, ht = w->param.i_height;
    const size_t frame_bytes = (size_t)wd * ht * 3 / 2;
    if (!f) { w->rc = -1; return NULL; }
    x264_b200_set_gop_seed(w->first_gop & 0xffff, w->first_frame);
    w->param.i_frame_total = w->n_frames;
    x264_t *h = x264_encoder_open(&w->param);
    if (!h) { w->rc = -1; fclose(f); return NULL; }
    x264_picture_alloc(&pic, X264_CSP_I420, wd, ht);
    for (int i = 0; i < w->n_frames && w->rc == 0; i++) {
        if (fseeko(f, (off_t)(w->first_frame + i) * frame_bytes, SEEK_SET) || fread(pic.img.plane[0], 1, (size_t)wd * ht, f) != (size_t)wd * ht ||
            fread(pic.img.plane[1], 1, (size_t)wd * ht / 4, f) != (size_t)wd * ht / 4 || fread(pic.img.plane[2], 1, (size_t)wd * ht / 4, f) != (size_t)wd * ht / 4) {
            w->rc = -2;
            break;
        }
        pic.i_pts = (int64_t)i * w->param.i_fps_den;
        pic.i_type = X264_TYPE_AUTO;
        pic.i_qpplus1 = 0;
        if (x264_encoder_encode(h, &nal, &i_nal, &pic, &pic_out) < 0) { w->rc = -3; break; }
        put_nals(w, nal, i_nal);
    }
    while (w->rc == 0) { /* flush delayed B-frames (S/x264.c:870-874) */
        if (x264_encoder_encode(h, &nal, &i_nal, NULL, &pic_out) < 0) { w->rc = -3; break; }
        if (!i_nal) break;
        put_nals(w, nal, i_nal);
    }
    x264_picture_clean(&pic);
    fclose(f);
    x264_b200_report();
    /* no x264_encoder_close: it would free the cost tables the other instances are using (see the header comment) */
    return NULL;
}

/* build the process-wide cost tables of every qp before any thread can race on them: a two-frame 16x16 encode per qp with the plain
 * reference code (hooks off for this thread) reaches x264_mb_analyse_load_costs for that qp */
static void build_cost_tables(const x264_param_t *base)
{
    x264_b200_disable_for_this_thread();
    for (int qp = 0; qp < 52; qp++) {
        x264_param_t p = *base;
        x264_picture_t pic, out;
        x264_nal_t *nal;
        int i_nal;
        p.i_width = p.i_height = 16;
        p.i_threads = 1;
        p.i_log_level = X264_LOG_NONE;
        p.rc.i_rc_method = X264_RC_CQP; p.rc.i_qp_constant = qp; p.rc.f_ip_factor = p.rc.f_pb_factor = 1.f;
        p.i_bframe = 0; p.i_keyint_max = 250; p.i_keyint_min = 25;
        p.analyse.b_psnr = p.analyse.b_ssim = 0;
        x264_t *h = x264_encoder_open(&p);
        if (!h) continue;
        x264_picture_alloc(&pic, X264_CSP_I420, 16, 16);
        for (int i = 0; i < 2; i++) {
            memset(pic.img.plane[0], 60 + 90 * i, 256); memset(pic.img.plane[1], 128, 64); memset(pic.img.plane[2], 128, 64);
            pic.img.plane[0][37 * i] ^= 0x5a; /* not a skip */
            pic.i_pts = i; pic.i_type = X264_TYPE_AUTO; pic.i_qpplus1 = 0;
            x264_encoder_encode(h, &nal, &i_nal, &pic, &out);
        }
        x264_picture_clean(&pic);
        /* not closed, see above */
    }
}

static void *warm_main(void *arg) { (void)arg; x264_b200_warm_device(); return NULL; }

static int is_flag(const char *name)
{
    static const char *flags[] = { "8x8dct", "weightb", "mixed-refs", "b-pyramid", "interlaced", "aud", "progress", "quiet", "verbose", "non-deterministic",
                                   "pre-scenecut", "bime", "b-rdo", "sps-id-unused", 0 };
    if (!strncmp(name, "no-", 3)) return 1;
    for (int i = 0; flags[i]; i++) if (!strcmp(name, flags[i])) return 1;
    return 0;
}

int main(int argc, char **argv)
{
    x264_param_t param;
    const char *out_path = NULL, *in_path = NULL, *res = NULL;
    int workers = 4, frames = 0, quiet = 0;
    x264_param_default(&param);
    for (int i = 1; i < argc; i++) {
        const char *a = argv[i];
        if (!strcmp(a, "-o") && i + 1 < argc) out_path = argv[++i];
        else if (!strcmp(a, "--workers") && i + 1 < argc) workers = atoi(argv[++i]);
        else if (!strcmp(a, "--frames") && i + 1 < argc) frames = atoi(argv[++i]);
        else if (!strcmp(a, "--quiet-front-end")) quiet = 1;
        else if (!strncmp(a, "--", 2)) {
            const char *name = a + 2, *val = NULL;
            if (!is_flag(name) && i + 1 < argc) val = argv[++i];
            if (x264_param_parse(&param, name, val)) { fprintf(stderr, "x264_b200_gops: bad option --%s %s\n", name, val ? val : ""); return 2; }
        } else if (!in_path) in_path = a;
        else res = a;
    }
    if (!out_path || !in_path || !res || sscanf(res, "%dx%d", &param.i_width, &param.i_height) != 2 || param.i_keyint_max <= 0 || param.i_keyint_max >= 1 << 20) {
        fprintf(stderr, "usage: x264_b200_gops [x264 options] --keyint K [--workers T] [--frames N] -o out.264 in.yuv WIDTHxHEIGHT\n");
        return 2;
    }
    const int keyint = param.i_keyint_max;
    param.i_keyint_min = keyint;          /* --min-keyint K --scenecut -1: IDR positions independent of the content */
    param.i_scenecut_threshold = -1;
    param.i_threads = 1;                  /* the bit-exact configuration; parallelism comes from the GOP runs */
    FILE *f = fopen(in_path, "rb");
    if (!f) { fprintf(stderr, "x264_b200_gops: cannot open %s\n", in_path); return 2; }
    fseeko(f, 0, SEEK_END);
    const int in_frames = (int)(ftello(f) / ((off_t)param.i_width * param.i_height * 3 / 2));
    fclose(f);
    if (frames <= 0 || frames > in_frames) frames = in_frames;
    const int n_gops = (frames + keyint - 1) / keyint;
    if (workers > n_gops) workers = n_gops;
    if (workers < 1) { fprintf(stderr, "x264_b200_gops: no frames\n"); return 2; }

    /* the device's context comes up on a helper thread while this one builds the cost tables */
    pthread_t warm;
    const int warming = pthread_create(&warm, NULL, warm_main, NULL) == 0;
    build_cost_tables(&param);
    if (warming) pthread_join(warm, NULL);
    worker_t *w = calloc(workers, sizeof(*w));
    const double t0 = now_s();
    for (int k = 0, gop = 0; k < workers; k++) { /* consecutive runs, sizes differing by at most one GOP */
        const int n = n_gops / workers + (k < n_gops % workers);
        w[k].index = k; w[k].first_gop = gop; w[k].first_frame = gop * keyint;
        w[k].n_frames = (gop + n) * keyint <= frames ? n * keyint : frames - gop * keyint;
        w[k].param = param; w[k].in_path = in_path;
        gop += n;
        pthread_create(&w[k].th, NULL, worker_main, &w[k]);
    }
    int rc = 0;
    for (int k = 0; k < workers; k++) { pthread_join(w[k].th, NULL); if (w[k].rc) { fprintf(stderr, "x264_b200_gops: worker %d failed (%d)\n", k, w[k].rc); rc = 1; } }
    const double t1 = now_s();
    if (rc) return rc;
    FILE *o = fopen(out_path, "wb");
    if (!o) { fprintf(stderr, "x264_b200_gops: cannot open %s\n", out_path); return 2; }
    for (int k = 0; k < workers; k++) {
        const uint8_t *p = w[k].out;
        size_t n = w[k].out_size;
        if (k && n > 5 && !memcmp(p, "\0\0\0\1", 4) && (p[4] & 0x1f) == 6) { /* the version SEI belongs to frame 0 of the whole stream only */
            size_t j = 4;
            while (j + 4 <= n && memcmp(p + j, "\0\0\0\1", 4)) j++;
            p += j; n -= j;
        }
        fwrite(p, 1, n, o);
    }
    fclose(o);
    if (!quiet) fprintf(stderr, "encoded %d frames, %.2f fps, %d encoder threads on one device\n", frames, frames / (t1 - t0), workers);
    fflush(stderr);
    _exit(0); /* instances are deliberately not closed; skip their (process-wide) teardown */
}
