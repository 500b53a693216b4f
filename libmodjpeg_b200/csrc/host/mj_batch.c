/*
 * mj_batch.c -- mj_compose_batch: the host pipeline around the kernels for many JPEGs and one dropon
 * (SURVEY 8f rank 1).  What the reference does per image in one thread
 *     mj_read_jpeg_from_memory -> mj_compose -> mj_write_jpeg_to_memory      (src/image.c:33, src/compose.c:33, src/image.c:136)
 * is split here into
 *     entropy decode   host libjpeg, a pool of threads, one image per task
 *     dropon compile   K1, once per image geometry of the batch (the reference recompiles per image, src/compose.c:155-177)
 *     blend            K2, ONE launch per window of images, working in place on a page-locked slab that holds the
 *                      region under the dropon of every image of the window (zero-copy over PCIe, mjx_compose_batch_host)
 *     entropy encode   host libjpeg, the same pool
 * Entropy coding is serial per image and stays on the CPU (north_star); the pool is what scales it.
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "mj_private.h"

typedef struct {
    int            n;
    const mj_blob_t *in;
    mj_blob_t      *out;
    int            *status;
    mj_jpeg_t      *jp;        /* decoded images of the current window, index i - w0 */
    int             w0, w1;    /* current window */
    int             phase;     /* 1 decode, 2 stage in, 3 stage out + encode */
    int             next;      /* next task of the phase (guarded by lock) */
    pthread_mutex_t lock;
    /* phase 2 / 3 */
    const int      *group;     /* indices (absolute) of the images composed in this round */
    int             ngroup;
    char           *slab;      /* page-locked staging: ngroup regions of region_bytes */
    size_t          region_bytes;
    size_t          comp_off[MJX_MAX_COMPONENTS];
    int             wb[MJX_MAX_COMPONENTS], hb[MJX_MAX_COMPONENTS];
    int             ncomp;
    mjx_geometry_t  g;
    int             write_options;
    double          phase_s[5]; /* MJ_BATCH_TRACE=1: seconds per phase ([0] = K2) */
} batch_t;

static int take(batch_t *b, int limit) {
    pthread_mutex_lock(&b->lock);
    int k = b->next < limit ? b->next++ : -1;
    pthread_mutex_unlock(&b->lock);
    return k;
}

/* copy the rows under the dropon between libjpeg's virtual arrays and the image's slab region */
static int stage_rows(batch_t *b, int slot, mj_jpeg_t *m, int to_slab) {
    mjp_trap_t *trap = mjp_image_trap(m);
    trap->armed = 1;
    if(setjmp(trap->escape)) {
        trap->armed = 0;
        return MJ_ERR_DECODE_JPEG;
    }
    char *region = b->slab + (size_t)slot * b->region_bytes;
    for(int c = 0; c < b->ncomp; c++) {
        jpeg_component_info *ci = &m->cinfo.comp_info[c];
        const unsigned       x0 = (unsigned)(b->g.block_x * ci->h_samp_factor), y0 = (unsigned)(b->g.block_y * ci->v_samp_factor);
        const size_t         wbytes = (size_t)b->wb[c] * 128;
        for(int l = 0; l < b->hb[c]; l++) {
            JBLOCKARRAY ba = (*m->cinfo.mem->access_virt_barray)((j_common_ptr)&m->cinfo, m->coef[c], y0 + (unsigned)l, 1, TRUE);
            char       *row = (char *)&ba[0][x0][0], *st = region + b->comp_off[c] + (size_t)l * wbytes;
            if(to_slab) memcpy(st, row, wbytes);
            else memcpy(row, st, wbytes);
        }
    }
    trap->armed = 0;
    return MJ_OK;
}

static void *worker(void *arg) {
    batch_t *b = (batch_t *)arg;
    for(;;) {
        if(b->phase == 1) {
            int k = take(b, b->w1 - b->w0);
            if(k < 0) break;
            const int i = b->w0 + k;
            mj_init_jpeg(&b->jp[k]);
            b->status[i] = (b->in[i].data == NULL) ? MJ_ERR_NULL_DATA : mj_read_jpeg_from_memory(&b->jp[k], b->in[i].data, b->in[i].len, 0);
        }
        else if(b->phase == 2) {
            int s = take(b, b->ngroup);
            if(s < 0) break;
            const int i = b->group[s];
            int       rv = stage_rows(b, s, &b->jp[i - b->w0], 1);
            if(rv != MJ_OK) b->status[i] = rv;
        }
        else if(b->phase == 3) {
            int s = take(b, b->ngroup);
            if(s < 0) break;
            const int i = b->group[s];
            if(b->status[i] == MJ_OK) {
                int rv = stage_rows(b, s, &b->jp[i - b->w0], 0);
                if(rv != MJ_OK) b->status[i] = rv;
            }
        }
        else { /* 4: encode + free, every image of the window */
            int k = take(b, b->w1 - b->w0);
            if(k < 0) break;
            const int i = b->w0 + k;
            if(b->status[i] == MJ_OK) b->status[i] = mj_write_jpeg_to_memory(&b->jp[k], &b->out[i].data, &b->out[i].len, b->write_options);
            mj_free_jpeg(&b->jp[k]);
        }
    }
    return NULL;
}

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void run_phase(batch_t *b, int phase, int nthreads, pthread_t *th) {
    const double t0 = now_s();
    b->phase = phase;
    b->next = 0;
    int started = 1;
    for(int t = 1; t < nthreads; t++) {
        if(pthread_create(&th[started], NULL, worker, b) != 0) break; /* fewer helpers: the tasks are taken by whoever runs */
        started++;
    }
    worker(b); /* the calling thread works too */
    for(int t = 1; t < started; t++) pthread_join(th[t], NULL);
    b->phase_s[phase] += now_s() - t0;
}

static int same_geometry(const mj_jpeg_t *a, const mj_jpeg_t *b) {
    if(a->width != b->width || a->height != b->height || a->cinfo.jpeg_color_space != b->cinfo.jpeg_color_space ||
       a->cinfo.num_components != b->cinfo.num_components)
        return 0;
    for(int c = 0; c < a->cinfo.num_components; c++)
        if(a->cinfo.comp_info[c].h_samp_factor != b->cinfo.comp_info[c].h_samp_factor ||
           a->cinfo.comp_info[c].v_samp_factor != b->cinfo.comp_info[c].v_samp_factor)
            return 0;
    return 1;
}

/* the pipeline on the calling thread's device */
static int batch_on_device(int n, const mj_blob_t *in, mj_blob_t *out, int *status, mj_dropon_t *d, unsigned int align, int offset_x,
                           int offset_y, int write_options, int nthreads) {
    const int compose = d->blend != MJ_BLEND_NONE && d->image != NULL && d->alpha != NULL;
    mjx_ctx  *ctx = compose ? mjx_host_ctx() : NULL;
    if(compose && ctx == NULL) return MJ_ERR_DEVICE;

    int window = 4 * nthreads;
    if(window > 256) window = 256;
    if(window > n) window = n;
    batch_t b;
    memset(&b, 0, sizeof(b));
    b.n = n, b.in = in, b.out = out, b.status = status, b.write_options = write_options;
    pthread_mutex_init(&b.lock, NULL);
    b.jp = (mj_jpeg_t *)calloc((size_t)window, sizeof(mj_jpeg_t));
    int       *group = (int *)malloc(sizeof(int) * (size_t)window);
    char      *done = (char *)malloc((size_t)window);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    mjx_host_image_t *items = (mjx_host_image_t *)malloc(sizeof(mjx_host_image_t) * (size_t)window);
    int            result = MJ_OK;
    mjx_dropon    *cd = NULL;
    mjx_layout_t   cd_layout;
    mjx_geometry_t cd_g;
    memset(&cd_layout, 0, sizeof(cd_layout));
    memset(&cd_g, 0, sizeof(cd_g));
    if(b.jp == NULL || group == NULL || done == NULL || th == NULL || items == NULL) {
        result = MJ_ERR_MEMORY;
        goto out;
    }

    for(b.w0 = 0; b.w0 < n; b.w0 = b.w1) {
        b.w1 = b.w0 + window < n ? b.w0 + window : n;
        run_phase(&b, 1, nthreads, th); /* entropy decode */

        /* compose the window group by group (images sharing one geometry share one compiled dropon and one launch) */
        memset(done, 0, (size_t)window);
        for(int k0 = 0; compose && k0 < b.w1 - b.w0; k0++) {
            if(done[k0] || status[b.w0 + k0] != MJ_OK) continue;
            mj_jpeg_t *ref = &b.jp[k0];
            b.ngroup = 0;
            for(int k = k0; k < b.w1 - b.w0; k++)
                if(!done[k] && status[b.w0 + k] == MJ_OK && same_geometry(ref, &b.jp[k])) {
                    group[b.ngroup++] = b.w0 + k;
                    done[k] = 1;
                }
            mjx_geometry(ref->width, ref->height, ref->sampling.h_factor, ref->sampling.v_factor, d->width, d->height, align, offset_x,
                         offset_y, &b.g);
            if(!b.g.visible) continue; /* dropon entirely off these images (reference: src/compose.c:136) */
            mjx_layout_t layout;
            int          rv = mjx_jpeg_layout(ref, &layout);
            /* one compiled dropon is kept across groups and windows while layout and placement repeat */
            if(rv == MJ_OK && cd != NULL && (memcmp(&layout, &cd_layout, sizeof(layout)) != 0 || memcmp(&b.g, &cd_g, sizeof(b.g)) != 0)) {
                mjx_dropon_free(cd);
                cd = NULL;
            }
            if(rv == MJ_OK && cd == NULL) {
                cd_layout = layout;
                cd_g = b.g;
                rv = mjx_dropon_compile(ctx, &cd, d->image, d->alpha, d->width, d->height, d->colorspace, &layout, b.g.blockoffset_x,
                                        b.g.blockoffset_y, b.g.crop_x, b.g.crop_y, b.g.crop_w, b.g.crop_h, 0);
                if(rv == MJX_ERR_UNSUPPORTED) fprintf(stderr, "Unsupported color conversion request\n");
                rv = mjp_map_error(rv);
            }
            if(rv == MJ_OK) {
                b.ncomp = layout.ncomp;
                b.region_bytes = 0;
                for(int c = 0; c < b.ncomp; c++) {
                    mjx_dropon_dims(cd, c, &b.wb[c], &b.hb[c]);
                    b.comp_off[c] = b.region_bytes;
                    b.region_bytes += ((size_t)b.wb[c] * (size_t)b.hb[c] * 128 + 255) & ~(size_t)255;
                }
                void *slab = NULL;
                rv = mjp_map_error(mjx_ctx_pinned_scratch(ctx, b.region_bytes * (size_t)b.ngroup, &slab));
                b.slab = (char *)slab;
            }
            if(rv == MJ_OK) {
                b.group = group;
                run_phase(&b, 2, nthreads, th); /* rows under the dropon -> page-locked slab */
                for(int s = 0; s < b.ngroup; s++) {
                    mj_jpeg_t *m = &b.jp[group[s] - b.w0];
                    memset(&items[s], 0, sizeof(items[s]));
                    for(int c = 0; c < b.ncomp; c++) {
                        items[s].plane[c] = (int16_t *)(b.slab + (size_t)s * b.region_bytes + b.comp_off[c]);
                        items[s].stride_blocks[c] = items[s].wreal[c] = b.wb[c];
                        items[s].rows[c] = items[s].hreal[c] = b.hb[c];
                        items[s].q[c] = m->cinfo.comp_info[c].quant_table ? m->cinfo.comp_info[c].quant_table->quantval : NULL;
                        if(items[s].q[c] == NULL) status[group[s]] = MJ_ERR_NULL_DATA;
                    }
                }
                int ok = 1;
                for(int s = 0; s < b.ngroup; s++) ok &= status[group[s]] == MJ_OK;
                if(ok) {
                    const double tk = now_s();
                    rv = mjx_compose_batch_host(ctx, items, b.ngroup, cd, 0, 0); /* K2: one launch for the group */
                    b.phase_s[0] += now_s() - tk;
                    if(rv != MJX_OK) fprintf(stderr, "libmodjpeg (B200): batch compose failed: %s\n", mjx_ctx_last_error(ctx));
                    rv = mjp_map_error(rv);
                }
                else rv = MJ_ERR_NULL_DATA;
                if(rv == MJ_OK) run_phase(&b, 3, nthreads, th); /* slab -> libjpeg's arrays */
            }
            if(rv != MJ_OK)
                for(int s = 0; s < b.ngroup; s++)
                    if(status[group[s]] == MJ_OK) status[group[s]] = rv;
        }
        run_phase(&b, 4, nthreads, th); /* entropy encode + free */
    }
    if(getenv("MJ_BATCH_TRACE") != NULL)
        fprintf(stderr, "mj_compose_batch: %d images, %d threads: decode %.3f s, stage-in %.3f s, K2 %.3f s, stage-out %.3f s, encode %.3f s\n", n,
                nthreads, b.phase_s[1], b.phase_s[2], b.phase_s[0], b.phase_s[3], b.phase_s[4]);
out:
    if(cd != NULL) mjx_dropon_free(cd);
    free(b.jp);
    free(group);
    free(done);
    free(th);
    free(items);
    pthread_mutex_destroy(&b.lock);
    return result;
}

/* ---- several devices from one process -------------------------------------------------------------------------------------
 * Images are independent (reference: src/compose.c:256-339 keeps no state between blocks), so the batch is cut into
 * one contiguous slice per device; every slice runs the pipeline above on its own group of host threads, whose engine contexts
 * (stream, staging pools, page-locked slab) and compiled dropon live on that device.  No data crosses between devices. */
static int g_devices = 0; /* 0: $MJX_DEVICES ("all" or a count), default 1 */

void mj_batch_set_devices(int devices) { __atomic_store_n(&g_devices, devices < 0 ? 0 : devices, __ATOMIC_RELAXED); }

static int batch_devices(void) {
    int want = __atomic_load_n(&g_devices, __ATOMIC_RELAXED);
    if(want == 0) {
        const char *e = getenv("MJX_DEVICES");
        if(e == NULL || !*e) return 1;
        want = strcmp(e, "all") == 0 ? 1 << 20 : atoi(e);
    }
    const int have = mjx_device_count();
    if(want > have) want = have;
    return want < 1 ? 1 : want;
}

typedef struct {
    int              device, n, nthreads, rv;
    const mj_blob_t *in;
    mj_blob_t       *out;
    int             *status;
    mj_dropon_t     *d;
    unsigned int     align;
    int              offset_x, offset_y, write_options;
} slice_t;

static void *slice_main(void *arg) {
    slice_t *s = (slice_t *)arg;
    mjx_host_set_device(s->device); /* this thread's context (created inside) lives on the slice's device */
    s->rv = batch_on_device(s->n, s->in, s->out, s->status, s->d, s->align, s->offset_x, s->offset_y, s->write_options, s->nthreads);
    return NULL;
}

int mj_compose_batch(int n, const mj_blob_t *in, mj_blob_t *out, int *status, mj_dropon_t *d, unsigned int align, int offset_x,
                     int offset_y, int write_options, int nthreads) {
    if(n < 0 || (n > 0 && (in == NULL || out == NULL || status == NULL)) || d == NULL) return MJ_ERR_NULL_DATA;
    if(nthreads < 1) nthreads = 1;
    if(nthreads > 256) nthreads = 256;
    for(int i = 0; i < n; i++) {
        out[i].data = NULL;
        out[i].len = 0;
        status[i] = MJ_OK;
    }
    if(n == 0) return MJ_OK;
    int ndev = batch_devices();
    if(ndev > nthreads) ndev = nthreads;
    if(ndev > n) ndev = n;
    if(ndev <= 1) return batch_on_device(n, in, out, status, d, align, offset_x, offset_y, write_options, nthreads);
    slice_t   sl[64];
    pthread_t th[64];
    char      joined[64];
    if(ndev > 64) ndev = 64;
    int result = MJ_OK, lo = 0;
    for(int k = 0; k < ndev; k++) {
        const int cnt = n / ndev + (k < n % ndev ? 1 : 0);
        slice_t  *s = &sl[k];
        s->device = k, s->n = cnt, s->in = in + lo, s->out = out + lo, s->status = status + lo, s->d = d, s->align = align;
        s->offset_x = offset_x, s->offset_y = offset_y, s->write_options = write_options, s->rv = MJ_OK;
        s->nthreads = nthreads / ndev + (k < nthreads % ndev ? 1 : 0);
        lo += cnt;
        joined[k] = pthread_create(&th[k], NULL, slice_main, s) == 0;
        if(!joined[k]) /* no thread to be had: the caller does this slice itself, on its own device */
            s->rv = batch_on_device(s->n, s->in, s->out, s->status, s->d, s->align, s->offset_x, s->offset_y, s->write_options, s->nthreads);
    }
    for(int k = 0; k < ndev; k++) {
        if(joined[k]) pthread_join(th[k], NULL);
        if(sl[k].rv != MJ_OK && result == MJ_OK) result = sl[k].rv;
    }
    return result;
}
