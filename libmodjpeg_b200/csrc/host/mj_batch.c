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
 * Entropy DEcoding is serial per image and stays on the CPU (north_star); the pool is what scales it.  For plain baseline
 * output (write_options == 0; MJX_GPU_HUFFMAN=0 switches it off) the second half runs on the device instead
 * (group_on_device): the window's whole planes go up once, K2 blends them in HBM, K4 (k4_huffman.cu) codes the scans, and
 * only the entropy-coded segments -- a twentieth of the planes -- come back; the pool puts libjpeg's markers in front.
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "mj_private.h"

/* device buffers of a window in flight (pieces of the context's device scratch) */
typedef struct {
    void *planes, *segs, *sizes, *descs, *input, *dstatus;
} devbufs_t;

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
    double          phase_s[9]; /* MJ_BATCH_TRACE=1: seconds per phase ([0] = K2 / waiting for the device) */
    /* phases 5 / 6: whole planes to the device, files from device-coded segments (MJX_GPU_HUFFMAN=1) */
    size_t          image_bytes, plane_off[MJX_MAX_COMPONENTS];
    int             stride[MJX_MAX_COMPONENTS], hreal[MJX_MAX_COMPONENTS], wreal[MJX_MAX_COMPONENTS];
    const unsigned int *seg_size; /* per slot: bytes of the segment in the slot's slab region, 0xFFFFFFFF: not coded */
    size_t          seg_cap;      /* bytes per segment in the device slab */
    char           *written;      /* per window image: the output file exists already */
    int            *group_buf;    /* storage of `group` */
    /* phases 7 / 8: the window never leaves the device (K5 decodes it there): header-only reads, entropy-coded segments to the slab */
    size_t         *ent_off;      /* per window image: where its entropy-coded segment starts in the input */
    mjx_scan_t     *in_scan;      /* per window image: its scan with the file's tables */
    char           *eligible;     /* per window image: header read, the device can decode the scan */
    size_t         *seg_in_off;   /* per slot: offset of the segment in the slab */
    int             full;         /* this window was queued by window_enqueue_full */
    int             out_by_offset; /* phase 6: the slot's segment lies at seg_in_off[s], not at s * image_bytes */
    int             vrows[MJX_MAX_COMPONENTS];
    uint32_t       *dec_status;   /* per slot: K5's verdict (in the slab, behind seg_size) */
    mjx_ctx        *ctx;          /* the context (stream, slab, device scratch) this window's device work was queued on */
    devbufs_t       dv;
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

/* copy every real row of every plane between libjpeg's arrays and the image's slab region (whole image, not only the region
 * under the dropon: the device also codes the file, k4_huffman.cu) */
static int stage_planes(batch_t *b, int slot, mj_jpeg_t *m, int to_slab) {
    mjp_trap_t *trap = mjp_image_trap(m);
    trap->armed = 1;
    if(setjmp(trap->escape)) {
        trap->armed = 0;
        return MJ_ERR_DECODE_JPEG;
    }
    char *region = b->slab + (size_t)slot * b->image_bytes;
    for(int c = 0; c < b->ncomp; c++) {
        const size_t wbytes = (size_t)b->stride[c] * 128;
        for(int l = 0; l < b->hreal[c]; l++) {
            JBLOCKARRAY ba = (*m->cinfo.mem->access_virt_barray)((j_common_ptr)&m->cinfo, m->coef[c], (unsigned)l, 1, TRUE);
            char       *row = (char *)&ba[0][0][0], *st = region + b->plane_off[c] + (size_t)l * wbytes;
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
        else if(b->phase == 7) { /* markers only: the frame, the tables, where the entropy-coded segment starts */
            int k = take(b, b->w1 - b->w0);
            if(k < 0) break;
            const int i = b->w0 + k;
            mj_init_jpeg(&b->jp[k]);
            b->eligible[k] = 0;
            if(b->in[i].data == NULL) {
                b->status[i] = MJ_ERR_NULL_DATA;
                continue;
            }
            const int rv = mjp_read_header_only(&b->jp[k], b->in[i].data, b->in[i].len, &b->ent_off[k], &b->in_scan[k]);
            if(rv == MJ_OK) b->eligible[k] = 1;
            else if(rv != MJ_ERR_UNSUPPORTED_FILETYPE) b->status[i] = rv; /* (another kind of JPEG: the window takes the ordinary path) */
        }
        else if(b->phase == 8) { /* entropy-coded segments -> page-locked slab */
            int s = take(b, b->ngroup);
            if(s < 0) break;
            const int i = b->group[s], k = i - b->w0;
            memcpy(b->slab + b->seg_in_off[s], b->in[i].data + b->ent_off[k], b->in[i].len - b->ent_off[k]);
        }
        else if(b->phase == 5) { /* whole planes -> page-locked slab */
            int s = take(b, b->ngroup);
            if(s < 0) break;
            const int i = b->group[s];
            int       rv = stage_planes(b, s, &b->jp[i - b->w0], 1);
            if(rv != MJ_OK) b->status[i] = rv;
        }
        else if(b->phase == 6) { /* the file: libjpeg's markers + the segment the device coded + EOI */
            int s = take(b, b->ngroup);
            if(s < 0) break;
            const int i = b->group[s];
            if(b->status[i] != MJ_OK || b->seg_size[s] == 0xFFFFFFFFu) continue;
            unsigned char *head = NULL;
            size_t         head_len = 0;
            mjx_scan_t     scan;
            int            rv = mjp_scan_headers(&b->jp[i - b->w0], &head, &head_len, &scan);
            if(rv == MJ_OK)
                rv = mjp_assemble_file(&b->out[i].data, &b->out[i].len, head, head_len,
                                       (const unsigned char *)b->slab + (b->out_by_offset ? b->seg_in_off[s] : (size_t)s * b->image_bytes), b->seg_size[s]);
            free(head);
            if(rv == MJ_OK) b->written[i - b->w0] = 1;
            else b->status[i] = rv;
        }
        else { /* 4: encode + free, every image of the window */
            int k = take(b, b->w1 - b->w0);
            if(k < 0) break;
            const int i = b->w0 + k;
            if(b->status[i] == MJ_OK && !(b->written && b->written[k]))
                b->status[i] = mj_write_jpeg_to_memory(&b->jp[k], &b->out[i].data, &b->out[i].len, b->write_options);
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


/* the device buffers of a window: one ctx-owned, grow-only scratch (it outlives the call: a server that sends batch after batch
 * allocates once), cut into the pieces a window needs */
static int dev_layout(mjx_ctx *ctx, devbufs_t *v, size_t planes, size_t segs, size_t sizes, size_t descs, size_t input, size_t dstatus) {
    const size_t need[6] = {planes, segs, sizes, descs, input, dstatus};
    size_t       off[6], total = 0;
    for(int k = 0; k < 6; k++) {
        off[k] = total;
        total += (need[k] + 255) & ~(size_t)255;
    }
    void *base = NULL;
    int   rv = mjx_ctx_device_scratch(ctx, total, &base);
    if(rv != MJX_OK) return rv;
    v->planes = (char *)base + off[0], v->segs = (char *)base + off[1], v->sizes = (char *)base + off[2];
    v->descs = (char *)base + off[3], v->input = (char *)base + off[4], v->dstatus = (char *)base + off[5];
    return MJX_OK;
}

/* One group (images of one geometry) entirely on the device: whole planes up, K2 in HBM, K4 codes the scans, only the
 * entropy-coded segments come back; the workers put libjpeg's markers in front of them.  Images the device cannot code
 * (a coefficient outside the baseline tables, a segment beyond the slab) get their planes back and are left to libjpeg.
 * Two halves, so that the device works on one window while the pool decodes the next:
 *   group_enqueue   planes -> page-locked slab (pool), then H2D, K2, K4 and the read of the segment sizes are QUEUED on the
 *                   ctx stream; returns without waiting.  MJ_OK: queued; another code: nothing was queued and the ordinary
 *                   path should take the group.
 *   group_finish    waits for the stream, fetches the segments, writes the files (pool), hands uncoded images back. */
static int group_enqueue(batch_t *b, mjx_ctx *ctx, devbufs_t *v, mjx_dropon *cd, int nthreads, pthread_t *th) {
    mj_jpeg_t *ref = &b->jp[b->group[0] - b->w0];
    mjx_scan_t scan;
    {
        unsigned char *head = NULL;
        size_t         head_len = 0;
        const int      rv = mjp_scan_headers(ref, &head, &head_len, &scan);
        free(head);
        if(rv != MJ_OK) return rv;
    }
    b->image_bytes = 0;
    for(int c = 0; c < b->ncomp; c++) {
        const jpeg_component_info *ci = &ref->cinfo.comp_info[c];
        b->stride[c] = (int)mjp_virtual_width(ci);
        b->wreal[c] = (int)ci->width_in_blocks;
        b->hreal[c] = (int)ci->height_in_blocks;
        b->plane_off[c] = b->image_bytes;
        b->image_bytes += ((size_t)b->stride[c] * (size_t)b->hreal[c] * 128 + 255) & ~(size_t)255;
    }
    /* a segment as long as a quarter of the coefficients is a file of ~4 bits per coefficient: beyond what quality 100 produces */
    size_t cap = (b->image_bytes / 4 + 255) & ~(size_t)255;
    if(cap < 65536) cap = 65536;
    if(cap > b->image_bytes) cap = b->image_bytes;
    b->seg_cap = cap;
    const size_t ng = (size_t)b->ngroup;
    void        *slab = NULL;
    /* slab: the images' regions, then the descriptors, then the segment sizes the device reports */
    int rv = mjx_ctx_pinned_scratch(ctx, b->image_bytes * ng + sizeof(mjx_image_desc_t) * ng + 4 * ng + 256, &slab);
    if(rv == MJX_OK) rv = dev_layout(ctx, v, b->image_bytes * ng, cap * ng, 4 * ng, sizeof(mjx_image_desc_t) * ng, 0, 0);
    if(rv != MJX_OK) return mjp_map_error(rv);
    b->slab = (char *)slab;
    run_phase(b, 5, nthreads, th); /* whole planes -> page-locked slab */

    mjx_image_desc_t *descs = (mjx_image_desc_t *)(b->slab + b->image_bytes * ng);
    b->seg_size = (unsigned int *)((char *)descs + sizeof(mjx_image_desc_t) * ng);
    memset(descs, 0, sizeof(mjx_image_desc_t) * ng);
    for(int s = 0; s < b->ngroup; s++) {
        mj_jpeg_t *m = &b->jp[b->group[s] - b->w0];
        for(int c = 0; c < b->ncomp; c++) {
            descs[s].plane[c] = (uint64_t)(uintptr_t)((char *)v->planes + (size_t)s * b->image_bytes + b->plane_off[c]);
            descs[s].stride_blocks[c] = b->stride[c];
            descs[s].rows[c] = b->hreal[c];
            descs[s].wreal[c] = b->wreal[c];
            descs[s].hreal[c] = b->hreal[c];
            const JQUANT_TBL *qt = m->cinfo.comp_info[c].quant_table;
            if(qt == NULL) b->status[b->group[s]] = MJ_ERR_NULL_DATA;
            else memcpy(descs[s].q[c], qt->quantval, 128);
        }
    }
    for(int s = 0; s < b->ngroup; s++)
        if(b->status[b->group[s]] != MJ_OK) return MJ_ERR_NULL_DATA; /* (the caller marks the group) */

    rv = mjx_copy_h2d(ctx, v->planes, b->slab, b->image_bytes * ng);
    if(rv == MJX_OK) rv = mjx_copy_h2d(ctx, v->descs, descs, sizeof(mjx_image_desc_t) * ng);
    if(rv == MJX_OK) rv = mjx_compose_batch_device(ctx, (const mjx_image_desc_t *)v->descs, b->ngroup, cd, b->g.block_x, b->g.block_y);
    if(rv == MJX_OK) rv = mjx_huffman_encode_batch_device(ctx, (const mjx_image_desc_t *)v->descs, b->ngroup, &scan, v->segs, cap, (uint32_t *)v->sizes);
    if(rv == MJX_OK) rv = mjx_copy_d2h(ctx, (void *)b->seg_size, v->sizes, 4 * ng);
    if(rv != MJX_OK) {
        mjx_ctx_sync(ctx);
        fprintf(stderr, "libmodjpeg (B200): device batch failed: %s\n", mjx_ctx_last_error(ctx));
        return mjp_map_error(rv);
    }
    return MJ_OK;
}

static void group_finish(batch_t *b, mjx_ctx *ctx, devbufs_t *v, int nthreads, pthread_t *th) {
    const double tw = now_s();
    int          rv = mjx_ctx_sync(ctx);
    /* the segments, each at the start of its image's slab region (the planes there have been uploaded); an image that was not
     * coded gets its composed planes back instead */
    for(int s = 0; rv == MJX_OK && s < b->ngroup; s++) {
        if(b->seg_size[s] != 0xFFFFFFFFu) rv = mjx_copy_d2h(ctx, b->slab + (size_t)s * b->image_bytes, (char *)v->segs + (size_t)s * b->seg_cap, b->seg_size[s]);
        else rv = mjx_copy_d2h(ctx, b->slab + (size_t)s * b->image_bytes, (char *)v->planes + (size_t)s * b->image_bytes, b->image_bytes);
    }
    if(rv == MJX_OK) rv = mjx_ctx_sync(ctx);
    b->phase_s[0] += now_s() - tw; /* what the host WAITED for the device (copies + K2 + K4 beyond what the next window's decode hid) */
    if(rv != MJX_OK) {
        fprintf(stderr, "libmodjpeg (B200): device batch failed: %s\n", mjx_ctx_last_error(ctx));
        for(int s = 0; s < b->ngroup; s++)
            if(b->status[b->group[s]] == MJ_OK) b->status[b->group[s]] = mjp_map_error(rv);
        return;
    }
    run_phase(b, 6, nthreads, th); /* files */
    for(int s = 0; s < b->ngroup; s++) { /* not coded on the device: composed planes back into libjpeg's arrays, phase 4 encodes */
        const int i = b->group[s];
        if(b->seg_size[s] == 0xFFFFFFFFu && b->status[i] == MJ_OK) {
            const int r2 = stage_planes(b, s, &b->jp[i - b->w0], 0);
            if(r2 != MJ_OK) b->status[i] = r2;
        }
    }
}

/* A window that never leaves the device: its images share one geometry and one set of Huffman tables and the device can decode
 * their scans (K5).  Only JPEG bytes cross PCIe, in both directions:
 *   window_enqueue_full   entropy-coded segments -> slab (pool) -> HBM, K5 decodes them into the planes, K2 blends, K4 codes
 *                         the scans; queued on the ctx stream, returns without waiting
 *   window_finish_full    waits, fetches the output segments, writes the files (pool); an image K5 or K4 handed back goes
 *                         through the ordinary calls on the host (mj_read_jpeg_from_memory -> mj_compose -> mj_write_jpeg_to_memory)
 * Returns MJ_OK when queued; another code when the window is not of that kind (nothing was queued, the header-only objects are
 * freed) and the ordinary path has to read it. */
static void finish_window(batch_t *w, mj_dropon_t *d, unsigned int align, int offset_x, int offset_y, int nthreads, pthread_t *th);

static int window_enqueue_full(batch_t *b, mjx_ctx *ctx, devbufs_t *v, mjx_dropon **cd, mjx_layout_t *cd_layout, mjx_geometry_t *cd_g, mj_dropon_t *d,
                               unsigned int align, int offset_x, int offset_y, int nthreads, pthread_t *th, batch_t **in_flight) {
    const int nw = b->w1 - b->w0;
    int       rv = MJ_OK, ref_k = -1;
    b->ngroup = 0;
    for(int k = 0; k < nw; k++) {
        if(b->status[b->w0 + k] != MJ_OK) continue; /* undecodable: reported, not part of the group */
        if(!b->eligible[k]) rv = MJ_ERR_UNSUPPORTED_FILETYPE;
        else if(ref_k < 0) ref_k = k;
        else if(!same_geometry(&b->jp[ref_k], &b->jp[k]) || memcmp(&b->in_scan[ref_k], &b->in_scan[k], sizeof(mjx_scan_t)) != 0) rv = MJ_ERR_UNSUPPORTED_FILETYPE;
        b->group_buf[b->ngroup++] = b->w0 + k;
    }
    mj_jpeg_t *ref = ref_k >= 0 ? &b->jp[ref_k] : NULL;
    mjx_scan_t out_scan;
    if(rv == MJ_OK && ref == NULL) rv = MJ_ERR_UNSUPPORTED_FILETYPE;
    if(rv == MJ_OK) {
        mjx_geometry(ref->width, ref->height, ref->sampling.h_factor, ref->sampling.v_factor, d->width, d->height, align, offset_x, offset_y, &b->g);
        if(!b->g.visible) rv = MJ_ERR_UNSUPPORTED_FILETYPE; /* nothing to blend: the ordinary path transcodes */
    }
    if(rv == MJ_OK) {
        unsigned char *head = NULL;
        size_t         head_len = 0;
        rv = mjp_scan_headers(ref, &head, &head_len, &out_scan);
        free(head);
    }
    mjx_layout_t layout;
    if(rv == MJ_OK) rv = mjp_layout_of(ref, &layout);
    if(rv == MJ_OK && layout.ncomp > MJX_MAX_COMPONENTS) rv = MJ_ERR_UNSUPPORTED_FILETYPE;
    if(rv == MJ_OK && *cd != NULL && (memcmp(&layout, cd_layout, sizeof(layout)) != 0 || memcmp(&b->g, cd_g, sizeof(b->g)) != 0)) {
        if(*in_flight != NULL) { /* the other window's blend still reads the compiled dropon */
            finish_window(*in_flight, d, align, offset_x, offset_y, nthreads, th);
            *in_flight = NULL;
        }
        mjx_dropon_free(*cd);
        *cd = NULL;
    }
    if(rv == MJ_OK && *cd == NULL) {
        *cd_layout = layout;
        *cd_g = b->g;
        const int crv = mjx_dropon_compile(ctx, cd, d->image, d->alpha, d->width, d->height, d->colorspace, &layout, b->g.blockoffset_x, b->g.blockoffset_y,
                                           b->g.crop_x, b->g.crop_y, b->g.crop_w, b->g.crop_h, 0);
        if(crv != MJX_OK) rv = MJ_ERR_UNSUPPORTED_FILETYPE; /* the ordinary path reports it per image */
    }
    size_t in_total = 0;
    if(rv == MJ_OK) {
        b->ncomp = layout.ncomp;
        b->image_bytes = 0;
        for(int c = 0; c < b->ncomp; c++) {
            const jpeg_component_info *ci = &ref->cinfo.comp_info[c];
            /* (width_in_blocks / height_in_blocks are set by jpeg_read_header; the arrays libjpeg would allocate are rounded up
             * to the sampling factors, and so are the planes K5 fills) */
            b->stride[c] = (int)mjp_virtual_width(ci);
            b->wreal[c] = (int)ci->width_in_blocks;
            b->hreal[c] = (int)ci->height_in_blocks;
            b->vrows[c] = (int)mjp_virtual_height(ci);
            b->plane_off[c] = b->image_bytes;
            b->image_bytes += ((size_t)b->stride[c] * (size_t)b->vrows[c] * 128 + 255) & ~(size_t)255;
        }
        size_t cap = (b->image_bytes / 4 + 255) & ~(size_t)255;
        if(cap < 65536) cap = 65536;
        if(cap > b->image_bytes) cap = b->image_bytes;
        b->seg_cap = cap;
        for(int s = 0; s < b->ngroup; s++) {
            const int i = b->group_buf[s], k = i - b->w0;
            b->seg_in_off[s] = in_total;
            in_total += (b->in[i].len - b->ent_off[k] + 255) & ~(size_t)255;
            if(b->in[i].len - b->ent_off[k] > 0x1ffffff0u) rv = MJ_ERR_UNSUPPORTED_FILETYPE;
        }
    }
    const size_t ng = (size_t)b->ngroup;
    const size_t tail = sizeof(mjx_image_desc_t) * ng + 8 * ng + 512; /* descriptors, output sizes, K5's verdicts */
    if(rv == MJ_OK) {
        void *slab = NULL;
        int   mrv = mjx_ctx_pinned_scratch(ctx, in_total + tail, &slab);
        if(mrv == MJX_OK) mrv = dev_layout(ctx, v, b->image_bytes * ng, b->seg_cap * ng, 4 * ng, sizeof(mjx_image_desc_t) * ng, in_total + 256, 4 * ng);
        if(mrv != MJX_OK) rv = MJ_ERR_UNSUPPORTED_FILETYPE; /* no room for the window on the device: the ordinary path */
        b->slab = (char *)slab;
    }
    if(rv != MJ_OK) {
        for(int k = 0; k < nw; k++) mj_free_jpeg(&b->jp[k]);
        for(int k = 0; k < nw; k++)
            if(b->status[b->w0 + k] != MJ_OK && b->in[b->w0 + k].data != NULL) b->status[b->w0 + k] = MJ_OK; /* the full read decides again */
        return rv;
    }
    b->group = b->group_buf;
    run_phase(b, 8, nthreads, th); /* entropy-coded segments -> page-locked slab */

    mjx_image_desc_t *descs = (mjx_image_desc_t *)(b->slab + in_total);
    b->seg_size = (unsigned int *)((char *)descs + sizeof(mjx_image_desc_t) * ng);
    b->dec_status = (uint32_t *)(b->seg_size + ng);
    memset(descs, 0, sizeof(mjx_image_desc_t) * ng);
    uint64_t *offs = (uint64_t *)malloc(8 * ng);
    uint32_t *lens = (uint32_t *)malloc(4 * ng);
    if(offs == NULL || lens == NULL) {
        free(offs);
        free(lens);
        for(int k = 0; k < nw; k++) mj_free_jpeg(&b->jp[k]);
        return MJ_ERR_MEMORY;
    }
    for(int s = 0; s < b->ngroup; s++) {
        const int  i = b->group_buf[s], k = i - b->w0;
        mj_jpeg_t *m = &b->jp[k];
        offs[s] = (uint64_t)b->seg_in_off[s];
        lens[s] = (uint32_t)(b->in[i].len - b->ent_off[k]);
        for(int c = 0; c < b->ncomp; c++) {
            descs[s].plane[c] = (uint64_t)(uintptr_t)((char *)v->planes + (size_t)s * b->image_bytes + b->plane_off[c]);
            descs[s].stride_blocks[c] = b->stride[c];
            descs[s].rows[c] = b->vrows[c];
            descs[s].wreal[c] = b->wreal[c];
            descs[s].hreal[c] = b->hreal[c];
            memcpy(descs[s].q[c], m->cinfo.quant_tbl_ptrs[m->cinfo.comp_info[c].quant_tbl_no]->quantval, 128);
        }
    }
    int mrv = mjx_copy_h2d(ctx, v->input, b->slab, in_total);
    if(mrv == MJX_OK) mrv = mjx_copy_h2d(ctx, v->descs, descs, sizeof(mjx_image_desc_t) * ng);
    if(mrv == MJX_OK) mrv = mjx_huffman_decode_batch_device(ctx, v->input, offs, lens, b->ngroup, &b->in_scan[ref_k], (const mjx_image_desc_t *)v->descs, (uint32_t *)v->dstatus);
    if(mrv == MJX_OK) mrv = mjx_compose_batch_device(ctx, (const mjx_image_desc_t *)v->descs, b->ngroup, *cd, b->g.block_x, b->g.block_y);
    if(mrv == MJX_OK) mrv = mjx_huffman_encode_batch_device(ctx, (const mjx_image_desc_t *)v->descs, b->ngroup, &out_scan, v->segs, b->seg_cap, (uint32_t *)v->sizes);
    if(mrv == MJX_OK) mrv = mjx_copy_d2h(ctx, (void *)b->seg_size, v->sizes, 4 * ng);
    if(mrv == MJX_OK) mrv = mjx_copy_d2h(ctx, (void *)b->dec_status, v->dstatus, 4 * ng);
    free(offs);
    free(lens);
    if(mrv != MJX_OK) {
        mjx_ctx_sync(ctx);
        fprintf(stderr, "libmodjpeg (B200): device batch failed: %s\n", mjx_ctx_last_error(ctx));
        for(int k = 0; k < nw; k++) mj_free_jpeg(&b->jp[k]);
        return mjp_map_error(mrv);
    }
    b->full = 1;
    return MJ_OK;
}

static void window_finish_full(batch_t *b, mjx_ctx *ctx, devbufs_t *v, mj_dropon_t *d, unsigned int align, int offset_x, int offset_y, int nthreads,
                               pthread_t *th) {
    const double tw = now_s();
    int          rv = mjx_ctx_sync(ctx);
    /* output segments, packed into the slab (the input segments there have been uploaded) */
    size_t at = 0;
    for(int s = 0; rv == MJX_OK && s < b->ngroup; s++) {
        unsigned int *sz = (unsigned int *)&b->seg_size[s];
        if(b->dec_status[s] != 0u) *sz = 0xFFFFFFFFu; /* not decoded: whatever was coded from its planes is void */
        if(*sz == 0xFFFFFFFFu) continue;
        b->seg_in_off[s] = at; /* (reused: where the slot's OUTPUT segment lies) */
        at += ((size_t)*sz + 15) & ~(size_t)15;
    }
    if(rv == MJX_OK) {
        /* descriptors / sizes / verdicts live behind the input segments: keep them clear of the output */
        const size_t room = (size_t)((const char *)b->seg_size - b->slab) - sizeof(mjx_image_desc_t) * (size_t)b->ngroup;
        char        *dst = b->slab;
        void        *big = NULL;
        if(at > room) { /* more output than input (a large dropon over small files): a slab of its own */
            big = malloc(at);
            if(big == NULL) rv = MJX_ERR_MEMORY;
            dst = (char *)big;
        }
        for(int s = 0; rv == MJX_OK && s < b->ngroup; s++)
            if(b->seg_size[s] != 0xFFFFFFFFu) rv = mjx_copy_d2h(ctx, dst + b->seg_in_off[s], (char *)v->segs + (size_t)s * b->seg_cap, b->seg_size[s]);
        if(rv == MJX_OK) rv = mjx_ctx_sync(ctx);
        b->phase_s[0] += now_s() - tw;
        if(rv == MJX_OK) {
            char *keep = b->slab;
            b->slab = dst;
            b->out_by_offset = 1;
            run_phase(b, 6, nthreads, th); /* files */
            b->out_by_offset = 0;
            b->slab = keep;
        }
        free(big);
    }
    if(rv != MJX_OK) {
        fprintf(stderr, "libmodjpeg (B200): device batch failed: %s\n", mjx_ctx_last_error(ctx));
        for(int s = 0; s < b->ngroup; s++)
            if(b->status[b->group[s]] == MJ_OK && !b->written[b->group[s] - b->w0]) b->status[b->group[s]] = mjp_map_error(rv);
    }
    /* what the device handed back: the ordinary calls, one image at a time */
    for(int s = 0; s < b->ngroup; s++) {
        const int i = b->group[s], k = i - b->w0;
        if(b->status[i] != MJ_OK || b->written[k]) continue;
        if(rv != MJX_OK) continue;
        mj_jpeg_t m;
        mj_init_jpeg(&m);
        int r2 = mj_read_jpeg_from_memory(&m, b->in[i].data, b->in[i].len, 0);
        if(r2 == MJ_OK) r2 = mj_compose(&m, d, align, offset_x, offset_y);
        if(r2 == MJ_OK) r2 = mj_write_jpeg_to_memory(&m, &b->out[i].data, &b->out[i].len, 0);
        mj_free_jpeg(&m);
        b->status[i] = r2;
        if(r2 == MJ_OK) b->written[k] = 1;
    }
    for(int k = 0; k < b->w1 - b->w0; k++) mj_free_jpeg(&b->jp[k]);
    b->full = 0;
}

/* collect a window whose device work is queued: files written, images freed */
static void finish_window(batch_t *w, mj_dropon_t *d, unsigned int align, int offset_x, int offset_y, int nthreads, pthread_t *th) {
    if(w->full) window_finish_full(w, w->ctx, &w->dv, d, align, offset_x, offset_y, nthreads, th);
    else {
        group_finish(w, w->ctx, &w->dv, nthreads, th);
        run_phase(w, 4, nthreads, th); /* free; host encode of what the device handed back */
    }
}

/* the pipeline on the calling thread's device */
static int batch_on_device(int n, const mj_blob_t *in, mj_blob_t *out, int *status, mj_dropon_t *d, unsigned int align, int offset_x,
                           int offset_y, int write_options, int nthreads) {
    const int compose = d->blend != MJ_BLEND_NONE && d->image != NULL && d->alpha != NULL;
    const double t_call = now_s();
    mjx_ctx     *ctx = compose ? mjx_host_ctx() : NULL;
    if(compose && ctx == NULL) return MJ_ERR_DEVICE;

    /* a plain baseline file wanted (and MJX_GPU_HUFFMAN not 0): planes, blend and entropy coding of a group stay on the device;
     * and unless MJX_GPU_DECODE=0, windows the device can also DEcode never leave it (window_enqueue_full) */
    const int on_device = compose && write_options == 0 && mjp_gpu_huffman_mode() != 0;
    int       full_mode = on_device;
    {
        const char *e = getenv("MJX_GPU_DECODE");
        if(e != NULL && *e == '0') full_mode = 0;
    }
    int window = 4 * nthreads;
    if(window > 256) window = 256;
    if(full_mode && window < 256) window = 256; /* one CTA per image decodes: the device wants a few hundred images at a time */
    {
        const char *e = getenv("MJX_BATCH_WINDOW"); /* images per window (16 .. 4096) */
        if(e != NULL && atoi(e) >= 16 && atoi(e) <= 4096) window = atoi(e);
    }
    if(window > n) window = n;
    /* two window states: while the device works on one window (queued by group_enqueue), the pool decodes the next */
    batch_t B[2];
    memset(B, 0, sizeof(B));
    int result = MJ_OK;
    for(int k = 0; k < 2; k++) {
        batch_t *b = &B[k];
        b->n = n, b->in = in, b->out = out, b->status = status, b->write_options = write_options;
        pthread_mutex_init(&b->lock, NULL);
        b->jp = (mj_jpeg_t *)calloc((size_t)window, sizeof(mj_jpeg_t));
        b->group_buf = (int *)malloc(sizeof(int) * (size_t)window);
        b->written = (char *)calloc((size_t)window, 1);
        b->ent_off = (size_t *)malloc(sizeof(size_t) * (size_t)window);
        b->seg_in_off = (size_t *)malloc(sizeof(size_t) * (size_t)window);
        b->eligible = (char *)calloc((size_t)window, 1);
        b->in_scan = full_mode ? (mjx_scan_t *)malloc(sizeof(mjx_scan_t) * (size_t)window) : NULL;
        if(b->jp == NULL || b->group_buf == NULL || b->written == NULL || b->ent_off == NULL || b->seg_in_off == NULL || b->eligible == NULL ||
           (full_mode && b->in_scan == NULL))
            result = MJ_ERR_MEMORY;
    }
    char      *done = (char *)malloc((size_t)window);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    mjx_host_image_t *items = (mjx_host_image_t *)malloc(sizeof(mjx_host_image_t) * (size_t)window);
    mjx_dropon    *cd = NULL;
    mjx_layout_t   cd_layout;
    mjx_geometry_t cd_g;
    batch_t       *pend[2] = {NULL, NULL}; /* windows whose device work is queued and whose files are still to be written */
    int            cur = 0;
    mjx_ctx       *ctx2 = NULL;
    memset(&cd_layout, 0, sizeof(cd_layout));
    memset(&cd_g, 0, sizeof(cd_g));
    if(result != MJ_OK || done == NULL || th == NULL || items == NULL) {
        result = MJ_ERR_MEMORY;
        goto out;
    }
    if(full_mode && n > window) ctx2 = mjp_host_ctx2(); /* (NULL: one window in flight) */

    /* Two windows may be in flight: pend[k] = the window of state B[k] whose device work is queued.  Windows alternate between
     * the states -- and, where the device decodes them, between the thread's two contexts, so that the device works on one
     * window while the host collects the window before it and prepares the next. */
    for(int w0 = 0; w0 < n;) {
        const int si = cur;
        batch_t  *bp = &B[si];
        if(pend[si] != NULL) { /* two windows ago: its state, context and buffers are needed again */
            finish_window(pend[si], d, align, offset_x, offset_y, nthreads, th);
            pend[si] = NULL;
        }
        bp->w0 = w0;
        bp->w1 = w0 + window < n ? w0 + window : n;
        w0 = bp->w1;
        if(full_mode) {
            bp->ctx = (ctx2 != NULL && si == 1) ? ctx2 : ctx;
            if(pend[si ^ 1] != NULL && pend[si ^ 1]->ctx == bp->ctx) { /* (one context only: one window in flight) */
                finish_window(pend[si ^ 1], d, align, offset_x, offset_y, nthreads, th);
                pend[si ^ 1] = NULL;
            }
            run_phase(bp, 7, nthreads, th); /* markers only */
            memset(bp->written, 0, (size_t)window);
            if(window_enqueue_full(bp, bp->ctx, &bp->dv, &cd, &cd_layout, &cd_g, d, align, offset_x, offset_y, nthreads, th, &pend[si ^ 1]) == MJ_OK) {
                pend[si] = bp;
                cur ^= 1;
                continue;
            }
            /* not a window of that kind: read it in full */
        }
        bp->ctx = ctx; /* the ordinary path lives on the thread's first context */
        run_phase(bp, 1, nthreads, th); /* entropy decode -- the device works on the window before meanwhile */
        if(pend[si ^ 1] != NULL) {
            finish_window(pend[si ^ 1], d, align, offset_x, offset_y, nthreads, th);
            pend[si ^ 1] = NULL;
        }
        devbufs_t *const dvp = &bp->dv;

        /* compose the window group by group (images sharing one geometry share one compiled dropon and one launch) */
        memset(done, 0, (size_t)window);
        memset(bp->written, 0, (size_t)window);
        int deferred = 0;
        for(int k0 = 0; compose && k0 < bp->w1 - bp->w0; k0++) {
            if(done[k0] || status[bp->w0 + k0] != MJ_OK) continue;
            mj_jpeg_t *ref = &bp->jp[k0];
            int       *group = bp->group_buf;
            bp->ngroup = 0;
            for(int k = k0; k < bp->w1 - bp->w0; k++)
                if(!done[k] && status[bp->w0 + k] == MJ_OK && same_geometry(ref, &bp->jp[k])) {
                    group[bp->ngroup++] = bp->w0 + k;
                    done[k] = 1;
                }
            mjx_geometry(ref->width, ref->height, ref->sampling.h_factor, ref->sampling.v_factor, d->width, d->height, align, offset_x,
                         offset_y, &bp->g);
            if(!bp->g.visible) continue; /* dropon entirely off these images (reference: src/compose.c:136) */
            mjx_layout_t layout;
            int          rv = mjx_jpeg_layout(ref, &layout);
            /* one compiled dropon is kept across groups and windows while layout and placement repeat */
            if(rv == MJ_OK && cd != NULL && (memcmp(&layout, &cd_layout, sizeof(layout)) != 0 || memcmp(&bp->g, &cd_g, sizeof(bp->g)) != 0)) {
                mjx_dropon_free(cd);
                cd = NULL;
            }
            if(rv == MJ_OK && cd == NULL) {
                cd_layout = layout;
                cd_g = bp->g;
                rv = mjx_dropon_compile(ctx, &cd, d->image, d->alpha, d->width, d->height, d->colorspace, &layout, bp->g.blockoffset_x,
                                        bp->g.blockoffset_y, bp->g.crop_x, bp->g.crop_y, bp->g.crop_w, bp->g.crop_h, 0);
                if(rv == MJX_ERR_UNSUPPORTED) fprintf(stderr, "Unsupported color conversion request\n");
                rv = mjp_map_error(rv);
            }
            if(rv == MJ_OK && on_device) {
                bp->ncomp = layout.ncomp;
                bp->group = group;
                if(group_enqueue(bp, ctx, dvp, cd, nthreads, th) == MJ_OK) {
                    /* the last group of the window is left running while the next window is decoded; a group with others
                     * behind it in this window is finished at once (its slab and device buffers are needed again) */
                    int more = 0;
                    for(int k = k0 + 1; k < bp->w1 - bp->w0; k++) more |= !done[k] && status[bp->w0 + k] == MJ_OK;
                    if(more) group_finish(bp, ctx, dvp, nthreads, th);
                    else deferred = 1;
                    continue;
                }
                /* (not a file the device codes, or no memory for it: the ordinary path below) */
                int bad = 0;
                for(int s = 0; s < bp->ngroup; s++) bad |= status[group[s]] != MJ_OK;
                if(bad) rv = MJ_ERR_NULL_DATA;
            }
            if(rv == MJ_OK) {
                bp->ncomp = layout.ncomp;
                bp->region_bytes = 0;
                for(int c = 0; c < bp->ncomp; c++) {
                    mjx_dropon_dims(cd, c, &bp->wb[c], &bp->hb[c]);
                    bp->comp_off[c] = bp->region_bytes;
                    bp->region_bytes += ((size_t)bp->wb[c] * (size_t)bp->hb[c] * 128 + 255) & ~(size_t)255;
                }
                void *slab = NULL;
                rv = mjp_map_error(mjx_ctx_pinned_scratch(ctx, bp->region_bytes * (size_t)bp->ngroup, &slab));
                bp->slab = (char *)slab;
            }
            if(rv == MJ_OK) {
                bp->group = group;
                run_phase(bp, 2, nthreads, th); /* rows under the dropon -> page-locked slab */
                for(int s = 0; s < bp->ngroup; s++) {
                    mj_jpeg_t *m = &bp->jp[group[s] - bp->w0];
                    memset(&items[s], 0, sizeof(items[s]));
                    for(int c = 0; c < bp->ncomp; c++) {
                        items[s].plane[c] = (int16_t *)(bp->slab + (size_t)s * bp->region_bytes + bp->comp_off[c]);
                        items[s].stride_blocks[c] = items[s].wreal[c] = bp->wb[c];
                        items[s].rows[c] = items[s].hreal[c] = bp->hb[c];
                        items[s].q[c] = m->cinfo.comp_info[c].quant_table ? m->cinfo.comp_info[c].quant_table->quantval : NULL;
                        if(items[s].q[c] == NULL) status[group[s]] = MJ_ERR_NULL_DATA;
                    }
                }
                int ok = 1;
                for(int s = 0; s < bp->ngroup; s++) ok &= status[group[s]] == MJ_OK;
                if(ok) {
                    const double tk = now_s();
                    rv = mjx_compose_batch_host(ctx, items, bp->ngroup, cd, 0, 0); /* K2: one launch for the group */
                    bp->phase_s[0] += now_s() - tk;
                    if(rv != MJX_OK) fprintf(stderr, "libmodjpeg (B200): batch compose failed: %s\n", mjx_ctx_last_error(ctx));
                    rv = mjp_map_error(rv);
                }
                else rv = MJ_ERR_NULL_DATA;
                if(rv == MJ_OK) run_phase(bp, 3, nthreads, th); /* slab -> libjpeg's arrays */
            }
            if(rv != MJ_OK)
                for(int s = 0; s < bp->ngroup; s++)
                    if(status[group[s]] == MJ_OK) status[group[s]] = rv;
        }
        if(deferred) {
            bp->full = 0;
            pend[si] = bp; /* its files are written after the next window's decode */
            cur ^= 1;
        }
        else run_phase(bp, 4, nthreads, th); /* entropy encode + free */
    }
    for(int k = 0; k < 2; k++) { /* the older window first */
        batch_t *w = pend[cur ^ k];
        if(w != NULL) finish_window(w, d, align, offset_x, offset_y, nthreads, th);
        pend[cur ^ k] = NULL;
    }
    if(getenv("MJ_BATCH_TRACE") != NULL) {
        double ps[9];
        for(int k = 0; k < 9; k++) ps[k] = B[0].phase_s[k] + B[1].phase_s[k];
        fprintf(stderr, "mj_compose_batch: %d images, %d threads, %.3f s: headers %.3f s, decode %.3f s, stage-in %.3f s, %s %.3f s, stage-out %.3f s, %s %.3f s\n", n,
                nthreads, now_s() - t_call, ps[7], ps[1], ps[2] + ps[5] + ps[8], on_device ? "waited for the device (copies + K2 + K4 not hidden by the next decode)" : "K2", ps[0], ps[3],
                on_device ? "files (markers + segment) + host encode" : "encode", ps[4] + ps[6]);
    }
out:
    if(ctx != NULL) mjx_ctx_sync(ctx);
    if(ctx2 != NULL) mjx_ctx_sync(ctx2);
    if(cd != NULL) mjx_dropon_free(cd);
    for(int k = 0; k < 2; k++) {
        free(B[k].jp);
        free(B[k].group_buf);
        free(B[k].written);
        free(B[k].ent_off);
        free(B[k].seg_in_off);
        free(B[k].eligible);
        free(B[k].in_scan);
        pthread_mutex_destroy(&B[k].lock);
    }
    free(done);
    free(th);
    free(items);
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
