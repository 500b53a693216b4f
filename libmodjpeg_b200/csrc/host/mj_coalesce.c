/*
 * mj_coalesce.c -- request coalescer: concurrent mj_compose calls that stamp the SAME dropon the same way are gathered
 * into one K2 launch (the caller it is made for is a request server such as the nginx filter,
 * reference: README.md:406-408, and src/contrib/modjpeg.c:79-99 is what each request does).
 *
 * Opt-in (mj_coalesce_configure, or MJX_COALESCE=1 in the environment) because it trades latency for throughput: the first
 * request of a batch -- the leader -- waits up to `wait_us` for others to join.  Every request stages the rows under the
 * dropon into its region of one page-locked slab (in parallel, outside the lock); the leader then runs ONE
 * mjx_compose_batch_host over the slab (the kernel works on the host memory directly, only touched blocks cross PCIe) and
 * wakes the others; every request copies its region back into its own libjpeg arrays.  The compiled dropon is shared by
 * all requests and kept while the key (dropon buffers + dimensions, target layout, placement remainder) repeats, so K1
 * runs once per key instead of once per request.  Results are byte-identical to unbatched calls: the same kernels see the
 * same blocks (tests/test_gpu_parity.py::test_coalesced_compose_equals_unbatched).
 */
#include <errno.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "mj_private.h"

#define MJ_COALESCE_MAX 256

typedef struct {
    const void    *image, *alpha;
    int            width, height, colorspace, blend;
    unsigned long  generation;
    mjx_layout_t   layout;
    mjx_geometry_t g;
} ckey_t;

static struct {
    pthread_mutex_t mu;
    pthread_cond_t  cv; /* one condition for every state change: batches are small, wake-ups cheap */
    int             enabled, max_batch, wait_us;
    /* compiled dropon of the current key (survives batches) */
    int             have_key;
    ckey_t          key;
    mjx_dropon     *cd;
    int             ncomp, wb[MJX_MAX_COMPONENTS], hb[MJX_MAX_COMPONENTS];
    size_t          comp_off[MJX_MAX_COMPONENTS], region_bytes;
    /* page-locked slab, max_batch regions */
    char           *slab;
    size_t          slab_bytes;
    /* the batch in flight */
    int             open;     /* a batch exists (joinable until `closed`) */
    int             closed;   /* the leader stopped admitting */
    int             count, staged, done, left;
    int             rv;       /* result of the launch */
    mjx_host_image_t items[MJ_COALESCE_MAX];
    int             item_ok[MJ_COALESCE_MAX];
    unsigned long   batches, requests; /* statistics */
} C = {PTHREAD_MUTEX_INITIALIZER, PTHREAD_COND_INITIALIZER, -1, 32, 200};

extern unsigned long mjp_dropon_generation;

void mj_coalesce_configure(int enable, int max_batch, int wait_us) {
    pthread_mutex_lock(&C.mu);
    C.enabled = enable ? 1 : 0;
    if(max_batch > 0) C.max_batch = max_batch < MJ_COALESCE_MAX ? max_batch : MJ_COALESCE_MAX;
    if(wait_us >= 0) C.wait_us = wait_us;
    pthread_mutex_unlock(&C.mu);
}

void mj_coalesce_stats(unsigned long *batches, unsigned long *requests) {
    pthread_mutex_lock(&C.mu);
    if(batches) *batches = C.batches;
    if(requests) *requests = C.requests;
    pthread_mutex_unlock(&C.mu);
}

int mjp_coalesce_enabled(void) {
    if(C.enabled < 0) {
        pthread_mutex_lock(&C.mu);
        if(C.enabled < 0) {
            const char *e = getenv("MJX_COALESCE");
            C.enabled = (e != NULL && atoi(e) == 1) ? 1 : 0;
            if((e = getenv("MJX_COALESCE_MAX")) != NULL && atoi(e) > 0) C.max_batch = atoi(e) < MJ_COALESCE_MAX ? atoi(e) : MJ_COALESCE_MAX;
            if((e = getenv("MJX_COALESCE_WAIT_US")) != NULL && atoi(e) >= 0) C.wait_us = atoi(e);
        }
        pthread_mutex_unlock(&C.mu);
    }
    return C.enabled;
}

static int key_equal(const ckey_t *a, const ckey_t *b) {
    return a->image == b->image && a->alpha == b->alpha && a->width == b->width && a->height == b->height && a->colorspace == b->colorspace &&
           a->blend == b->blend && a->generation == b->generation && memcmp(&a->layout, &b->layout, sizeof(a->layout)) == 0 &&
           a->g.blockoffset_x == b->g.blockoffset_x && a->g.blockoffset_y == b->g.blockoffset_y && a->g.crop_x == b->g.crop_x &&
           a->g.crop_y == b->g.crop_y && a->g.crop_w == b->g.crop_w && a->g.crop_h == b->g.crop_h;
}

/* rows under the dropon: libjpeg's arrays <-> the request's slab region (one row is live at a time) */
static int move_rows(mj_jpeg_t *m, const mjx_geometry_t *g, char *region, int to_slab) {
    mjp_trap_t *trap = mjp_image_trap(m);
    trap->armed = 1;
    if(setjmp(trap->escape)) {
        trap->armed = 0;
        return MJ_ERR_DECODE_JPEG;
    }
    for(int c = 0; c < C.ncomp; c++) {
        jpeg_component_info *ci = &m->cinfo.comp_info[c];
        const unsigned       x0 = (unsigned)(g->block_x * ci->h_samp_factor), y0 = (unsigned)(g->block_y * ci->v_samp_factor);
        const size_t         wbytes = (size_t)C.wb[c] * 128;
        if(ci->quant_table == NULL || x0 + (unsigned)C.wb[c] > mjp_virtual_width(ci) || y0 + (unsigned)C.hb[c] > mjp_virtual_height(ci)) {
            trap->armed = 0;
            return MJ_ERR_DROPON_DIMENSIONS;
        }
        for(int l = 0; l < C.hb[c]; l++) {
            JBLOCKARRAY ba = (*m->cinfo.mem->access_virt_barray)((j_common_ptr)&m->cinfo, m->coef[c], y0 + (unsigned)l, 1, TRUE);
            char       *row = (char *)&ba[0][x0][0], *st = region + C.comp_off[c] + (size_t)l * wbytes;
            if(to_slab) memcpy(st, row, wbytes);
            else memcpy(row, st, wbytes);
        }
    }
    trap->armed = 0;
    return MJ_OK;
}

/* returns 1 when the request was served here (*result holds mj_compose's return value), 0 when the caller should take
 * the ordinary path */
int mjp_coalesce_compose(mj_jpeg_t *m, mj_dropon_t *d, const mjx_layout_t *layout, const mjx_geometry_t *g, int *result) {
    mjx_ctx *ctx = mjx_host_ctx();
    if(ctx == NULL) return 0;
    ckey_t key;
    memset(&key, 0, sizeof(key));
    key.image = d->image, key.alpha = d->alpha, key.width = d->width, key.height = d->height, key.colorspace = d->colorspace, key.blend = d->blend;
    key.generation = __atomic_load_n(&mjp_dropon_generation, __ATOMIC_RELAXED);
    key.layout = *layout;
    key.g = *g;

    pthread_mutex_lock(&C.mu);
    /* a batch that no longer admits (or one for another key) has to drain first */
    while(C.open && (C.closed || C.count >= C.max_batch || !key_equal(&key, &C.key))) pthread_cond_wait(&C.cv, &C.mu);
    int slot, leader = 0;
    if(!C.open) {
        /* open a batch: compile the dropon for this key unless the last batch used the same one */
        if(!C.have_key || !key_equal(&key, &C.key)) {
            if(C.cd != NULL) mjx_dropon_free(C.cd);
            C.cd = NULL;
            C.have_key = 0;
            int rv = mjx_dropon_compile(ctx, &C.cd, d->image, d->alpha, d->width, d->height, d->colorspace, layout, g->blockoffset_x,
                                        g->blockoffset_y, g->crop_x, g->crop_y, g->crop_w, g->crop_h, 0);
            if(rv == MJX_OK) rv = mjx_ctx_sync(ctx); /* other threads' contexts will use it */
            if(rv != MJX_OK) {
                pthread_mutex_unlock(&C.mu);
                return 0; /* the ordinary path reports the error */
            }
            C.key = key;
            C.have_key = 1;
            C.ncomp = layout->ncomp;
            C.region_bytes = 0;
            for(int c = 0; c < C.ncomp; c++) {
                mjx_dropon_dims(C.cd, c, &C.wb[c], &C.hb[c]);
                C.comp_off[c] = C.region_bytes;
                C.region_bytes += ((size_t)C.wb[c] * (size_t)C.hb[c] * 128 + 255) & ~(size_t)255;
            }
        }
        const size_t need = C.region_bytes * (size_t)C.max_batch;
        if(need > C.slab_bytes) {
            if(C.slab != NULL) mjx_host_free(ctx, C.slab); /* portable page-locked memory: any context may release it */
            C.slab = NULL, C.slab_bytes = 0;
            void *p = NULL;
            if(mjx_host_alloc(ctx, &p, need) != MJX_OK) {
                pthread_mutex_unlock(&C.mu);
                return 0;
            }
            C.slab = (char *)p, C.slab_bytes = need;
        }
        C.open = 1, C.closed = 0, C.count = 0, C.staged = 0, C.done = 0, C.left = 0, C.rv = MJX_OK;
        leader = 1;
    }
    slot = C.count++;
    C.requests++;
    const mjx_geometry_t gg = *g; /* (block_x / block_y are the request's own: the key only fixes what the compile depends on) */
    pthread_cond_broadcast(&C.cv); /* the leader watches the count */
    pthread_mutex_unlock(&C.mu);

    /* stage my rows (parallel across the batch's threads) */
    char *region = C.slab + (size_t)slot * C.region_bytes;
    int   st = move_rows(m, &gg, region, 1);
    mjx_host_image_t it;
    memset(&it, 0, sizeof(it));
    for(int c = 0; c < C.ncomp; c++) {
        it.plane[c] = (int16_t *)(region + C.comp_off[c]);
        it.stride_blocks[c] = it.wreal[c] = C.wb[c];
        it.rows[c] = it.hreal[c] = C.hb[c];
        it.q[c] = m->cinfo.comp_info[c].quant_table ? m->cinfo.comp_info[c].quant_table->quantval : NULL;
        if(it.q[c] == NULL && st == MJ_OK) st = MJ_ERR_NULL_DATA;
    }

    pthread_mutex_lock(&C.mu);
    C.items[slot] = it;
    C.item_ok[slot] = st == MJ_OK;
    C.staged++;
    pthread_cond_broadcast(&C.cv);
    if(leader) {
        /* admit until the batch is full or the wait is over, then wait for the admitted to finish staging */
        struct timespec until;
        clock_gettime(CLOCK_REALTIME, &until);
        until.tv_nsec += (long)C.wait_us * 1000L;
        until.tv_sec += until.tv_nsec / 1000000000L;
        until.tv_nsec %= 1000000000L;
        while(C.count < C.max_batch)
            if(pthread_cond_timedwait(&C.cv, &C.mu, &until) == ETIMEDOUT) break;
        C.closed = 1;
        while(C.staged < C.count) pthread_cond_wait(&C.cv, &C.mu);
        const int n = C.count;
        pthread_mutex_unlock(&C.mu);
        /* one launch over the requests that staged cleanly (compacted in place; slots keep their regions) */
        mjx_host_image_t run[MJ_COALESCE_MAX];
        int              nrun = 0;
        for(int s = 0; s < n; s++)
            if(C.item_ok[s]) run[nrun++] = C.items[s];
        int rv = nrun > 0 ? mjx_compose_batch_host(ctx, run, nrun, C.cd, 0, 0) : MJX_OK;
        if(rv != MJX_OK) fprintf(stderr, "libmodjpeg (B200): coalesced compose failed: %s\n", mjx_ctx_last_error(ctx));
        pthread_mutex_lock(&C.mu);
        C.rv = rv;
        C.done = 1;
        C.batches++;
        pthread_cond_broadcast(&C.cv);
    }
    else
        while(!C.done) pthread_cond_wait(&C.cv, &C.mu);
    const int rv = C.rv;
    pthread_mutex_unlock(&C.mu);

    if(st == MJ_OK && rv == MJX_OK) st = move_rows(m, &gg, region, 0); /* my region back into my image */
    else if(st == MJ_OK) st = mjp_map_error(rv);

    pthread_mutex_lock(&C.mu);
    if(++C.left == C.count) { /* last one out: the slab is free again */
        C.open = 0;
        pthread_cond_broadcast(&C.cv);
    }
    pthread_mutex_unlock(&C.mu);
    *result = st;
    return 1;
}
