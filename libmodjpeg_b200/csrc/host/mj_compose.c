/*
 * mj_compose.c -- mj_compose of the host boundary (reference: src/compose.c:33-180).
 * Host work: argument checks, the placement arithmetic (mjx_geometry), collecting libjpeg's row
 * pointers for the region under the dropon.  Device work: K1 compiles the dropon for this
 * image's colour space / sampling / block offset (mjx_dropon_compile), K2 blends it into the
 * staged region (mjx_compose_rows_host).  Only rows under the dropon are touched or copied.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mj_private.h"

/*
 * Compiled-dropon cache (opt-in: MJX_DROPON_CACHE=1 in the environment).  The reference compiles the dropon again on
 * every mj_compose (src/compose.c:155-177); a server that puts ONE logo on many images pays K1 + the pixel upload per
 * call for nothing.  With the cache on, the calling thread keeps its last few compiled dropons keyed by everything the
 * compile depends on: the dropon's buffers and dimensions, the target layout and the placement remainder.  The key
 * holds the buffer ADDRESSES, not their content, so the contract is: do not edit d->image / d->alpha in place after
 * mj_read_dropon_* (re-reading or freeing a dropon is fine: mj_free_dropon bumps the generation below).
 */
#define MJ_CACHE_SLOTS 4
typedef struct {
    const void    *image, *alpha;
    int            width, height, colorspace, blend;
    unsigned long  generation;
    mjx_layout_t   layout;
    mjx_geometry_t g; /* only the fields the compile uses are compared */
    mjx_dropon    *cd;
    unsigned long  last_use;
} cache_slot_t;

static __thread cache_slot_t  t_cache[MJ_CACHE_SLOTS];
static __thread unsigned long t_clock;
unsigned long                 mjp_dropon_generation = 1; /* bumped by mj_free_dropon / mj_read_dropon_* (mj_dropon.c) */

static int cache_enabled(void) {
    static int on = -1;
    if(on < 0) {
        const char *e = getenv("MJX_DROPON_CACHE");
        on = (e != NULL && atoi(e) == 1) ? 1 : 0;
    }
    return on;
}

static int slot_matches(const cache_slot_t *s, const mj_dropon_t *d, const mjx_layout_t *L, const mjx_geometry_t *g) {
    return s->cd != NULL && s->image == d->image && s->alpha == d->alpha && s->width == d->width && s->height == d->height &&
           s->colorspace == d->colorspace && s->blend == d->blend && s->generation == __atomic_load_n(&mjp_dropon_generation, __ATOMIC_RELAXED) &&
           memcmp(&s->layout, L, sizeof(*L)) == 0 && s->g.blockoffset_x == g->blockoffset_x && s->g.blockoffset_y == g->blockoffset_y &&
           s->g.crop_x == g->crop_x && s->g.crop_y == g->crop_y && s->g.crop_w == g->crop_w && s->g.crop_h == g->crop_h;
}

/* called when the calling thread's engine context goes away (mj_device.c) */
void mjp_compose_cache_clear(void) {
    for(int i = 0; i < MJ_CACHE_SLOTS; i++) {
        if(t_cache[i].cd != NULL) mjx_dropon_free(t_cache[i].cd);
        memset(&t_cache[i], 0, sizeof(t_cache[i]));
    }
}

/* returns a compiled dropon for (d, layout, g): from the cache, or freshly compiled (and cached when enabled);
 * *owned = 1 when the caller has to free it */
static int get_compiled(mjx_ctx *ctx, mj_dropon_t *d, const mjx_layout_t *L, const mjx_geometry_t *g, mjx_dropon **out, int *owned) {
    *owned = 1;
    if(cache_enabled()) {
        for(int i = 0; i < MJ_CACHE_SLOTS; i++)
            if(slot_matches(&t_cache[i], d, L, g)) {
                t_cache[i].last_use = ++t_clock;
                *out = t_cache[i].cd;
                *owned = 0;
                return MJX_OK;
            }
    }
    int rv = mjx_dropon_compile(ctx, out, d->image, d->alpha, d->width, d->height, d->colorspace, L, g->blockoffset_x, g->blockoffset_y,
                                g->crop_x, g->crop_y, g->crop_w, g->crop_h, 0);
    if(rv != MJX_OK || !cache_enabled()) return rv;
    int victim = 0;
    for(int i = 1; i < MJ_CACHE_SLOTS; i++)
        if(t_cache[i].cd == NULL || (t_cache[victim].cd != NULL && t_cache[i].last_use < t_cache[victim].last_use)) victim = i;
    if(t_cache[victim].cd != NULL) mjx_dropon_free(t_cache[victim].cd);
    cache_slot_t *s = &t_cache[victim];
    s->image = d->image, s->alpha = d->alpha, s->width = d->width, s->height = d->height, s->colorspace = d->colorspace, s->blend = d->blend;
    s->generation = __atomic_load_n(&mjp_dropon_generation, __ATOMIC_RELAXED);
    s->layout = *L;
    s->g = *g;
    s->cd = *out;
    s->last_use = ++t_clock;
    *owned = 0;
    return MJX_OK;
}

int mj_compose(mj_jpeg_t *m, mj_dropon_t *d, unsigned int align, int offset_x, int offset_y) {
    if(m == NULL || d == NULL) return MJ_ERR_NULL_DATA;
    if(d->blend == MJ_BLEND_NONE) return MJ_OK; /* fully transparent: nothing to do (compose.c:38) */
    if(m->coef == NULL || d->image == NULL || d->alpha == NULL) return MJ_ERR_NULL_DATA;

    mjx_geometry_t g;
    mjx_geometry(m->width, m->height, m->sampling.h_factor, m->sampling.v_factor, d->width, d->height, align, offset_x, offset_y, &g);
    if(!g.visible) return MJ_OK; /* dropon entirely off the image (compose.c:136) */

    mjx_layout_t layout;
    int          rv = mjx_jpeg_layout(m, &layout);
    if(rv != MJ_OK) return rv;

    mjx_ctx *ctx = mjx_host_ctx();
    if(ctx == NULL) return MJ_ERR_DEVICE;

    if(mjp_coalesce_enabled()) { /* opt-in: concurrent calls with the same dropon share one launch (mj_coalesce.c) */
        int result = MJ_OK;
        if(mjp_coalesce_compose(m, d, &layout, &g, &result)) return result;
    }

    mjx_dropon *cd = NULL;
    int         cd_owned = 1;
    rv = get_compiled(ctx, d, &layout, &g, &cd, &cd_owned);
    if(rv != MJX_OK) {
        if(rv == MJX_ERR_UNSUPPORTED) fprintf(stderr, "Unsupported color conversion request\n"); /* libjpeg's words, as the reference prints them */
        else if(rv == MJX_ERR_DEVICE) fprintf(stderr, "libmodjpeg (B200): %s\n", mjx_ctx_last_error(ctx));
        return mjp_map_error(rv);
    }

    /* row pointers of the region under the dropon, per component (compose.c:264-274) */
    const int ncomp = layout.ncomp;
    int16_t **rows[MJX_MAX_COMPONENTS] = {NULL, NULL, NULL, NULL};
    const uint16_t *q[MJX_MAX_COMPONENTS] = {NULL, NULL, NULL, NULL};
    mjp_trap_t *trap = mjp_image_trap(m);
    int         result = MJ_OK;

    trap->armed = 1;
    if(setjmp(trap->escape)) {
        result = MJ_ERR_DECODE_JPEG;
        goto done;
    }
    for(int c = 0; c < ncomp; c++) {
        jpeg_component_info *ci = &m->cinfo.comp_info[c];
        int                  wb = 0, hb = 0;
        mjx_dropon_dims(cd, c, &wb, &hb);
        const unsigned x0 = (unsigned)(g.block_x * ci->h_samp_factor), y0 = (unsigned)(g.block_y * ci->v_samp_factor);
        if(ci->quant_table == NULL || x0 + (unsigned)wb > mjp_virtual_width(ci) || y0 + (unsigned)hb > mjp_virtual_height(ci)) {
            result = MJ_ERR_DROPON_DIMENSIONS; /* cannot happen for geometry produced above */
            goto done;
        }
        q[c] = ci->quant_table->quantval;
        rows[c] = (int16_t **)malloc(sizeof(int16_t *) * (size_t)(hb > 0 ? hb : 1));
        if(rows[c] == NULL) {
            result = MJ_ERR_MEMORY;
            goto done;
        }
        for(int l = 0; l < hb; l++) {
            JBLOCKARRAY ba = (*m->cinfo.mem->access_virt_barray)((j_common_ptr)&m->cinfo, m->coef[c], y0 + (unsigned)l, 1, TRUE);
            rows[c][l] = &ba[0][x0][0];
        }
    }
    rv = mjx_compose_rows_host(ctx, ncomp, (int16_t *const *const *)rows, q, cd);
    if(rv != MJX_OK) {
        fprintf(stderr, "libmodjpeg (B200): compose failed: %s\n", mjx_ctx_last_error(ctx));
        result = mjp_map_error(rv);
    }

done:
    trap->armed = 0;
    for(int c = 0; c < MJX_MAX_COMPONENTS; c++) free(rows[c]);
    if(cd_owned) mjx_dropon_free(cd);
    return result;
}
