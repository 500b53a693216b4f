/*
 * mj_compose.c -- mj_compose of the host boundary (reference: src/compose.c:33-180).
 * Host work: argument checks, the placement arithmetic (mjx_geometry), collecting libjpeg's row
 * pointers for the region under the dropon.  Device work: K1 compiles the dropon for this
 * image's colour space / sampling / block offset (mjx_dropon_compile), K2 blends it into the
 * staged region (mjx_compose_rows_host).  Only rows under the dropon are touched or copied.
 */
#include <stdio.h>
#include <stdlib.h>

#include "mj_private.h"

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

    mjx_dropon *cd = NULL;
    rv = mjx_dropon_compile(ctx, &cd, d->image, d->alpha, d->width, d->height, d->colorspace, &layout, g.blockoffset_x,
                            g.blockoffset_y, g.crop_x, g.crop_y, g.crop_w, g.crop_h, 0);
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
    mjx_dropon_free(cd);
    return result;
}
