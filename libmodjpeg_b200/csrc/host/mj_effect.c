/*
 * mj_effect.c -- the four coefficient effects of the host boundary
 * (reference: src/effect.c:28-222).  Each call builds a short step list for K3 and runs it
 * through mjx_effects_rows_host on the REAL blocks of the affected components; the colour-space
 * gates and return codes are the reference's.
 */
#include <stdio.h>
#include <stdlib.h>

#include "mj_private.h"

static int run_effect(mj_jpeg_t *m, const mjx_effect_op_t *ops, int nops) {
    const int ncomp = m->cinfo.num_components;
    if(ncomp > MJX_MAX_COMPONENTS) return MJ_ERR_UNSUPPORTED_COLORSPACE;
    mjx_ctx *ctx = mjx_host_ctx();
    if(ctx == NULL) return MJ_ERR_DEVICE;

    int16_t       **rows[MJX_MAX_COMPONENTS] = {NULL, NULL, NULL, NULL};
    const uint16_t *q[MJX_MAX_COMPONENTS] = {NULL, NULL, NULL, NULL};
    int             wreal[MJX_MAX_COMPONENTS] = {0}, hreal[MJX_MAX_COMPONENTS] = {0};
    int             used[MJX_MAX_COMPONENTS] = {0};
    for(int i = 0; i < nops; i++) used[ops[i].comp] = 1;

    mjp_trap_t *trap = mjp_image_trap(m);
    int         result = MJ_OK;
    trap->armed = 1;
    if(setjmp(trap->escape)) {
        result = MJ_ERR_DECODE_JPEG;
        goto done;
    }
    for(int c = 0; c < ncomp; c++) {
        if(!used[c]) continue;
        jpeg_component_info *ci = &m->cinfo.comp_info[c];
        if(ci->quant_table == NULL) {
            result = MJ_ERR_NULL_DATA;
            goto done;
        }
        q[c] = ci->quant_table->quantval;
        wreal[c] = (int)ci->width_in_blocks;
        hreal[c] = (int)ci->height_in_blocks;
        rows[c] = (int16_t **)malloc(sizeof(int16_t *) * (size_t)(hreal[c] > 0 ? hreal[c] : 1));
        if(rows[c] == NULL) {
            result = MJ_ERR_MEMORY;
            goto done;
        }
        for(int l = 0; l < hreal[c]; l++) {
            JBLOCKARRAY ba = (*m->cinfo.mem->access_virt_barray)((j_common_ptr)&m->cinfo, m->coef[c], (JDIMENSION)l, 1, TRUE);
            rows[c][l] = &ba[0][0][0];
        }
    }
    {
        int rv = mjx_effects_rows_host(ctx, ncomp, (int16_t *const *const *)rows, wreal, hreal, q, ops, nops);
        if(rv != MJX_OK) {
            fprintf(stderr, "libmodjpeg (B200): effect failed: %s\n", mjx_ctx_last_error(ctx));
            result = mjp_map_error(rv);
        }
    }
done:
    trap->armed = 0;
    for(int c = 0; c < MJX_MAX_COMPONENTS; c++) free(rows[c]);
    return result;
}

int mj_effect_grayscale(mj_jpeg_t *m) {
    if(m == NULL || m->coef == NULL) return MJ_ERR_NULL_DATA;
    if(m->cinfo.jpeg_color_space != JCS_YCbCr) return MJ_OK; /* effect.c:39 */
    mjx_effect_op_t ops[MJX_MAX_COMPONENTS];
    int             n = 0;
    for(int c = 1; c < m->cinfo.num_components && c < MJX_MAX_COMPONENTS; c++) {
        ops[n].op = MJX_FX_ZERO, ops[n].comp = c, ops[n].value = 0;
        n++;
    }
    return n ? run_effect(m, ops, n) : MJ_OK;
}

int mj_effect_pixelate(mj_jpeg_t *m) {
    if(m == NULL || m->coef == NULL) return MJ_ERR_NULL_DATA;
    mjx_effect_op_t ops[MJX_MAX_COMPONENTS];
    int             n = 0;
    for(int c = 0; c < m->cinfo.num_components && c < MJX_MAX_COMPONENTS; c++) {
        ops[n].op = MJX_FX_PIXELATE, ops[n].comp = c, ops[n].value = 0;
        n++;
    }
    return n ? run_effect(m, ops, n) : MJ_OK;
}

int mj_effect_tint(mj_jpeg_t *m, int cb_value, int cr_value) {
    if(m == NULL || m->coef == NULL) return MJ_ERR_NULL_DATA;
    if(m->cinfo.jpeg_color_space != JCS_YCbCr) return MJ_OK; /* effect.c:126 */
    mjx_effect_op_t ops[2];
    int             n = 0;
    if(cb_value != 0) {
        ops[n].op = MJX_FX_ADD_DC, ops[n].comp = 1, ops[n].value = cb_value;
        n++;
    }
    if(cr_value != 0) {
        ops[n].op = MJX_FX_ADD_DC, ops[n].comp = 2, ops[n].value = cr_value;
        n++;
    }
    return n ? run_effect(m, ops, n) : MJ_OK; /* both zero: effect.c:130 */
}

int mj_effect_luminance(mj_jpeg_t *m, int value) {
    if(m == NULL || m->coef == NULL) return MJ_ERR_NULL_DATA;
    if(m->cinfo.jpeg_color_space != JCS_YCbCr) return MJ_OK; /* effect.c:195 */
    mjx_effect_op_t op;
    op.op = MJX_FX_ADD_DC, op.comp = 0, op.value = value;
    return run_effect(m, &op, 1);
}
