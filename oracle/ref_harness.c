/*
 * ref_harness.c -- TEST INFRASTRUCTURE.  Plane access for an mj_jpeg_t, whichever library
 * filled it (the reference build in oracle/_ref or the product): copies libjpeg's virtual
 * coefficient arrays to/from flat int16 [rows][cols][64] buffers so Python can compare them.
 * Works on both because the struct layout is the public one (reference: src/libmodjpeg.h:99-107)
 * and all access goes through cinfo.mem->access_virt_barray like the reference's own loops
 * (reference: src/compose.c:269, src/effect.c:48).
 */
#include <string.h>

#include "libmodjpeg.h"

int mjh_sizeof_jpeg(void) { return (int)sizeof(mj_jpeg_t); }
int mjh_sizeof_dropon(void) { return (int)sizeof(mj_dropon_t); }
int mjh_sizeof_decompress(void) { return (int)sizeof(struct jpeg_decompress_struct); }
int mjh_sizeof_compress(void) { return (int)sizeof(struct jpeg_compress_struct); }
int mjh_sizeof_component(void) { return (int)sizeof(jpeg_component_info); }
int mjh_sizeof_error_mgr(void) { return (int)sizeof(struct jpeg_error_mgr); }

int mjh_offsetof_coef(void) { return (int)offsetof(mj_jpeg_t, coef); }
int mjh_offsetof_width(void) { return (int)offsetof(mj_jpeg_t, width); }
int mjh_offsetof_sampling(void) { return (int)offsetof(mj_jpeg_t, sampling); }

/* info[0..3] = ncomp, jpeg_color_space, width, height; info[4..] = max_h, max_v */
int mjh_image_info(mj_jpeg_t *m, int *info) {
    if(m == NULL || m->coef == NULL) return -1;
    info[0] = m->cinfo.num_components;
    info[1] = (int)m->cinfo.jpeg_color_space;
    info[2] = m->width;
    info[3] = m->height;
    info[4] = m->cinfo.max_h_samp_factor;
    info[5] = m->cinfo.max_v_samp_factor;
    return 0;
}

static unsigned round_up(unsigned a, unsigned b) { return ((a + b - 1) / b) * b; }

/* info = width_in_blocks, height_in_blocks, h_samp, v_samp, virt_width, virt_height */
int mjh_comp_info(mj_jpeg_t *m, int c, int *info) {
    if(m == NULL || m->coef == NULL || c < 0 || c >= m->cinfo.num_components) return -1;
    jpeg_component_info *ci = &m->cinfo.comp_info[c];
    info[0] = (int)ci->width_in_blocks;
    info[1] = (int)ci->height_in_blocks;
    info[2] = ci->h_samp_factor;
    info[3] = ci->v_samp_factor;
    /* jdcoefct.c allocates the arrays rounded up to the sampling factors */
    info[4] = (int)round_up(ci->width_in_blocks, (unsigned)ci->h_samp_factor);
    info[5] = (int)round_up(ci->height_in_blocks, (unsigned)ci->v_samp_factor);
    return 0;
}

int mjh_qtable(mj_jpeg_t *m, int c, unsigned short *q) {
    if(m == NULL || m->coef == NULL || c < 0 || c >= m->cinfo.num_components) return -1;
    if(m->cinfo.comp_info[c].quant_table == NULL) return -2;
    memcpy(q, m->cinfo.comp_info[c].quant_table->quantval, 64 * sizeof(unsigned short));
    return 0;
}

static int plane_copy(mj_jpeg_t *m, int c, short *buf, int to_buf) {
    int info[6];
    if(mjh_comp_info(m, c, info) != 0) return -1;
    int vw = info[4], vh = info[5];
    for(int r = 0; r < vh; r++) {
        JBLOCKARRAY rows = (*m->cinfo.mem->access_virt_barray)((j_common_ptr)&m->cinfo, m->coef[c], (JDIMENSION)r, 1, TRUE);
        if(to_buf) memcpy(buf + (size_t)r * vw * 64, rows[0], (size_t)vw * 128);
        else memcpy(rows[0], buf + (size_t)r * vw * 64, (size_t)vw * 128);
    }
    return 0;
}

int mjh_export_plane(mj_jpeg_t *m, int c, short *dst) { return plane_copy(m, c, dst, 1); }
int mjh_import_plane(mj_jpeg_t *m, int c, const short *src) { return plane_copy(m, c, (short *)src, 0); }
