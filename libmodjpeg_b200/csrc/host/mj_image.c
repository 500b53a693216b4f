/*
 * mj_image.c -- JPEG coefficient I/O of the host boundary: mj_init_jpeg, mj_free_jpeg,
 * mj_read_jpeg_from_memory/_file, mj_write_jpeg_to_memory/_file, and the flat plane accessors
 * of mjx_host.h.  Same observable behaviour as reference: src/image.c:33-255 (return codes,
 * marker preservation, output bytes); entropy coding stays on host libjpeg (north_star).
 */
#include <stdlib.h>
#include <string.h>

#include "mj_private.h"

void mj_init_jpeg(mj_jpeg_t *m) {
    if(m != NULL) memset(m, 0, sizeof(*m));
}

void mj_free_jpeg(mj_jpeg_t *m) {
    if(m == NULL) return;
    /* a zeroed cinfo (mem == NULL) is accepted by jpeg_destroy */
    jpeg_destroy_decompress(&m->cinfo);
    mj_init_jpeg(m);
}

mjp_trap_t *mjp_image_trap(mj_jpeg_t *m) { return (mjp_trap_t *)m->cinfo.err; }

int mj_read_jpeg_from_memory(mj_jpeg_t *m, const unsigned char *memory, size_t len, size_t max_pixel) {
    if(m == NULL || memory == NULL || len == 0) return MJ_ERR_NULL_DATA;

    mj_free_jpeg(m);

    /* bootstrap with a trap on the stack, then move trap + source manager into libjpeg's
     * permanent pool so that m->cinfo.err / m->cinfo.src never dangle (the reference leaves
     * them pointing at dead stack slots, src/image.c:44-47) */
    mjp_trap_t boot;
    mjp_trap_init(&boot);
    m->cinfo.err = &boot.base;
    boot.armed = 1;
    if(setjmp(boot.escape)) {
        jpeg_destroy_decompress(&m->cinfo);
        mj_init_jpeg(m);
        return MJ_ERR_DECODE_JPEG;
    }
    jpeg_create_decompress(&m->cinfo);

    mjp_trap_t   *trap = (mjp_trap_t *)(*m->cinfo.mem->alloc_small)((j_common_ptr)&m->cinfo, JPOOL_PERMANENT, sizeof(mjp_trap_t));
    mjp_memsrc_t *src = (mjp_memsrc_t *)(*m->cinfo.mem->alloc_small)((j_common_ptr)&m->cinfo, JPOOL_PERMANENT, sizeof(mjp_memsrc_t));
    mjp_trap_init(trap);
    m->cinfo.err = &trap->base;
    trap->armed = 1;
    if(setjmp(trap->escape)) {
        jpeg_destroy_decompress(&m->cinfo);
        mj_init_jpeg(m);
        return MJ_ERR_DECODE_JPEG;
    }
    mjp_memsrc_init(src, memory, len);
    m->cinfo.src = &src->base;

    /* keep COM and APP0..APP15 so that they can be written back (reference: src/image.c:67-72) */
    jpeg_save_markers(&m->cinfo, JPEG_COM, 0xFFFF);
    for(int k = 0; k < 16; k++) jpeg_save_markers(&m->cinfo, JPEG_APP0 + k, 0xFFFF);

    jpeg_read_header(&m->cinfo, TRUE);
    m->width = (int)m->cinfo.image_width;
    m->height = (int)m->cinfo.image_height;

    int rv = MJ_OK;
    if(max_pixel != 0 && (size_t)m->width * (size_t)m->height > max_pixel) rv = MJ_ERR_IMAGE_SIZE;
    else if(m->cinfo.jpeg_color_space != JCS_GRAYSCALE && m->cinfo.jpeg_color_space != JCS_RGB &&
            m->cinfo.jpeg_color_space != JCS_YCbCr)
        rv = MJ_ERR_UNSUPPORTED_COLORSPACE; /* CMYK / YCCK (reference: src/image.c:84-92) */
    if(rv != MJ_OK) {
        jpeg_destroy_decompress(&m->cinfo);
        mj_init_jpeg(m);
        return rv;
    }

    m->coef = jpeg_read_coefficients(&m->cinfo); /* Huffman / arithmetic decode of every scan */

    mj_sampling_t *s = &m->sampling;
    s->max_h_samp_factor = m->cinfo.max_h_samp_factor;
    s->max_v_samp_factor = m->cinfo.max_v_samp_factor;
    s->h_factor = s->max_h_samp_factor * DCTSIZE;
    s->v_factor = s->max_v_samp_factor * DCTSIZE;
    for(int c = 0; c < m->cinfo.num_components && c < 4; c++) {
        s->samp_factor[c].h_samp_factor = m->cinfo.comp_info[c].h_samp_factor;
        s->samp_factor[c].v_samp_factor = m->cinfo.comp_info[c].v_samp_factor;
    }
    /* all input has been consumed: the caller may release `memory` now */
    src->data = NULL;
    src->size = 0;
    trap->armed = 0;
    return MJ_OK;
}

int mj_read_jpeg_from_file(mj_jpeg_t *m, const char *filename, size_t max_pixel) {
    if(m == NULL) return MJ_ERR_NULL_DATA;
    unsigned char *buffer = NULL;
    size_t         len = 0;
    int            rv = mjp_read_whole_file(&buffer, &len, filename);
    if(rv != MJ_OK) return rv;
    rv = mj_read_jpeg_from_memory(m, buffer, len, max_pixel);
    free(buffer);
    return rv;
}

int mj_write_jpeg_to_memory(mj_jpeg_t *m, unsigned char **memory, size_t *len, int options) {
    if(m == NULL || memory == NULL || len == NULL || m->coef == NULL) return MJ_ERR_NULL_DATA;

    struct jpeg_compress_struct out;
    mjp_trap_t                  trap;
    mjp_memdst_t                dst;
    mjp_trap_t                 *itrap = mjp_image_trap(m);

    mjp_memdst_init(&dst);
    mjp_trap_init(&trap);
    out.err = &trap.base;
    trap.armed = 1;
    itrap->armed = 1;
    /* an error may be raised through either object (the compressor, or the decompressor that
     * owns the coefficient arrays); both land in the same cleanup */
    if(setjmp(trap.escape)) goto failed;
    if(setjmp(itrap->escape)) goto failed;
    jpeg_create_compress(&out);
    out.dest = &dst.base;

    jpeg_copy_critical_parameters(&m->cinfo, &out);
    out.optimize_coding = (options & MJ_OPTION_OPTIMIZE) ? TRUE : FALSE;
    if(options & MJ_OPTION_PROGRESSIVE) jpeg_simple_progression(&out);
    else out.scan_info = NULL;
    out.arith_code = (options & MJ_OPTION_ARITHMETRIC) ? TRUE : FALSE;

    jpeg_write_coefficients(&out, m->coef);
    /* re-emit every saved marker after the headers libjpeg wrote itself -- this duplicates the
     * JFIF APP0 exactly like the reference does (src/image.c:196-200), keeping output bytes identical */
    for(jpeg_saved_marker_ptr mk = m->cinfo.marker_list; mk != NULL; mk = mk->next)
        jpeg_write_marker(&out, mk->marker, mk->data, mk->data_length);
    jpeg_finish_compress(&out);
    jpeg_destroy_compress(&out);
    itrap->armed = 0;

    *memory = dst.data;
    *len = dst.length;
    return MJ_OK;

failed:
    jpeg_destroy_compress(&out);
    free(dst.data);
    itrap->armed = 0;
    return MJ_ERR_ENCODE_JPEG;
}

int mj_write_jpeg_to_file(mj_jpeg_t *m, char *filename, int options) {
    if(m == NULL) return MJ_ERR_NULL_DATA;
    if(filename == NULL) return MJ_ERR_FILEIO;
    FILE *fp = fopen(filename, "wb");
    if(fp == NULL) return MJ_ERR_FILEIO;
    unsigned char *buffer = NULL;
    size_t         len = 0;
    int            rv = mj_write_jpeg_to_memory(m, &buffer, &len, options);
    if(rv == MJ_OK && fwrite(buffer, 1, len, fp) != len) rv = MJ_ERR_FILEIO;
    if(fclose(fp) != 0 && rv == MJ_OK) rv = MJ_ERR_FILEIO;
    free(buffer);
    return rv;
}

/* ---- mjx_host.h: flat plane access -------------------------------------------------------- */

int mjx_jpeg_image_info(mj_jpeg_t *m, int *info) {
    if(m == NULL || m->coef == NULL || info == NULL) return MJ_ERR_NULL_DATA;
    info[0] = m->cinfo.num_components;
    info[1] = (int)m->cinfo.jpeg_color_space;
    info[2] = m->width;
    info[3] = m->height;
    info[4] = m->cinfo.max_h_samp_factor;
    info[5] = m->cinfo.max_v_samp_factor;
    return MJ_OK;
}

int mjx_jpeg_component_info(mj_jpeg_t *m, int c, int *info) {
    if(m == NULL || m->coef == NULL || info == NULL || c < 0 || c >= m->cinfo.num_components) return MJ_ERR_NULL_DATA;
    const jpeg_component_info *ci = &m->cinfo.comp_info[c];
    info[0] = (int)ci->width_in_blocks;
    info[1] = (int)ci->height_in_blocks;
    info[2] = ci->h_samp_factor;
    info[3] = ci->v_samp_factor;
    info[4] = (int)mjp_virtual_width(ci);
    info[5] = (int)mjp_virtual_height(ci);
    return MJ_OK;
}

int mjx_jpeg_qtable(mj_jpeg_t *m, int c, unsigned short *q64) {
    if(m == NULL || m->coef == NULL || q64 == NULL || c < 0 || c >= m->cinfo.num_components) return MJ_ERR_NULL_DATA;
    if(m->cinfo.comp_info[c].quant_table == NULL) return MJ_ERR_NULL_DATA;
    memcpy(q64, m->cinfo.comp_info[c].quant_table->quantval, 64 * sizeof(unsigned short));
    return MJ_OK;
}

int mjx_jpeg_layout(mj_jpeg_t *m, mjx_layout_t *layout) {
    if(m == NULL || m->coef == NULL || layout == NULL) return MJ_ERR_NULL_DATA;
    memset(layout, 0, sizeof(*layout));
    layout->colorspace = (int)m->cinfo.jpeg_color_space;
    layout->ncomp = m->cinfo.num_components;
    if(layout->ncomp > MJX_MAX_COMPONENTS) return MJ_ERR_UNSUPPORTED_COLORSPACE;
    for(int c = 0; c < layout->ncomp; c++) {
        layout->h_samp[c] = m->cinfo.comp_info[c].h_samp_factor;
        layout->v_samp[c] = m->cinfo.comp_info[c].v_samp_factor;
    }
    return MJ_OK;
}

static int plane_io(mj_jpeg_t *m, int c, short *flat, int export_it) {
    if(m == NULL || m->coef == NULL || flat == NULL || c < 0 || c >= m->cinfo.num_components) return MJ_ERR_NULL_DATA;
    const jpeg_component_info *ci = &m->cinfo.comp_info[c];
    const unsigned             vw = mjp_virtual_width(ci), vh = mjp_virtual_height(ci);
    mjp_trap_t                *trap = mjp_image_trap(m);
    trap->armed = 1;
    if(setjmp(trap->escape)) {
        trap->armed = 0;
        return MJ_ERR_DECODE_JPEG;
    }
    for(unsigned r = 0; r < vh; r++) {
        JBLOCKARRAY rows = (*m->cinfo.mem->access_virt_barray)((j_common_ptr)&m->cinfo, m->coef[c], r, 1, TRUE);
        short      *line = flat + (size_t)r * vw * DCTSIZE2;
        if(export_it) memcpy(line, rows[0], (size_t)vw * sizeof(JBLOCK));
        else memcpy(rows[0], line, (size_t)vw * sizeof(JBLOCK));
    }
    trap->armed = 0;
    return MJ_OK;
}

int mjx_jpeg_export_plane(mj_jpeg_t *m, int c, short *dst) { return plane_io(m, c, dst, 1); }
int mjx_jpeg_import_plane(mj_jpeg_t *m, int c, const short *src) { return plane_io(m, c, (short *)src, 0); }
