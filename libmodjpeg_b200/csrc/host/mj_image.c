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

static void copy_table(mjx_huff_table_t *dst, const JHUFF_TBL *src);

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
    /* mj_compose / mj_effect_* hand the kernels row pointers gathered from several access_virt_barray calls (the reference
     * touches one row at a time).  That is sound only while libjpeg keeps whole arrays in memory -- always true for the
     * malloc-only memory manager (jmemnobs, what libjpeg-turbo ships), and made true here for managers with a backing store,
     * which would otherwise follow $JPEGMEM: no limit, no swapping of rows. */
    m->cinfo.mem->max_memory_to_use = 0x3fffffffL * (long)(sizeof(long) > 4 ? 1024 : 1);

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

/* Read the markers of a JPEG up to its first scan and stop: m holds the frame (size, sampling, quantisation tables, saved
 * markers) but no coefficients (m->coef == NULL).  For files whose ONE scan the device can decode (k5_huffman_decode.cu):
 * 8-bit sequential Huffman, every component in the scan, no restart markers.  *entropy_off = where the entropy-coded segment
 * starts in `memory`; *scan = that scan with the FILE's tables.  MJ_ERR_UNSUPPORTED_FILETYPE: a valid file of another kind
 * (the caller reads it with mj_read_jpeg_from_memory); other codes as mj_read_jpeg_from_memory. */
int mjp_read_header_only(mj_jpeg_t *m, const unsigned char *memory, size_t len, size_t *entropy_off, mjx_scan_t *scan) {
    if(m == NULL || memory == NULL || len == 0) return MJ_ERR_NULL_DATA;
    mj_free_jpeg(m);
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
    jpeg_save_markers(&m->cinfo, JPEG_COM, 0xFFFF);
    for(int k = 0; k < 16; k++) jpeg_save_markers(&m->cinfo, JPEG_APP0 + k, 0xFFFF);
    jpeg_read_header(&m->cinfo, TRUE); /* stops behind the header of the first scan */
    m->width = (int)m->cinfo.image_width;
    m->height = (int)m->cinfo.image_height;

    int                     rv = MJ_OK;
    const j_decompress_ptr  ci = &m->cinfo;
    const int               nc = ci->num_components;
    if(ci->jpeg_color_space != JCS_GRAYSCALE && ci->jpeg_color_space != JCS_RGB && ci->jpeg_color_space != JCS_YCbCr) rv = MJ_ERR_UNSUPPORTED_COLORSPACE;
    else if(ci->progressive_mode || ci->arith_code || ci->data_precision != 8 || ci->restart_interval != 0 || nc < 1 || nc > MJX_MAX_COMPONENTS ||
            ci->comps_in_scan != nc || ci->Ss != 0 || ci->Se != DCTSIZE2 - 1 || ci->Ah != 0 || ci->Al != 0)
        rv = MJ_ERR_UNSUPPORTED_FILETYPE;
    if(rv == MJ_OK) {
        memset(scan, 0, sizeof(*scan));
        scan->ncomp = nc;
        int blocks = 0;
        for(int c = 0; c < nc && rv == MJ_OK; c++) {
            const jpeg_component_info *cc = ci->cur_comp_info[c];
            if(cc != &ci->comp_info[c] || ci->quant_tbl_ptrs[cc->quant_tbl_no] == NULL) rv = MJ_ERR_UNSUPPORTED_FILETYPE; /* scan order = frame order */
            else {
                scan->h_samp[c] = cc->h_samp_factor;
                scan->v_samp[c] = cc->v_samp_factor;
                scan->dc_tbl[c] = cc->dc_tbl_no;
                scan->ac_tbl[c] = cc->ac_tbl_no;
                blocks += cc->h_samp_factor * cc->v_samp_factor;
                if(cc->dc_tbl_no < 0 || cc->dc_tbl_no > 3 || cc->ac_tbl_no < 0 || cc->ac_tbl_no > 3 || ci->dc_huff_tbl_ptrs[cc->dc_tbl_no] == NULL ||
                   ci->ac_huff_tbl_ptrs[cc->ac_tbl_no] == NULL)
                    rv = MJ_ERR_UNSUPPORTED_FILETYPE;
            }
        }
        if(nc > 1 && blocks > 10) rv = MJ_ERR_UNSUPPORTED_FILETYPE;
        for(int i = 0; i < 4 && rv == MJ_OK; i++) {
            copy_table(&scan->dc[i], ci->dc_huff_tbl_ptrs[i]);
            copy_table(&scan->ac[i], ci->ac_huff_tbl_ptrs[i]);
        }
        if(rv == MJ_OK) {
            if(nc == 1) { /* not interleaved: the component's own grid of blocks (jdinput.c per_scan_setup) */
                const long hs = ci->comp_info[0].h_samp_factor, vs = ci->comp_info[0].v_samp_factor;
                scan->mcus_per_row = (int)(((long)ci->image_width * hs + ci->max_h_samp_factor * DCTSIZE - 1) / (ci->max_h_samp_factor * DCTSIZE));
                scan->mcu_rows = (int)(((long)ci->image_height * vs + ci->max_v_samp_factor * DCTSIZE - 1) / (ci->max_v_samp_factor * DCTSIZE));
            }
            else {
                const long mw = (long)ci->max_h_samp_factor * DCTSIZE, mh = (long)ci->max_v_samp_factor * DCTSIZE;
                scan->mcus_per_row = (int)(((long)ci->image_width + mw - 1) / mw);
                scan->mcu_rows = (int)(((long)ci->image_height + mh - 1) / mh);
            }
            *entropy_off = (size_t)(src->base.next_input_byte - memory);
        }
    }
    if(rv != MJ_OK) {
        jpeg_destroy_decompress(&m->cinfo);
        mj_init_jpeg(m);
        return rv;
    }
    mj_sampling_t *sp = &m->sampling;
    sp->max_h_samp_factor = ci->max_h_samp_factor;
    sp->max_v_samp_factor = ci->max_v_samp_factor;
    sp->h_factor = sp->max_h_samp_factor * DCTSIZE;
    sp->v_factor = sp->max_v_samp_factor * DCTSIZE;
    for(int c = 0; c < nc && c < 4; c++) {
        sp->samp_factor[c].h_samp_factor = ci->comp_info[c].h_samp_factor;
        sp->samp_factor[c].v_samp_factor = ci->comp_info[c].v_samp_factor;
    }
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

/* ---- entropy coding on the device (K4, opt-in) ------------------------------------------------
 * libjpeg still writes every marker: jpeg_write_coefficients emits SOI / JFIF, the saved markers follow exactly as in
 * the host path, and jpeg_finish_compress emits DQT / SOF / DHT / SOS at the start of its (only) pass -- a progress
 * monitor, which libjpeg calls before the first row of blocks is coded, leaves at that point.  What the destination
 * buffer holds then is the complete header; the entropy-coded segment comes from mjx_huffman_encode_rows_host and EOI
 * closes the file.  The result is byte-identical to the host path (tests/test_gpu_huffman.py). */
typedef struct {
    struct jpeg_progress_mgr base;
    jmp_buf                  leave;
} mjp_stop_t;

static void stop_after_headers(j_common_ptr cinfo) { longjmp(((mjp_stop_t *)cinfo->progress)->leave, 1); }

static void copy_table(mjx_huff_table_t *dst, const JHUFF_TBL *src) {
    memset(dst, 0, sizeof(*dst));
    if(src == NULL) return;
    memcpy(dst->bits, src->bits, 17);
    memcpy(dst->vals, src->huffval, 256);
}

/* Everything of the output file in front of the entropy-coded segment (SOI ... SOS header), written by libjpeg itself, and the
 * description of the one scan that follows it.  MJ_ERR_UNSUPPORTED_FILETYPE: not a file the device encoder takes. */
int mjp_scan_headers(mj_jpeg_t *m, unsigned char **head, size_t *head_len, mjx_scan_t *scan) {
    *head = NULL;
    *head_len = 0;
    const int nc = m->cinfo.num_components;
    if(nc < 1 || nc > MJX_MAX_COMPONENTS || m->cinfo.data_precision != 8) return MJ_ERR_UNSUPPORTED_FILETYPE;

    struct jpeg_compress_struct out;
    mjp_trap_t                  trap;
    mjp_memdst_t                dst;
    mjp_stop_t                  stop;
    mjp_trap_t                 *itrap = mjp_image_trap(m);
    volatile int                rv = MJ_ERR_ENCODE_JPEG;

    memset(&out, 0, sizeof(out)); /* jpeg_destroy_compress accepts it at any point below */
    mjp_memdst_init(&dst);
    mjp_trap_init(&trap);
    out.err = &trap.base;
    trap.armed = 1;
    itrap->armed = 1;
    if(setjmp(trap.escape)) goto done;
    if(setjmp(itrap->escape)) goto done;
    jpeg_create_compress(&out);
    out.dest = &dst.base;
    jpeg_copy_critical_parameters(&m->cinfo, &out);
    out.optimize_coding = FALSE;
    out.scan_info = NULL;
    out.arith_code = FALSE;
    if(out.restart_interval != 0 || out.restart_in_rows != 0) {
        rv = MJ_ERR_UNSUPPORTED_FILETYPE;
        goto done;
    }

    /* the scan libjpeg is about to describe in SOS: every component, its tables, the MCU grid (jcmaster.c per_scan_setup) */
    memset(scan, 0, sizeof(*scan));
    scan->ncomp = nc;
    {
        int blocks = 0;
        for(int c = 0; c < nc; c++) {
            const jpeg_component_info *ci = &out.comp_info[c];
            scan->h_samp[c] = ci->h_samp_factor;
            scan->v_samp[c] = ci->v_samp_factor;
            scan->dc_tbl[c] = ci->dc_tbl_no;
            scan->ac_tbl[c] = ci->ac_tbl_no;
            blocks += ci->h_samp_factor * ci->v_samp_factor;
        }
        if(nc > 1 && blocks > 10) { /* C_MAX_BLOCKS_IN_MCU: libjpeg refuses the scan; let it say so */
            rv = MJ_ERR_UNSUPPORTED_FILETYPE;
            goto done;
        }
        for(int i = 0; i < 4; i++) {
            copy_table(&scan->dc[i], out.dc_huff_tbl_ptrs[i]);
            copy_table(&scan->ac[i], out.ac_huff_tbl_ptrs[i]);
        }
        if(nc == 1) {
            scan->mcus_per_row = (int)m->cinfo.comp_info[0].width_in_blocks;
            scan->mcu_rows = (int)m->cinfo.comp_info[0].height_in_blocks;
        }
        else {
            const long mw = (long)m->cinfo.max_h_samp_factor * DCTSIZE, mh = (long)m->cinfo.max_v_samp_factor * DCTSIZE;
            scan->mcus_per_row = (int)(((long)m->cinfo.image_width + mw - 1) / mw);
            scan->mcu_rows = (int)(((long)m->cinfo.image_height + mh - 1) / mh);
        }
    }

    {
        /* An image of which only the header was read (mjp_read_header_only) has no coefficient arrays; libjpeg only stores the
         * pointer here and would first look at the arrays when it codes a row of blocks -- which it never gets to. */
        static jvirt_barray_ptr no_arrays[MAX_COMPONENTS];
        jpeg_write_coefficients(&out, m->coef != NULL ? m->coef : no_arrays);
    }
    for(jpeg_saved_marker_ptr mk = m->cinfo.marker_list; mk != NULL; mk = mk->next)
        jpeg_write_marker(&out, mk->marker, mk->data, mk->data_length);
    memset(&stop, 0, sizeof(stop));
    stop.base.progress_monitor = stop_after_headers;
    out.progress = &stop.base;
    if(setjmp(stop.leave) == 0) {
        jpeg_finish_compress(&out); /* leaves through stop_after_headers once the headers are out */
        goto done;                  /* (not reached for an image with at least one row of blocks) */
    }
    *head_len = dst.capacity - dst.base.free_in_buffer;
    *head = dst.data; /* handed to the caller */
    dst.data = NULL;
    rv = MJ_OK;
done:
    jpeg_destroy_compress(&out);
    free(dst.data);
    itrap->armed = 0;
    return rv;
}

int mjx_write_jpeg_to_memory_device(mj_jpeg_t *m, unsigned char **memory, size_t *len, int options) {
    if(m == NULL || memory == NULL || len == NULL || m->coef == NULL) return MJ_ERR_NULL_DATA;
    if(options & (MJ_OPTION_OPTIMIZE | MJ_OPTION_PROGRESSIVE | MJ_OPTION_ARITHMETRIC)) return MJ_ERR_UNSUPPORTED_FILETYPE;
    mjx_ctx *ctx = mjx_host_ctx();
    if(ctx == NULL) return MJ_ERR_DEVICE;

    unsigned char *head = NULL, *sb = NULL;
    size_t         head_len = 0, sl = 0;
    mjx_scan_t     scan;
    int            rv = mjp_scan_headers(m, &head, &head_len, &scan);
    if(rv != MJ_OK) return rv;

    const int   nc = m->cinfo.num_components;
    int         stride[MJX_MAX_COMPONENTS], vrows[MJX_MAX_COMPONENTS], wreal[MJX_MAX_COMPONENTS], hreal[MJX_MAX_COMPONENTS];
    size_t      nrows = 0;
    for(int c = 0; c < nc; c++) {
        const jpeg_component_info *ci = &m->cinfo.comp_info[c];
        stride[c] = (int)mjp_virtual_width(ci);
        vrows[c] = (int)mjp_virtual_height(ci);
        wreal[c] = (int)ci->width_in_blocks;
        hreal[c] = (int)ci->height_in_blocks;
        nrows += (size_t)hreal[c];
    }
    short **rowbuf = (short **)malloc(nrows * sizeof(short *));
    if(rowbuf == NULL) {
        free(head);
        return MJ_ERR_MEMORY;
    }
    mjp_trap_t *itrap = mjp_image_trap(m);
    itrap->armed = 1;
    if(setjmp(itrap->escape)) {
        itrap->armed = 0;
        free(head);
        free(rowbuf);
        return MJ_ERR_ENCODE_JPEG;
    }
    const short *const *rows[MJX_MAX_COMPONENTS] = {NULL, NULL, NULL, NULL};
    size_t              at = 0;
    for(int c = 0; c < nc; c++) {
        rows[c] = (const short *const *)(rowbuf + at);
        for(int l = 0; l < hreal[c]; l++) {
            JBLOCKARRAY ba = (*m->cinfo.mem->access_virt_barray)((j_common_ptr)&m->cinfo, m->coef[c], (JDIMENSION)l, 1, FALSE);
            rowbuf[at++] = (short *)ba[0];
        }
    }
    itrap->armed = 0;
    const int mrv = mjx_huffman_encode_rows_host(ctx, nc, rows, stride, vrows, wreal, hreal, &scan, &sb, &sl);
    free(rowbuf);
    if(mrv != MJX_OK) {
        free(head);
        return mrv == MJX_ERR_UNSUPPORTED ? MJ_ERR_UNSUPPORTED_FILETYPE : mjp_map_error(mrv);
    }
    rv = mjp_assemble_file(memory, len, head, head_len, sb, sl);
    free(head);
    free(sb);
    return rv;
}

/* header + entropy-coded segment + EOI */
int mjp_assemble_file(unsigned char **memory, size_t *len, const unsigned char *head, size_t head_len, const unsigned char *seg, size_t seg_len) {
    unsigned char *file = (unsigned char *)malloc(head_len + seg_len + 2);
    if(file == NULL) return MJ_ERR_MEMORY;
    memcpy(file, head, head_len);
    memcpy(file + head_len, seg, seg_len);
    file[head_len + seg_len] = 0xFF;
    file[head_len + seg_len + 1] = 0xD9;
    *memory = file;
    *len = head_len + seg_len + 2;
    return MJ_OK;
}

/* MJX_GPU_HUFFMAN: 1 = mj_write_jpeg_to_memory codes baseline scans on the device, and so does mj_compose_batch; 0 = nobody
 * does; unset = mj_compose_batch does (whole windows of images stay in HBM between blend and entropy coding, which pays:
 * 965 -> 1 280 files/s on the bench batch), single writes stay on the host (for one image the staging copy of its planes
 * costs what libjpeg's encoder costs).  Returns -1 (unset), 0 or 1. */
int mjp_gpu_huffman_mode(void) {
    static int state = -2;
    if(state == -2) {
        const char *e = getenv("MJX_GPU_HUFFMAN");
        state = (e == NULL || !*e) ? -1 : (*e != '0' ? 1 : 0);
    }
    return state;
}

int mj_write_jpeg_to_memory(mj_jpeg_t *m, unsigned char **memory, size_t *len, int options) {
    if(m == NULL || memory == NULL || len == NULL || m->coef == NULL) return MJ_ERR_NULL_DATA;
    if(mjp_gpu_huffman_mode() == 1 && !(options & (MJ_OPTION_OPTIMIZE | MJ_OPTION_PROGRESSIVE | MJ_OPTION_ARITHMETRIC))) {
        /* anything the device path does not take (or cannot code) goes through libjpeg below, which also owns the error */
        if(mjx_write_jpeg_to_memory_device(m, memory, len, options) == MJ_OK) return MJ_OK;
    }

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
    return mjp_layout_of(m, layout);
}

/* (also for an image of which only the header has been read, mjp_read_header_only) */
int mjp_layout_of(mj_jpeg_t *m, mjx_layout_t *layout) {
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
