/*
 * mj_jpegio.c -- libjpeg plumbing of the host boundary: error trap, memory source, growing
 * memory destination, JPEG -> raw decode, whole-file read.
 * Same role as reference: src/jpeg.c:34-109 and src/image.c:349-490, written against the
 * classic libjpeg API.  Entropy coding stays on the host by design (north_star).
 */
#include <stdlib.h>
#include <string.h>

#include <jerror.h>

#include "mj_private.h"

/* ---- error trap ---------------------------------------------------------------------- */

static void trap_error_exit(j_common_ptr cinfo) {
    mjp_trap_t *t = (mjp_trap_t *)cinfo->err;
    /* like the reference, let libjpeg print its message to stderr first */
    (*cinfo->err->output_message)(cinfo);
    if(t->armed) longjmp(t->escape, 1);
    /* no public call on the stack to catch it (the caller drove libjpeg directly through
     * m->cinfo): behave like libjpeg's stock error_exit */
    jpeg_destroy(cinfo);
    exit(EXIT_FAILURE);
}

void mjp_trap_init(mjp_trap_t *t) {
    jpeg_std_error(&t->base);
    t->base.error_exit = trap_error_exit;
    t->armed = 0;
}

/* ---- memory source --------------------------------------------------------------------- */

static void src_init(j_decompress_ptr cinfo) {
    mjp_memsrc_t *s = (mjp_memsrc_t *)cinfo->src;
    s->base.next_input_byte = s->data;
    s->base.bytes_in_buffer = s->size;
}

static boolean src_fill(j_decompress_ptr cinfo) {
    /* the whole file was handed over at once; running dry means a truncated stream.
     * Feed an EOI so libjpeg terminates cleanly with a warning instead of spinning. */
    static const JOCTET eoi[2] = {0xFF, JPEG_EOI};
    cinfo->src->next_input_byte = eoi;
    cinfo->src->bytes_in_buffer = 2;
    return TRUE;
}

static void src_skip(j_decompress_ptr cinfo, long n) {
    struct jpeg_source_mgr *s = cinfo->src;
    if(n <= 0) return;
    if((size_t)n > s->bytes_in_buffer) n = (long)s->bytes_in_buffer;
    s->next_input_byte += n;
    s->bytes_in_buffer -= (size_t)n;
}

static void src_term(j_decompress_ptr cinfo) { (void)cinfo; }

void mjp_memsrc_init(mjp_memsrc_t *s, const unsigned char *data, size_t size) {
    memset(s, 0, sizeof(*s));
    s->data = data;
    s->size = size;
    s->base.init_source = src_init;
    s->base.fill_input_buffer = src_fill;
    s->base.skip_input_data = src_skip;
    s->base.resync_to_restart = jpeg_resync_to_restart;
    s->base.term_source = src_term;
}

/* ---- growing memory destination --------------------------------------------------------- */

#define MJP_DST_FIRST 4096

static void dst_init(j_compress_ptr cinfo) {
    mjp_memdst_t *d = (mjp_memdst_t *)cinfo->dest;
    d->data = (unsigned char *)malloc(MJP_DST_FIRST);
    if(d->data == NULL) ERREXIT1(cinfo, JERR_OUT_OF_MEMORY, 0);
    d->capacity = MJP_DST_FIRST;
    d->base.next_output_byte = d->data;
    d->base.free_in_buffer = d->capacity;
}

static boolean dst_grow(j_compress_ptr cinfo) {
    /* libjpeg calls this only when the buffer is completely full; double it */
    mjp_memdst_t  *d = (mjp_memdst_t *)cinfo->dest;
    size_t         bigger = d->capacity * 2;
    unsigned char *p = (unsigned char *)realloc(d->data, bigger);
    if(p == NULL) ERREXIT1(cinfo, JERR_OUT_OF_MEMORY, 0);
    d->data = p;
    d->base.next_output_byte = p + d->capacity;
    d->base.free_in_buffer = bigger - d->capacity;
    d->capacity = bigger;
    return TRUE;
}

static void dst_term(j_compress_ptr cinfo) {
    mjp_memdst_t *d = (mjp_memdst_t *)cinfo->dest;
    d->length = d->capacity - d->base.free_in_buffer;
}

void mjp_memdst_init(mjp_memdst_t *d) {
    memset(d, 0, sizeof(*d));
    d->base.init_destination = dst_init;
    d->base.empty_output_buffer = dst_grow;
    d->base.term_destination = dst_term;
}

/* ---- helpers ---------------------------------------------------------------------------- */

unsigned mjp_virtual_width(const jpeg_component_info *ci) {
    unsigned h = (unsigned)ci->h_samp_factor;
    return (ci->width_in_blocks + h - 1) / h * h;
}

unsigned mjp_virtual_height(const jpeg_component_info *ci) {
    unsigned v = (unsigned)ci->v_samp_factor;
    return (ci->height_in_blocks + v - 1) / v * v;
}

int mjp_map_error(int rv) {
    switch(rv) {
        case MJX_OK: return MJ_OK;
        case MJX_ERR_MEMORY: return MJ_ERR_MEMORY;
        case MJX_ERR_ARG: return MJ_ERR_NULL_DATA;
        case MJX_ERR_UNSUPPORTED: return MJ_ERR_ENCODE_JPEG;
        default: return MJ_ERR_DEVICE;
    }
}

/* JPEG in memory -> interleaved samples in the wanted colourspace (dropons stored as JPEG:
 * role of reference src/image.c:377-448) */
int mjp_decode_to_raw(unsigned char **raw, int *width, int *height, int want_colorspace, const unsigned char *memory, size_t len) {
    struct jpeg_decompress_struct cinfo;
    mjp_trap_t                    trap;
    mjp_memsrc_t                  src;
    unsigned char *volatile       pixels = NULL;

    *raw = NULL;
    mjp_trap_init(&trap);
    cinfo.err = &trap.base;
    trap.armed = 1;
    if(setjmp(trap.escape)) {
        jpeg_destroy_decompress(&cinfo);
        free(pixels);
        return MJ_ERR_DECODE_JPEG;
    }
    jpeg_create_decompress(&cinfo);
    mjp_memsrc_init(&src, memory, len);
    cinfo.src = &src.base;
    jpeg_read_header(&cinfo, TRUE);

    if(want_colorspace == MJ_COLORSPACE_RGB) cinfo.out_color_space = JCS_RGB;
    else if(want_colorspace == MJ_COLORSPACE_YCC) cinfo.out_color_space = JCS_YCbCr;
    else if(want_colorspace == MJ_COLORSPACE_GRAYSCALE) cinfo.out_color_space = JCS_GRAYSCALE;
    else {
        jpeg_destroy_decompress(&cinfo);
        return MJ_ERR_UNSUPPORTED_COLORSPACE;
    }
    jpeg_start_decompress(&cinfo);

    size_t stride = (size_t)cinfo.output_width * (size_t)cinfo.output_components;
    pixels = (unsigned char *)calloc(stride ? stride * cinfo.output_height : 1, 1);
    if(pixels == NULL) {
        jpeg_destroy_decompress(&cinfo);
        return MJ_ERR_MEMORY;
    }
    while(cinfo.output_scanline < cinfo.output_height) {
        JSAMPROW line = pixels + (size_t)cinfo.output_scanline * stride;
        jpeg_read_scanlines(&cinfo, &line, 1);
    }
    *width = (int)cinfo.output_width;
    *height = (int)cinfo.output_height;
    jpeg_finish_decompress(&cinfo);
    jpeg_destroy_decompress(&cinfo);
    *raw = pixels;
    return MJ_OK;
}

int mjp_read_whole_file(unsigned char **buffer, size_t *len, const char *filename) {
    *buffer = NULL;
    *len = 0;
    if(filename == NULL) return MJ_ERR_NULL_DATA;
    FILE *fp = fopen(filename, "rb");
    if(fp == NULL) return MJ_ERR_FILEIO;
    if(fseek(fp, 0, SEEK_END) != 0) {
        fclose(fp);
        return MJ_ERR_FILEIO;
    }
    long size = ftell(fp);
    if(size < 0 || fseek(fp, 0, SEEK_SET) != 0) {
        fclose(fp);
        return MJ_ERR_FILEIO;
    }
    unsigned char *p = (unsigned char *)malloc((size_t)size + 1);
    if(p == NULL) {
        fclose(fp);
        return MJ_ERR_MEMORY;
    }
    size_t got = fread(p, 1, (size_t)size, fp);
    fclose(fp);
    if(got != (size_t)size) {
        free(p);
        return MJ_ERR_FILEIO;
    }
    p[size] = 0;
    *buffer = p;
    *len = (size_t)size;
    return MJ_OK;
}
