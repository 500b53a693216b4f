/*
 * mj_private.h -- internals of the host boundary (libmodjpeg.so): libjpeg plumbing shared by
 * the public entry points.  Host code only moves bytes and talks to libjpeg; all coefficient
 * arithmetic is done by the kernels behind mjx.h.
 */
#ifndef MJ_PRIVATE_H
#define MJ_PRIVATE_H

#include <setjmp.h>
#include <stddef.h>

#include "libmodjpeg.h"
#include "mjx.h"
#include "mjx_host.h"

/* error manager that turns libjpeg's error_exit into a longjmp
 * (role of reference: src/jpeg.h:30-34, src/jpeg.c:34-40) */
typedef struct {
    struct jpeg_error_mgr base;
    jmp_buf               escape;
    int                   armed; /* escape is valid only while a public call is on the stack */
} mjp_trap_t;

void mjp_trap_init(mjp_trap_t *t);

/* memory source over a caller-owned buffer (role of reference: src/jpeg.c:82-109) */
typedef struct {
    struct jpeg_source_mgr base;
    const unsigned char   *data;
    size_t                 size;
} mjp_memsrc_t;

void mjp_memsrc_init(mjp_memsrc_t *s, const unsigned char *data, size_t size);

/* growing malloc destination; the finished buffer is handed to the caller
 * (role of reference: src/jpeg.c:42-80) */
typedef struct {
    struct jpeg_destination_mgr base;
    unsigned char              *data;
    size_t                      capacity;
    size_t                      length; /* valid after term_destination */
} mjp_memdst_t;

void mjp_memdst_init(mjp_memdst_t *d);

/* The trap and the source manager of a decoded image live in libjpeg's permanent pool of
 * m->cinfo, so m->cinfo.err / .src stay valid for the lifetime of the mj_jpeg_t. */
mjp_trap_t *mjp_image_trap(mj_jpeg_t *m);

/* libjpeg allocates coefficient arrays rounded up to the sampling factors (jdcoefct.c) */
unsigned mjp_virtual_width(const jpeg_component_info *ci);
unsigned mjp_virtual_height(const jpeg_component_info *ci);

/* decode a JPEG in memory to interleaved 8-bit samples (for dropons given as JPEG files) */
int mjp_decode_to_raw(unsigned char **raw, int *width, int *height, int want_colorspace, const unsigned char *memory, size_t len);

int mjp_read_whole_file(unsigned char **buffer, size_t *len, const char *filename);

/* frees the calling thread's cached compiled dropons (mj_compose.c, MJX_DROPON_CACHE=1) */
void mjp_compose_cache_clear(void);

/* request coalescer (mj_coalesce.c): 1 when mj_compose should try it; mjp_coalesce_compose returns 1 when it served the
 * request (*result = mj_compose's return value), 0 when the ordinary path has to */
int mjp_coalesce_enabled(void);
int mjp_coalesce_compose(mj_jpeg_t *m, mj_dropon_t *d, const mjx_layout_t *layout, const mjx_geometry_t *g, int *result);

/* entropy coding on the device (mj_image.c): $MJX_GPU_HUFFMAN; the file in front of the entropy-coded segment as libjpeg writes
 * it (malloc()ed) + the description of the scan; header + segment + EOI as one malloc()ed file */
int mjp_gpu_huffman_mode(void); /* -1 unset, 0 off, 1 on */
int mjp_scan_headers(mj_jpeg_t *m, unsigned char **head, size_t *head_len, mjx_scan_t *scan);
int mjp_assemble_file(unsigned char **memory, size_t *len, const unsigned char *head, size_t head_len, const unsigned char *seg, size_t seg_len);

int mjp_read_header_only(mj_jpeg_t *m, const unsigned char *memory, size_t len, size_t *entropy_off, mjx_scan_t *scan);
int mjp_layout_of(mj_jpeg_t *m, mjx_layout_t *layout);

mjx_ctx *mjp_host_ctx2(void); /* a second context of the calling thread, same device (mj_device.c) */

/* MJX_* -> MJ_* */
int mjp_map_error(int mjx_rv);

#endif
