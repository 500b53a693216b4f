/*
 * mjx_host.h -- additive host-side entry points of libmodjpeg.so (beyond the 16 functions of
 * libmodjpeg.h): flat access to the coefficient planes of a decoded JPEG, so that a batch host
 * (bench.py, a server front end) can move them to HBM and drive the kernel-level C-ABI (mjx.h)
 * itself.  They go through cinfo.mem->access_virt_barray exactly like the reference's own loops
 * (reference: src/compose.c:269, src/effect.c:48).  They expect an mj_jpeg_t filled by this
 * library's mj_read_jpeg_* (its error trap lives in the struct's libjpeg pool).
 */
#ifndef MJX_HOST_H
#define MJX_HOST_H

#include "libmodjpeg.h"
#include "mjx.h"

#ifdef __cplusplus
extern "C" {
#endif

/* info[0..5] = num_components, jpeg_color_space, width, height, max_h_samp_factor, max_v_samp_factor */
int mjx_jpeg_image_info(mj_jpeg_t *m, int *info);
/* info[0..5] = width_in_blocks, height_in_blocks, h_samp_factor, v_samp_factor, virtual width, virtual height */
int mjx_jpeg_component_info(mj_jpeg_t *m, int c, int *info);
int mjx_jpeg_qtable(mj_jpeg_t *m, int c, unsigned short *q64);
/* copy a whole plane ([virtual height][virtual width][64] int16) out of / into libjpeg's arrays */
int mjx_jpeg_export_plane(mj_jpeg_t *m, int c, short *dst);
int mjx_jpeg_import_plane(mj_jpeg_t *m, int c, const short *src);
/* the target layout of a decoded JPEG, as K1 wants it */
int mjx_jpeg_layout(mj_jpeg_t *m, mjx_layout_t *layout);
/* mj_write_jpeg_to_memory with the entropy coding done on the device (K4, mjx_huffman_encode_rows_host): libjpeg writes
 * the markers, the kernels the scan; byte-identical to mj_write_jpeg_to_memory(m, .., 0).  Baseline only: MJ_ERR_UNSUPPORTED_FILETYPE
 * for MJ_OPTION_OPTIMIZE / _PROGRESSIVE / _ARITHMETRIC and for coefficients Huffman tables cannot code.  With MJX_GPU_HUFFMAN=1 in
 * the environment mj_write_jpeg_to_memory takes this path by itself and falls back to libjpeg where it does not apply. */
int mjx_write_jpeg_to_memory_device(mj_jpeg_t *m, unsigned char **memory, size_t *len, int options);
/* the calling thread's engine context (created on first use; device from $MJX_DEVICE, default 0).
 * Returns NULL and prints one line to stderr when no CUDA device is usable. */
mjx_ctx *mjx_host_ctx(void);
/* device of the calling thread's context, for threads that have not made a compute call yet (-1: $MJX_DEVICE or 0) */
void mjx_host_set_device(int device);

#ifdef __cplusplus
}
#endif

#endif
