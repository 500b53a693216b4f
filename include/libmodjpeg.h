/*
 * libmodjpeg.h -- public C API of the B200-native libmodjpeg drop-in.
 *
 * Same 16 entry points, constants, error codes and caller-visible struct layouts as the
 * reference's public header (reference: src/libmodjpeg.h:33-149), so a program written
 * against the reference (src/contrib/modjpeg.c, the nginx filter) recompiles and links
 * unchanged.  What is different is behind the API: mj_compose() and the mj_effect_*()
 * calls run on an NVIDIA B200 (sm_100a) through the kernel-level C-ABI in mjx.h; entropy
 * decode/encode stays on host libjpeg (jpeg_read_coefficients / jpeg_write_coefficients).
 *
 * There is no CPU fallback: if no CUDA device can be initialised the compute entry points
 * print one line to stderr and return MJ_ERR_DEVICE (an additive error code, 10).
 *
 * Like the reference, mj_jpeg_t embeds struct jpeg_decompress_struct by value, so callers
 * and the library must be compiled against the same <jpeglib.h> (here: the ABI-62 header in
 * third_party/jpeg62, matching the libjpeg-turbo 3.1.x runtime in this image).
 */
#ifndef _LIBMODJPEG_H_
#define _LIBMODJPEG_H_

/* stdio.h must precede jpeglib.h (size_t, FILE) */
#include <stdio.h>
#include <jpeglib.h>

#ifdef __cplusplus
extern "C" {
#endif

/* reference: src/libmodjpeg.h:33-36 (the reference header still says 1.0.0) */
#define MJ_LIB_VERSION_MAJOR   1
#define MJ_LIB_VERSION_MINOR   0
#define MJ_LIB_VERSION_RELEASE 0
#define MJ_LIB_VERSION         10000

/* raw dropon pixel formats accepted by mj_read_dropon_from_raw (reference: :38-43) */
#define MJ_COLORSPACE_RGB        1
#define MJ_COLORSPACE_RGBA       2
#define MJ_COLORSPACE_GRAYSCALE  3
#define MJ_COLORSPACE_GRAYSCALEA 4
#define MJ_COLORSPACE_YCC        5
#define MJ_COLORSPACE_YCCA       6

/* placement bits for mj_compose (reference: :45-49) */
#define MJ_ALIGN_LEFT   (1 << 0)
#define MJ_ALIGN_RIGHT  (1 << 1)
#define MJ_ALIGN_TOP    (1 << 2)
#define MJ_ALIGN_BOTTOM (1 << 3)
#define MJ_ALIGN_CENTER (1 << 4)

/* blend values (reference: :51-53) */
#define MJ_BLEND_NONUNIFORM -1
#define MJ_BLEND_NONE       0
#define MJ_BLEND_FULL       255

/* mj_write_jpeg_* options (reference: :55-58; the misspelling is the reference's) */
#define MJ_OPTION_NONE        0
#define MJ_OPTION_OPTIMIZE    (1 << 0)
#define MJ_OPTION_PROGRESSIVE (1 << 1)
#define MJ_OPTION_ARITHMETRIC (1 << 2)

/* return codes (reference: :60-69) */
#define MJ_OK                         0
#define MJ_ERR_MEMORY                 1
#define MJ_ERR_NULL_DATA              2
#define MJ_ERR_DROPON_DIMENSIONS      3
#define MJ_ERR_UNSUPPORTED_COLORSPACE 4
#define MJ_ERR_DECODE_JPEG            5
#define MJ_ERR_ENCODE_JPEG            6
#define MJ_ERR_FILEIO                 7
#define MJ_ERR_IMAGE_SIZE             8
#define MJ_ERR_UNSUPPORTED_FILETYPE   9
/* additive: the CUDA device / extension is missing or a kernel launch failed */
#define MJ_ERR_DEVICE                 10

/* sampling description of a decoded JPEG (reference: :71-84) */
typedef struct {
    int h_samp_factor;
    int v_samp_factor;
} mj_samplingfactor_t;

typedef struct {
    int max_h_samp_factor;
    int max_v_samp_factor;

    int h_factor; /* MCU width in pixels  = max_h_samp_factor * 8 */
    int v_factor; /* MCU height in pixels = max_v_samp_factor * 8 */

    mj_samplingfactor_t samp_factor[4];
} mj_sampling_t;

/* a decoded JPEG: libjpeg state + coefficient arrays (reference: :99-107; 696 bytes) */
typedef struct {
    struct jpeg_decompress_struct cinfo;
    jvirt_barray_ptr             *coef;

    int width;
    int height;

    mj_sampling_t sampling;
} mj_jpeg_t;

/* an overlay: pixels and alpha, both stored with 3 bytes per pixel (reference: :109-118) */
typedef struct {
    unsigned char *image;
    unsigned char *alpha;

    int width;
    int height;
    int colorspace; /* MJ_COLORSPACE_RGB, _YCC or _GRAYSCALE after reading */

    int blend; /* 0..255 uniform, or MJ_BLEND_NONUNIFORM when the pixels carry alpha */
} mj_dropon_t;

/* dropon ingest -- host side (reference: src/dropon.c:33-323,578-604) */
void mj_init_dropon(mj_dropon_t *d);
int  mj_read_dropon_from_raw(mj_dropon_t *d, const unsigned char *rawdata, unsigned int colorspace, int width, int height, short blend);
int  mj_read_dropon_from_memory(mj_dropon_t *d, const unsigned char *memory, size_t len, const unsigned char *maskmemory, size_t masklen, short blend);
int  mj_read_dropon_from_file(mj_dropon_t *d, const char *filename, const char *maskfilename, short blend);

/* JPEG coefficient I/O -- host libjpeg (reference: src/image.c:33-255) */
void mj_init_jpeg(mj_jpeg_t *m);
int  mj_read_jpeg_from_memory(mj_jpeg_t *m, const unsigned char *memory, size_t len, size_t max_pixel);
int  mj_read_jpeg_from_file(mj_jpeg_t *m, const char *filename, size_t max_pixel);

/* DCT-domain compositing -- B200 kernels K1 + K2 (reference: src/compose.c:33-180) */
int mj_compose(mj_jpeg_t *m, mj_dropon_t *d, unsigned int align, int offset_x, int offset_y);

int mj_write_jpeg_to_memory(mj_jpeg_t *m, unsigned char **memory, size_t *len, int options);
int mj_write_jpeg_to_file(mj_jpeg_t *m, char *filename, int options);

void mj_free_jpeg(mj_jpeg_t *m);
void mj_free_dropon(mj_dropon_t *d);

/* coefficient effects -- B200 kernel K3 (reference: src/effect.c:28-222) */
int mj_effect_grayscale(mj_jpeg_t *m);
int mj_effect_pixelate(mj_jpeg_t *m);
int mj_effect_tint(mj_jpeg_t *m, int cb_value, int cr_value);
int mj_effect_luminance(mj_jpeg_t *m, int value);

/* ---- additive: batch pipeline (not in the reference) ------------------------------------
 * n JPEGs in memory, one dropon, one placement: entropy decode and encode on a pool of `nthreads` host threads,
 * the dropon compiled once per image geometry (K1) and blended by ONE K2 launch per window of images.  Per image
 * this is mj_read_jpeg_from_memory + mj_compose + mj_write_jpeg_to_memory (same results); out[i].data is
 * malloc()ed for the caller, status[i] is that image's MJ_* code.  The return value reports batch-level failures
 * only (arguments, memory, no device). */
typedef struct {
    unsigned char *data;
    size_t         len;
} mj_blob_t;

int mj_compose_batch(int n, const mj_blob_t *in, mj_blob_t *out, int *status, mj_dropon_t *d, unsigned int align, int offset_x,
                     int offset_y, int write_options, int nthreads);

#ifdef __cplusplus
}
#endif

#endif
