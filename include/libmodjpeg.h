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
#ifndef MODJPEG_B200_LIBMODJPEG_H
#define MODJPEG_B200_LIBMODJPEG_H

/* stdio.h must precede jpeglib.h (size_t, FILE) */
#include <stdio.h>
#include <jpeglib.h>

#ifdef __cplusplus
extern "C" {
#endif

/* API level implemented: the reference's header says 1.0.0 (reference: src/libmodjpeg.h:33-36) */
#define MJ_LIB_VERSION_MAJOR 1
#define MJ_LIB_VERSION_MINOR 0
#define MJ_LIB_VERSION_RELEASE 0
#define MJ_LIB_VERSION (MJ_LIB_VERSION_MAJOR * 10000 + MJ_LIB_VERSION_MINOR * 100 + MJ_LIB_VERSION_RELEASE)

/* The numeric values below are the reference's ABI (reference: src/libmodjpeg.h:38-69). */

enum { /* pixel formats of mj_read_dropon_from_raw: odd = no alpha byte, even = trailing alpha byte */
    MJ_COLORSPACE_RGB = 1,
    MJ_COLORSPACE_RGBA,
    MJ_COLORSPACE_GRAYSCALE,
    MJ_COLORSPACE_GRAYSCALEA,
    MJ_COLORSPACE_YCC,
    MJ_COLORSPACE_YCCA
};

enum { /* placement bits of mj_compose; neither LEFT nor RIGHT (TOP nor BOTTOM) set means centred on that axis */
    MJ_ALIGN_LEFT = 0x01,
    MJ_ALIGN_RIGHT = 0x02,
    MJ_ALIGN_TOP = 0x04,
    MJ_ALIGN_BOTTOM = 0x08,
    MJ_ALIGN_CENTER = 0x10
};

enum { /* mj_dropon_t.blend */
    MJ_BLEND_NONUNIFORM = -1, /* per-pixel alpha */
    MJ_BLEND_NONE = 0,        /* fully transparent: mj_compose is a no-op */
    MJ_BLEND_FULL = 255       /* fully opaque */
};

enum { /* flags of mj_write_jpeg_* (ARITHMETRIC is the reference's spelling) */
    MJ_OPTION_NONE = 0,
    MJ_OPTION_OPTIMIZE = 0x1,
    MJ_OPTION_PROGRESSIVE = 0x2,
    MJ_OPTION_ARITHMETRIC = 0x4
};

enum { /* return codes */
    MJ_OK = 0,
    MJ_ERR_MEMORY,
    MJ_ERR_NULL_DATA,
    MJ_ERR_DROPON_DIMENSIONS,
    MJ_ERR_UNSUPPORTED_COLORSPACE,
    MJ_ERR_DECODE_JPEG,
    MJ_ERR_ENCODE_JPEG,
    MJ_ERR_FILEIO,
    MJ_ERR_IMAGE_SIZE,
    MJ_ERR_UNSUPPORTED_FILETYPE,
    MJ_ERR_DEVICE /* 10, additive: no usable CUDA device, or a kernel launch failed */
};

/* The reference spells these constants as macros; callers that test them with #ifdef keep compiling the same way. */
#define MJ_COLORSPACE_RGB MJ_COLORSPACE_RGB
#define MJ_COLORSPACE_RGBA MJ_COLORSPACE_RGBA
#define MJ_COLORSPACE_GRAYSCALE MJ_COLORSPACE_GRAYSCALE
#define MJ_COLORSPACE_GRAYSCALEA MJ_COLORSPACE_GRAYSCALEA
#define MJ_COLORSPACE_YCC MJ_COLORSPACE_YCC
#define MJ_COLORSPACE_YCCA MJ_COLORSPACE_YCCA
#define MJ_ALIGN_LEFT MJ_ALIGN_LEFT
#define MJ_ALIGN_RIGHT MJ_ALIGN_RIGHT
#define MJ_ALIGN_TOP MJ_ALIGN_TOP
#define MJ_ALIGN_BOTTOM MJ_ALIGN_BOTTOM
#define MJ_ALIGN_CENTER MJ_ALIGN_CENTER
#define MJ_BLEND_NONUNIFORM MJ_BLEND_NONUNIFORM
#define MJ_BLEND_NONE MJ_BLEND_NONE
#define MJ_BLEND_FULL MJ_BLEND_FULL
#define MJ_OPTION_NONE MJ_OPTION_NONE
#define MJ_OPTION_OPTIMIZE MJ_OPTION_OPTIMIZE
#define MJ_OPTION_PROGRESSIVE MJ_OPTION_PROGRESSIVE
#define MJ_OPTION_ARITHMETRIC MJ_OPTION_ARITHMETRIC
#define MJ_OK MJ_OK
#define MJ_ERR_MEMORY MJ_ERR_MEMORY
#define MJ_ERR_NULL_DATA MJ_ERR_NULL_DATA
#define MJ_ERR_DROPON_DIMENSIONS MJ_ERR_DROPON_DIMENSIONS
#define MJ_ERR_UNSUPPORTED_COLORSPACE MJ_ERR_UNSUPPORTED_COLORSPACE
#define MJ_ERR_DECODE_JPEG MJ_ERR_DECODE_JPEG
#define MJ_ERR_ENCODE_JPEG MJ_ERR_ENCODE_JPEG
#define MJ_ERR_FILEIO MJ_ERR_FILEIO
#define MJ_ERR_IMAGE_SIZE MJ_ERR_IMAGE_SIZE
#define MJ_ERR_UNSUPPORTED_FILETYPE MJ_ERR_UNSUPPORTED_FILETYPE
#define MJ_ERR_DEVICE MJ_ERR_DEVICE

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

/* ---- overlay ingest: host side, one-time (reference: src/dropon.c:33-323,578-604) ---------------------------
 * mj_init_dropon() zeroes a caller-allocated struct (required before first use); the readers release whatever
 * the struct held, then fill it.  `pixels` of the raw reader is copied. */
void mj_init_dropon(mj_dropon_t *dropon);
void mj_free_dropon(mj_dropon_t *dropon);
int  mj_read_dropon_from_raw(mj_dropon_t         *dropon,
                             const unsigned char *pixels,      /* width * height * (1 | 2 | 3 | 4) bytes */
                             unsigned int         pixel_format, /* MJ_COLORSPACE_* */
                             int                  width,
                             int                  height,
                             short                blend);
int  mj_read_dropon_from_memory(mj_dropon_t         *dropon,
                                const unsigned char *file_bytes, /* a JPEG or PNG file image */
                                size_t               nbytes,
                                const unsigned char *mask_bytes, /* optional grayscale JPEG used as alpha (JPEG dropons) */
                                size_t               mask_nbytes,
                                short                blend);
int  mj_read_dropon_from_file(mj_dropon_t *dropon, const char *path, const char *mask_path, short blend);

/* ---- JPEG coefficient I/O: host libjpeg entropy decode / encode (reference: src/image.c:33-255) ------------
 * max_pixel: refuse images with more pixels (0 = no limit).  mj_write_jpeg_to_memory() hands a malloc()ed
 * buffer to the caller. */
void mj_init_jpeg(mj_jpeg_t *jpeg);
void mj_free_jpeg(mj_jpeg_t *jpeg);
int  mj_read_jpeg_from_memory(mj_jpeg_t *jpeg, const unsigned char *file_bytes, size_t nbytes, size_t max_pixel);
int  mj_read_jpeg_from_file(mj_jpeg_t *jpeg, const char *path, size_t max_pixel);
int  mj_write_jpeg_to_memory(mj_jpeg_t *jpeg, unsigned char **out_bytes, size_t *out_nbytes, int options);
int  mj_write_jpeg_to_file(mj_jpeg_t *jpeg, char *path, int options);

/* ---- DCT-domain compositing: kernels K1 + K2 on the GPU (reference: src/compose.c:33-180) ------------------
 * Blends `dropon` into the coefficients of `jpeg` at the place given by the MJ_ALIGN_* bits plus a pixel
 * offset; returns when jpeg->coef holds the result. */
int mj_compose(mj_jpeg_t *jpeg, mj_dropon_t *dropon, unsigned int align, int offset_x, int offset_y);

/* ---- coefficient effects: kernel K3 on the GPU (reference: src/effect.c:28-222) ----------------------------
 * grayscale / tint / luminance act on YCbCr images only (MJ_OK and no change otherwise). */
int mj_effect_grayscale(mj_jpeg_t *jpeg);                 /* drop all chroma */
int mj_effect_pixelate(mj_jpeg_t *jpeg);                  /* keep only the DC of every block */
int mj_effect_tint(mj_jpeg_t *jpeg, int cb_add, int cr_add); /* add to the chroma DCs */
int mj_effect_luminance(mj_jpeg_t *jpeg, int y_add);      /* add to the luma DC */

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
/* GPUs mj_compose_batch spreads a batch over (images are independent: one contiguous slice, one group of host threads, one
 * compiled dropon per device; nothing crosses between devices).  0 = take $MJX_DEVICES ("all" or a count), default 1. */
void mj_batch_set_devices(int devices);

/* ---- additive: request coalescer (not in the reference) ---------------------------------
 * For request servers (the nginx filter's shape: many threads, one image per call, one shared logo).  When enabled,
 * concurrent mj_compose calls with the same dropon, target layout and placement remainder are gathered -- for at most
 * `wait_us` microseconds or `max_batch` requests -- into ONE kernel launch over a shared page-locked slab, and the
 * dropon is compiled once per key instead of once per call.  Results are byte-identical to unbatched calls; a call
 * still returns only when its image holds the result.  Off by default (it adds up to wait_us of latency to a lone
 * request); MJX_COALESCE=1 [MJX_COALESCE_MAX, MJX_COALESCE_WAIT_US] in the environment switches it on without code.
 * max_batch <= 0 / wait_us < 0 keep the current values (defaults 32 and 200). */
void mj_coalesce_configure(int enable, int max_batch, int wait_us);
void mj_coalesce_stats(unsigned long *batches, unsigned long *requests); /* launches made / requests served so far */

#ifdef __cplusplus
}
#endif

#endif
