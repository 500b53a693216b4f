/*
 * mjx.h -- kernel-level C-ABI of the B200 (sm_100a) compositing engine, libmjx.so.
 *
 * This is the drop-in boundary for the hot path: plain pointers, sizes and ints, no torch or
 * C++ types.  Each entry point names the reference code it replaces.  The host library
 * (libmodjpeg.so, include/libmodjpeg.h) calls these from mj_compose() / mj_effect_*(); a
 * batch host (bench.py, a server) calls the *_batch_device / *_batch_host forms directly.
 * INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Conventions
 *   - A coefficient plane is int16 [rows][stride_blocks][64], natural order (index 8*v + u),
 *     exactly libjpeg's JBLOCKROW layout (reference: src/compose.c:269-274).
 *   - All calls are asynchronous on the context's stream unless they take host pointers;
 *     host-pointer calls return when the host buffers hold the result (the reference's
 *     synchronous semantics).
 *   - Return value: MJX_OK or an MJX_ERR_* code; mjx_ctx_last_error() has the CUDA text.
 *   - There is no CPU fallback.  Without a usable CUDA device every compute entry point
 *     returns MJX_ERR_DEVICE.
 */
#ifndef MJX_H
#define MJX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MJX_OK              0
#define MJX_ERR_MEMORY      1  /* == MJ_ERR_MEMORY */
#define MJX_ERR_ARG         2  /* == MJ_ERR_NULL_DATA */
#define MJX_ERR_UNSUPPORTED 6  /* == MJ_ERR_ENCODE_JPEG: a dropon->target conversion libjpeg rejects */
#define MJX_ERR_DEVICE      10 /* == MJ_ERR_DEVICE: no device, or a CUDA call failed */

#define MJX_MAX_COMPONENTS 4

/* dropon pixel formats after ingest (== MJ_COLORSPACE_RGB / _GRAYSCALE / _YCC) */
#define MJX_CS_RGB       1
#define MJX_CS_GRAYSCALE 3
#define MJX_CS_YCC       5

/* block classes of a compiled dropon */
#define MJX_CLS_T      0
#define MJX_CLS_U      1
#define MJX_CLS_OPAQUE 2
#define MJX_CLS_G      3

typedef struct mjx_ctx    mjx_ctx;    /* one device + stream + staging pools; not shareable between threads */
typedef struct mjx_dropon mjx_dropon; /* a compiled dropon resident in HBM; immutable, shareable between ctxs on one device */

/* component layout of the target JPEG: what mj_compile_dropon receives as
 * (J_COLOR_SPACE colorspace, mj_sampling_t *sampling) (reference: src/dropon.c:325) */
typedef struct {
    int colorspace; /* J_COLOR_SPACE: 1 GRAYSCALE, 2 RGB, 3 YCbCr */
    int ncomp;
    int h_samp[MJX_MAX_COMPONENTS];
    int v_samp[MJX_MAX_COMPONENTS];
} mjx_layout_t;

/* result of the placement arithmetic of mj_compose (reference: src/compose.c:42-172) */
typedef struct {
    int visible; /* 0: nothing of the dropon lies on the image (reference returns MJ_OK, compose.c:136) */
    int crop_x, crop_y, crop_w, crop_h;
    int blockoffset_x, blockoffset_y;
    int block_x, block_y; /* origin on the image in MCUs */
} mjx_geometry_t;

/* one image of a device-resident batch; an array of these lives in device memory, 16-byte aligned (the kernels
 * fetch the quantisation tables with 128-bit loads) */
typedef struct {
    uint64_t plane[MJX_MAX_COMPONENTS];         /* device address of block (0,0) of each component plane */
    int32_t  stride_blocks[MJX_MAX_COMPONENTS]; /* blocks per plane row (libjpeg's virtual width) */
    int32_t  rows[MJX_MAX_COMPONENTS];          /* plane rows in blocks (virtual height) */
    int32_t  wreal[MJX_MAX_COMPONENTS];         /* width_in_blocks  (effects touch real blocks only) */
    int32_t  hreal[MJX_MAX_COMPONENTS];         /* height_in_blocks */
    uint16_t q[MJX_MAX_COMPONENTS][64];         /* quantisation tables, natural order */
} mjx_image_desc_t;

/* the same for host-resident planes (contiguous [rows][stride_blocks][64] per component) */
typedef struct {
    int16_t        *plane[MJX_MAX_COMPONENTS];
    int32_t         stride_blocks[MJX_MAX_COMPONENTS];
    int32_t         rows[MJX_MAX_COMPONENTS];
    int32_t         wreal[MJX_MAX_COMPONENTS];
    int32_t         hreal[MJX_MAX_COMPONENTS];
    const uint16_t *q[MJX_MAX_COMPONENTS];
} mjx_host_image_t;

/* one step of an effect pipeline (K3).  Steps are applied in order to every block. */
#define MJX_FX_ZERO     1 /* all 64 coefficients of component `comp` := 0   (mj_effect_grayscale, src/effect.c:28) */
#define MJX_FX_PIXELATE 2 /* coefficients 1..63 of component `comp` := 0    (mj_effect_pixelate,  src/effect.c:70) */
#define MJX_FX_ADD_DC   3 /* DC of `comp`: dequantise, += value, clamp +-2047, requantise (tint/luminance, src/effect.c:116,185) */
typedef struct {
    int op;
    int comp;
    int value;
} mjx_effect_op_t;

/* ---- context ---------------------------------------------------------------------- */
int         mjx_device_count(void);
int         mjx_ctx_create(mjx_ctx **ctx, int device);
void        mjx_ctx_destroy(mjx_ctx *ctx);
int         mjx_ctx_set_stream(mjx_ctx *ctx, void *cuda_stream); /* borrow a caller-owned cudaStream_t; the handle is used as is (0 = the legacy default stream) */
int         mjx_ctx_use_own_stream(mjx_ctx *ctx);                /* back to the stream the ctx created */
/* strict = 1: K2 runs as one kernel that reproduces the reference's int16 wrap-around on out-of-range
 * products (adversarial streams); default 0: the fast kernels, identical on every encoder-produced JPEG */
int         mjx_ctx_set_strict(mjx_ctx *ctx, int strict);
/* class G blocks (non-uniform alpha) of batches of >= 256 images: the blend runs as one tensor-core product per dropon
 * block and 128 images (tcgen05, fp16 operands, fp32 accumulation; libmodjpeg_b200/csrc/k2_generic_op.cu) -- what
 * mj_compose_with_mask + mj_convolve compute per block is linear in the image block (reference: src/compose.c:289-312),
 * and its 64 x 64 operator is built once per (compiled dropon, quantisation tables) and cached in the dropon.  The integer
 * operand needs every coefficient in the baseline range [-1024, 1023] (DC 11 bits, AC 10 bits: what ITU-T T.81 allows an
 * 8-bit JPEG to carry) and quantiser values <= 255.
 * mode 1 (default): the kernel checks the range per block and hands what is outside it, and images whose tables differ
 * from the first image's, to the fp32 kernel; mode 2: no range check, the caller vouches for it; mode 0: fp32 kernel only.
 * Env MJX_K2_TC sets the initial mode.  The cache (16 or 24 KB per class G block, MJX_K2_OP_MAX_MB caps it, default 4096)
 * belongs to the first ctx that uses the dropon this way; other ctxs run the fp32 kernel with it. */
int         mjx_ctx_set_tensor_core(mjx_ctx *ctx, int mode);
/* smallest batch that takes the tensor-core kernel (>= 256; default 1025): the kernel deals batches of 64 images to three
 * groups of warps, so a batch of a few hundred images leaves groups idle and the fp32 kernel is the faster one. */
int         mjx_ctx_set_tensor_core_min_images(mjx_ctx *ctx, int n);
/* fp16 pieces per operator entry: 2 (22 significant bits) is the only value the kernel is built for -- it reproduces the
 * reference on every test image (tests/test_gpu_tensor_core.py); the call exists so that a build with more pieces stays
 * source compatible. */
int         mjx_ctx_set_operator_pieces(mjx_ctx *ctx, int pieces);
/* on = 1 (default): mjx_compose_batch_host runs K2 directly on page-locked (GPU-addressable) host planes, so only
 * the blocks the dropon touches cross PCIe; 0: always stage the region under the dropon through device memory */
int         mjx_ctx_set_zero_copy(mjx_ctx *ctx, int on);
/* device self-test: K2's per-image reciprocal tables (MUFU.RCP based) give trunc(a / q) exactly for every q in
 * [1, 65535] and every |a| <= 2^17; *mismatches receives the number of (a, q) pairs that do not (expected 0) */
int         mjx_selftest_reciprocal(mjx_ctx *ctx, long long *mismatches);
/* profiling aid: which fast-path K2 kernels run -- bit 0 the OPAQUE/U kernel, bit 1 the G kernel (default 3 = both) */
int         mjx_ctx_set_class_mask(mjx_ctx *ctx, int mask);
/* 1 (default): batches of >= 64 images run the OPAQUE/U kernel beside the G kernel (a low-priority side stream of the
 * ctx, joined back into the ctx stream before the call's work counts as done); 0: one after the other */
int         mjx_ctx_set_overlap(mjx_ctx *ctx, int on);
void       *mjx_ctx_stream(mjx_ctx *ctx);
int         mjx_ctx_sync(mjx_ctx *ctx);
const char *mjx_ctx_last_error(mjx_ctx *ctx);
long long   mjx_ctx_kernel_launches(mjx_ctx *ctx); /* kernels launched through this ctx so far */

/* device / pinned-host memory for callers without their own allocator */
int  mjx_device_alloc(mjx_ctx *ctx, void **ptr, size_t bytes);
void mjx_device_free(mjx_ctx *ctx, void *ptr);
int  mjx_host_alloc(mjx_ctx *ctx, void **ptr, size_t bytes); /* page-locked */
void mjx_host_free(mjx_ctx *ctx, void *ptr);
/* a grow-only page-locked scratch buffer owned by the ctx (valid until the next call that asks for more, or destroy) */
int  mjx_ctx_pinned_scratch(mjx_ctx *ctx, size_t bytes, void **ptr);
/* the same in device memory: one grow-only buffer owned by the ctx (freed with it); growing waits for the ctx's streams and
 * loses the contents */
int  mjx_ctx_device_scratch(mjx_ctx *ctx, size_t bytes, void **ptr);
int  mjx_copy_h2d(mjx_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes); /* async on the ctx stream */
int  mjx_copy_d2h(mjx_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes);

/* ---- A1: placement arithmetic, host only (replaces src/compose.c:42-172) ------------- */
void mjx_geometry(int image_width, int image_height, int h_factor, int v_factor, int dropon_width, int dropon_height,
                  unsigned int align, int offset_x, int offset_y, mjx_geometry_t *out);

/* ---- K1: dropon compile (replaces mj_compile_dropon, src/dropon.c:325-576, and the two
 *      libjpeg encode + decode round trips it makes through src/image.c:257-347) -------- */
int  mjx_dropon_compile(mjx_ctx *ctx, mjx_dropon **out,
                        const uint8_t *image3, const uint8_t *alpha3, /* mj_dropon_t.image / .alpha: 3 bytes per pixel */
                        int width, int height, int dropon_colorspace, /* MJX_CS_* */
                        const mjx_layout_t *layout,
                        int blockoffset_x, int blockoffset_y, int crop_x, int crop_y, int crop_w, int crop_h,
                        int pixels_on_device); /* 0: image3/alpha3 are host pointers; 1: device pointers */
/* build a compiled dropon from host-side coefficient planes D[c], W[c] ([hb][wb][64] int16, W with
 * DC already += 1024) -- lets K2 be driven with a dropon compiled elsewhere (e.g. by the oracle) */
int  mjx_dropon_from_coefficients(mjx_ctx *ctx, mjx_dropon **out, const mjx_layout_t *layout,
                                  const int *wb, const int *hb, const int16_t *const *D, const int16_t *const *W);
void mjx_dropon_free(mjx_dropon *d);
int  mjx_dropon_ncomp(const mjx_dropon *d);
int  mjx_dropon_dims(const mjx_dropon *d, int comp, int *wb, int *hb);
long long mjx_dropon_blocks(const mjx_dropon *d); /* all components */
/* copy component `comp` back: D, W int16 [hb][wb][64]; cls uint8 [hb][wb] (any may be NULL) */
int  mjx_dropon_download(mjx_ctx *ctx, const mjx_dropon *d, int comp, int16_t *D, int16_t *W, uint8_t *cls);
/* the generic-class work list: n = mjx_dropon_generic_slots() entries (comp << 30 | row << 15 | col; every component
 * starts on a multiple of 32, padding slots hold 0xffffffff) and, in list order, Ds = D * IDCT prescale (natural order)
 * and A = pixel-domain alpha / 255 stored row-paired: float (8i + k)*2 + h = A[row 2i + h][col k]; 64 floats per slot
 * (any pointer may be NULL) */
int  mjx_dropon_generic_slots(const mjx_dropon *d);
int  mjx_dropon_download_generic(mjx_ctx *ctx, const mjx_dropon *d, uint32_t *list, float *Ds, float *A);
/* counts[MJX_CLS_*] summed over all components */
int  mjx_dropon_class_counts(mjx_ctx *ctx, const mjx_dropon *d, long long counts[4]);

/* ---- K2: masked blend (replaces mj_compose_with_mask + mj_convolve, src/compose.c:237-342,
 *      src/convolve.c:29-1099) ---------------------------------------------------------- */
/* n images resident in HBM, one compiled dropon at MCU position (block_x, block_y) on each */
int mjx_compose_batch_device(mjx_ctx *ctx, const mjx_image_desc_t *items_dev, int n, const mjx_dropon *d,
                             int block_x, int block_y);
/* n images in host memory.  Page-locked planes (mjx_host_alloc / cudaHostRegister): one launch working directly on
 * host memory (zero-copy, see mjx_ctx_set_zero_copy); pageable planes: the region under the dropon is staged H2D,
 * blended and staged back through a 3-stream pipeline.  Returns when the host planes hold the result.
 * Unlike mjx_compose_batch_device -- which skips, per image, the dropon blocks that fall outside that image's planes -- this
 * call wants the dropon's whole region inside every plane and returns MJX_ERR_ARG otherwise (the staged path moves the region
 * as one rectangle); mjx_geometry() crops a dropon to the image, so callers that place dropons with it never see the error. */
int mjx_compose_batch_host(mjx_ctx *ctx, const mjx_host_image_t *items, int n, const mjx_dropon *d,
                           int block_x, int block_y);
/* one image given as libjpeg row pointers: rows[c][l] -> block (block_y*v_c + l, block_x*h_c) of component c,
 * l in [0, hb_c) (what access_virt_barray returns, reference: src/compose.c:269) */
int mjx_compose_rows_host(mjx_ctx *ctx, int ncomp, int16_t *const *const *rows, const uint16_t *const *q,
                          const mjx_dropon *d);

/* ---- K4: Huffman coding of a baseline scan on the device -----------------------------------------
 *      replaces, for mj_write_jpeg_to_memory (reference: src/image.c:120-209), the entropy encoder that
 *      jpeg_write_coefficients / jpeg_finish_compress run on the host (libjpeg jctrans.c compress_output, jchuff.c
 *      encode_mcu_huff): sequential DCT, Huffman tables as given (not optimised), no restart markers, ONE scan holding
 *      every component.  Byte-identical to libjpeg's entropy-coded segment. */
typedef struct {
    uint8_t bits[17];  /* bits[k] = number of codes of length k, k = 1..16 (JHUFF_TBL.bits, the DHT segment's list) */
    uint8_t vals[256]; /* the symbols in order of increasing code length (JHUFF_TBL.huffval) */
} mjx_huff_table_t;
typedef struct {
    int32_t          ncomp;                                          /* components of the scan = of the image, in order */
    int32_t          h_samp[MJX_MAX_COMPONENTS], v_samp[MJX_MAX_COMPONENTS]; /* (ignored when ncomp == 1: not interleaved) */
    int32_t          dc_tbl[MJX_MAX_COMPONENTS], ac_tbl[MJX_MAX_COMPONENTS]; /* which of dc[] / ac[] a component uses */
    int32_t          mcus_per_row, mcu_rows;                         /* ncomp == 1: the component's width / height in blocks */
    mjx_huff_table_t dc[4], ac[4];                                   /* unused tables: all zero */
} mjx_scan_t;
/* n images resident in HBM (planes as for K2; wreal / hreal of the descriptors say where the encoder's dummy blocks start).
 * Image i's segment -- byte-stuffed, last byte padded with 1-bits, no EOI -- goes to out_dev + i * out_stride, its length to
 * sizes_dev[i]; 0xFFFFFFFF there means "not coded": a coefficient the tables cannot code (DC difference beyond 11 bits, AC
 * beyond 10: libjpeg's JERR_BAD_DCT_COEF), a symbol without a code, or a segment longer than out_stride.  Asynchronous on
 * the ctx stream. */
int mjx_huffman_encode_batch_device(mjx_ctx *ctx, const mjx_image_desc_t *items_dev, int n, const mjx_scan_t *scan,
                                    void *out_dev, size_t out_stride, uint32_t *sizes_dev);
/* one image as libjpeg row pointers: rows[c][l] -> block (l, 0) of component c, l in [0, hreal[c]), stride_blocks[c] blocks
 * per row.  *out is malloc()ed (the caller frees it).  MJX_ERR_UNSUPPORTED: not codable, see above. */
int mjx_huffman_encode_rows_host(mjx_ctx *ctx, int ncomp, const int16_t *const *const *rows, const int *stride_blocks,
                                 const int *vrows, const int *wreal, const int *hreal, const mjx_scan_t *scan,
                                 unsigned char **out, size_t *len);

/* ---- K5: Huffman decoding of a baseline scan on the device (second half) ----------------------------
 *      replaces, for mj_read_jpeg_from_memory (reference: src/image.c:33-118), the entropy decoder jpeg_read_coefficients
 *      runs on the host (libjpeg jdhuff.c decode_mcu): 8-bit sequential DCT, ONE scan with every component, no restart
 *      markers.  data_dev + offsets[i] .. + lengths[i]: image i's entropy-coded segment as it stands in the file (byte
 *      stuffing included, from behind the SOS header; bytes behind the last MCU are ignored).  offsets / lengths are HOST
 *      arrays.  The planes the descriptors name are overwritten (padding blocks included, like libjpeg's arrays); wreal /
 *      hreal / q of the descriptors are not looked at.  status_dev[i] != 0: not decoded (a bit pattern that is no code, a run
 *      past coefficient 63, a stream with fewer blocks than the frame) -- let libjpeg read that image.  Asynchronous. */
int mjx_huffman_decode_batch_device(mjx_ctx *ctx, const void *data_dev, const uint64_t *offsets, const uint32_t *lengths, int n,
                                    const mjx_scan_t *scan, const mjx_image_desc_t *items_dev, uint32_t *status_dev);

/* ---- K3: coefficient effects (replaces src/effect.c:28-222) --------------------------- */
int mjx_effects_batch_device(mjx_ctx *ctx, const mjx_image_desc_t *items_dev, int n, int ncomp,
                             const mjx_effect_op_t *ops, int nops);
/* one image as libjpeg row pointers: rows[c][l] -> block (l, 0), l in [0, hreal[c]); only the
 * components named by ops are staged */
int mjx_effects_rows_host(mjx_ctx *ctx, int ncomp, int16_t *const *const *rows, const int *wreal, const int *hreal,
                          const uint16_t *const *q, const mjx_effect_op_t *ops, int nops);

#ifdef __cplusplus
}
#endif

#endif /* MJX_H */
