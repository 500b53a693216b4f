/*
 * mj_oracle.c -- CPU restatement of libmodjpeg's DCT-domain compositing hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this.  The product
 * (libmodjpeg_b200, libmjx.so / libmodjpeg.so) never links, loads or calls it.
 *
 * Parity pin: every function here is checked bit-for-bit against the UNMODIFIED reference
 * compiled as oracle/_ref/libmodjpeg_ref.so (tests/test_oracle_vs_ref.py) and against the
 * golden vectors in tests/golden/ (generated from that build by tests/golden/make_golden.py,
 * plus the reference's own README fixture image.jpg + dropon.png -> image_dropon.jpg).
 *
 * It restates, in plain C on flat arrays (no libjpeg):
 *   A1  mj_compose placement/crop geometry              reference: src/compose.c:42-172
 *   A2  mj_compile_dropon, image half                   reference: src/dropon.c:325-389,430-495
 *       (+ what libjpeg-turbo 3.1.x does inside mj_encode_raw_to_jpeg_memory,
 *        src/image.c:257-347: colour convert, downsample, islow FDCT, quality-100 quantise)
 *   A3  mj_compile_dropon, alpha half                   reference: src/dropon.c:391-427,497-576
 *   A4  mj_compose_with_mask per-block blend            reference: src/compose.c:237-342
 *   A5  mj_convolve                                     reference: src/convolve.c:29-1099
 *   A7-A9 mj_effect_grayscale/pixelate/tint/luminance   reference: src/effect.c:28-222
 *
 * libjpeg (the reference's un-vendored dependency, CMakeLists.txt:13) is pinned to
 * libjpeg-turbo 3.1.4.1, the only libjpeg in this image; its published integer algorithms
 * (jccolor.c rgb_ycc_convert, jcsample.c h2v1/h2v2/int_downsample, jfdctint.c jpeg_fdct_islow,
 * jcdctmgr.c quantize) are restated below from their documented arithmetic.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (no -ffast-math, no -march=native: the
 * float/double evaluation order below must round exactly like the reference built with -O2).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MJO_CS_RGB       1 /* MJ_COLORSPACE_RGB       (dropon pixel formats, after ingest) */
#define MJO_CS_GRAYSCALE 3 /* MJ_COLORSPACE_GRAYSCALE */
#define MJO_CS_YCC       5 /* MJ_COLORSPACE_YCC       */

#define MJO_JCS_GRAYSCALE 1 /* J_COLOR_SPACE values of the target JPEG */
#define MJO_JCS_RGB       2
#define MJO_JCS_YCbCr     3

#define MJO_ALIGN_LEFT   (1 << 0)
#define MJO_ALIGN_RIGHT  (1 << 1)
#define MJO_ALIGN_TOP    (1 << 2)
#define MJO_ALIGN_BOTTOM (1 << 3)

#define MJO_OK              0
#define MJO_ERR_MEMORY      1
#define MJO_ERR_NULL_DATA   2
#define MJO_ERR_ENCODE_JPEG 6 /* what the reference returns for unsupported conversions */

/* ------------------------------------------------------------------------------------ */
/* A1: geometry of mj_compose (reference: src/compose.c:42-172)                           */
/* ------------------------------------------------------------------------------------ */

typedef struct {
    int visible; /* 0: nothing to do (reference returns MJ_OK at compose.c:136) */
    int crop_x, crop_y, crop_w, crop_h;
    int blockoffset_x, blockoffset_y;
    int block_x, block_y; /* in MCUs */
} mjo_geometry_t;

void mjo_geometry(int img_w, int img_h, int h_factor, int v_factor, int d_w, int d_h,
                  unsigned int align, int offset_x, int offset_y, mjo_geometry_t *g) {
    int position_x, position_y;

    /* compose.c:57-68 */
    if(align & MJO_ALIGN_LEFT) position_x = 0;
    else if(align & MJO_ALIGN_RIGHT) position_x = img_w - d_w;
    else position_x = img_w / 2 - d_w / 2;
    position_x += offset_x;

    /* compose.c:71-82 */
    if(align & MJO_ALIGN_TOP) position_y = 0;
    else if(align & MJO_ALIGN_BOTTOM) position_y = img_h - d_h;
    else position_y = img_h / 2 - d_h / 2;
    position_y += offset_y;

    /* compose.c:87-109 */
    int crop_x = position_x < 0 ? -position_x : 0;
    int crop_w = d_w - crop_x;
    if(crop_x > d_w) crop_w = 0;
    else if(position_x > img_w) crop_w = 0;
    else if(position_x + crop_x + crop_w > img_w) crop_w = img_w - crop_x - position_x;

    /* compose.c:111-133 */
    int crop_y = position_y < 0 ? -position_y : 0;
    int crop_h = d_h - crop_y;
    if(crop_y > d_h) crop_h = 0;
    else if(position_y > img_h) crop_h = 0;
    else if(position_y + crop_y + crop_h > img_h) crop_h = img_h - crop_y - position_y;

    memset(g, 0, sizeof(*g));
    g->crop_x = crop_x;
    g->crop_y = crop_y;
    g->crop_w = crop_w;
    g->crop_h = crop_h;
    if(crop_w == 0 || crop_h == 0) { /* compose.c:136 */
        g->visible = 0;
        return;
    }
    g->visible = 1;

    /* compose.c:144-151: C remainder keeps the sign of the dividend, then clamp */
    g->blockoffset_x = position_x % h_factor;
    if(g->blockoffset_x < 0) g->blockoffset_x = 0;
    g->blockoffset_y = position_y % v_factor;
    if(g->blockoffset_y < 0) g->blockoffset_y = 0;

    /* compose.c:163-172: truncating division, then clamp */
    g->block_x = position_x / h_factor;
    g->block_y = position_y / v_factor;
    if(g->block_x < 0) g->block_x = 0;
    if(g->block_y < 0) g->block_y = 0;
}

/* ------------------------------------------------------------------------------------ */
/* A2/A3: dropon compile (reference: src/dropon.c:325-576 + libjpeg-turbo integer pipeline) */
/* ------------------------------------------------------------------------------------ */

typedef struct {
    int colorspace; /* target J_COLOR_SPACE: 1 gray, 2 RGB, 3 YCbCr */
    int ncomp;
    int h[4], v[4]; /* sampling factors of the target JPEG's components */
} mjo_layout_t;

static int max_of(const int *a, int n) {
    int m = a[0];
    for(int i = 1; i < n; i++)
        if(a[i] > m) m = a[i];
    return m;
}

/* padded canvas size: dropon.c:340-350 */
void mjo_canvas_dims(const mjo_layout_t *L, int blockoffset_x, int blockoffset_y, int crop_w, int crop_h,
                     int *width, int *height) {
    int hf = max_of(L->h, L->ncomp) * 8, vf = max_of(L->v, L->ncomp) * 8;
    int w = crop_w + blockoffset_x;
    int pad = w % hf;
    if(pad != 0) w += hf - pad;
    int h = crop_h + blockoffset_y;
    pad = h % vf;
    if(pad != 0) h += vf - pad;
    *width = w;
    *height = h;
}

/* per-component dims of the compiled dropon in blocks (what libjpeg reports as
 * width_in_blocks/height_in_blocks for an MCU-aligned image; dropon.c:458-459) */
void mjo_compiled_dims(const mjo_layout_t *L, int blockoffset_x, int blockoffset_y, int crop_w, int crop_h,
                       int *wb, int *hb) {
    int W, H;
    mjo_canvas_dims(L, blockoffset_x, blockoffset_y, crop_w, crop_h, &W, &H);
    int mh = max_of(L->h, L->ncomp), mv = max_of(L->v, L->ncomp);
    for(int c = 0; c < L->ncomp; c++) {
        wb[c] = (W / (mh * 8)) * L->h[c];
        hb[c] = (H / (mv * 8)) * L->v[c];
    }
}

/* libjpeg-turbo jccolor.c, 16-bit fixed point (SCALEBITS 16) */
static inline int cc_y(int r, int g, int b) { return (19595 * r + 38470 * g + 7471 * b + 32768) >> 16; }
static inline int cc_cb(int r, int g, int b) { return (-11059 * r - 21709 * g + 32768 * b + (128 << 16) + 32767) >> 16; }
static inline int cc_cr(int r, int g, int b) { return (32768 * r - 27439 * g - 5329 * b + (128 << 16) + 32767) >> 16; }

/* jfdctint.c jpeg_fdct_islow: CONST_BITS 13, PASS1_BITS 2.  In place on 64 ints, rows first. */
#define FIX_0_298631336 2446
#define FIX_0_390180644 3196
#define FIX_0_541196100 4433
#define FIX_0_765366865 6270
#define FIX_0_899976223 7373
#define FIX_1_175875602 9633
#define FIX_1_501321110 12299
#define FIX_1_847759065 15137
#define FIX_1_961570560 16069
#define FIX_2_053119869 16819
#define FIX_2_562915447 20995
#define FIX_3_072711026 25172
#define DESCALE(x, n) (((x) + (1 << ((n)-1))) >> (n))

static void fdct_islow_1d(int *d, int stride, int pass) {
    int d0 = d[0], d1 = d[stride], d2 = d[2 * stride], d3 = d[3 * stride];
    int d4 = d[4 * stride], d5 = d[5 * stride], d6 = d[6 * stride], d7 = d[7 * stride];
    int tmp0 = d0 + d7, tmp7 = d0 - d7, tmp1 = d1 + d6, tmp6 = d1 - d6;
    int tmp2 = d2 + d5, tmp5 = d2 - d5, tmp3 = d3 + d4, tmp4 = d3 - d4;
    int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    int sh = pass == 0 ? 13 - 2 : 13 + 2;

    if(pass == 0) {
        d[0] = (tmp10 + tmp11) * 4;
        d[4 * stride] = (tmp10 - tmp11) * 4;
    }
    else {
        d[0] = DESCALE(tmp10 + tmp11, 2);
        d[4 * stride] = DESCALE(tmp10 - tmp11, 2);
    }
    int z1 = (tmp12 + tmp13) * FIX_0_541196100;
    d[2 * stride] = DESCALE(z1 + tmp13 * FIX_0_765366865, sh);
    d[6 * stride] = DESCALE(z1 + tmp12 * (-FIX_1_847759065), sh);

    z1 = tmp4 + tmp7;
    int z2 = tmp5 + tmp6, z3 = tmp4 + tmp6, z4 = tmp5 + tmp7;
    int z5 = (z3 + z4) * FIX_1_175875602;
    tmp4 *= FIX_0_298631336;
    tmp5 *= FIX_2_053119869;
    tmp6 *= FIX_3_072711026;
    tmp7 *= FIX_1_501321110;
    z1 *= -FIX_0_899976223;
    z2 *= -FIX_2_562915447;
    z3 *= -FIX_1_961570560;
    z4 *= -FIX_0_390180644;
    z3 += z5;
    z4 += z5;
    d[7 * stride] = DESCALE(tmp4 + z1 + z3, sh);
    d[5 * stride] = DESCALE(tmp5 + z2 + z4, sh);
    d[3 * stride] = DESCALE(tmp6 + z2 + z3, sh);
    d[1 * stride] = DESCALE(tmp7 + z1 + z4, sh);
}

/* 8x8 samples (already level-shifted by -128) -> quantised coefficients for q == 1
 * (quality 100 forced baseline, image.c:327): sign(x) * ((|x| + 4) >> 3). */
static void fdct_quant_q1(const int *samples, int16_t *out) {
    int w[64];
    memcpy(w, samples, sizeof(w));
    for(int r = 0; r < 8; r++) fdct_islow_1d(w + 8 * r, 1, 0);
    for(int c = 0; c < 8; c++) fdct_islow_1d(w + c, 8, 1);
    for(int i = 0; i < 64; i++) {
        int t = w[i];
        out[i] = (int16_t)(t < 0 ? -((-t + 4) >> 3) : ((t + 4) >> 3));
    }
}

/* Build the full-resolution sample plane of component c for one canvas, i.e. what libjpeg's
 * colour converter hands to the downsampler.  `src3` is a W x H x 3 canvas (dropon.c:352-369:
 * zero-filled, crop copied in at the block offset).  in_cs is the colourspace libjpeg was told
 * the canvas has.  Returns 0, or MJO_ERR_ENCODE_JPEG for a conversion libjpeg rejects. */
static int convert_component(const uint8_t *src3, int W, int H, int in_cs, int target_cs, int c, uint8_t *plane) {
    size_t n = (size_t)W * H;
    if(in_cs == MJO_CS_GRAYSCALE) {
        /* image.c:295-297,331: the 3-byte canvas is handed to libjpeg as ONE byte per pixel with
         * row_stride = width, so sample (x,y) is byte y*W + x of the canvas (the reference's
         * documented-as-buggy behaviour for grayscale dropons; kept because it is what it does). */
        if(target_cs != MJO_JCS_GRAYSCALE) return MJO_ERR_ENCODE_JPEG;
        memcpy(plane, src3, n);
        return 0;
    }
    if(in_cs == MJO_CS_RGB) {
        if(target_cs == MJO_JCS_RGB) {
            for(size_t i = 0; i < n; i++) plane[i] = src3[3 * i + c];
            return 0;
        }
        if(target_cs == MJO_JCS_YCbCr || target_cs == MJO_JCS_GRAYSCALE) {
            for(size_t i = 0; i < n; i++) {
                int r = src3[3 * i], g = src3[3 * i + 1], b = src3[3 * i + 2];
                plane[i] = (uint8_t)(c == 0 ? cc_y(r, g, b) : c == 1 ? cc_cb(r, g, b) : cc_cr(r, g, b));
            }
            return 0;
        }
        return MJO_ERR_ENCODE_JPEG;
    }
    if(in_cs == MJO_CS_YCC) {
        if(target_cs == MJO_JCS_YCbCr || target_cs == MJO_JCS_GRAYSCALE) {
            for(size_t i = 0; i < n; i++) plane[i] = src3[3 * i + c]; /* gray: c == 0 */
            return 0;
        }
        return MJO_ERR_ENCODE_JPEG;
    }
    return MJO_ERR_ENCODE_JPEG;
}

/* libjpeg-turbo jcsample.c: fullsize copy, h2v1, h2v2 (alternating bias), generic box mean */
static void downsample(const uint8_t *full, int W, int H, int he, int ve, uint8_t *out) {
    int ow = W / he, oh = H / ve;
    if(he == 1 && ve == 1) {
        memcpy(out, full, (size_t)W * H);
        return;
    }
    if(he == 2 && ve == 1) {
        for(int y = 0; y < oh; y++)
            for(int x = 0; x < ow; x++) {
                const uint8_t *p = full + (size_t)y * W + 2 * x;
                out[(size_t)y * ow + x] = (uint8_t)((p[0] + p[1] + (x & 1)) >> 1);
            }
        return;
    }
    if(he == 2 && ve == 2) {
        for(int y = 0; y < oh; y++)
            for(int x = 0; x < ow; x++) {
                const uint8_t *p0 = full + (size_t)(2 * y) * W + 2 * x, *p1 = p0 + W;
                out[(size_t)y * ow + x] = (uint8_t)((p0[0] + p0[1] + p1[0] + p1[1] + 1 + (x & 1)) >> 2);
            }
        return;
    }
    int numpix = he * ve, numpix2 = numpix / 2;
    for(int y = 0; y < oh; y++)
        for(int x = 0; x < ow; x++) {
            int s = 0;
            for(int v = 0; v < ve; v++)
                for(int h = 0; h < he; h++) s += full[(size_t)(y * ve + v) * W + x * he + h];
            out[(size_t)y * ow + x] = (uint8_t)((s + numpix2) / numpix);
        }
}

/* canvas -> per-component int16 coefficient planes [hb][wb][64] (natural order) */
static int encode_canvas(const uint8_t *canvas, int W, int H, int in_cs, const mjo_layout_t *L,
                         int16_t *const *out, int dc_add) {
    int mh = max_of(L->h, L->ncomp), mv = max_of(L->v, L->ncomp);
    uint8_t *full = (uint8_t *)malloc((size_t)W * H), *ds = (uint8_t *)malloc((size_t)W * H);
    if(!full || !ds) {
        free(full);
        free(ds);
        return MJO_ERR_MEMORY;
    }
    for(int c = 0; c < L->ncomp; c++) {
        if(mh % L->h[c] || mv % L->v[c]) { /* jcsample.c: fractional sampling not implemented */
            free(full);
            free(ds);
            return MJO_ERR_ENCODE_JPEG;
        }
        int rv = convert_component(canvas, W, H, in_cs, L->colorspace, c, full);
        if(rv) {
            free(full);
            free(ds);
            return rv;
        }
        int he = mh / L->h[c], ve = mv / L->v[c];
        downsample(full, W, H, he, ve, ds);
        int ow = W / he, oh = H / ve, wb = ow / 8, hb = oh / 8;
        for(int by = 0; by < hb; by++)
            for(int bx = 0; bx < wb; bx++) {
                int s[64];
                for(int y = 0; y < 8; y++)
                    for(int x = 0; x < 8; x++) s[8 * y + x] = (int)ds[(size_t)(by * 8 + y) * ow + bx * 8 + x] - 128;
                int16_t *o = out[c] + ((size_t)by * wb + bx) * 64;
                fdct_quant_q1(s, o);
                o[0] = (int16_t)(o[0] + dc_add); /* dropon.c:542 for the alpha planes */
            }
    }
    free(full);
    free(ds);
    return 0;
}

/* mj_compile_dropon (dropon.c:325-428).  image3/alpha3: the dropon's 3-byte-per-pixel buffers
 * (mj_dropon_t.image / .alpha), dw x dh.  D[c], W[c]: caller-allocated [hb_c][wb_c][64] int16.
 * D = integer DCT coefficients of the overlay; W = those of the alpha mask with DC += 1024. */
int mjo_compile_dropon(const uint8_t *image3, const uint8_t *alpha3, int dw, int dh, int dropon_cs,
                       const mjo_layout_t *L, int blockoffset_x, int blockoffset_y,
                       int crop_x, int crop_y, int crop_w, int crop_h,
                       int16_t *const *D, int16_t *const *W) {
    (void)dh;
    /* jcmaster.c per_scan_setup: the interleaved scan libjpeg would write holds at most
     * C_MAX_BLOCKS_IN_MCU (10) blocks per MCU, else "Sampling factors too large" -> error 6 */
    if(L->ncomp > 1) {
        int blocks = 0;
        for(int c = 0; c < L->ncomp; c++) blocks += L->h[c] * L->v[c];
        if(blocks > 10) return MJO_ERR_ENCODE_JPEG;
    }
    int width, height;
    mjo_canvas_dims(L, blockoffset_x, blockoffset_y, crop_w, crop_h, &width, &height);
    uint8_t *data = (uint8_t *)calloc((size_t)3 * width * height, 1);
    if(!data) return MJO_ERR_MEMORY;

    for(int i = crop_y; i < crop_y + crop_h; i++) /* dropon.c:360-369 */
        memcpy(data + ((size_t)(i - crop_y + blockoffset_y) * width + blockoffset_x) * 3,
               image3 + ((size_t)i * dw + crop_x) * 3, (size_t)crop_w * 3);
    int rv = encode_canvas(data, width, height, dropon_cs, L, D, 0);
    if(rv) {
        free(data);
        return rv;
    }
    for(int i = crop_y; i < crop_y + crop_h; i++) /* dropon.c:391-400 (same buffer, not re-zeroed) */
        memcpy(data + ((size_t)(i - crop_y + blockoffset_y) * width + blockoffset_x) * 3,
               alpha3 + ((size_t)i * dw + crop_x) * 3, (size_t)crop_w * 3);
    int alpha_cs = L->colorspace == MJO_JCS_RGB ? MJO_CS_RGB : MJO_CS_YCC; /* dropon.c:411-414 */
    rv = encode_canvas(data, width, height, alpha_cs, L, W, 1024);
    free(data);
    return rv;
}

/* ------------------------------------------------------------------------------------ */
/* A3: alpha weights (reference: src/dropon.c:548-566)                                    */
/* ------------------------------------------------------------------------------------ */

void mjo_alpha_weights(const int16_t *Wc, float *w) {
    w[0] = (float)Wc[0] * (0.3535534 * 0.3535534 / 1020.0);
    for(int i = 1; i < 8; i++) w[i] = (float)Wc[i] * (0.3535534 * 0.5 / 1020.0);
    for(int i = 8; i < 64; i += 8) {
        w[i] = (float)Wc[i] * (0.5 * 0.3535534 / 1020.0);
        for(int j = 1; j < 8; j++) w[i + j] = (float)Wc[i + j] * (0.5 * 0.5 / 1020.0);
    }
}

/* ------------------------------------------------------------------------------------ */
/* A5: mj_convolve (reference: src/convolve.c:29-1099), table-driven restatement.         */
/*                                                                                        */
/* The 1 099 unrolled lines are y += w * (M_k (x) M_l) x with eight symmetric 8x8 matrices: */
/* M_0 = 2*I; for l >= 1 column 0 has one entry sqrt2 at row l, and column i >= 1 has +1   */
/* at row |i-l| (sqrt2 when that row is 0), +1 at row i+l if i+l < 8, -1 at row 16-i-l if  */
/* i+l > 8.  Every output row therefore has at most two terms.  Rounding is reproduced:   */
/* an expression containing the double constants 2.0 / M_SQRT2 is evaluated in double and */
/* rounded to float on the store; one without is evaluated in float (FLT_EVAL_METHOD 0).  */
/* ------------------------------------------------------------------------------------ */

#define MJO_SQRT2 1.41421356237309504880 /* M_SQRT2 */

typedef struct {
    int n;      /* number of terms (1 or 2) */
    int idx[2]; /* source index 0..7 */
    int kind[2]; /* 0: +1, 1: -1, 2: *sqrt2 (double), 3: *2.0 (double) */
} mjo_row_t;

static mjo_row_t g_rows[8][8];
static int g_rows_ready = 0;

static void build_rows(void) {
    for(int l = 0; l < 8; l++) {
        int M[8][8]; /* 0 none, 1:+1, 2:-1, 3:sqrt2, 4:2.0 */
        memset(M, 0, sizeof(M));
        if(l == 0) {
            for(int i = 0; i < 8; i++) M[i][i] = 4;
        }
        else {
            M[l][0] = 3;
            for(int i = 1; i < 8; i++) {
                int r = i > l ? i - l : l - i;
                M[r][i] = r == 0 ? 3 : 1;
                if(i + l < 8) M[i + l][i] = 1;
                else if(i + l > 8) M[16 - i - l][i] = 2;
            }
        }
        for(int j = 0; j < 8; j++) {
            mjo_row_t *R = &g_rows[l][j];
            R->n = 0;
            /* double-constant term first, as in the source text; addition order does not
             * change the IEEE result of a two-operand sum */
            for(int pass = 0; pass < 2; pass++)
                for(int i = 0; i < 8; i++) {
                    int m = M[j][i];
                    if(m == 0) continue;
                    int is_dbl = (m == 3 || m == 4);
                    if((pass == 0) != is_dbl) continue;
                    R->idx[R->n] = i;
                    R->kind[R->n] = m == 1 ? 0 : m == 2 ? 1 : m == 3 ? 2 : 3;
                    R->n++;
                }
        }
    }
    g_rows_ready = 1;
}

/* value of row R applied to v[0], v[stride], ...; *is_double tells which arithmetic was used */
static inline double row_eval(const mjo_row_t *R, const float *v, int stride, int *is_double) {
    if(R->kind[0] >= 2) {
        double a = (R->kind[0] == 2 ? MJO_SQRT2 : 2.0) * v[R->idx[0] * stride];
        if(R->n == 2) {
            if(R->kind[1] == 0) a = a + v[R->idx[1] * stride];
            else a = a - v[R->idx[1] * stride];
        }
        *is_double = 1;
        return a;
    }
    /* pure float expression; a leading -1 never occurs alone in the source */
    float a;
    if(R->n == 1) a = v[R->idx[0] * stride];
    else if(R->kind[0] == 0 && R->kind[1] == 0) a = v[R->idx[0] * stride] + v[R->idx[1] * stride];
    else if(R->kind[0] == 0) a = v[R->idx[0] * stride] - v[R->idx[1] * stride];
    else a = v[R->idx[1] * stride] - v[R->idx[0] * stride];
    *is_double = 0;
    return a;
}

void mjo_convolve(const float *x, float *y, float w, int k, int l) {
    if(!g_rows_ready) build_rows();
    float z[64];
    if(w == 0.0) return; /* convolve.c:32 */

    /* stage 1 (convolve.c:36-565): z = (I (x) M_l) x, along each row of 8 */
    for(int r = 0; r < 8; r++)
        for(int j = 0; j < 8; j++) {
            int dbl;
            double a = row_eval(&g_rows[l][j], x + 8 * r, 1, &dbl);
            z[8 * r + j] = (float)a;
        }
    /* stage 2 (convolve.c:567-1096): y += ((M_k (x) I) z) * w, along each column */
    for(int j = 0; j < 8; j++)
        for(int c = 0; c < 8; c++) {
            int dbl;
            double a = row_eval(&g_rows[k][j], z + c, 8, &dbl);
            if(dbl) y[8 * j + c] = (float)((double)y[8 * j + c] + a * (double)w);
            else y[8 * j + c] = y[8 * j + c] + (float)a * w;
        }
}

/* ------------------------------------------------------------------------------------ */
/* A4: mj_compose_with_mask (reference: src/compose.c:237-342)                            */
/* ------------------------------------------------------------------------------------ */

/* one block, in place on the image's int16 coefficients */
void mjo_compose_block(int16_t *I, const int16_t *Dc, const int16_t *Wc, const uint16_t *q) {
    float X[64], Y[64], w[64], Df[64];
    for(int i = 0; i < 64; i++) Df[i] = (float)Dc[i]; /* dropon.c:476-485 */
    mjo_alpha_weights(Wc, w);
    for(int i = 0; i < 64; i++) I[i] = (int16_t)(I[i] * q[i]); /* compose.c:277-286 */
    for(int i = 0; i < 64; i++) X[i] = Df[i] - I[i];           /* compose.c:289-298 */
    memset(Y, 0, sizeof(Y));
    for(int i = 0; i < 8; i++)                                  /* compose.c:303-312 */
        for(int j = 0; j < 8; j++) mjo_convolve(X, Y, w[8 * i + j], i, j);
    for(int i = 0; i < 64; i++) I[i] = (int16_t)(I[i] + (int)Y[i]); /* compose.c:315-324 */
    for(int i = 0; i < 64; i++) I[i] = (int16_t)(I[i] / q[i]);      /* compose.c:327-336 */
}

/* one component: plane is [rows][stride_blocks][64]; the compiled dropon's block (l,k) lands
 * on plane block (y0 + l, x0 + k) (compose.c:264-274) */
void mjo_compose_plane(int16_t *plane, int stride_blocks, int x0, int y0, const int16_t *Dp, const int16_t *Wp,
                       int wb, int hb, const uint16_t *q) {
    for(int l = 0; l < hb; l++)
        for(int k = 0; k < wb; k++)
            mjo_compose_block(plane + ((size_t)(y0 + l) * stride_blocks + x0 + k) * 64,
                              Dp + ((size_t)l * wb + k) * 64, Wp + ((size_t)l * wb + k) * 64, q);
}

/* ------------------------------------------------------------------------------------ */
/* A7-A9: effects (reference: src/effect.c:28-222), on the REAL blocks of one plane       */
/* ------------------------------------------------------------------------------------ */

void mjo_effect_zero_plane(int16_t *plane, int stride_blocks, int wreal, int hreal) { /* effect.c:43-64 */
    for(int l = 0; l < hreal; l++)
        for(int k = 0; k < wreal; k++) memset(plane + ((size_t)l * stride_blocks + k) * 64, 0, 128);
}

void mjo_effect_pixelate_plane(int16_t *plane, int stride_blocks, int wreal, int hreal) { /* effect.c:81-110 */
    for(int l = 0; l < hreal; l++)
        for(int k = 0; k < wreal; k++) memset(plane + ((size_t)l * stride_blocks + k) * 64 + 1, 0, 126);
}

void mjo_effect_add_dc_plane(int16_t *plane, int stride_blocks, int wreal, int hreal, uint16_t q0, int value) {
    for(int l = 0; l < hreal; l++) /* effect.c:135-155, 199-219 */
        for(int k = 0; k < wreal; k++) {
            int16_t *c = plane + ((size_t)l * stride_blocks + k) * 64;
            c[0] = (int16_t)(c[0] * q0);
            c[0] = (int16_t)(c[0] + value);
            if(c[0] > 2047) c[0] = 2047;
            else if(c[0] < -2047) c[0] = -2047;
            c[0] = (int16_t)(c[0] / q0);
        }
}
