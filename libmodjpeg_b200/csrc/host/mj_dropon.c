/*
 * mj_dropon.c -- dropon ingest of the host boundary: mj_init_dropon, mj_free_dropon,
 * mj_read_dropon_from_raw/_memory/_file.  One-time, tiny, host-side; keeps the
 * reference's storage contract (reference: src/dropon.c:33-323,578-604): `image` and `alpha`
 * are both width*height*3 bytes, alpha replicated over the three channels, and any
 * alpha-carrying format forces blend = MJ_BLEND_NONUNIFORM.  K1 consumes these two buffers.
 */
#include <stdlib.h>
#include <string.h>

#include "mj_private.h"

#ifdef WITH_LIBPNG
#include <png.h>
#endif

void mj_init_dropon(mj_dropon_t *d) {
    if(d != NULL) memset(d, 0, sizeof(*d));
}

extern unsigned long mjp_dropon_generation; /* mj_compose.c: invalidates cached compiled dropons */

void mj_free_dropon(mj_dropon_t *d) {
    if(d == NULL) return;
    if(d->image != NULL || d->alpha != NULL) __atomic_add_fetch(&mjp_dropon_generation, 1, __ATOMIC_RELAXED);
    free(d->image);
    free(d->alpha);
    mj_init_dropon(d);
}

int mj_read_dropon_from_raw(mj_dropon_t *d, const unsigned char *rawdata, unsigned int colorspace, int width, int height, short blend) {
    if(d == NULL) return MJ_ERR_NULL_DATA;
    mj_free_dropon(d);
    if(rawdata == NULL) return MJ_ERR_NULL_DATA;

    if(blend < MJ_BLEND_NONE) blend = MJ_BLEND_NONE;
    if(blend > MJ_BLEND_FULL) blend = MJ_BLEND_FULL;

    /* how the raw pixels are laid out: colour channels, then an optional alpha byte */
    int ncolor, has_alpha, stored;
    switch(colorspace) {
        case MJ_COLORSPACE_RGB: ncolor = 3, has_alpha = 0, stored = MJ_COLORSPACE_RGB; break;
        case MJ_COLORSPACE_RGBA: ncolor = 3, has_alpha = 1, stored = MJ_COLORSPACE_RGB; break;
        case MJ_COLORSPACE_YCC: ncolor = 3, has_alpha = 0, stored = MJ_COLORSPACE_YCC; break;
        case MJ_COLORSPACE_YCCA: ncolor = 3, has_alpha = 1, stored = MJ_COLORSPACE_YCC; break;
        case MJ_COLORSPACE_GRAYSCALE: ncolor = 1, has_alpha = 0, stored = MJ_COLORSPACE_GRAYSCALE; break;
        case MJ_COLORSPACE_GRAYSCALEA: ncolor = 1, has_alpha = 1, stored = MJ_COLORSPACE_GRAYSCALE; break;
        default: return MJ_ERR_UNSUPPORTED_COLORSPACE;
    }
    if(width < 0 || height < 0) return MJ_ERR_DROPON_DIMENSIONS;

    const size_t npixel = (size_t)width * (size_t)height;
    d->image = (unsigned char *)calloc(npixel ? 3 * npixel : 1, 1);
    d->alpha = (unsigned char *)calloc(npixel ? 3 * npixel : 1, 1);
    if(d->image == NULL || d->alpha == NULL) {
        mj_free_dropon(d);
        return MJ_ERR_MEMORY;
    }
    d->width = width;
    d->height = height;
    d->colorspace = stored;
    d->blend = has_alpha ? MJ_BLEND_NONUNIFORM : blend;

    const unsigned char *in = rawdata;
    const unsigned char  flat = (unsigned char)blend;
    for(size_t i = 0; i < npixel; i++) {
        unsigned char *px = d->image + 3 * i, *al = d->alpha + 3 * i;
        if(ncolor == 3) {
            px[0] = in[0], px[1] = in[1], px[2] = in[2];
        }
        else {
            px[0] = px[1] = px[2] = in[0];
        }
        in += ncolor;
        const unsigned char a = has_alpha ? *in++ : flat;
        al[0] = al[1] = al[2] = a;
    }
    return MJ_OK;
}

/* dropon stored as a JPEG, optionally with a second grayscale JPEG as mask
 * (role of reference src/dropon.c:101-161) */
static int dropon_from_jpeg(mj_dropon_t *d, const unsigned char *memory, size_t len, const unsigned char *maskmemory, size_t masklen, short blend) {
    unsigned char *rgb = NULL, *mask = NULL;
    int            w = 0, h = 0, mw = 0, mh = 0;
    int            rv = mjp_decode_to_raw(&rgb, &w, &h, MJ_COLORSPACE_RGB, memory, len);
    if(rv != MJ_OK) return rv;
    if(maskmemory == NULL || masklen == 0) {
        rv = mj_read_dropon_from_raw(d, rgb, MJ_COLORSPACE_RGB, w, h, blend);
        free(rgb);
        return rv;
    }
    rv = mjp_decode_to_raw(&mask, &mw, &mh, MJ_COLORSPACE_GRAYSCALE, maskmemory, masklen);
    if(rv != MJ_OK) {
        free(rgb);
        return rv;
    }
    if(mw != w || mh != h) {
        free(rgb);
        free(mask);
        return MJ_ERR_DROPON_DIMENSIONS;
    }
    const size_t   n = (size_t)w * (size_t)h;
    unsigned char *rgba = (unsigned char *)malloc(n ? 4 * n : 1);
    if(rgba == NULL) {
        free(rgb);
        free(mask);
        return MJ_ERR_MEMORY;
    }
    for(size_t i = 0; i < n; i++) {
        memcpy(rgba + 4 * i, rgb + 3 * i, 3);
        rgba[4 * i + 3] = mask[i];
    }
    rv = mj_read_dropon_from_raw(d, rgba, MJ_COLORSPACE_RGBA, w, h, blend);
    free(rgb);
    free(mask);
    free(rgba);
    return rv;
}

#ifdef WITH_LIBPNG
/* PNG dropon through libpng's simplified API: always decoded to 8-bit RGBA, so the alpha channel (or
 * tRNS) becomes the mask and blend is forced to MJ_BLEND_NONUNIFORM (role of reference src/dropon.c:163-201) */
static int dropon_from_png(mj_dropon_t *d, const unsigned char *memory, size_t len) {
    png_image image;
    memset(&image, 0, sizeof(image));
    image.version = PNG_IMAGE_VERSION;
    if(png_image_begin_read_from_memory(&image, memory, len) == 0) return MJ_ERR_FILEIO;
    if(image.width >= (2u << 16) || image.height >= (2u << 16)) {
        png_image_free(&image);
        return MJ_ERR_DROPON_DIMENSIONS;
    }
    image.format = PNG_FORMAT_RGBA;
    const size_t   nbytes = PNG_IMAGE_SIZE(image);
    unsigned char *rgba = (unsigned char *)malloc(nbytes > 0 ? nbytes : 1);
    if(rgba == NULL) {
        png_image_free(&image);
        return MJ_ERR_MEMORY;
    }
    if(png_image_finish_read(&image, NULL, rgba, 0, NULL) == 0) {
        free(rgba);
        png_image_free(&image);
        return MJ_ERR_FILEIO;
    }
    int rv = mj_read_dropon_from_raw(d, rgba, MJ_COLORSPACE_RGBA, (int)image.width, (int)image.height, MJ_BLEND_NONUNIFORM);
    free(rgba);
    png_image_free(&image);
    return rv;
}
#endif

int mj_read_dropon_from_memory(mj_dropon_t *d, const unsigned char *memory, size_t len, const unsigned char *maskmemory, size_t masklen, short blend) {
    if(d == NULL || memory == NULL || len < 8) return MJ_ERR_NULL_DATA;
    if(memory[0] == 0xFF && memory[1] == 0xD8 && memory[2] == 0xFF) /* JPEG SOI + marker */
        return dropon_from_jpeg(d, memory, len, maskmemory, masklen, blend);
#ifdef WITH_LIBPNG
    if(memory[0] == 0x89 && memory[1] == 'P' && memory[2] == 'N' && memory[3] == 'G' && memory[4] == 0x0d && memory[5] == 0x0a &&
       memory[6] == 0x1a && memory[7] == 0x0a)
        return dropon_from_png(d, memory, len);
#endif
    /* anything else (and PNG when built without libpng, like the reference, src/dropon.c:80-96) */
    return MJ_ERR_UNSUPPORTED_FILETYPE;
}

int mj_read_dropon_from_file(mj_dropon_t *d, const char *filename, const char *maskfilename, short blend) {
    if(d == NULL) return MJ_ERR_NULL_DATA;
    unsigned char *data = NULL, *mask = NULL;
    size_t         len = 0, masklen = 0;
    int            rv = mjp_read_whole_file(&data, &len, filename);
    if(rv != MJ_OK) return rv;
    if(maskfilename != NULL) {
        rv = mjp_read_whole_file(&mask, &masklen, maskfilename);
        if(rv != MJ_OK) {
            free(data);
            return rv;
        }
    }
    rv = mj_read_dropon_from_memory(d, data, len, mask, masklen, blend);
    free(data);
    free(mask);
    return rv;
}
