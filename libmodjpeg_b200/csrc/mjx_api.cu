// mjx_api.cu -- the extern "C" entry points of include/mjx.h: context and memory management,
// the compiled-dropon object, and the host-pointer (staged) forms of K2 / K3.
// No compute happens on the host here: the only host work is geometry (A1), argument checks and
// gathering/scattering libjpeg rows into page-locked staging memory.
#include <string.h>

#include <new>

#include "mjx_internal.cuh"
#include "mjx_math.cuh"
#include "k2_common.cuh"

namespace mjx {

int fail(mjx_ctx *ctx, cudaError_t e, const char *what) {
    if(ctx) {
        ctx->last_error = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    }
    cudaGetLastError(); // clear the sticky-less error state
    return e == cudaErrorMemoryAllocation ? MJX_ERR_MEMORY : MJX_ERR_DEVICE;
}

static int grow(mjx_ctx *ctx, void **p, size_t *have, size_t want, bool pinned) {
    if(*have >= want) return MJX_OK;
    size_t n = want + want / 4 + 4096;
    MJX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for(int i = 0; i < mjx_ctx::kPipe; i++)
        if(ctx->pipe[i]) MJX_CUDA(ctx, cudaStreamSynchronize(ctx->pipe[i]));
    if(*p) {
        if(pinned) cudaFreeHost(*p);
        else cudaFree(*p);
        *p = nullptr;
        *have = 0;
    }
    if(pinned) MJX_CUDA(ctx, cudaHostAlloc(p, n, cudaHostAllocDefault));
    else MJX_CUDA(ctx, cudaMalloc(p, n));
    *have = n;
    return MJX_OK;
}

int ensure_pin(mjx_ctx *ctx, size_t bytes) { return grow(ctx, &ctx->pin, &ctx->pin_bytes, bytes, true); }
int ensure_dev(mjx_ctx *ctx, size_t bytes) { return grow(ctx, &ctx->dev, &ctx->dev_bytes, bytes, false); }
int ensure_desc(mjx_ctx *ctx, size_t bytes) { return grow(ctx, &ctx->desc_dev, &ctx->desc_bytes, bytes, false); }
int ensure_scratch(mjx_ctx *ctx, size_t bytes) { return grow(ctx, &ctx->scratch, &ctx->scratch_bytes, bytes, false); }
int list_chunks(int total_blocks);

static int use_device(mjx_ctx *ctx) {
    if(!ctx) return MJX_ERR_ARG;
    MJX_CUDA(ctx, cudaSetDevice(ctx->device));
    return MJX_OK;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

} // namespace mjx

using namespace mjx;

extern "C" {

// ---------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------

int mjx_device_count(void) {
    int         n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if(e != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int mjx_ctx_create(mjx_ctx **out, int device) {
    if(!out) return MJX_ERR_ARG;
    *out = nullptr;
    if(device < 0 || device >= mjx_device_count()) return MJX_ERR_DEVICE;
    mjx_ctx *ctx = new(std::nothrow) mjx_ctx();
    if(!ctx) return MJX_ERR_MEMORY;
    ctx->device = device;
    cudaError_t e = cudaSetDevice(device);
    if(e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
    if(e == cudaSuccess) e = cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    if(e == cudaSuccess) {
        int lo = 0, hi = 0; // numerically larger = lower priority
        e = cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if(e == cudaSuccess) e = cudaStreamCreateWithPriority(&ctx->side_stream, cudaStreamNonBlocking, lo);
        if(e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->side_fork, cudaEventDisableTiming);
        if(e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->side_join, cudaEventDisableTiming);
    }
    if(e == cudaSuccess) {
        // compiled dropons come from the device's stream-ordered pool; keep freed blocks cached so that the
        // per-call compile of mj_compose (the reference recompiles per call too) never reaches the driver
        cudaMemPool_t pool;
        if(cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = 1ull << 30;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    if(e != cudaSuccess) {
        cudaGetLastError();
        delete ctx;
        return MJX_ERR_DEVICE;
    }
    ctx->stream = ctx->own_stream;
    if(const char *ev = getenv("MJX_L2_FETCH")) { // experiment: L2 fill granularity hint (32 / 64 / 128 bytes), device-wide
        cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(ev));
        cudaGetLastError();
    }
    if(const char *ev = getenv("MJX_K2_OVERLAP")) ctx->overlap = atoi(ev) != 0;
    if(const char *ev = getenv("MJX_K2_TC")) ctx->k2_tc = atoi(ev) < 0 ? 0 : (atoi(ev) > 2 ? 2 : atoi(ev));
    if(const char *ev = getenv("MJX_K2_OP_MAX_MB")) ctx->k2_op_max_bytes = (size_t)(atoll(ev) > 0 ? atoll(ev) : 0) << 20;
    *out = ctx;
    return MJX_OK;
}

void mjx_ctx_destroy(mjx_ctx *ctx) {
    if(!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for(int i = 0; i < mjx_ctx::kPipe; i++) {
        if(ctx->pipe[i]) {
            cudaStreamSynchronize(ctx->pipe[i]);
            cudaStreamDestroy(ctx->pipe[i]);
        }
    }
    if(ctx->pin) cudaFreeHost(ctx->pin);
    if(ctx->pin2) cudaFreeHost(ctx->pin2);
    if(ctx->dev) cudaFree(ctx->dev);
    if(ctx->desc_dev) cudaFree(ctx->desc_dev);
    if(ctx->scratch) cudaFree(ctx->scratch);
    if(ctx->huff) cudaFree(ctx->huff);
    if(ctx->dev2) cudaFree(ctx->dev2);
    if(ctx->side_stream) {
        cudaStreamSynchronize(ctx->side_stream);
        cudaStreamDestroy(ctx->side_stream);
    }
    if(ctx->side_fork) cudaEventDestroy(ctx->side_fork);
    if(ctx->side_join) cudaEventDestroy(ctx->side_join);
    if(ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

int mjx_ctx_set_stream(mjx_ctx *ctx, void *cuda_stream) {
    if(!ctx) return MJX_ERR_ARG;
    ctx->stream = (cudaStream_t)cuda_stream;
    return MJX_OK;
}

int mjx_ctx_use_own_stream(mjx_ctx *ctx) {
    if(!ctx) return MJX_ERR_ARG;
    ctx->stream = ctx->own_stream;
    return MJX_OK;
}

void *mjx_ctx_stream(mjx_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int mjx_ctx_sync(mjx_ctx *ctx) {
    int rv = use_device(ctx);
    if(rv) return rv;
    MJX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MJX_OK;
}

const char *mjx_ctx_last_error(mjx_ctx *ctx) { return ctx ? ctx->last_error.c_str() : "no context"; }

long long mjx_ctx_kernel_launches(mjx_ctx *ctx) { return ctx ? ctx->launches : 0; }

int mjx_device_alloc(mjx_ctx *ctx, void **ptr, size_t bytes) {
    int rv = use_device(ctx);
    if(rv) return rv;
    if(!ptr) return MJX_ERR_ARG;
    MJX_CUDA(ctx, cudaMalloc(ptr, bytes));
    return MJX_OK;
}

void mjx_device_free(mjx_ctx *ctx, void *ptr) {
    if(ctx && ptr && use_device(ctx) == MJX_OK) cudaFree(ptr);
}

int mjx_host_alloc(mjx_ctx *ctx, void **ptr, size_t bytes) {
    int rv = use_device(ctx);
    if(rv) return rv;
    if(!ptr) return MJX_ERR_ARG;
    MJX_CUDA(ctx, cudaHostAlloc(ptr, bytes, cudaHostAllocPortable));
    return MJX_OK;
}

void mjx_host_free(mjx_ctx *ctx, void *ptr) {
    if(ctx && ptr && use_device(ctx) == MJX_OK) cudaFreeHost(ptr);
}

int mjx_ctx_pinned_scratch(mjx_ctx *ctx, size_t bytes, void **ptr) {
    int rv = use_device(ctx);
    if(rv) return rv;
    if(!ptr) return MJX_ERR_ARG;
    if((rv = grow(ctx, &ctx->pin2, &ctx->pin2_bytes, bytes ? bytes : 1, true)) != MJX_OK) return rv;
    *ptr = ctx->pin2;
    return MJX_OK;
}

int mjx_ctx_device_scratch(mjx_ctx *ctx, size_t bytes, void **ptr) {
    int rv = use_device(ctx);
    if(rv) return rv;
    if(!ptr) return MJX_ERR_ARG;
    if((rv = grow(ctx, &ctx->dev2, &ctx->dev2_bytes, bytes ? bytes : 1, false)) != MJX_OK) return rv;
    *ptr = ctx->dev2;
    return MJX_OK;
}

int mjx_copy_h2d(mjx_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes) {
    int rv = use_device(ctx);
    if(rv) return rv;
    MJX_CUDA(ctx, cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return MJX_OK;
}

int mjx_copy_d2h(mjx_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes) {
    int rv = use_device(ctx);
    if(rv) return rv;
    MJX_CUDA(ctx, cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return MJX_OK;
}

// ---------------------------------------------------------------------------------------
// A1: placement arithmetic (reference: src/compose.c:42-172).  Pure integer host code.
// ---------------------------------------------------------------------------------------

static void place_axis(int image_len, int dropon_len, int at_start, int at_end, int offset, int factor, int *crop_from,
                       int *crop_len, int *blockoffset, int *block) {
    // where the dropon's first pixel falls on the image along this axis
    int pos = at_start ? 0 : (at_end ? image_len - dropon_len : image_len / 2 - dropon_len / 2);
    pos += offset;
    // part of the dropon that lies left of / above the image is cut away
    int from = pos < 0 ? -pos : 0;
    int len = dropon_len - from;
    if(from > dropon_len || pos > image_len) len = 0;
    else if(pos + from + len > image_len) len = image_len - from - pos;
    *crop_from = from;
    *crop_len = len;
    // pixels between the MCU boundary and the dropon's first pixel (C remainder, clamped)
    int bo = pos % factor;
    *blockoffset = bo < 0 ? 0 : bo;
    int b = pos / factor;
    *block = b < 0 ? 0 : b;
}

void mjx_geometry(int image_width, int image_height, int h_factor, int v_factor, int dropon_width, int dropon_height,
                  unsigned int align, int offset_x, int offset_y, mjx_geometry_t *g) {
    if(!g) return;
    memset(g, 0, sizeof(*g));
    if(h_factor <= 0 || v_factor <= 0) return;
    place_axis(image_width, dropon_width, (align & 1u) != 0, (align & 2u) != 0, offset_x, h_factor, &g->crop_x,
               &g->crop_w, &g->blockoffset_x, &g->block_x);
    place_axis(image_height, dropon_height, (align & 4u) != 0, (align & 8u) != 0, offset_y, v_factor, &g->crop_y,
               &g->crop_h, &g->blockoffset_y, &g->block_y);
    g->visible = (g->crop_w != 0 && g->crop_h != 0) ? 1 : 0;
    if(!g->visible) g->blockoffset_x = g->blockoffset_y = g->block_x = g->block_y = 0;
}

// ---------------------------------------------------------------------------------------
// compiled dropon
// ---------------------------------------------------------------------------------------

static int layout_check(const mjx_layout_t *L, int *max_h, int *max_v) {
    if(!L || L->ncomp < 1 || L->ncomp > MJX_MAX_COMPONENTS) return MJX_ERR_ARG;
    int mh = 0, mv = 0, blocks = 0;
    for(int c = 0; c < L->ncomp; c++) {
        if(L->h_samp[c] < 1 || L->h_samp[c] > 4 || L->v_samp[c] < 1 || L->v_samp[c] > 4) return MJX_ERR_ARG;
        if(L->h_samp[c] > mh) mh = L->h_samp[c];
        if(L->v_samp[c] > mv) mv = L->v_samp[c];
        blocks += L->h_samp[c] * L->v_samp[c];
    }
    for(int c = 0; c < L->ncomp; c++) // jcsample.c: only integral ratios are implemented
        if(mh % L->h_samp[c] || mv % L->v_samp[c]) return MJX_ERR_UNSUPPORTED;
    if(L->ncomp > 1 && blocks > 10) return MJX_ERR_UNSUPPORTED; // jcmaster.c: C_MAX_BLOCKS_IN_MCU
    // component count implied by the colour space (jcparam.c jpeg_set_colorspace)
    if(L->colorspace == 1 && L->ncomp != 1) return MJX_ERR_ARG;
    if((L->colorspace == 2 || L->colorspace == 3) && L->ncomp != 3) return MJX_ERR_ARG;
    if(L->colorspace < 1 || L->colorspace > 3) return MJX_ERR_UNSUPPORTED;
    *max_h = mh;
    *max_v = mv;
    return MJX_OK;
}

// allocate the slab and fill the views for the given per-component block dims
static int dropon_alloc(mjx_ctx *ctx, mjx_dropon **out, const mjx_layout_t *L, const int *wb, const int *hb) {
    mjx_dropon *d = new(std::nothrow) mjx_dropon();
    if(!d) return MJX_ERR_MEMORY;
    d->device = ctx->device;
    d->layout = *L;
    size_t off = 0, offD[4], offW[4], offM[4];
    int    start = 0;
    for(int c = 0; c < L->ncomp; c++) {
        size_t nb = (size_t)wb[c] * hb[c];
        offD[c] = off;
        off = align_up(off + nb * 128, 256);
        offW[c] = off;
        off = align_up(off + nb * 128, 256);
        offM[c] = off;
        off = align_up(off + nb * 4, 256);
        d->view.comp[c].wb = wb[c];
        d->view.comp[c].hb = hb[c];
        d->view.comp[c].hs = L->h_samp[c];
        d->view.comp[c].vs = L->v_samp[c];
        d->view.comp[c].start = start;
        start += (int)nb;
    }
    d->view.ncomp = L->ncomp;
    d->view.total_blocks = start;
    d->slab_bytes = off ? off : 256;
    cudaError_t e = cudaMallocAsync(&d->slab, d->slab_bytes, ctx->stream);
    if(e != cudaSuccess) {
        delete d;
        return fail(ctx, e, "cudaMallocAsync(compiled dropon)");
    }
    for(int c = 0; c < L->ncomp; c++) {
        d->D[c] = (int16_t *)((char *)d->slab + offD[c]);
        d->W[c] = (int16_t *)((char *)d->slab + offW[c]);
        d->meta[c] = (uint32_t *)((char *)d->slab + offM[c]);
        d->view.comp[c].D = d->D[c];
        d->view.comp[c].W = d->W[c];
        d->view.comp[c].meta = d->meta[c];
    }
    *out = d;
    return MJX_OK;
}

// second half of a compile: read the class counts back, size and fill the work lists and the
// compact generic-class arrays (k1_lists.cu).  Synchronises the stream (one-time per dropon).
static int dropon_finish(mjx_ctx *ctx, mjx_dropon *d) {
    const int NC = MJX_MAX_COMPONENTS * 4;
    int       rvs = ensure_scratch(ctx, 256 + NC * sizeof(unsigned long long));
    if(rvs) return rvs;
    unsigned long long *cnt_dev = reinterpret_cast<unsigned long long *>((char *)ctx->scratch + 256); // after K2's work counter
    cudaError_t         e = cudaMemsetAsync(cnt_dev, 0, NC * sizeof(unsigned long long), ctx->stream);
    if(e == cudaSuccess) e = launch_count_classes(ctx->stream, d, cnt_dev);
    ctx->launches += d->view.ncomp;
    unsigned long long h[MJX_MAX_COMPONENTS * 4] = {};
    if(e == cudaSuccess) e = cudaMemcpyAsync(h, cnt_dev, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream);
    if(e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if(e != cudaSuccess) return fail(ctx, e, "count classes");
    for(int i = 0; i < 4; i++) d->counts[i] = 0;
    // the generic list is padded so that every component starts on a tile (32-entry) boundary
    // both lists are padded so that every component starts on a tile (32-entry) boundary
    size_t n_simple = 0, n_generic = 0, s_blocks = 0, g_blocks = 0;
    for(int c = 0; c < d->view.ncomp; c++) {
        for(int i = 0; i < 4; i++) d->counts[i] += (long long)h[4 * c + i];
        const size_t ns = (size_t)(h[4 * c + MJX_CLS_U] + h[4 * c + MJX_CLS_OPAQUE]), ng = (size_t)h[4 * c + MJX_CLS_G];
        d->simple_pad[c] = (int)(n_simple - s_blocks);
        d->view.stile_start[c] = (int)(n_simple / 32);
        s_blocks += ns;
        n_simple = align_up(n_simple + ns, 32);
        d->generic_pad[c] = (int)(n_generic - g_blocks);
        d->view.gtile_start[c] = (int)(n_generic / 32);
        g_blocks += ng;
        n_generic = align_up(n_generic + ng, 32);
    }
    const size_t nchunks = (size_t)list_chunks(d->view.total_blocks);
    size_t       off = 0;
    const size_t off_chunks = off;
    off = align_up(off + nchunks * 2 * sizeof(uint32_t), 256);
    const size_t off_ls = off;
    off = align_up(off + n_simple * sizeof(uint32_t), 256);
    const size_t off_lg = off;
    off = align_up(off + n_generic * sizeof(uint32_t), 256);
    const size_t off_ds = off;
    off = align_up(off + n_generic * 256, 256);
    const size_t off_a = off;
    off = align_up(off + n_generic * 256, 256);
    const size_t off_ad = off;
    off = align_up(off + n_generic * 256, 256);
    d->slab2_bytes = off ? off : 256;
    e = cudaMallocAsync(&d->slab2, d->slab2_bytes, ctx->stream);
    if(e != cudaSuccess) return fail(ctx, e, "cudaMallocAsync(compiled dropon lists)");
    char *base = (char *)d->slab2;
    if(n_generic) { // padding slots: entry 0xffffffff, A = Ds = 0
        e = cudaMemsetAsync(base + off_lg, 0xff, n_generic * sizeof(uint32_t), ctx->stream);
        if(e == cudaSuccess) e = cudaMemsetAsync(base + off_ds, 0, n_generic * 768, ctx->stream);
        if(e != cudaSuccess) return fail(ctx, e, "clear generic list");
    }
    if(n_simple) {
        e = cudaMemsetAsync(base + off_ls, 0xff, n_simple * sizeof(uint32_t), ctx->stream);
        if(e != cudaSuccess) return fail(ctx, e, "clear simple list");
    }
    d->view.list_simple = (const uint32_t *)(base + off_ls);
    d->view.list_generic = (const uint32_t *)(base + off_lg);
    d->view.n_simple = (int)n_simple;
    d->view.n_generic = (int)n_generic;
    d->view.gDs = (const float *)(base + off_ds);
    d->view.gA = (const float *)(base + off_a);
    d->view.gAd = (const float *)(base + off_ad);
    int launches = 0;
    e = launch_build_lists(ctx->stream, d, (uint32_t *)(base + off_chunks), &launches);
    ctx->launches += launches;
    if(e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if(e != cudaSuccess) return fail(ctx, e, "build work lists");
    return MJX_OK;
}

int mjx_dropon_compile(mjx_ctx *ctx, mjx_dropon **out, const uint8_t *image3, const uint8_t *alpha3, int width,
                       int height, int dropon_colorspace, const mjx_layout_t *layout, int blockoffset_x,
                       int blockoffset_y, int crop_x, int crop_y, int crop_w, int crop_h, int pixels_on_device) {
    if(!out) return MJX_ERR_ARG;
    *out = nullptr;
    int rv = use_device(ctx);
    if(rv) return rv;
    if(!image3 || !alpha3 || width <= 0 || height <= 0) return MJX_ERR_ARG;
    if(crop_w <= 0 || crop_h <= 0 || crop_x < 0 || crop_y < 0 || crop_x + crop_w > width || crop_y + crop_h > height ||
       blockoffset_x < 0 || blockoffset_y < 0)
        return MJX_ERR_ARG;
    int max_h, max_v;
    rv = layout_check(layout, &max_h, &max_v);
    if(rv) return rv;
    // conversions libjpeg's colour converter implements (jccolor.c jinit_color_converter);
    // anything else makes the reference return MJ_ERR_ENCODE_JPEG (SURVEY 8b "Errors")
    const int t = layout->colorspace;
    const bool ok = (dropon_colorspace == MJX_CS_RGB) ||
                    (dropon_colorspace == MJX_CS_YCC && (t == 3 || t == 1)) ||
                    (dropon_colorspace == MJX_CS_GRAYSCALE && t == 1);
    if(!ok) return MJX_ERR_UNSUPPORTED;

    // canvas padded to whole MCUs (reference: src/dropon.c:340-350)
    const int hf = max_h * 8, vf = max_v * 8;
    int canvas_w = crop_w + blockoffset_x, canvas_h = crop_h + blockoffset_y;
    if(canvas_w % hf) canvas_w += hf - canvas_w % hf;
    if(canvas_h % vf) canvas_h += vf - canvas_h % vf;
    int wb[4], hb[4];
    for(int c = 0; c < layout->ncomp; c++) {
        wb[c] = canvas_w / hf * layout->h_samp[c];
        hb[c] = canvas_h / vf * layout->v_samp[c];
    }
    mjx_dropon *d = nullptr;
    rv = dropon_alloc(ctx, &d, layout, wb, hb);
    if(rv) return rv;

    const uint8_t *img_dev = image3, *alp_dev = alpha3;
    void          *tmp = nullptr;
    const size_t   npx = (size_t)width * height * 3;
    if(!pixels_on_device) {
        cudaError_t e = cudaMallocAsync(&tmp, 2 * npx, ctx->stream);
        if(e == cudaSuccess) e = cudaMemcpyAsync(tmp, image3, npx, cudaMemcpyHostToDevice, ctx->stream);
        if(e == cudaSuccess) e = cudaMemcpyAsync((char *)tmp + npx, alpha3, npx, cudaMemcpyHostToDevice, ctx->stream);
        if(e != cudaSuccess) {
            if(tmp) cudaFreeAsync(tmp, ctx->stream);
            mjx_dropon_free(d);
            return fail(ctx, e, "upload dropon pixels");
        }
        img_dev = (const uint8_t *)tmp;
        alp_dev = (const uint8_t *)tmp + npx;
    }
    cudaError_t e = launch_k1(ctx->stream, img_dev, alp_dev, width, height, dropon_colorspace, t, blockoffset_x,
                              blockoffset_y, crop_x, crop_y, crop_w, crop_h, canvas_w, canvas_h, max_h, max_v, d);
    ctx->launches++;
    if(tmp) {
        // stream-ordered: released once the kernel that reads it has run.  The source pixels are pageable host
        // memory, which cudaMemcpyAsync has already staged by the time it returned.
        cudaError_t e2 = cudaFreeAsync(tmp, ctx->stream);
        if(e == cudaSuccess) e = e2;
    }
    if(e != cudaSuccess) {
        mjx_dropon_free(d);
        return fail(ctx, e, "k1_compile_kernel");
    }
    if((rv = dropon_finish(ctx, d)) != MJX_OK) {
        mjx_dropon_free(d);
        return rv;
    }
    *out = d;
    return MJX_OK;
}

int mjx_dropon_from_coefficients(mjx_ctx *ctx, mjx_dropon **out, const mjx_layout_t *layout, const int *wb,
                                 const int *hb, const int16_t *const *D, const int16_t *const *W) {
    if(!out) return MJX_ERR_ARG;
    *out = nullptr;
    int rv = use_device(ctx);
    if(rv) return rv;
    if(!layout || !wb || !hb || !D || !W) return MJX_ERR_ARG;
    if(layout->ncomp < 1 || layout->ncomp > MJX_MAX_COMPONENTS) return MJX_ERR_ARG;
    for(int c = 0; c < layout->ncomp; c++)
        if(wb[c] < 0 || hb[c] < 0 || !D[c] || !W[c] || layout->h_samp[c] < 1 || layout->v_samp[c] < 1) return MJX_ERR_ARG;
    mjx_dropon *d = nullptr;
    rv = dropon_alloc(ctx, &d, layout, wb, hb);
    if(rv) return rv;
    cudaError_t e = cudaSuccess;
    for(int c = 0; c < layout->ncomp && e == cudaSuccess; c++) {
        size_t bytes = (size_t)wb[c] * hb[c] * 128;
        if(!bytes) continue;
        e = cudaMemcpyAsync(d->D[c], D[c], bytes, cudaMemcpyHostToDevice, ctx->stream);
        if(e == cudaSuccess) e = cudaMemcpyAsync(d->W[c], W[c], bytes, cudaMemcpyHostToDevice, ctx->stream);
    }
    if(e == cudaSuccess) {
        e = launch_classify(ctx->stream, d);
        ctx->launches += layout->ncomp;
    }
    if(e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream); // host sources may go away after return
    if(e != cudaSuccess) {
        mjx_dropon_free(d);
        return fail(ctx, e, "mjx_dropon_from_coefficients");
    }
    if((rv = dropon_finish(ctx, d)) != MJX_OK) {
        mjx_dropon_free(d);
        return rv;
    }
    *out = d;
    return MJX_OK;
}

void mjx_dropon_free(mjx_dropon *d) {
    if(!d) return;
    cudaSetDevice(d->device);
    // stream-ordered free on the calling thread's default stream: the caller guarantees that no launch still
    // uses the dropon (mj_compose synchronises before it frees; batch hosts free after their own sync), and the
    // pool only hands the block to another stream once this free has been reached
    if(d->slab) cudaFreeAsync(d->slab, cudaStreamPerThread);
    if(d->slab2) cudaFreeAsync(d->slab2, cudaStreamPerThread);
    if(d->op_slab) cudaFree(d->op_slab);
    delete d->op;
    cudaGetLastError();
    delete d;
}

int mjx_dropon_ncomp(const mjx_dropon *d) { return d ? d->view.ncomp : 0; }

int mjx_dropon_dims(const mjx_dropon *d, int comp, int *wb, int *hb) {
    if(!d || comp < 0 || comp >= d->view.ncomp) return MJX_ERR_ARG;
    if(wb) *wb = d->view.comp[comp].wb;
    if(hb) *hb = d->view.comp[comp].hb;
    return MJX_OK;
}

long long mjx_dropon_blocks(const mjx_dropon *d) { return d ? d->view.total_blocks : 0; }

int mjx_dropon_download(mjx_ctx *ctx, const mjx_dropon *d, int comp, int16_t *D, int16_t *W, uint8_t *cls) {
    int rv = use_device(ctx);
    if(rv) return rv;
    if(!d || comp < 0 || comp >= d->view.ncomp) return MJX_ERR_ARG;
    const size_t nb = (size_t)d->view.comp[comp].wb * d->view.comp[comp].hb;
    MJX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if(D) MJX_CUDA(ctx, cudaMemcpy(D, d->D[comp], nb * 128, cudaMemcpyDeviceToHost));
    if(W) MJX_CUDA(ctx, cudaMemcpy(W, d->W[comp], nb * 128, cudaMemcpyDeviceToHost));
    if(cls) {
        std::vector<uint32_t> m(nb);
        if(nb) MJX_CUDA(ctx, cudaMemcpy(m.data(), d->meta[comp], nb * 4, cudaMemcpyDeviceToHost));
        for(size_t i = 0; i < nb; i++) cls[i] = (uint8_t)meta_cls(m[i]);
    }
    return MJX_OK;
}

int mjx_dropon_class_counts(mjx_ctx *ctx, const mjx_dropon *d, long long counts[4]) {
    (void)ctx;
    if(!d || !counts) return MJX_ERR_ARG;
    for(int i = 0; i < 4; i++) counts[i] = d->counts[i];
    return MJX_OK;
}

int mjx_dropon_generic_slots(const mjx_dropon *d) { return d ? d->view.n_generic : 0; }

int mjx_dropon_download_generic(mjx_ctx *ctx, const mjx_dropon *d, uint32_t *list, float *Ds, float *A) {
    int rv = use_device(ctx);
    if(rv) return rv;
    if(!d) return MJX_ERR_ARG;
    const size_t n = (size_t)d->view.n_generic;
    MJX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if(n == 0) return MJX_OK;
    if(list) MJX_CUDA(ctx, cudaMemcpy(list, d->view.list_generic, n * 4, cudaMemcpyDeviceToHost));
    if(Ds) MJX_CUDA(ctx, cudaMemcpy(Ds, d->view.gDs, n * 256, cudaMemcpyDeviceToHost));
    if(A) MJX_CUDA(ctx, cudaMemcpy(A, d->view.gA, n * 256, cudaMemcpyDeviceToHost));
    return MJX_OK;
}

int mjx_selftest_reciprocal(mjx_ctx *ctx, long long *mismatches) {
    int rv = use_device(ctx);
    if(rv) return rv;
    if(!mismatches) return MJX_ERR_ARG;
    if((rv = ensure_scratch(ctx, 1024)) != MJX_OK) return rv;
    unsigned long long *cnt = reinterpret_cast<unsigned long long *>((char *)ctx->scratch + 512), h = 0;
    MJX_CUDA(ctx, cudaMemsetAsync(cnt, 0, sizeof(h), ctx->stream));
    cudaError_t e = launch_selftest_reciprocal(ctx->stream, cnt);
    ctx->launches += 2;
    if(e != cudaSuccess) return fail(ctx, e, "selftest_reciprocal_kernel");
    MJX_CUDA(ctx, cudaMemcpyAsync(&h, cnt, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    MJX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *mismatches = (long long)h;
    return MJX_OK;
}

int mjx_ctx_set_class_mask(mjx_ctx *ctx, int mask) {
    if(!ctx) return MJX_ERR_ARG;
    ctx->class_mask = mask & 3;
    return MJX_OK;
}

int mjx_ctx_set_overlap(mjx_ctx *ctx, int on) {
    if(!ctx) return MJX_ERR_ARG;
    ctx->overlap = on ? 1 : 0;
    return MJX_OK;
}

int mjx_ctx_set_zero_copy(mjx_ctx *ctx, int on) {
    if(!ctx) return MJX_ERR_ARG;
    ctx->zero_copy = on ? 1 : 0;
    return MJX_OK;
}

int mjx_ctx_set_tensor_core(mjx_ctx *ctx, int mode) {
    if(!ctx || mode < 0 || mode > 2) return MJX_ERR_ARG;
    ctx->k2_tc = mode;
    return MJX_OK;
}

int mjx_ctx_set_tensor_core_min_images(mjx_ctx *ctx, int n) {
    if(!ctx || n < kOpMinImages) return MJX_ERR_ARG;
    ctx->k2_op_min_images = n;
    return MJX_OK;
}

int mjx_ctx_set_operator_pieces(mjx_ctx *ctx, int pieces) {
    if(!ctx || pieces != 2) return MJX_ERR_ARG; // the kernel's shared-memory budget holds two pieces per operator
    ctx->k2_op_pieces = pieces;
    return MJX_OK;
}

int mjx_ctx_set_strict(mjx_ctx *ctx, int strict) {
    if(!ctx) return MJX_ERR_ARG;
    ctx->strict = strict ? 1 : 0;
    return MJX_OK;
}

// ---------------------------------------------------------------------------------------
// K2 entry points
// ---------------------------------------------------------------------------------------

} // extern "C"

// the dropon's operator cache for the tensor-core G kernel, if this ctx may use it (allocated on first use)
static const OpView *dropon_op_cache(mjx_ctx *ctx, const mjx_dropon *cd, int n) {
    mjx_dropon *d = const_cast<mjx_dropon *>(cd); // the cache is the one mutable part of a compiled dropon
    if(!ctx->k2_tc || ctx->strict || n < kOpMinImages || n < ctx->k2_op_min_images || d->view.n_generic <= 0) return nullptr;
    mjx_ctx *none = nullptr;
    if(d->op_owner.compare_exchange_strong(none, ctx)) {
        const size_t bytes = op_cache_bytes(d->view.n_generic, ctx->k2_op_pieces, nullptr);
        void        *slab = nullptr;
        if(bytes <= ctx->k2_op_max_bytes && cudaMalloc(&slab, bytes) == cudaSuccess) {
            OpView *v = new(std::nothrow) OpView();
            if(v) {
                v->B = (unsigned char *)slab;
                op_cache_bytes(d->view.n_generic, ctx->k2_op_pieces, v);
                // key = 0 never equals a real table: the first launch builds the operator
                const size_t tail = (size_t)((char *)v->key - (char *)slab);
                if(cudaMemsetAsync(v->key, 0, bytes - tail, ctx->stream) == cudaSuccess) {
                    d->op_slab = slab;
                    d->op_bytes = bytes;
                    d->op = v;
                }
                else delete v;
            }
            if(!d->op) cudaFree(slab);
        }
        cudaGetLastError(); // an allocation failure only means the fp32 kernel runs
    }
    return d->op_owner.load() == ctx ? d->op : nullptr;
}

// one K2 launch with the ctx's settings on stream `st`; `scratch` holds k2_scratch_bytes(n, view) bytes
static cudaError_t run_k2(mjx_ctx *ctx, cudaStream_t st, void *scratch, const mjx_image_desc_t *items_dev, int n, const mjx_dropon *d,
                          int block_x, int block_y, bool with_side, bool planes_in_hbm = true) {
    const K2Side side = {ctx->side_stream, ctx->side_fork, ctx->side_join};
    int          launches = 0;
    K2Launch     L;
    L.stream = st;
    L.scratch = scratch;
    L.strict = ctx->strict;
    L.sm_count = ctx->sm_count;
    L.class_mask = ctx->class_mask;
    L.tc = ctx->k2_tc;
    // the tensor-core kernel: batch entry points only, and only for planes in device memory -- its gather (one block per image and
    // request, past L1) is made for HBM; over PCIe the fp32 kernel's runs of neighbouring blocks move twice as much per second
    L.op = with_side && planes_in_hbm ? dropon_op_cache(ctx, d, n) : nullptr;
    L.side = with_side && ctx->overlap ? &side : nullptr;
    L.dev = &ctx->k2dev;
    L.launches = &launches;
    const cudaError_t e = launch_k2(L, items_dev, n, d->view, block_x, block_y);
    ctx->launches += launches;
    return e;
}

extern "C" {

int mjx_compose_batch_device(mjx_ctx *ctx, const mjx_image_desc_t *items_dev, int n, const mjx_dropon *d, int block_x,
                             int block_y) {
    int rv = use_device(ctx);
    if(rv) return rv;
    if(!items_dev || !d || n < 0 || block_x < 0 || block_y < 0) return MJX_ERR_ARG;
    if(d->device != ctx->device) return MJX_ERR_ARG;
    if((rv = ensure_scratch(ctx, k2_scratch_bytes(n, d->view))) != MJX_OK) return rv;
    const cudaError_t e = run_k2(ctx, ctx->stream, ctx->scratch, items_dev, n, d, block_x, block_y, true);
    if(e != cudaSuccess) return fail(ctx, e, "k2_compose_kernel");
    return MJX_OK;
}

int mjx_compose_rows_host(mjx_ctx *ctx, int ncomp, int16_t *const *const *rows, const uint16_t *const *q,
                          const mjx_dropon *d) {
    int rv = use_device(ctx);
    if(rv) return rv;
    if(!rows || !q || !d || ncomp != d->view.ncomp || d->device != ctx->device) return MJX_ERR_ARG;
    // staging layout: [desc][comp 0 ROI][comp 1 ROI]...
    size_t off[MJX_MAX_COMPONENTS], total = align_up(sizeof(mjx_image_desc_t), 256);
    for(int c = 0; c < ncomp; c++) {
        if(!rows[c] || !q[c]) return MJX_ERR_ARG;
        off[c] = total;
        total = align_up(total + (size_t)d->view.comp[c].wb * d->view.comp[c].hb * 128, 256);
    }
    if((rv = ensure_pin(ctx, total)) || (rv = ensure_dev(ctx, total))) return rv;
    char *pin = (char *)ctx->pin, *dev = (char *)ctx->dev;

    mjx_image_desc_t *desc = (mjx_image_desc_t *)pin;
    memset(desc, 0, sizeof(*desc));
    for(int c = 0; c < ncomp; c++) {
        const int wb = d->view.comp[c].wb, hb = d->view.comp[c].hb;
        desc->plane[c] = (uint64_t)(uintptr_t)(dev + off[c]);
        desc->stride_blocks[c] = wb;
        desc->rows[c] = hb;
        desc->wreal[c] = wb;
        desc->hreal[c] = hb;
        memcpy(desc->q[c], q[c], 128);
        for(int l = 0; l < hb; l++) {
            if(!rows[c][l]) return MJX_ERR_ARG;
            memcpy(pin + off[c] + (size_t)l * wb * 128, rows[c][l], (size_t)wb * 128);
        }
    }
    MJX_CUDA(ctx, cudaMemcpyAsync(dev, pin, total, cudaMemcpyHostToDevice, ctx->stream));
    // the staged region starts at the dropon's origin: MCU position (0, 0)
    if((rv = ensure_scratch(ctx, k2_scratch_bytes(1, d->view))) != MJX_OK) return rv;
    const cudaError_t e = run_k2(ctx, ctx->stream, ctx->scratch, (const mjx_image_desc_t *)dev, 1, d, 0, 0, false);
    if(e != cudaSuccess) return fail(ctx, e, "k2_compose_kernel");
    const size_t head = align_up(sizeof(mjx_image_desc_t), 256);
    MJX_CUDA(ctx, cudaMemcpyAsync(pin + head, dev + head, total - head, cudaMemcpyDeviceToHost, ctx->stream));
    MJX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for(int c = 0; c < ncomp; c++) {
        const int wb = d->view.comp[c].wb, hb = d->view.comp[c].hb;
        for(int l = 0; l < hb; l++) memcpy(rows[c][l], pin + off[c] + (size_t)l * wb * 128, (size_t)wb * 128);
    }
    return MJX_OK;
}

static int pipe_init(mjx_ctx *ctx) {
    for(int i = 0; i < mjx_ctx::kPipe; i++) {
        if(!ctx->pipe[i]) MJX_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->pipe[i], cudaStreamNonBlocking));
    }
    return MJX_OK;
}

int mjx_compose_batch_host(mjx_ctx *ctx, const mjx_host_image_t *items, int n, const mjx_dropon *d, int block_x,
                           int block_y) {
    int rv = use_device(ctx);
    if(rv) return rv;
    if(!items || !d || n < 0 || block_x < 0 || block_y < 0 || d->device != ctx->device) return MJX_ERR_ARG;
    if(n == 0) return MJX_OK;
    const int ncomp = d->view.ncomp;
    // validate that the dropon's region lies inside every plane
    for(int i = 0; i < n; i++)
        for(int c = 0; c < ncomp; c++) {
            const DropComp &dc = d->view.comp[c];
            if(!items[i].plane[c] || !items[i].q[c]) return MJX_ERR_ARG;
            if(block_x * dc.hs + dc.wb > items[i].stride_blocks[c] || block_y * dc.vs + dc.hb > items[i].rows[c])
                return MJX_ERR_ARG;
        }
    // Zero-copy path: when every plane lies in page-locked host memory the GPU can address (mjx_host_alloc,
    // cudaHostAlloc, cudaHostRegister), K2 runs directly on the host planes in ONE launch over the whole batch:
    // only the blocks the dropon touches cross PCIe (class G blocks are read and written, OPAQUE blocks are
    // written only, transparent blocks and everything outside the dropon never move).
    if(ctx->zero_copy) {
        bool mapped = true;
        for(int i = 0; i < n && mapped; i++)
            for(int c = 0; c < ncomp && mapped; c++) {
                cudaPointerAttributes at;
                if(cudaPointerGetAttributes(&at, items[i].plane[c]) != cudaSuccess) {
                    cudaGetLastError();
                    mapped = false;
                }
                else mapped = (at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged) && at.devicePointer != nullptr;
            }
        if(mapped) {
            const size_t dbytes = sizeof(mjx_image_desc_t) * (size_t)n;
            if((rv = ensure_pin(ctx, dbytes)) || (rv = ensure_desc(ctx, dbytes)) ||
               (rv = ensure_scratch(ctx, k2_scratch_bytes(n, d->view))))
                return rv;
            mjx_image_desc_t *desc = (mjx_image_desc_t *)ctx->pin;
            memset(desc, 0, dbytes);
            for(int i = 0; i < n; i++)
                for(int c = 0; c < ncomp; c++) {
                    cudaPointerAttributes at;
                    MJX_CUDA(ctx, cudaPointerGetAttributes(&at, items[i].plane[c]));
                    desc[i].plane[c] = (uint64_t)(uintptr_t)at.devicePointer;
                    desc[i].stride_blocks[c] = items[i].stride_blocks[c];
                    desc[i].rows[c] = items[i].rows[c];
                    desc[i].wreal[c] = items[i].wreal[c];
                    desc[i].hreal[c] = items[i].hreal[c];
                    memcpy(desc[i].q[c], items[i].q[c], 128);
                }
            MJX_CUDA(ctx, cudaMemcpyAsync(ctx->desc_dev, desc, dbytes, cudaMemcpyHostToDevice, ctx->stream));
            const cudaError_t e = run_k2(ctx, ctx->stream, ctx->scratch, (const mjx_image_desc_t *)ctx->desc_dev, n, d, block_x, block_y, true, false);
            if(e != cudaSuccess) return fail(ctx, e, "k2_compose_kernel");
            MJX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            return MJX_OK;
        }
    }
    // staged path -- per pipeline slot: one image's region in device memory; descriptors for all images in pinned memory
    size_t off[MJX_MAX_COMPONENTS], slot = 0;
    for(int c = 0; c < ncomp; c++) {
        off[c] = slot;
        slot = align_up(slot + (size_t)d->view.comp[c].wb * d->view.comp[c].hb * 128, 256);
    }
    const int    P = mjx_ctx::kPipe;
    const size_t desc_sz = align_up(sizeof(mjx_image_desc_t), 256);
    const size_t scr_sz = align_up(k2_scratch_bytes(1, d->view), 256);
    if((rv = pipe_init(ctx)) || (rv = ensure_dev(ctx, slot * P)) || (rv = ensure_desc(ctx, desc_sz * P)) || (rv = ensure_scratch(ctx, scr_sz * P)) ||
       (rv = ensure_pin(ctx, desc_sz * (size_t)n)))
        return rv;
    char *dev = (char *)ctx->dev, *pin = (char *)ctx->pin, *ddev = (char *)ctx->desc_dev;

    // the pipeline streams start after whatever is queued on the ctx stream
    cudaEvent_t start;
    MJX_CUDA(ctx, cudaEventCreateWithFlags(&start, cudaEventDisableTiming));
    MJX_CUDA(ctx, cudaEventRecord(start, ctx->stream));
    for(int s = 0; s < P; s++) MJX_CUDA(ctx, cudaStreamWaitEvent(ctx->pipe[s], start, 0));
    cudaEventDestroy(start);

    // an error in the middle of the pipeline: copies into the caller's planes are in flight on the other streams -- they must
    // not outlive the call
#define PIPE_CUDA(call)                                                        \
    do {                                                                       \
        cudaError_t e__ = (call);                                              \
        if(e__ != cudaSuccess) {                                               \
            for(int k__ = 0; k__ < P; k__++) cudaStreamSynchronize(ctx->pipe[k__]); \
            return mjx::fail(ctx, e__, #call);                                 \
        }                                                                      \
    } while(0)
    for(int i = 0; i < n; i++) {
        const int         s = i % P;
        cudaStream_t      st = ctx->pipe[s];
        char             *base = dev + slot * s;
        mjx_image_desc_t *desc = (mjx_image_desc_t *)(pin + desc_sz * i);
        memset(desc, 0, sizeof(*desc));
        for(int c = 0; c < ncomp; c++) {
            const DropComp &dc = d->view.comp[c];
            desc->plane[c] = (uint64_t)(uintptr_t)(base + off[c]);
            desc->stride_blocks[c] = dc.wb;
            desc->rows[c] = dc.hb;
            desc->wreal[c] = dc.wb;
            desc->hreal[c] = dc.hb;
            memcpy(desc->q[c], items[i].q[c], 128);
        }
        PIPE_CUDA(cudaMemcpyAsync(ddev + desc_sz * s, desc, sizeof(*desc), cudaMemcpyHostToDevice, st));
        for(int c = 0; c < ncomp; c++) {
            const DropComp &dc = d->view.comp[c];
            const size_t    sp = (size_t)items[i].stride_blocks[c] * 128, wbytes = (size_t)dc.wb * 128;
            const char     *src = (const char *)items[i].plane[c] + ((size_t)block_y * dc.vs * items[i].stride_blocks[c] + (size_t)block_x * dc.hs) * 128;
            if(sp == wbytes) PIPE_CUDA(cudaMemcpyAsync(base + off[c], src, wbytes * dc.hb, cudaMemcpyHostToDevice, st));
            else PIPE_CUDA(cudaMemcpy2DAsync(base + off[c], wbytes, src, sp, wbytes, dc.hb, cudaMemcpyHostToDevice, st));
        }
        const cudaError_t e = run_k2(ctx, st, (char *)ctx->scratch + scr_sz * s, (const mjx_image_desc_t *)(ddev + desc_sz * s), 1, d, 0, 0, false);
        if(e != cudaSuccess) {
            for(int k = 0; k < P; k++) cudaStreamSynchronize(ctx->pipe[k]); // copies into caller memory must not outlive the call
            return fail(ctx, e, "k2_compose_kernel");
        }
        for(int c = 0; c < ncomp; c++) {
            const DropComp &dc = d->view.comp[c];
            const size_t    sp = (size_t)items[i].stride_blocks[c] * 128, wbytes = (size_t)dc.wb * 128;
            char           *dst = (char *)items[i].plane[c] + ((size_t)block_y * dc.vs * items[i].stride_blocks[c] + (size_t)block_x * dc.hs) * 128;
            if(sp == wbytes) PIPE_CUDA(cudaMemcpyAsync(dst, base + off[c], wbytes * dc.hb, cudaMemcpyDeviceToHost, st));
            else PIPE_CUDA(cudaMemcpy2DAsync(dst, sp, base + off[c], wbytes, wbytes, dc.hb, cudaMemcpyDeviceToHost, st));
        }
    }
#undef PIPE_CUDA
    for(int s = 0; s < P; s++) MJX_CUDA(ctx, cudaStreamSynchronize(ctx->pipe[s]));
    return MJX_OK;
}

// ---------------------------------------------------------------------------------------
// K3 entry points
// ---------------------------------------------------------------------------------------

static int ops_check(int ncomp, const mjx_effect_op_t *ops, int nops) {
    if(nops < 0 || (nops > 0 && !ops)) return MJX_ERR_ARG;
    for(int i = 0; i < nops; i++) {
        if(ops[i].comp < 0 || ops[i].comp >= ncomp) return MJX_ERR_ARG;
        if(ops[i].op != MJX_FX_ZERO && ops[i].op != MJX_FX_PIXELATE && ops[i].op != MJX_FX_ADD_DC) return MJX_ERR_ARG;
    }
    return MJX_OK;
}

int mjx_effects_batch_device(mjx_ctx *ctx, const mjx_image_desc_t *items_dev, int n, int ncomp,
                             const mjx_effect_op_t *ops, int nops) {
    int rv = use_device(ctx);
    if(rv) return rv;
    if(!items_dev || n < 0 || ncomp < 1 || ncomp > MJX_MAX_COMPONENTS) return MJX_ERR_ARG;
    if((rv = ops_check(ncomp, ops, nops))) return rv;
    int         launches = 0;
    cudaError_t e = launch_k3(ctx->stream, items_dev, n, ncomp, ops, nops, &launches);
    ctx->launches += launches;
    if(e != cudaSuccess) return fail(ctx, e, "k3 effects kernel");
    return MJX_OK;
}

int mjx_effects_rows_host(mjx_ctx *ctx, int ncomp, int16_t *const *const *rows, const int *wreal, const int *hreal,
                          const uint16_t *const *q, const mjx_effect_op_t *ops, int nops) {
    int rv = use_device(ctx);
    if(rv) return rv;
    if(!rows || !wreal || !hreal || !q || ncomp < 1 || ncomp > MJX_MAX_COMPONENTS) return MJX_ERR_ARG;
    if((rv = ops_check(ncomp, ops, nops))) return rv;
    bool   used[MJX_MAX_COMPONENTS] = {}, need_read[MJX_MAX_COMPONENTS] = {};
    for(int i = 0; i < nops; i++) {
        const int c = ops[i].comp;
        if(!used[c]) need_read[c] = ops[i].op != MJX_FX_ZERO; // a leading ZERO makes the old content irrelevant
        used[c] = true;
    }
    bool dc_only = nops > 0;
    for(int i = 0; i < nops; i++) dc_only = dc_only && ops[i].op == MJX_FX_ADD_DC;
    if(dc_only) {
        // tint / luminance: only the DC of every block is read and written, so only the DCs are staged -- the host
        // gathers them (one strided pass over the rows, like the reference's own loop), 2 bytes per block cross PCIe
        size_t doff[MJX_MAX_COMPONENTS], dtotal = 0;
        for(int c = 0; c < ncomp; c++) {
            if(!used[c]) continue;
            if(!rows[c] || !q[c] || wreal[c] < 0 || hreal[c] < 0) return MJX_ERR_ARG;
            doff[c] = dtotal;
            dtotal = align_up(dtotal + (size_t)wreal[c] * hreal[c] * sizeof(int16_t), 256);
        }
        if(dtotal == 0) return MJX_OK;
        if((rv = ensure_pin(ctx, dtotal)) || (rv = ensure_dev(ctx, dtotal))) return rv;
        char *pin = (char *)ctx->pin, *dev = (char *)ctx->dev;
        for(int c = 0; c < ncomp; c++) {
            if(!used[c]) continue;
            int16_t *dst = (int16_t *)(pin + doff[c]);
            for(int l = 0; l < hreal[c]; l++) {
                if(!rows[c][l]) return MJX_ERR_ARG;
                const int16_t *row = rows[c][l];
                for(int k = 0; k < wreal[c]; k++) dst[(size_t)l * wreal[c] + k] = row[(size_t)k * 64];
            }
        }
        MJX_CUDA(ctx, cudaMemcpyAsync(dev, pin, dtotal, cudaMemcpyHostToDevice, ctx->stream));
        for(int c = 0; c < ncomp; c++) {
            if(!used[c]) continue;
            cudaError_t e = launch_k3_dc_compact(ctx->stream, (int16_t *)(dev + doff[c]), wreal[c] * hreal[c], (int)q[c][0], c, ops, nops);
            ctx->launches++;
            if(e != cudaSuccess) return fail(ctx, e, "k3 effects kernel");
        }
        MJX_CUDA(ctx, cudaMemcpyAsync(pin, dev, dtotal, cudaMemcpyDeviceToHost, ctx->stream));
        MJX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for(int c = 0; c < ncomp; c++) {
            if(!used[c]) continue;
            const int16_t *src = (const int16_t *)(pin + doff[c]);
            for(int l = 0; l < hreal[c]; l++) {
                int16_t *row = rows[c][l];
                for(int k = 0; k < wreal[c]; k++) row[(size_t)k * 64] = src[(size_t)l * wreal[c] + k];
            }
        }
        return MJX_OK;
    }
    // Rewrite pipelines (a ZERO or PIXELATE step): every AC coefficient ends as 0, so the only input a component can
    // still need is its DCs (PIXELATE first) -- those are gathered compactly (2 bytes per block up); the whole
    // planes come back (128 bytes per block down).  A component's list that contains no such step but ADD_DCs only
    // cannot occur here together with a rewrite of another component in the reference's API; it is handled by
    // staging its rows fully.
    bool   rewrites[MJX_MAX_COMPONENTS] = {};
    for(int i = 0; i < nops; i++)
        if(ops[i].op == MJX_FX_ZERO || ops[i].op == MJX_FX_PIXELATE) rewrites[ops[i].comp] = true;
    size_t off[MJX_MAX_COMPONENTS], dcoff[MJX_MAX_COMPONENTS], total = align_up(sizeof(mjx_image_desc_t), 256);
    for(int c = 0; c < ncomp; c++) {
        if(!used[c]) continue;
        if(!rows[c] || !q[c] || wreal[c] < 0 || hreal[c] < 0) return MJX_ERR_ARG;
        off[c] = total;
        total = align_up(total + (size_t)wreal[c] * hreal[c] * 128, 256);
        dcoff[c] = total;
        if(rewrites[c] && need_read[c]) total = align_up(total + (size_t)wreal[c] * hreal[c] * sizeof(int16_t), 256);
    }
    if((rv = ensure_pin(ctx, total)) || (rv = ensure_dev(ctx, total))) return rv;
    char             *pin = (char *)ctx->pin, *dev = (char *)ctx->dev;
    mjx_image_desc_t *desc = (mjx_image_desc_t *)pin;
    memset(desc, 0, sizeof(*desc));
    const size_t   head = align_up(sizeof(mjx_image_desc_t), 256);
    const int16_t *dc_dev[MJX_MAX_COMPONENTS] = {};
    for(int c = 0; c < ncomp; c++) {
        if(!used[c]) continue;
        desc->plane[c] = (uint64_t)(uintptr_t)(dev + off[c]);
        desc->stride_blocks[c] = wreal[c];
        desc->rows[c] = hreal[c];
        desc->wreal[c] = wreal[c];
        desc->hreal[c] = hreal[c];
        memcpy(desc->q[c], q[c], 128);
        if(!need_read[c]) continue;
        for(int l = 0; l < hreal[c]; l++)
            if(!rows[c][l]) return MJX_ERR_ARG;
        if(rewrites[c]) { // DCs only
            int16_t *dst = (int16_t *)(pin + dcoff[c]);
            for(int l = 0; l < hreal[c]; l++)
                for(int k = 0; k < wreal[c]; k++) dst[(size_t)l * wreal[c] + k] = rows[c][l][(size_t)k * 64];
            dc_dev[c] = (const int16_t *)(dev + dcoff[c]);
        }
        else
            for(int l = 0; l < hreal[c]; l++) memcpy(pin + off[c] + (size_t)l * wreal[c] * 128, rows[c][l], (size_t)wreal[c] * 128);
    }
    MJX_CUDA(ctx, cudaMemcpyAsync(dev, pin, head, cudaMemcpyHostToDevice, ctx->stream));
    for(int c = 0; c < ncomp; c++) {
        if(!used[c] || !need_read[c]) continue;
        if(rewrites[c])
            MJX_CUDA(ctx, cudaMemcpyAsync(dev + dcoff[c], pin + dcoff[c], (size_t)wreal[c] * hreal[c] * sizeof(int16_t), cudaMemcpyHostToDevice, ctx->stream));
        else
            MJX_CUDA(ctx, cudaMemcpyAsync(dev + off[c], pin + off[c], (size_t)wreal[c] * hreal[c] * 128, cudaMemcpyHostToDevice, ctx->stream));
    }
    int         launches = 0;
    cudaError_t e = launch_k3(ctx->stream, (const mjx_image_desc_t *)dev, 1, ncomp, ops, nops, &launches, dc_dev);
    ctx->launches += launches;
    if(e != cudaSuccess) return fail(ctx, e, "k3 effects kernel");
    for(int c = 0; c < ncomp; c++)
        if(used[c])
            MJX_CUDA(ctx, cudaMemcpyAsync(pin + off[c], dev + off[c], (size_t)wreal[c] * hreal[c] * 128, cudaMemcpyDeviceToHost, ctx->stream));
    MJX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for(int c = 0; c < ncomp; c++)
        if(used[c])
            for(int l = 0; l < hreal[c]; l++) memcpy(rows[c][l], pin + off[c] + (size_t)l * wreal[c] * 128, (size_t)wreal[c] * 128);
    return MJX_OK;
}

} // extern "C"
