/*
 * mj_device.c -- the calling thread's engine context.  The reference has no global state and
 * is re-entrant per object (SURVEY 8b "Threading"); to keep that, every host thread gets its
 * own mjx_ctx (stream + staging pools), created on first use and destroyed with the thread.
 * There is no CPU fallback: without a CUDA device the compute entry points fail loudly.
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>

#include "mj_private.h"

static pthread_key_t  g_key;
static __thread int   t_device = -1; /* mjx_host_set_device: this thread's device, -1 = $MJX_DEVICE or 0 */
static pthread_once_t g_once = PTHREAD_ONCE_INIT;

static void ctx_destructor(void *p) {
    mjp_compose_cache_clear(); /* compiled dropons cached by this thread's mj_compose calls */
    mjx_ctx_destroy((mjx_ctx *)p);
}
static void make_key(void) { pthread_key_create(&g_key, ctx_destructor); }

mjx_ctx *mjx_host_ctx(void) {
    pthread_once(&g_once, make_key);
    mjx_ctx *ctx = (mjx_ctx *)pthread_getspecific(g_key);
    if(ctx != NULL) return ctx;
    int         device = 0;
    const char *env = getenv("MJX_DEVICE");
    if(t_device >= 0) device = t_device;
    else if(env != NULL && *env) device = atoi(env);
    int rv = mjx_ctx_create(&ctx, device);
    if(rv != MJX_OK || ctx == NULL) {
        fprintf(stderr, "libmodjpeg (B200): no usable CUDA device %d (%d devices visible) - mj_compose/mj_effect_* need the GPU engine, there is no CPU fallback\n",
                device, mjx_device_count());
        return NULL;
    }
    pthread_setspecific(g_key, ctx);
    return ctx;
}

/* A second context of the calling thread on the same device (own stream, own staging pools): mj_compose_batch keeps two windows
 * in flight, one per context.  NULL when it cannot be had -- the caller then works with one. */
static pthread_key_t  g_key2;
static pthread_once_t g_once2 = PTHREAD_ONCE_INIT;
static void           ctx2_destructor(void *p) { mjx_ctx_destroy((mjx_ctx *)p); }
static void           make_key2(void) { pthread_key_create(&g_key2, ctx2_destructor); }

mjx_ctx *mjp_host_ctx2(void) {
    if(mjx_host_ctx() == NULL) return NULL;
    pthread_once(&g_once2, make_key2);
    mjx_ctx *ctx = (mjx_ctx *)pthread_getspecific(g_key2);
    if(ctx != NULL) return ctx;
    int         device = 0;
    const char *env = getenv("MJX_DEVICE");
    if(t_device >= 0) device = t_device;
    else if(env != NULL && *env) device = atoi(env);
    if(mjx_ctx_create(&ctx, device) != MJX_OK || ctx == NULL) return NULL;
    pthread_setspecific(g_key2, ctx);
    return ctx;
}

/* The device the calling thread's context is created on (before its first compute call; a thread that already has a
 * context keeps it).  mj_compose_batch uses it to give each device its own group of host threads. */
void mjx_host_set_device(int device) { t_device = device; }
