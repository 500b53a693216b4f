// k2_compose.cu -- K2, the masked blend: dequantise, blend with the compiled dropon, requantise
// to the image's own tables.  Replaces mj_compose_with_mask + mj_convolve
// (reference: src/compose.c:237-342, src/convolve.c:29-1099).
//
// The class K1 gave every dropon block decides what is touched:
//   T       nothing: never listed, never loaded, never stored
//   OPAQUE  store trunc(D / q); the image block is not read                       (bit-exact)
//   U       one fp32 multiply per coefficient on the dequantised difference        (bit-exact)
//   G       pixel-domain blend  Y = DCT2(A * IDCT2(D - I*q)),  A = IDCT2(W)/255 precomputed by K1:
//           the closed form of the reference's 64 sparse DCT-domain products, within +-1
//           quantisation step of it.
//
// Two kernels per call (one launch each, any number of images; side by side for batches of >= 64
// images, see launch_k2):
//   k2_simple_kernel   OPAQUE/U list; 8 lanes per block (lane r = row r, one 128-bit access per
//                      plane), D row kept in registers while the lanes walk 16 images whose quant
//                      tables were converted once per CTA into shared memory; no global load in the
//                      image loop, opaque rows reused across images with identical tables; bound by
//                      its writes (6 TB/s write-only)
//   k2_generic_kernel  G list; one thread owns one block (all 64 coefficients in registers as 32
//                      fp32 pairs, so the 2-D transforms need no shuffles), arithmetic on packed
//                      fp32 (FADD2 / FMUL2 / FFMA2).  A CTA of 4 warps shares a tile of 32 list
//                      entries (A and Ds in shared memory) and streams the images through a
//                      cp.async double buffer per warp; shared-memory blocks are padded so the
//                      thread-per-block 128-bit accesses are conflict-free at immediate offsets;
//                      results go back through the same buffer so global stores are coalesced.
//                      Work (tile x image chunk) is claimed from an atomic counter by persistent
//                      CTAs, 3 per SM.  Details at the kernel.
//   The arithmetic stays in the fp32 pipe: on sm_100a F2I/FRND issue at 1/8 rate and I2F.S16/SHFL
//   at 1/4 (profiles/microbench/ubench.txt), so truncation and int16 packing use magic-number adds.
//
// k2_strict_kernel is the first-generation single kernel (8 lanes per block, warp-shuffle
// transposes, integer requantisation).  It reproduces the reference's int16 wrap-around on
// out-of-range products exactly and is selected with mjx_ctx_set_strict(); the fast kernels equal
// it whenever |I*q| and the blended value stay inside int16, i.e. for every JPEG a conforming
// encoder writes.
//
// Roofline: HBM.  Algorithmic bytes per block: T 0, OPAQUE 128 (write), U/G 256 (read + write),
// plus the compiled dropon once per launch (L2 / shared-memory resident across the images).
#include <stdlib.h>

#include "k2_common.cuh"

namespace mjx {

// =========================================================================================
// strict kernel (all classes, exact int16 wrap-around)
// =========================================================================================

struct StrictParams {
    DropView                drop;
    const mjx_image_desc_t *items;
    int                     block_x, block_y; // dropon origin on the image, in MCUs
};

static constexpr int kThreads = 256;
static constexpr int kBlocksPerCta = kThreads / 8;

__global__ void __launch_bounds__(kThreads) k2_strict_kernel(const StrictParams p) {
    const int r = threadIdx.x & 7;
    const int b = blockIdx.x * kBlocksPerCta + (threadIdx.x >> 3);
    if(b >= p.drop.total_blocks) return;

    int c = 0;
#pragma unroll
    for(int i = 1; i < MJX_MAX_COMPONENTS; i++)
        if(i < p.drop.ncomp && b >= p.drop.comp[i].start) c = i;
    const DropComp &dc = p.drop.comp[c];
    const int       bi = b - dc.start;

    const uint32_t meta = __ldg(dc.meta + bi);
    const uint32_t cls = meta_cls(meta);

    const mjx_image_desc_t &im = p.items[blockIdx.y];
    const int l = bi / dc.wb, k = bi - l * dc.wb;
    const int row = p.block_y * dc.vs + l, col = p.block_x * dc.hs + k;
    if(row >= im.rows[c] || col >= im.stride_blocks[c]) return;

    int16_t       *ip = reinterpret_cast<int16_t *>(im.plane[c]) + ((size_t)row * im.stride_blocks[c] + col) * 64 + r * 8;
    const int16_t *dp = dc.D + (size_t)bi * 64 + r * 8;

    int   q[8], D[8], out[8];
    float rq[8];
    {
        Row8 qr = ld_row_keep(&im.q[c][r * 8]);
#pragma unroll
        for(int i = 0; i < 8; i++) {
            q[i] = (int)((qr.w[i >> 1] >> ((i & 1) * 16)) & 0xffffu);
            rq[i] = quant_rcp(q[i]);
        }
    }
    row_unpack(ld_row_keep(dp), D);

    if(cls == CLS_OPAQUE) {
#pragma unroll
        for(int i = 0; i < 8; i++) out[i] = tdiv(D[i], rq[i]);
        st_row_stream(ip, row_pack(out));
        return;
    }

    int I[8];
    row_unpack(ld_row_stream(ip), I);

    if(cls == CLS_T) {
        // all-zero alpha: the reference still dequantises and requantises the block in place (src/compose.c:277-286, 327-336), which
        // is the identity unless I*q leaves int16 -- the case this kernel exists for
        bool changed = false;
#pragma unroll
        for(int i = 0; i < 8; i++) {
            out[i] = wrap16(tdiv(wrap16(I[i] * q[i]), rq[i]));
            changed = changed || out[i] != I[i];
        }
        if(changed) st_row_stream(ip, row_pack(out));
        return;
    }

    if(cls == CLS_U) {
        const float w4 = uniform_w4(meta_wdc(meta));
#pragma unroll
        for(int i = 0; i < 8; i++) out[i] = blend_uniform(I[i], D[i], q[i], rq[i], w4);
        st_row_stream(ip, row_pack(out));
        return;
    }

    const unsigned mask = group_mask();
    int            W[8], deq[8];
    row_unpack(ld_row_keep(dc.W + (size_t)bi * 64 + r * 8), W);

    float       x[8], a[8];
    const float pr = c_inv_scale[r];
    {
        const float isc[8] = MJX_INV_SCALE_INIT;
#pragma unroll
        for(int i = 0; i < 8; i++) {
            deq[i] = wrap16(I[i] * q[i]);
            const float s = pr * isc[i];
            x[i] = (float)(D[i] - deq[i]) * s;
            a[i] = (float)W[i] * (s * (1.0f / 255.0f));
        }
    }
    idct8(x);
    idct8(a);
    transpose8(x, r, mask);
    transpose8(a, r, mask);
    idct8(x);
    idct8(a);
#pragma unroll
    for(int i = 0; i < 8; i++) x[i] *= a[i];
    fdct8(x);
    transpose8(x, r, mask);
    fdct8(x);
    {
        const float fr = c_fwd_scale[r];
        const float fsc[8] = MJX_FWD_SCALE_INIT;
#pragma unroll
        for(int i = 0; i < 8; i++) {
            const float Y = x[i] * (fr * fsc[i]);
            out[i] = tdiv(wrap16(deq[i] + f2i_trunc(Y)), rq[i]);
        }
    }
    st_row_stream(ip, row_pack(out));
}

// =========================================================================================
// fast path, kernel 1: OPAQUE and U blocks
// =========================================================================================
//
// These classes are pure streaming (OPAQUE: 128 B written per block, the image is never read; U: 128 B
// read + written, one multiply per coefficient), so the kernel is built for memory parallelism, not
// arithmetic: 8 lanes per block (lane r = row r, one 128-bit access per plane), a CTA owns one tile of
// 32 list entries (one component, the list is padded per component) and walks a chunk of kSimpleImages
// images with the lane's D row in registers.  The chunk's quantisation tables are converted once per CTA
// into shared memory (q and the biased reciprocal as floats), so the per-image path is
// 2-4 conflict-free LDS.128 + 4 packed-fp32 pair operations + one coalesced 128-bit store -- and for an
// opaque block whose image has the same table and geometry as the one before, just the store.
// Arithmetic: tdiv_pair / uniform_pair (mjx_math.cuh), bit-exact with the reference.

static constexpr int kSimpleImages = 16; // images walked by one CTA of the simple kernel

// THREADS = 256: one CTA per tile.  THREADS = 128: two CTAs per tile, 64 registers x 128 threads -- small enough to sit
// beside the three resident CTAs of the G kernel on every SM (10 240 registers are left there), which is how the two
// kernels run concurrently (launch_k2).
template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS) k2_simple_kernel(const FastParams p) {
    __shared__ __align__(16) float s_q[kSimpleImages][64], s_rq[kSimpleImages][64];
    constexpr int kPerTile = kThreads / THREADS; // CTAs per tile
    const int     r = threadIdx.x & 7, t = (threadIdx.x >> 3) + (blockIdx.x % kPerTile) * (THREADS / 8);
    const int     tile = blockIdx.x / kPerTile;
    int       c = 0; // tiles hold one component only
#pragma unroll
    for(int i = 1; i < MJX_MAX_COMPONENTS; i++)
        if(i < p.drop.ncomp && tile >= p.drop.stile_start[i]) c = i;
    const int i0 = blockIdx.y * kSimpleImages, ni = min(p.n - i0, kSimpleImages);

    // float tables of the chunk's images for this component: one 128-bit load = 8 entries per thread (THREADS >= 8 per
    // image), so that the CTA's start-up is one memory latency (it is exposed when a single small CTA runs beside the G
    // kernel).  The same threads note which images carry the SAME table as their predecessor in the chunk -- batches
    // from one encoder setting do throughout -- because trunc(D / q) of an opaque block is then the same row again.
    __shared__ unsigned s_same[4]; // per table-converting warp: bit g = image g's table equals image g - 1's
    {
        const int k = threadIdx.x;
        bool      eq = false;
        if(k < ni * 8) {
            const uint4 w = __ldg(reinterpret_cast<const uint4 *>(&p.items[i0 + (k >> 3)].q[c][(k & 7) * 8]));
            if(k >= 8) {
                const uint4 wp = __ldg(reinterpret_cast<const uint4 *>(&p.items[i0 + (k >> 3) - 1].q[c][(k & 7) * 8]));
                eq = w.x == wp.x && w.y == wp.y && w.z == wp.z && w.w == wp.w;
            }
            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
            // row r of a table lives as two float4: entries 0..3 at [r * 4], entries 4..7 at [32 + r * 4], so that the 8
            // lanes of a block read 128 contiguous bytes per access (rows side by side at stride 32 B were a 2-way conflict)
            float *dq = &s_q[k >> 3][(k & 7) * 4], *drq = &s_rq[k >> 3][(k & 7) * 4];
            float  q[8];
#pragma unroll
            for(int j = 0; j < 4; j++) q[2 * j] = fmaxf((float)(ww[j] & 0xffffu), 1.0f), q[2 * j + 1] = fmaxf((float)(ww[j] >> 16), 1.0f);
            *reinterpret_cast<float4 *>(dq) = make_float4(q[0], q[1], q[2], q[3]);
            *reinterpret_cast<float4 *>(dq + 32) = make_float4(q[4], q[5], q[6], q[7]);
            *reinterpret_cast<float4 *>(drq) = make_float4(quant_rcp_fast(q[0]), quant_rcp_fast(q[1]), quant_rcp_fast(q[2]), quant_rcp_fast(q[3]));
            *reinterpret_cast<float4 *>(drq + 32) = make_float4(quant_rcp_fast(q[4]), quant_rcp_fast(q[5]), quant_rcp_fast(q[6]), quant_rcp_fast(q[7]));
        }
        const unsigned b = __ballot_sync(0xffffffffu, eq); // 8 lanes per image, 4 images per warp
        if((threadIdx.x & 31) == 0 && threadIdx.x < 128) {
            unsigned m = 0;
#pragma unroll
            for(int g = 0; g < 4; g++)
                if(((b >> (8 * g)) & 0xffu) == 0xffu) m |= 1u << ((threadIdx.x >> 5) * 4 + g);
            s_same[threadIdx.x >> 5] = m;
        }
    }

    const uint32_t  e = __ldg(p.drop.list_simple + tile * 32 + t);
    const bool      valid = e != 0xffffffffu;
    const DropComp &dc = p.drop.comp[c];
    const size_t    bi = valid ? (size_t)entry_row(e) * dc.wb + entry_col(e) : 0;
    const uint32_t  meta = valid ? __ldg(dc.meta + bi) : 0u;
    const bool      opaque = meta_cls(meta) == CLS_OPAQUE;
    const float     w4 = uniform_w4(meta_wdc(meta));
    const int       row = p.block_y * dc.vs + entry_row(e), col = p.block_x * dc.hs + entry_col(e);
    F2              D[4];
    {
        const Row8 dr = valid ? ld_row_keep(dc.D + bi * 64 + r * 8) : Row8{{0u, 0u, 0u, 0u}};
#pragma unroll
        for(int j = 0; j < 4; j++) D[j] = f2((float)row_get(dr, 2 * j), (float)row_get(dr, 2 * j + 1));
    }
    // lane k of every warp fetches plane pointer / stride / rows of image k of the chunk; the per-image values are
    // then broadcast by shuffle, so no global load sits on the per-image path (it matters when only one small CTA of
    // this kernel is resident per SM, beside the G kernel)
    const int          lane = threadIdx.x & 31;
    unsigned long long d_plane = 0;
    int                d_stride = 0, d_rows = 0;
    if(lane < ni) {
        const mjx_image_desc_t &im = p.items[i0 + lane];
        d_plane = im.plane[c];
        d_stride = im.stride_blocks[c];
        d_rows = im.rows[c];
    }
    __syncthreads();
    const unsigned same = s_same[0] | s_same[1] | s_same[2] | s_same[3];
    // likewise for the plane geometry: images of one size share the block's byte offset, only the plane base changes
    const int      p_stride = __shfl_up_sync(0xffffffffu, d_stride, 1), p_rows = __shfl_up_sync(0xffffffffu, d_rows, 1);
    const unsigned gsame = __ballot_sync(0xffffffffu, lane > 0 && lane < ni && p_stride == d_stride && p_rows == d_rows);

    Row8   out = {{0u, 0u, 0u, 0u}};
    bool   have = false; // `out` is trunc(D / q) for the table of the image before this one
    bool   on = false;   // the block lies on the image (for the current geometry)
    size_t off = 0;      // byte offset of the lane's row in the plane (for the current geometry)
    for(int k = 0; k < ni; k++) {
        const unsigned long long plane = __shfl_sync(0xffffffffu, d_plane, k);
        if(!((gsame >> k) & 1u)) { // warp-uniform
            const int stride = __shfl_sync(0xffffffffu, d_stride, k), rows = __shfl_sync(0xffffffffu, d_rows, k);
            on = valid && row < rows && col < stride;
            off = (((size_t)row * stride + col) * 64 + r * 8) * sizeof(int16_t);
        }
        have = have && ((same >> k) & 1u);
        if(!on) continue;
        int16_t *ip = reinterpret_cast<int16_t *>(plane + off);
        if(opaque) { // trunc(D / q), the image block is not read; same table as before: same row again
            if(!have) {
                const float4 ra = *reinterpret_cast<const float4 *>(&s_rq[k][r * 4]), rb = *reinterpret_cast<const float4 *>(&s_rq[k][32 + r * 4]);
                out.w[0] = tdiv_pair(D[0], f2(ra.x, ra.y));
                out.w[1] = tdiv_pair(D[1], f2(ra.z, ra.w));
                out.w[2] = tdiv_pair(D[2], f2(rb.x, rb.y));
                out.w[3] = tdiv_pair(D[3], f2(rb.z, rb.w));
                have = true;
            }
        }
        else {
            const float4 ra = *reinterpret_cast<const float4 *>(&s_rq[k][r * 4]), rb = *reinterpret_cast<const float4 *>(&s_rq[k][32 + r * 4]);
            const Row8   in = ld_row_stream(ip);
            const float4 qa = *reinterpret_cast<const float4 *>(&s_q[k][r * 4]), qb = *reinterpret_cast<const float4 *>(&s_q[k][32 + r * 4]);
            out.w[0] = uniform_pair(f2((float)row_get(in, 0), (float)row_get(in, 1)), D[0], f2(qa.x, qa.y), f2(ra.x, ra.y), w4);
            out.w[1] = uniform_pair(f2((float)row_get(in, 2), (float)row_get(in, 3)), D[1], f2(qa.z, qa.w), f2(ra.z, ra.w), w4);
            out.w[2] = uniform_pair(f2((float)row_get(in, 4), (float)row_get(in, 5)), D[2], f2(qb.x, qb.y), f2(rb.x, rb.y), w4);
            out.w[3] = uniform_pair(f2((float)row_get(in, 6), (float)row_get(in, 7)), D[3], f2(qb.z, qb.w), f2(rb.z, rb.w), w4);
        }
        st_row_stream(ip, out); // evict-first: measured 5 % faster per step than default-policy stores
    }
}

// =========================================================================================
// fast path, kernel 2: G blocks -- thread per block, packed fp32 (FADD2 / FMUL2 / FFMA2)
// =========================================================================================
//
// Work item = (tile of 32 list entries of ONE component, chunk of images).  A CTA of 4 warps
// claims items from an atomic counter; the tile's A and Ds (2 x 8 KB) are staged once per item
// in shared memory and shared by the 4 warps, which split the chunk's images between them.
// Per warp and image: cp.async the 32 image blocks (4 KB) + the raw quantisation table (128 B)
// of the NEXT image while the current one is computed; lane t owns block t of the tile:
//
//   load     x = Ds - I * (q * prescale)                        pairs P: (row r; cols 2j, 2j+1)
//   IDCT     columns on P (packed), last butterfly stage scalar -> pairs Q: (rows 2i, 2i+1; col k)
//            rows on Q (packed)
//   blend    x *= A                                             A stored Q-paired by K1
//   FDCT     rows on Q, last stage scalar -> P; columns on P (packed)
//   requant  requant_pair() on P, int16 results written over the staged block, then the warp
//            writes the 32 blocks back with coalesced 128-bit stores.
//
// The two pairings make every 1-D pass a pure SIMD2 computation; switching pairing costs 8 extra
// scalar adds per pass instead of a register transpose.  ~1.5 k issue slots per block (the
// scalar fp32 version needed ~2.3 k); measured balance and what limits it: DESIGN.md 4.2.

// kRedo: second pass behind the tensor-core kernel -- only the blocks whose bit is set in p.redo_mask are processed
// (coefficients outside the baseline range, images with other quantisation tables: what that kernel left alone);
// exits at once when the count is zero.
template <int kGWarps, int kMinCtas, bool kRedo>
__global__ void __launch_bounds__(kGWarps * 32, kMinCtas) k2_generic_kernel(const FastParams p) {
    constexpr int kGThreads = kGWarps * 32;
    if(kRedo && *reinterpret_cast<volatile const unsigned int *>(p.redo_count) == 0u) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ int s_item;
    const int      lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // per-lane bases; everything below is base + immediate
    const unsigned char *myA = smem_raw + lane * kF32Stride, *myD = smem_raw + kTileHalf + lane * kF32Stride;
    unsigned char *ws = smem_raw + kTileBytes + warp * kWarpBytes;
    float         *tab = reinterpret_cast<float *>(ws + kGStages * kStageBytes);
    const unsigned ws32 = smem_u32(ws), tile32 = smem_u32(smem_raw);

    const int ntiles = p.drop.n_generic >> 5; // the list is padded to whole single-component tiles
    const int nchunks = (p.n + p.images_per_item - 1) / p.images_per_item;
    const int nitems = ntiles * nchunks;
    // IDCT prescale of the two table entries (2*lane, 2*lane + 1) this lane converts per image
    const float pre0 = c_inv_scale[lane >> 2] * c_inv_scale[(2 * lane) & 7], pre1 = c_inv_scale[lane >> 2] * c_inv_scale[((2 * lane) & 7) + 1];

    int  cur_tile = -1;
    int  my_c = 0, my_row = 0, my_col = 0, tile_c = 0;
    bool my_valid = false;

    for(;;) {
        __syncthreads(); // every warp is done with the previous item's tile (and with s_item)
        if(threadIdx.x == 0) s_item = (int)atomicAdd(p.counter + (kRedo ? 2 : 0), 1u);
        __syncthreads();
        const int item = s_item;
        if(item >= nitems) break;
        const int  tile = item / nchunks, chunk = item - tile * nchunks;
        const bool new_tile = tile != cur_tile;
        if(new_tile) {
            cur_tile = tile;
            const uint32_t e = __ldg(p.drop.list_generic + tile * 32 + lane);
            my_valid = e != 0xffffffffu;
            if(my_valid) {
                const DropComp &dc = p.drop.comp[entry_comp(e)];
                my_c = entry_comp(e);
                my_row = p.block_y * dc.vs + entry_row(e);
                my_col = p.block_x * dc.hs + entry_col(e);
            }
            tile_c = 0; // tiles hold one component only (the list is padded per component)
#pragma unroll
            for(int c = 1; c < MJX_MAX_COMPONENTS; c++)
                if(c < p.drop.ncomp && tile >= p.drop.gtile_start[c]) tile_c = c;
            const float *ga = p.drop.gA + (size_t)tile * 32 * 64, *gd = p.drop.gDs + (size_t)tile * 32 * 64;
            // 512 chunks of 16 B per array; thread -> chunk (tid & 15) of blocks (tid >> 4) + 8j
            const unsigned tdst = tile32 + (threadIdx.x >> 4) * kF32Stride + (threadIdx.x & 15) * 16;
#pragma unroll
            for(int j = 0; j < 512 / kGThreads; j++) {
                cp_async16(tdst + j * (kGThreads / 16) * kF32Stride, ga + (j * kGThreads + threadIdx.x) * 4);
                cp_async16(tdst + kTileHalf + j * (kGThreads / 16) * kF32Stride, gd + (j * kGThreads + threadIdx.x) * 4);
            }
        }

        const int i0 = chunk * p.images_per_item + warp, i1 = min(p.n, (chunk + 1) * p.images_per_item);

        // issue the loads of image `img` into stage `st` (addresses of the 32 blocks first)
        // this warp's images of the item are i0 + k * kGWarps: lane k fetches image k's plane pointer,
        // stride and height once per item; the per-image values are then broadcast by shuffle, so no
        // global load sits on the per-image path
        unsigned long long d_plane = 0;
        int                d_stride = 0, d_rows = 0;
        if(i0 + lane * kGWarps < i1) {
            const mjx_image_desc_t &im = p.items[i0 + lane * kGWarps];
            d_plane = im.plane[tile_c];
            d_stride = im.stride_blocks[tile_c];
            d_rows = im.rows[tile_c];
        }

        // issue the loads of the warp's k-th image into stage `st` (addresses of the 32 blocks first)
        auto prefetch = [&](int k, int st) {
            const unsigned long long plane = __shfl_sync(0xffffffffu, d_plane, k);
            const int                stride = __shfl_sync(0xffffffffu, d_stride, k), rows = __shfl_sync(0xffffffffu, d_rows, k);
            unsigned char           *sb = ws + st * kStageBytes;
            unsigned long long      *addr = reinterpret_cast<unsigned long long *>(sb + kInBytes + kQRawBytes);
            unsigned long long       a = 0;
            if(my_valid && my_row < rows && my_col < stride) a = plane + ((unsigned long long)my_row * stride + my_col) * 128ull;
            if(kRedo) {
                const unsigned m = __ldg(p.redo_mask + (size_t)cur_tile * p.n + (i0 + k * kGWarps));
                if(!((m >> lane) & 1u)) a = 0;
            }
            addr[lane] = a;
            const bool all_there = __all_sync(0xffffffffu, a != 0); // the usual case: every block of the tile lies on this image
            __syncwarp();
            // lane -> chunk (lane & 7) of blocks (lane >> 3) + 4j
            const unsigned            dst = ws32 + st * kStageBytes + (lane >> 3) * kInStride + (lane & 7) * 16;
            const unsigned long long *ap = addr + (lane >> 3);
            const unsigned            coff = (lane & 7) * 16;
            unsigned long long        b[8];
#pragma unroll
            for(int j = 0; j < 8; j++) b[j] = ap[4 * j];
            if(all_there) {
#pragma unroll
                for(int j = 0; j < 8; j++) cp_async16(dst + j * 4 * kInStride, reinterpret_cast<const void *>(b[j] + coff));
            }
            else {
#pragma unroll
                for(int j = 0; j < 8; j++)
                    cp_async16(dst + j * 4 * kInStride, reinterpret_cast<const void *>((b[j] ? b[j] : (unsigned long long)(uintptr_t)p.items) + coff), b[j] ? 16u : 0u);
            }
            if(lane < 8) cp_async16(ws32 + st * kStageBytes + kInBytes + lane * 16, reinterpret_cast<const char *>(&p.items[i0 + k * kGWarps].q[tile_c][0]) + lane * 16);
        };

        if(i0 < i1) prefetch(0, 0);
        cp_async_commit();
        if(new_tile) {
            cp_async_wait<0>();
            __syncthreads(); // the tile (copied by all threads) is visible to all
        }
        for(int img = i0, it = 0; img < i1; img += kGWarps, it++) {
            const int st = it & 1;
            if(img + kGWarps < i1) prefetch(it + 1, st ^ 1);
            cp_async_commit();
            cp_async_wait<1>(); // everything but the newest group: image `img` has landed
            __syncwarp();

            unsigned char            *sb = ws + st * kStageBytes;
            unsigned char            *my_in = sb + lane * kInStride;
            const unsigned long long *addr = reinterpret_cast<const unsigned long long *>(sb + kInBytes + kQRawBytes);
            {   // float tables of this image: entries 2*lane, 2*lane + 1
                const uint32_t qw = *reinterpret_cast<const uint32_t *>(sb + kInBytes + lane * 4);
                const float    q0 = fmaxf((float)(qw & 0xffffu), 1.0f), q1 = fmaxf((float)(qw >> 16), 1.0f);
                *reinterpret_cast<float2 *>(tab + 2 * lane) = make_float2(q0 * pre0, q1 * pre1);
                *reinterpret_cast<float2 *>(tab + 64 + 2 * lane) = make_float2(q0, q1);
                *reinterpret_cast<float2 *>(tab + 128 + 2 * lane) = make_float2(quant_rcp_fast(q0), quant_rcp_fast(q1));
            }
            __syncwarp();

            if(addr[lane] != 0) {
                F2 x[32], y[32];
#pragma unroll
                for(int r = 0; r < 8; r++) {
                    const uint4  w = *reinterpret_cast<const uint4 *>(my_in + r * 16);
                    const float4 d0 = *reinterpret_cast<const float4 *>(myD + r * 32);
                    const float4 d1 = *reinterpret_cast<const float4 *>(myD + r * 32 + 16);
                    const float4 s0 = *reinterpret_cast<const float4 *>(tab + r * 8);
                    const float4 s1 = *reinterpret_cast<const float4 *>(tab + r * 8 + 4);
                    // (D - I*q) * prescale
                    x[4 * r + 0] = fma2(s16pair(w.x), f2(-s0.x, -s0.y), f2(d0.x, d0.y));
                    x[4 * r + 1] = fma2(s16pair(w.y), f2(-s0.z, -s0.w), f2(d0.z, d0.w));
                    x[4 * r + 2] = fma2(s16pair(w.z), f2(-s1.x, -s1.y), f2(d1.x, d1.y));
                    x[4 * r + 3] = fma2(s16pair(w.w), f2(-s1.z, -s1.w), f2(d1.z, d1.w));
                }
#pragma unroll
                for(int j = 0; j < 4; j++) idct8p_cols_to_rowpairs(x, y, j);
#pragma unroll
                for(int i = 0; i < 4; i++) idct8p<1>(y + 8 * i);
#pragma unroll
                for(int c = 0; c < 16; c++) {
                    const float4 a = *reinterpret_cast<const float4 *>(myA + c * 16);
                    y[2 * c] = mul2(y[2 * c], f2(a.x, a.y));
                    y[2 * c + 1] = mul2(y[2 * c + 1], f2(a.z, a.w));
                }
#pragma unroll
                for(int i = 0; i < 4; i++) fdct8p_rowpairs_to_cols(y, x, i);
#pragma unroll
                for(int j = 0; j < 4; j++) fdct8p<4>(x + j);
#pragma unroll
                for(int r = 0; r < 8; r++) {
                    const uint4  w = *reinterpret_cast<const uint4 *>(my_in + r * 16);
                    const float4 q0 = *reinterpret_cast<const float4 *>(tab + 64 + r * 8);
                    const float4 q1 = *reinterpret_cast<const float4 *>(tab + 64 + r * 8 + 4);
                    const float4 r0 = *reinterpret_cast<const float4 *>(tab + 128 + r * 8);
                    const float4 r1 = *reinterpret_cast<const float4 *>(tab + 128 + r * 8 + 4);
                    uint4        o;
                    o.x = requant_pair(x[4 * r + 0], c_fwd2.v[4 * r + 0], s16pair(w.x), f2(q0.x, q0.y), f2(r0.x, r0.y));
                    o.y = requant_pair(x[4 * r + 1], c_fwd2.v[4 * r + 1], s16pair(w.y), f2(q0.z, q0.w), f2(r0.z, r0.w));
                    o.z = requant_pair(x[4 * r + 2], c_fwd2.v[4 * r + 2], s16pair(w.z), f2(q1.x, q1.y), f2(r1.x, r1.y));
                    o.w = requant_pair(x[4 * r + 3], c_fwd2.v[4 * r + 3], s16pair(w.w), f2(q1.z, q1.w), f2(r1.z, r1.w));
                    *reinterpret_cast<uint4 *>(my_in + r * 16) = o;
                }
            }
            __syncwarp();
            // coalesced write-back: 8 lanes per block, lane -> chunk (lane & 7) of blocks (lane >> 3) + 4j
            {
                const unsigned char      *src = sb + (lane >> 3) * kInStride + (lane & 7) * 16;
                const unsigned long long *ap = addr + (lane >> 3);
                const unsigned            coff = (lane & 7) * 16;
                unsigned long long        b[8];
                uint4              v[8];
#pragma unroll
                for(int j = 0; j < 8; j++) {
                    b[j] = ap[4 * j];
                    v[j] = *reinterpret_cast<const uint4 *>(src + j * 4 * kInStride);
                }
#pragma unroll
                for(int j = 0; j < 8; j++)
                    if(b[j]) __stcs(reinterpret_cast<uint4 *>(b[j] + coff), v[j]);
            }
            __syncwarp();
        }
        cp_async_wait<0>();
    }
}

// =========================================================================================
// self-test: the MUFU-based reciprocal tables divide exactly
// =========================================================================================

// one CTA per quantiser value q; its threads sweep a = -2^17 .. 2^17 in pairs through tdiv_pair()
__global__ void __launch_bounds__(256) selftest_reciprocal_kernel(unsigned long long *mismatches, int q_first) {
    const int   q = q_first + blockIdx.x;
    const float rq = quant_rcp_fast((float)q);
    unsigned    bad = 0;
    for(int a = -(1 << 17) + 2 * (int)threadIdx.x; a <= (1 << 17); a += 2 * 256) {
        const int      b = a + 1 <= (1 << 17) ? a + 1 : a;
        const uint32_t pk = tdiv_pair(f2((float)a, (float)b), f2(rq, rq));
        bad += (int16_t)(pk & 0xffffu) != (int16_t)(a / q);
        bad += (int16_t)(pk >> 16) != (int16_t)(b / q);
    }
    if(bad) atomicAdd(mismatches, (unsigned long long)bad);
}

cudaError_t launch_selftest_reciprocal(cudaStream_t s, unsigned long long *mismatches_dev) {
    for(int q0 = 1; q0 <= 65535; q0 += 32768) {
        const int cnt = 65535 - q0 + 1 < 32768 ? 65535 - q0 + 1 : 32768;
        selftest_reciprocal_kernel<<<cnt, 256, 0, s>>>(mismatches_dev, q0);
        cudaError_t e = cudaGetLastError();
        if(e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// =========================================================================================
// launcher
// =========================================================================================

static size_t redo_mask_bytes(int n, const DropView &view) { return ((size_t)(view.n_generic / 32) * (size_t)(n > 0 ? n : 0) * 4 + 255) / 256 * 256; }
// [0] work counter of the G kernel, [1] redo count, [2] work counter of the redo pass; from byte 256 on, for batches the
// tensor-core kernel may serve: the redo mask (one word per tile and image), then its address table (16 bytes per
// component and image)
size_t k2_scratch_bytes(int n, const DropView &view) {
    if(n < kOpMinImages || view.n_generic == 0) return 256;
    return 256 + redo_mask_bytes(n, view) + (size_t)MJX_MAX_COMPONENTS * n * 16;
}

cudaError_t launch_k2(const K2Launch &L, const mjx_image_desc_t *items_dev, int n, const DropView &view, int block_x, int block_y) {
    if(n <= 0 || view.total_blocks <= 0) return cudaSuccess;
    cudaStream_t s = L.stream;
    int         *launches = L.launches;
    cudaError_t  e;
    if(L.strict) {
        StrictParams p;
        p.drop = view;
        p.block_x = block_x;
        p.block_y = block_y;
        const unsigned gx = (unsigned)((view.total_blocks + kBlocksPerCta - 1) / kBlocksPerCta);
        for(int first = 0; first < n; first += 65535) {
            const int cnt = n - first < 65535 ? n - first : 65535;
            p.items = items_dev + first;
            k2_strict_kernel<<<dim3(gx, (unsigned)cnt), kThreads, 0, s>>>(p);
            if((e = cudaGetLastError()) != cudaSuccess) return e;
            if(launches) (*launches)++;
        }
        return cudaSuccess;
    }
    if(view.n_simple == 0 && view.n_generic == 0) return cudaSuccess;

    // fp32 generic kernel: 4 warps per CTA, 3 CTAs per SM (142 registers, 60 KB shared memory).  8 warps x 2 CTAs at 128
    // registers was measured slower (1.89 vs 1.72 ms): the spills cost more than the extra warps hide.
    constexpr int g_warps = 4;
    K2Dev         local_dev;
    K2Dev        *dev = L.dev ? L.dev : &local_dev;
    if(!dev->g_attr) { // function attributes are per device: cached in the ctx, not in a process-wide static
        if((e = cudaFuncSetAttribute(k2_generic_kernel<g_warps, 3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem(g_warps))) != cudaSuccess) return e;
        if((e = cudaFuncSetAttribute(k2_generic_kernel<g_warps, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem(g_warps))) != cudaSuccess) return e;
        int occ = 0;
        if((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k2_generic_kernel<g_warps, 3, false>, g_warps * 32, g_smem(g_warps))) != cudaSuccess) return e;
        dev->g_ctas_per_sm = occ > 0 ? occ : 1;
        dev->g_attr = true;
    }
    const int ctas_per_sm = dev->g_ctas_per_sm;
    const bool use_op = L.tc != 0 && L.op != nullptr && n >= kOpMinImages && view.n_generic > 0 && (L.class_mask & 2);
    FastParams p;
    p.drop = view;
    p.items = items_dev;
    p.counter = reinterpret_cast<unsigned int *>(L.scratch);
    p.redo_count = p.counter + 1;
    p.redo_mask = p.counter + 64;
    p.n = n;
    p.block_x = block_x;
    p.block_y = block_y;
    p.images_per_item = n < 24 * g_warps ? n : 24 * g_warps; // per warp: <= 32 (one descriptor per lane); 16..32 measured within 1.5 %

    // Both classes present and a batch large enough to fill the machine: the two kernels run side by side.  The G kernel
    // is launched first; the OPAQUE/U kernel (write bound) follows on the low-priority side stream in its 128-thread
    // shape, one CTA of which fits into the registers the G kernel leaves free, and fills the idle issue slots and HBM
    // bandwidth.
    const bool both = view.n_simple > 0 && view.n_generic > 0 && (L.class_mask & 3) == 3;
    const bool overlap = both && L.side && L.side->stream && n >= 64;
    auto       launch_simple = [&](cudaStream_t st, bool small) -> cudaError_t {
        const unsigned tiles = (unsigned)(view.n_simple / 32);
        for(int first = 0; first < n; first += 65535 * kSimpleImages) {
            const int cnt = n - first < 65535 * kSimpleImages ? n - first : 65535 * kSimpleImages;
            FastParams q = p;
            q.items = items_dev + first;
            q.n = cnt;
            const unsigned gy = (unsigned)((cnt + kSimpleImages - 1) / kSimpleImages);
            if(small) k2_simple_kernel<128><<<dim3(tiles * 2, gy), 128, 0, st>>>(q);
            else k2_simple_kernel<kThreads><<<dim3(tiles, gy), kThreads, 0, st>>>(q);
            const cudaError_t le = cudaGetLastError();
            if(le != cudaSuccess) return le;
            if(launches) (*launches)++;
        }
        return cudaSuccess;
    };
    auto launch_generic = [&]() -> cudaError_t {
        cudaError_t ge;
        const int   sms = L.sm_count > 0 ? L.sm_count : 148;
        if(use_op) {
            // tensor-core kernel; the fp32 kernel follows as a redo pass over what that kernel left alone (coefficients
            // outside the baseline range when checked, images with other quantisation tables) and exits at once when
            // there is nothing
            if((ge = cudaMemsetAsync(p.counter, 0, 256 + redo_mask_bytes(n, view), s)) != cudaSuccess) return ge;
            OpParams op;
            op.drop = view;
            op.op = *L.op;
            op.items = items_dev;
            op.table = reinterpret_cast<uint4 *>(reinterpret_cast<char *>(L.scratch) + 256 + redo_mask_bytes(n, view));
            op.redo_mask = p.redo_mask;
            op.redo_count = p.redo_count;
            op.n = n;
            op.block_x = block_x;
            op.block_y = block_y;
            if((ge = launch_k2_generic_op(s, op, sms, L.tc == 1, dev->op_attr, launches)) != cudaSuccess) return ge;
            const long long nitems = (long long)(view.n_generic / 32) * ((n + p.images_per_item - 1) / p.images_per_item);
            if(nitems > 0x7fffffffLL) return cudaErrorInvalidValue;
            const int ctas = nitems < (long long)sms * ctas_per_sm ? (int)nitems : sms * ctas_per_sm;
            k2_generic_kernel<g_warps, 3, true><<<ctas, g_warps * 32, g_smem(g_warps), s>>>(p);
            if((ge = cudaGetLastError()) != cudaSuccess) return ge;
            if(launches) (*launches)++;
            return cudaSuccess;
        }
        if((ge = cudaMemsetAsync(p.counter, 0, sizeof(unsigned int), s)) != cudaSuccess) return ge;
        const long long nitems = (long long)(view.n_generic / 32) * ((n + p.images_per_item - 1) / p.images_per_item);
        if(nitems > 0x7fffffffLL) return cudaErrorInvalidValue;
        const int ctas = nitems < (long long)sms * ctas_per_sm ? (int)nitems : sms * ctas_per_sm;
        k2_generic_kernel<g_warps, 3, false><<<ctas, g_warps * 32, g_smem(g_warps), s>>>(p);
        if((ge = cudaGetLastError()) != cudaSuccess) return ge;
        if(launches) (*launches)++;
        return cudaSuccess;
    };
    if(overlap) {
        if((e = cudaEventRecord(L.side->fork, s)) != cudaSuccess) return e;
        if((e = cudaStreamWaitEvent(L.side->stream, L.side->fork, 0)) != cudaSuccess) return e;
        if((e = launch_generic()) != cudaSuccess) return e;
        if((e = launch_simple(L.side->stream, true)) != cudaSuccess) return e;
        if((e = cudaEventRecord(L.side->join, L.side->stream)) != cudaSuccess) return e;
        if((e = cudaStreamWaitEvent(s, L.side->join, 0)) != cudaSuccess) return e;
        return cudaSuccess;
    }
    if(view.n_simple > 0 && (L.class_mask & 1))
        if((e = launch_simple(s, false)) != cudaSuccess) return e;
    if(view.n_generic > 0 && (L.class_mask & 2))
        if((e = launch_generic()) != cudaSuccess) return e;
    return cudaSuccess;
}

} // namespace mjx
