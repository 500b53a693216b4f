// k4_huffman.cu -- K4: Huffman coding of a baseline sequential scan on the device.
// Replaces, for the reference's mj_write_jpeg_to_memory (src/image.c:120-209), the entropy encoder that
// jpeg_write_coefficients / jpeg_finish_compress run on the host (libjpeg jctrans.c compress_output + jchuff.c
// encode_mcu_huff / encode_one_block): non-optimised tables, no restart markers, one scan with every component.
//
// The serial part of Huffman coding is only WHERE a block's bits go; what they are depends on the block and on the DC of
// its predecessor, which lies in the coefficient planes and not in the bit stream.  So:
//   k4_block_kernel<false>  thread per block in scan order (MCU by MCU; inside an MCU component by component, rows,
//                           columns): number of bits of the block -- DC difference category + bits, run/size symbols,
//                           ZRL, EOB
//   k4_scan_kernel          exclusive prefix sum per image (one CTA per image): bit offset of every block, total
//   k4_block_kernel<true>   thread per block: the same walk again, codes shifted into a 64-bit accumulator and written as
//                           32-bit words at the block's offset (first and last word by atomicOr: neighbours share them)
//   k4_stuff_kernel<false>  thread per 64 bytes of the stream: bytes equal to 0xFF (each needs a stuffed 0x00 behind it)
//   k4_scan_kernel          again: output position of every 64-byte piece
//   k4_stuff_kernel<true>   the bytes, stuffed, at their final place; the last byte padded with 1-bits (jchuff.c flush_bits)
// A warp fetches its 32 blocks in whole 128-byte lines and hands them to their threads through shared memory; the scans move
// 128 bits per access.
// Blocks past a component's real width / height (the MCU grid is rounded up) are the encoder's dummy blocks: no AC, DC
// equal to the previous block's, i.e. a zero difference (jctrans.c compress_output) -- whatever the plane holds there.
// The result is byte-identical to libjpeg's entropy-coded segment (tests/test_gpu_huffman.py compares whole files with
// mj_write_jpeg_to_memory on the host path).  A coefficient the baseline tables cannot code (DC difference beyond 11
// bits, AC beyond 10 -- libjpeg raises JERR_BAD_DCT_COEF), a symbol without a code, or a segment larger than the
// caller's buffer gives size 0xFFFFFFFF for that image: the caller lets libjpeg do it (and raise its error).
#include <stdlib.h>
#include <string.h>

#include "mjx_internal.cuh"

namespace mjx {

static constexpr int kHuffThreads = 256;
static constexpr int kScanThreads = 1024;
static constexpr int kStuffChunk = 64; // stream bytes per thread of the stuffing kernels

struct HuffParams {
    const mjx_image_desc_t *items;
    int                     n;
    int                     ncomp, blocks_per_mcu, mcus_per_row, mcu_rows, nblk;
    int                     h[MJX_MAX_COMPONENTS], v[MJX_MAX_COMPONENTS];
    int                     dc_tbl[MJX_MAX_COMPONENTS], ac_tbl[MJX_MAX_COMPONENTS];
    signed char             bcomp[16], bidx[16]; // per block of the MCU: its component, its index among that component's blocks
    const uint32_t         *tables;              // [8][256] code | length << 16: 0..3 DC, 4..7 AC
    uint32_t               *len;                 // [n][len_stride] bits per block, then bit offsets
    size_t                  len_stride;
    uint32_t               *total_bits;          // [n]
    uint32_t               *bits;                // [n][bits_stride] the stream as big-endian-in-register words
    size_t                  bits_stride;
    uint32_t               *ff;                  // [n][ff_stride] 0xFF bytes per piece, then their prefix sum
    size_t                  ff_stride;
    uint32_t               *total_ff;            // [n]
    uint32_t               *status;              // [n] 0: fine; bit 0 coefficient / symbol not codable, bit 1 does not fit
    unsigned char          *out;
    size_t                  out_stride;
    uint32_t               *sizes;
};

// where block `t` of the scan lies
struct BlockPos {
    int  c, row, col;
    int  mrow, mcol, k; // MCU position, index among the component's blocks of the MCU (row-major)
    bool real;
};

__device__ __forceinline__ BlockPos block_pos(const HuffParams &p, const mjx_image_desc_t &im, int t) {
    BlockPos  b;
    const int mcu = t / p.blocks_per_mcu, bi = t - mcu * p.blocks_per_mcu;
    b.c = p.bcomp[bi];
    b.k = p.bidx[bi];
    b.mrow = mcu / p.mcus_per_row;
    b.mcol = mcu - b.mrow * p.mcus_per_row;
    const int h = p.h[b.c];
    b.row = b.mrow * p.v[b.c] + b.k / h;
    b.col = b.mcol * h + b.k % h;
    b.real = b.row < im.hreal[b.c] && b.col < im.wreal[b.c];
    return b;
}

__device__ __forceinline__ const int16_t *block_ptr(const mjx_image_desc_t &im, int c, int row, int col) {
    return reinterpret_cast<const int16_t *>(im.plane[c]) + ((size_t)row * im.stride_blocks[c] + col) * 64;
}

// DC the encoder sees for block k of component c in MCU (mrow, mcol): a dummy block repeats its predecessor's
// (block 0 of an MCU is always a real block)
__device__ __forceinline__ int effective_dc(const HuffParams &p, const mjx_image_desc_t &im, int c, int mrow, int mcol, int k) {
    const int h = p.h[c], v = p.v[c];
    for(; k > 0; k--) {
        const int row = mrow * v + k / h, col = mcol * h + k % h;
        if(row < im.hreal[c] && col < im.wreal[c]) return (int)__ldg(block_ptr(im, c, row, col));
    }
    return (int)__ldg(block_ptr(im, c, mrow * v, mcol * h));
}

// DC of the block coded before block b in its component (0 at the start of the scan: no restart intervals)
__device__ __forceinline__ int predecessor_dc(const HuffParams &p, const mjx_image_desc_t &im, const BlockPos &b) {
    if(b.k > 0) return effective_dc(p, im, b.c, b.mrow, b.mcol, b.k - 1);
    if(b.mcol > 0) return effective_dc(p, im, b.c, b.mrow, b.mcol - 1, p.h[b.c] * p.v[b.c] - 1);
    if(b.mrow > 0) return effective_dc(p, im, b.c, b.mrow - 1, p.mcus_per_row - 1, p.h[b.c] * p.v[b.c] - 1);
    return 0;
}

// the walk over one block (jchuff.c encode_one_block): put(code, length) for every code word, bits of the value appended
template <class Put>
__device__ __forceinline__ uint32_t code_block(const uint32_t (&w)[32], bool real, int diff, const uint32_t *dct, const uint32_t *act, Put &put) {
    uint32_t err = 0;
    {
        int t = diff, t2 = diff;
        if(t < 0) t = -t, t2--;
        const int nb = 32 - __clz(t);
        if(nb > 11) return 1u;
        const uint32_t e = dct[nb];
        if((e >> 16) == 0) return 1u;
        put(((e & 0xffffu) << nb) | ((uint32_t)t2 & ((1u << nb) - 1u)), (int)(e >> 16) + nb);
    }
    if(!real) { // dummy block: end of block at once
        const uint32_t e = act[0];
        if((e >> 16) == 0) return 1u;
        put(e & 0xffffu, (int)(e >> 16));
        return 0u;
    }
    int r = 0;
    // zigzag position -> natural index (libjpeg jutils.c jpeg_natural_order; ITU-T T.81 figure A.6); the loop is unrolled, so
    // every coefficient is a fixed half of a fixed register
#pragma unroll
    for(int k = 1; k < 64; k++) {
        constexpr unsigned char zz[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                          41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                          30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
        const int      nat = zz[k];
        const uint32_t word = w[nat >> 1];
        const int      val = (nat & 1) ? ((int)word >> 16) : (int)(short)(word & 0xffffu);
        if(val == 0) {
            r++;
            continue;
        }
        while(r > 15) {
            const uint32_t z = act[0xF0];
            if((z >> 16) == 0) err = 1u;
            put(z & 0xffffu, (int)(z >> 16));
            r -= 16;
        }
        int t = val, t2 = val;
        if(t < 0) t = -t, t2--;
        const int nb = 32 - __clz(t);
        if(nb > 10) return 1u;
        const uint32_t e = act[(r << 4) + nb];
        if((e >> 16) == 0) err = 1u;
        put(((e & 0xffffu) << nb) | ((uint32_t)t2 & ((1u << nb) - 1u)), (int)(e >> 16) + nb);
        r = 0;
    }
    if(r > 0) {
        const uint32_t e = act[0];
        if((e >> 16) == 0) err = 1u;
        put(e & 0xffffu, (int)(e >> 16));
    }
    return err;
}

// the block's 64 coefficients as 32 words (natural order); a dummy block is not read
__device__ __forceinline__ void load_block(const mjx_image_desc_t &im, const BlockPos &b, uint32_t (&w)[32]) {
    if(b.real) {
        const uint4 *src = reinterpret_cast<const uint4 *>(block_ptr(im, b.c, b.row, b.col));
#pragma unroll
        for(int i = 0; i < 8; i++) {
            const uint4 x = __ldg(src + i);
            w[4 * i] = x.x, w[4 * i + 1] = x.y, w[4 * i + 2] = x.z, w[4 * i + 3] = x.w;
        }
    }
    else {
#pragma unroll
        for(int i = 0; i < 32; i++) w[i] = 0u;
    }
}

template <bool kEmit>
__global__ void __launch_bounds__(kHuffThreads) k4_block_kernel(const HuffParams p) {
    __shared__ uint32_t s_tab[8 * 256];
    // the CTA's blocks pass through shared memory: a warp fetches its 32 blocks with eight instructions that each take whole
    // 128-byte lines (8 lanes per block), then every thread picks up its own block -- a thread reading its 128 bytes from
    // global memory by itself costs 32 lines per instruction.  One 16-byte chunk of padding per block keeps both sides free of
    // bank conflicts.
    __shared__ uint4 s_blk[kHuffThreads * 9];
    for(int i = threadIdx.x; i < 8 * 256; i += kHuffThreads) s_tab[i] = __ldg(p.tables + i);
    __syncthreads();
    const int img = blockIdx.y, t = blockIdx.x * kHuffThreads + threadIdx.x;
    if(kEmit && p.status[img] != 0u) return; // (the whole CTA: one image per blockIdx.y)
    const mjx_image_desc_t &im = p.items[img];
    const bool              valid = t < p.nblk;
    BlockPos                b;
    if(valid) b = block_pos(p, im, t);
    else b.c = 0, b.row = b.col = b.mrow = b.mcol = b.k = 0, b.real = false;
    uint32_t w[32];
    {
        const int                lane = threadIdx.x & 31, wbase = threadIdx.x & ~31;
        const unsigned long long mine = b.real ? (unsigned long long)(uintptr_t)block_ptr(im, b.c, b.row, b.col) : 0ull;
#pragma unroll
        for(int i = 0; i < 8; i++) {
            const int                sl = 4 * i + (lane >> 3);
            const unsigned long long q = __shfl_sync(0xffffffffu, mine, sl);
            const uint4              x = q ? __ldg(reinterpret_cast<const uint4 *>(q) + (lane & 7)) : make_uint4(0u, 0u, 0u, 0u);
            s_blk[(wbase + sl) * 9 + (lane & 7)] = x;
        }
        __syncwarp();
#pragma unroll
        for(int i = 0; i < 8; i++) {
            const uint4 x = s_blk[threadIdx.x * 9 + i];
            w[4 * i] = x.x, w[4 * i + 1] = x.y, w[4 * i + 2] = x.z, w[4 * i + 3] = x.w;
        }
    }
    if(!valid) return;
    int diff = 0;
    if(b.real) diff = (int)(short)(w[0] & 0xffffu) - predecessor_dc(p, im, b);
    const uint32_t *dct = s_tab + 256 * p.dc_tbl[b.c], *act = s_tab + 256 * (4 + p.ac_tbl[b.c]);
    if(!kEmit) {
        uint32_t bits = 0;
        auto     put = [&](uint32_t, int len) { bits += (uint32_t)len; };
        const uint32_t err = code_block(w, b.real, diff, dct, act, put);
        p.len[(size_t)img * p.len_stride + t] = err ? 0u : bits;
        if(err) atomicOr(p.status + img, 1u);
    }
    else {
        const uint32_t off = p.len[(size_t)img * p.len_stride + t];
        uint32_t      *wp = p.bits + (size_t)img * p.bits_stride + (off >> 5);
        uint64_t       acc = 0;
        int            nacc = (int)(off & 31u); // the word's leading bits belong to the blocks before
        bool           first = true;
        auto           put = [&](uint32_t code, int len) {
            acc = (acc << len) | code;
            nacc += len;
            if(nacc >= 32) {
                const uint32_t word = (uint32_t)(acc >> (nacc - 32));
                if(first) atomicOr(wp, word), first = false;
                else *wp = word;
                wp++;
                nacc -= 32;
                acc &= (1ull << nacc) - 1ull;
            }
        };
        code_block(w, b.real, diff, dct, act, put);
        if(nacc > 0) atomicOr(wp, (uint32_t)(acc << (32 - nacc)));
    }
}

// exclusive prefix sum of m values per image, in place; one CTA per image.  total > limit raises status bit 1.
__global__ void __launch_bounds__(kScanThreads) k4_scan_kernel(uint32_t *data, size_t stride, const uint32_t *count_src, int count_shift, int m_fixed,
                                                               uint32_t *totals, uint32_t *status, unsigned long long limit) {
    __shared__ unsigned long long s_warp[kScanThreads / 32];
    const int                     img = blockIdx.x;
    // m: either fixed, or derived from another per-image total (the stream's pieces: (bits + 7) / 8 bytes in 64-byte pieces)
    int m = m_fixed;
    if(count_src) {
        // an image that is not being coded (its stream would not fit the slab, or a coefficient cannot be coded) has no pieces
        // to count -- and its bit total may describe more pieces than its row holds
        if(status[img] != 0u) return;
        const unsigned long long nbytes = ((unsigned long long)count_src[img] + 7ull) >> 3;
        m = (int)((nbytes + (1ull << count_shift) - 1ull) >> count_shift);
    }
    if((size_t)m > stride) m = (int)stride;
    // tiles of 4 x kScanThreads values: every thread four consecutive ones (one 128-bit access each way), a CTA-wide scan of
    // the threads' sums per tile, the running total carried from tile to tile (the rows are padded to multiples of 64 values,
    // so a whole 128-bit access never leaves the row; what lies behind m counts as zero)
    uint4             *d4 = reinterpret_cast<uint4 *>(data + (size_t)img * stride);
    const int          lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long carry = 0;
    for(int base = 0; base < m; base += 4 * kScanThreads) {
        const int i0 = base + 4 * (int)threadIdx.x;
        uint4     v = make_uint4(0u, 0u, 0u, 0u);
        if(i0 < m) {
            v = d4[i0 >> 2];
            if(i0 + 1 >= m) v.y = 0u;
            if(i0 + 2 >= m) v.z = 0u;
            if(i0 + 3 >= m) v.w = 0u;
        }
        const unsigned long long s = (unsigned long long)v.x + v.y + v.z + v.w;
        unsigned long long       x = s;
#pragma unroll
        for(int o = 1; o < 32; o <<= 1) {
            const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
            if(lane >= o) x += y;
        }
        __syncthreads(); // (s_warp of the previous tile has been read)
        if(lane == 31) s_warp[warp] = x;
        __syncthreads();
        if(warp == 0) {
            unsigned long long y = s_warp[lane];
#pragma unroll
            for(int o = 1; o < 32; o <<= 1) {
                const unsigned long long z = __shfl_up_sync(0xffffffffu, y, o);
                if(lane >= o) y += z;
            }
            s_warp[lane] = y;
        }
        __syncthreads();
        const unsigned long long run = carry + x - s + (warp > 0 ? s_warp[warp - 1] : 0ull);
        if(i0 < m) d4[i0 >> 2] = make_uint4((uint32_t)run, (uint32_t)(run + v.x), (uint32_t)(run + v.x + v.y), (uint32_t)(run + v.x + v.y + v.z));
        carry += s_warp[kScanThreads / 32 - 1];
    }
    if(threadIdx.x == 0) {
        totals[img] = carry > 0xffffffffull ? 0xffffffffu : (uint32_t)carry;
        if(carry > limit) atomicOr(status + img, 2u);
    }
}

// byte i of the padded stream
__device__ __forceinline__ uint32_t stream_byte(const uint32_t *words, uint32_t i, uint32_t nbytes, uint32_t total_bits) {
    uint32_t b = (words[i >> 2] >> (24 - 8 * (i & 3u))) & 0xffu;
    if(i == nbytes - 1u && (total_bits & 7u)) b |= (1u << (8 - (total_bits & 7u))) - 1u; // 1-bits up to the byte boundary
    return b;
}

template <bool kWrite>
__global__ void __launch_bounds__(kHuffThreads) k4_stuff_kernel(const HuffParams p) {
    const int img = blockIdx.y, j = blockIdx.x * kHuffThreads + threadIdx.x;
    if(p.status[img] != 0u) {
        if(kWrite && j == 0) p.sizes[img] = 0xffffffffu;
        return;
    }
    const uint32_t total = p.total_bits[img], nbytes = (total + 7u) >> 3;
    const uint32_t lo = (uint32_t)j * kStuffChunk;
    if(kWrite) {
        const unsigned long long size = (unsigned long long)nbytes + p.total_ff[img];
        const bool               fits = size <= (unsigned long long)p.out_stride;
        if(j == 0) p.sizes[img] = fits ? (uint32_t)size : 0xffffffffu;
        if(!fits) return;
    }
    if(lo >= nbytes) return;
    const uint32_t  hi = min(nbytes, lo + kStuffChunk);
    const uint32_t *words = p.bits + (size_t)img * p.bits_stride;
    if(!kWrite) {
        uint32_t cnt = 0;
        for(uint32_t i = lo; i < hi; i++) cnt += stream_byte(words, i, nbytes, total) == 0xffu;
        p.ff[(size_t)img * p.ff_stride + j] = cnt;
    }
    else {
        unsigned char *o = p.out + (size_t)img * p.out_stride + lo + p.ff[(size_t)img * p.ff_stride + j];
        for(uint32_t i = lo; i < hi; i++) {
            const uint32_t b = stream_byte(words, i, nbytes, total);
            *o++ = (unsigned char)b;
            if(b == 0xffu) *o++ = 0;
        }
    }
}

// code and length of every symbol from the table as a DHT segment states it (ITU-T T.81 Annex C; libjpeg
// jchuff.c jpeg_make_c_derived_tbl)
static bool derive_table(const mjx_huff_table_t &t, uint32_t *out256) {
    unsigned char size[257];
    unsigned      code[257];
    int           p = 0;
    for(int l = 1; l <= 16; l++) {
        const int cnt = t.bits[l];
        if(p + cnt > 256) return false;
        for(int i = 0; i < cnt; i++) size[p++] = (unsigned char)l;
    }
    size[p] = 0;
    const int last = p;
    unsigned  c = 0;
    int       si = size[0];
    p = 0;
    while(size[p]) {
        while(size[p] == si) code[p++] = c++;
        if(c > (1u << si)) return false;
        c <<= 1;
        si++;
    }
    for(int i = 0; i < 256; i++) out256[i] = 0u;
    for(p = 0; p < last; p++) out256[t.vals[p]] = code[p] | ((uint32_t)size[p] << 16);
    return true;
}

} // namespace mjx

using namespace mjx;

// scratch of one call: per image the block lengths, the word stream, the piece counts; then totals, status, tables
struct HuffLayout {
    size_t len_stride, bits_stride, ff_stride;
    size_t o_len, o_bits, o_ff, o_tot, o_totff, o_status, o_tab, bytes;
};

static HuffLayout huff_layout(int n, int nblk, size_t cap_bytes) {
    HuffLayout L;
    L.len_stride = ((size_t)nblk + 63) / 64 * 64;
    L.bits_stride = (cap_bytes + 3) / 4 + 2; // one word of slack: the last block's tail
    L.bits_stride = (L.bits_stride + 63) / 64 * 64;
    L.ff_stride = ((cap_bytes + kStuffChunk - 1) / kStuffChunk + 63) / 64 * 64;
    size_t off = 0;
    auto   take = [&](size_t b) {
        const size_t o = off;
        off = (off + b + 255) / 256 * 256;
        return o;
    };
    L.o_len = take((size_t)n * L.len_stride * 4);
    L.o_bits = take((size_t)n * L.bits_stride * 4);
    L.o_ff = take((size_t)n * L.ff_stride * 4);
    L.o_tot = take((size_t)n * 4);
    L.o_totff = take((size_t)n * 4);
    L.o_status = take((size_t)n * 4);
    L.o_tab = take(8 * 256 * 4);
    L.bytes = off;
    return L;
}

static int ensure_huff(mjx_ctx *ctx, size_t bytes) {
    if(ctx->huff_bytes >= bytes) return MJX_OK;
    MJX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if(ctx->huff) cudaFree(ctx->huff);
    ctx->huff = nullptr;
    ctx->huff_bytes = 0;
    const size_t want = bytes + bytes / 8 + 4096;
    MJX_CUDA(ctx, cudaMalloc(&ctx->huff, want));
    ctx->huff_bytes = want;
    return MJX_OK;
}

extern "C" {

int mjx_huffman_encode_batch_device(mjx_ctx *ctx, const mjx_image_desc_t *items_dev, int n, const mjx_scan_t *scan, void *out_dev,
                                    size_t out_stride, uint32_t *sizes_dev) {
    if(!ctx) return MJX_ERR_ARG;
    MJX_CUDA(ctx, cudaSetDevice(ctx->device));
    if(!items_dev || !scan || !out_dev || !sizes_dev || n < 0 || out_stride == 0) return MJX_ERR_ARG;
    if(n == 0) return MJX_OK;
    if(scan->ncomp < 1 || scan->ncomp > MJX_MAX_COMPONENTS || scan->mcus_per_row < 1 || scan->mcu_rows < 1) return MJX_ERR_ARG;
    HuffParams p;
    memset(&p, 0, sizeof(p));
    p.items = items_dev;
    p.n = n;
    p.ncomp = scan->ncomp;
    p.mcus_per_row = scan->mcus_per_row;
    p.mcu_rows = scan->mcu_rows;
    int bpm = 0;
    for(int c = 0; c < scan->ncomp; c++) {
        // a scan of one component is not interleaved: its MCU is one block whatever the sampling factors say
        const int h = scan->ncomp == 1 ? 1 : scan->h_samp[c], v = scan->ncomp == 1 ? 1 : scan->v_samp[c];
        if(h < 1 || v < 1 || h > 4 || v > 4 || scan->dc_tbl[c] < 0 || scan->dc_tbl[c] > 3 || scan->ac_tbl[c] < 0 || scan->ac_tbl[c] > 3) return MJX_ERR_ARG;
        p.h[c] = h, p.v[c] = v;
        p.dc_tbl[c] = scan->dc_tbl[c], p.ac_tbl[c] = scan->ac_tbl[c];
        for(int k = 0; k < h * v; k++) {
            if(bpm >= 10) return MJX_ERR_UNSUPPORTED; // C_MAX_BLOCKS_IN_MCU
            p.bcomp[bpm] = (signed char)c, p.bidx[bpm] = (signed char)k;
            bpm++;
        }
    }
    p.blocks_per_mcu = bpm;
    const long long nblk = (long long)scan->mcus_per_row * scan->mcu_rows * bpm;
    if(nblk > 0x7fffffffLL / 2) return MJX_ERR_UNSUPPORTED;
    p.nblk = (int)nblk;
    if(out_stride > 0xfffffff0ull / 2) return MJX_ERR_ARG;
    // the unstuffed stream is never longer than the stuffed one: out_stride bounds both
    const HuffLayout L = huff_layout(n, p.nblk, out_stride);
    int              rv = ensure_huff(ctx, L.bytes);
    if(rv) return rv;
    char *base = (char *)ctx->huff;
    p.len = (uint32_t *)(base + L.o_len), p.len_stride = L.len_stride;
    p.bits = (uint32_t *)(base + L.o_bits), p.bits_stride = L.bits_stride;
    p.ff = (uint32_t *)(base + L.o_ff), p.ff_stride = L.ff_stride;
    p.total_bits = (uint32_t *)(base + L.o_tot);
    p.total_ff = (uint32_t *)(base + L.o_totff);
    p.status = (uint32_t *)(base + L.o_status);
    uint32_t *tab_dev = (uint32_t *)(base + L.o_tab);
    p.tables = tab_dev;
    p.out = (unsigned char *)out_dev;
    p.out_stride = out_stride;
    p.sizes = sizes_dev;

    uint32_t tab[8 * 256];
    for(int i = 0; i < 4; i++) {
        if(!derive_table(scan->dc[i], tab + 256 * i) || !derive_table(scan->ac[i], tab + 256 * (4 + i))) return MJX_ERR_ARG;
    }
    cudaStream_t s = ctx->stream;
    // (pageable source: the copy is staged by the driver before the call returns)
    MJX_CUDA(ctx, cudaMemcpyAsync(tab_dev, tab, sizeof(tab), cudaMemcpyHostToDevice, s));
    MJX_CUDA(ctx, cudaMemsetAsync(base + L.o_tot, 0, L.o_tab - L.o_tot, s)); // totals, status
    MJX_CUDA(ctx, cudaMemsetAsync(p.bits, 0, (size_t)n * L.bits_stride * 4, s));
    for(int first = 0; first < n; first += 65535) {
        const int  cnt = n - first < 65535 ? n - first : 65535;
        HuffParams q = p;
        q.items = items_dev + first;
        q.n = cnt;
        q.len += (size_t)first * L.len_stride, q.bits += (size_t)first * L.bits_stride, q.ff += (size_t)first * L.ff_stride;
        q.total_bits += first, q.total_ff += first, q.status += first, q.sizes += first;
        q.out += (size_t)first * out_stride;
        const dim3 gb((unsigned)((p.nblk + kHuffThreads - 1) / kHuffThreads), (unsigned)cnt);
        const dim3 gs((unsigned)((L.ff_stride + kHuffThreads - 1) / kHuffThreads), (unsigned)cnt);
        k4_block_kernel<false><<<gb, kHuffThreads, 0, s>>>(q);
        k4_scan_kernel<<<cnt, kScanThreads, 0, s>>>(q.len, L.len_stride, nullptr, 0, p.nblk, q.total_bits, q.status, (unsigned long long)out_stride * 8ull);
        k4_block_kernel<true><<<gb, kHuffThreads, 0, s>>>(q);
        k4_stuff_kernel<false><<<gs, kHuffThreads, 0, s>>>(q);
        k4_scan_kernel<<<cnt, kScanThreads, 0, s>>>(q.ff, L.ff_stride, q.total_bits, 6, 0, q.total_ff, q.status, (unsigned long long)out_stride);
        k4_stuff_kernel<true><<<gs, kHuffThreads, 0, s>>>(q);
        ctx->launches += 6;
        const cudaError_t e = cudaGetLastError();
        if(e != cudaSuccess) return fail(ctx, e, "k4_huffman kernels");
    }
    return MJX_OK;
}

int mjx_huffman_encode_rows_host(mjx_ctx *ctx, int ncomp, const int16_t *const *const *rows, const int *stride_blocks, const int *vrows,
                                 const int *wreal, const int *hreal, const mjx_scan_t *scan, unsigned char **out, size_t *len) {
    if(!ctx) return MJX_ERR_ARG;
    MJX_CUDA(ctx, cudaSetDevice(ctx->device));
    if(!rows || !stride_blocks || !vrows || !wreal || !hreal || !scan || !out || !len || ncomp < 1 || ncomp > MJX_MAX_COMPONENTS) return MJX_ERR_ARG;
    *out = nullptr;
    *len = 0;
    // staging: [desc][planes...] in pinned memory and on the device, then the output slab and one size word behind it
    size_t off[MJX_MAX_COMPONENTS], total = 256;
    for(int c = 0; c < ncomp; c++) {
        if(!rows[c] || stride_blocks[c] < wreal[c] || vrows[c] < hreal[c] || wreal[c] < 1 || hreal[c] < 1) return MJX_ERR_ARG;
        off[c] = total;
        total = (total + (size_t)hreal[c] * stride_blocks[c] * 128 + 255) / 256 * 256;
    }
    const size_t cap = total; // as many bytes as the coefficients themselves: beyond that libjpeg takes over
    int          rv;
    if((rv = ensure_pin(ctx, total + 256)) || (rv = ensure_dev(ctx, total + cap + 256))) return rv;
    char *pin = (char *)ctx->pin, *dev = (char *)ctx->dev;
    mjx_image_desc_t hd;
    memset(&hd, 0, sizeof(hd));
    for(int c = 0; c < ncomp; c++) {
        hd.plane[c] = (uint64_t)(uintptr_t)(dev + off[c]);
        hd.stride_blocks[c] = stride_blocks[c];
        hd.rows[c] = hreal[c];
        hd.wreal[c] = wreal[c];
        hd.hreal[c] = hreal[c];
        for(int l = 0; l < hreal[c]; l++) {
            if(!rows[c][l]) return MJX_ERR_ARG;
            memcpy(pin + off[c] + (size_t)l * stride_blocks[c] * 128, rows[c][l], (size_t)stride_blocks[c] * 128);
        }
    }
    if((rv = ensure_desc(ctx, sizeof(hd)))) return rv;
    MJX_CUDA(ctx, cudaMemcpyAsync(ctx->desc_dev, &hd, sizeof(hd), cudaMemcpyHostToDevice, ctx->stream));
    MJX_CUDA(ctx, cudaMemcpyAsync(dev + 256, pin + 256, total - 256, cudaMemcpyHostToDevice, ctx->stream));
    unsigned char *out_dev = (unsigned char *)dev + total;
    uint32_t      *size_dev = (uint32_t *)(dev + total + cap);
    rv = mjx_huffman_encode_batch_device(ctx, (const mjx_image_desc_t *)ctx->desc_dev, 1, scan, out_dev, cap, size_dev);
    if(rv) return rv;
    uint32_t size = 0;
    MJX_CUDA(ctx, cudaMemcpyAsync(&size, size_dev, 4, cudaMemcpyDeviceToHost, ctx->stream));
    MJX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if(size == 0xffffffffu) return MJX_ERR_UNSUPPORTED;
    // the segment is smaller than the planes, so it fits the pinned pool the planes were staged in
    if(size) {
        MJX_CUDA(ctx, cudaMemcpyAsync(pin, out_dev, size, cudaMemcpyDeviceToHost, ctx->stream));
        MJX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    unsigned char *buf = (unsigned char *)malloc(size ? size : 1);
    if(!buf) return MJX_ERR_MEMORY;
    memcpy(buf, pin, size);
    *out = buf;
    *len = size;
    return MJX_OK;
}

} // extern "C"
